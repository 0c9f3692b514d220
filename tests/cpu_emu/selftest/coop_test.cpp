#include <cuda_runtime.h>
struct GridBar { unsigned count, gen; };
static inline void grid_barrier(GridBar *gb) {
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned g = __atomic_load_n(&gb->gen, __ATOMIC_ACQUIRE);
		if (atomicAdd(&gb->count, 1u) == gridDim.x - 1) {
			__atomic_store_n(&gb->count, 0u, __ATOMIC_RELAXED);
			__atomic_store_n(&gb->gen, g + 1, __ATOMIC_RELEASE);
		} else {
			while (__atomic_load_n(&gb->gen, __ATOMIC_ACQUIRE) == g) __nanosleep(32);
		}
	}
	__syncthreads();
}
void k_coop(int *data, GridBar *gb, int rounds) {
	unsigned char *smem = (unsigned char *)emu::dynamic_smem();
	int *mine = (int *)smem;
	for (int r = 0; r < rounds; r++) {
		mine[threadIdx.x] = data[(blockIdx.x + 1) % gridDim.x * blockDim.x + threadIdx.x]; /* read the neighbour CTA's slice */
		grid_barrier(gb);
		data[blockIdx.x * blockDim.x + threadIdx.x] = mine[threadIdx.x] + 1;             /* write my own slice */
		grid_barrier(gb);
	}
}
int main() {
	const int G = 5, T = 96, R = 7;
	int *d; GridBar *gb;
	cudaMalloc(&d, G * T * sizeof(int)); cudaMalloc(&gb, sizeof(GridBar));
	memset(gb, 0, sizeof *gb);
	for (int i = 0; i < G * T; i++) d[i] = i / T * 1000;
	cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
	cudaLaunchConfig_t cfg = {dim3(G), dim3(T), T * sizeof(int), nullptr, at, 1};
	cudaLaunchKernelEx(&cfg, k_coop, d, gb, R);
	int bad = 0;
	for (int c = 0; c < G; c++) for (int t = 0; t < T; t++) bad += d[c * T + t] != ((c + R) % G) * 1000 + R;
	printf("coop test: %s\n", bad ? "FAILED" : "ok");
	return bad != 0;
}
