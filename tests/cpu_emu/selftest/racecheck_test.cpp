#include <cuda_runtime.h>
void k_racy(int *out) {
	int *buf = (int *)emu::dynamic_smem();
	int t = threadIdx.x;
	buf[t] = t;
	out[blockIdx.x * blockDim.x + t] = buf[(t + 1) % blockDim.x]; /* missing __syncthreads() */
	__syncthreads();
}
void k_ok(int *out) {
	int *buf = (int *)emu::dynamic_smem();
	int t = threadIdx.x;
	buf[t] = t;
	__syncthreads();
	out[blockIdx.x * blockDim.x + t] = buf[(t + 1) % blockDim.x];
	__syncthreads();
}
extern "C" unsigned long emu_racecheck_count(void);
int main() {
	int *g; cudaMalloc(&g, 2048 * sizeof(int));
	EMU_LAUNCH(k_ok, 2, 1024, 4096, nullptr, g);
	unsigned long a = emu_racecheck_count();
	EMU_LAUNCH(k_racy, 2, 1024, 4096, nullptr, g);
	unsigned long b = emu_racecheck_count();
	printf("ok kernel: %lu hazards, racy kernel: %lu hazards\n", a, b - a);
	return !(a == 0 && b > a);
}
