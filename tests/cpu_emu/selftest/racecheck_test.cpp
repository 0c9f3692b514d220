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
/* two CTAs: the last one to take a ticket reads what the other wrote -- ordered by
 * fence + barrier + atomic (the pattern of a "last CTA" epilogue or a grid barrier) */
void k_ticket(int *data, unsigned *ticket, int *sum) {
	int t = threadIdx.x;
	data[blockIdx.x * blockDim.x + t] = t;
	__threadfence();
	__syncthreads();
	__shared__ int last;
	if (t == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (last) sum[t] = data[t] + data[blockDim.x + t];
}
/* the same without the ticket: whichever CTA runs second reads unordered data */
void k_no_ticket(int *data, int *sum) {
	int t = threadIdx.x;
	data[blockIdx.x * blockDim.x + t] = t;
	__syncthreads();
	if (blockIdx.x == 1) sum[t] = data[t] + data[blockDim.x + t];
}
extern "C" unsigned long emu_racecheck_count(void);
int main() {
	int *g; cudaMalloc(&g, 2048 * sizeof(int));
	EMU_LAUNCH(k_ok, 2, 1024, 4096, nullptr, g);
	unsigned long a = emu_racecheck_count();
	EMU_LAUNCH(k_racy, 2, 1024, 4096, nullptr, g);
	unsigned long b = emu_racecheck_count();
	printf("ok kernel: %lu hazards, racy kernel: %lu hazards\n", a, b - a);
	unsigned *ticket; int *sum;
	cudaMalloc(&ticket, sizeof(unsigned)); cudaMalloc(&sum, 1024 * sizeof(int));
	*ticket = 0;
	EMU_LAUNCH(k_ticket, 2, 1024, 0, nullptr, g, ticket, sum);
	unsigned long c = emu_racecheck_count();
	EMU_LAUNCH(k_no_ticket, 2, 1024, 0, nullptr, g, sum);
	unsigned long d = emu_racecheck_count();
	printf("ticket kernel: %lu hazards, no-ticket kernel: %lu hazards\n", c - b, d - c);
	return !(a == 0 && b > a && c == b && d > c);
}
