// l1tex_probe.cu -- what do the global-side instructions of k_sweep cost on the l1tex
// data pipe?  One tiny kernel per instruction kind; run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,\
//       l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,\
//       smsp__inst_executed.sum,gpu__time_duration.sum
// and divide the wavefronts by the warp-level requests.  (profiles/r01g_sweep_ncu.md: the global
// side of k_sweep costs 12.3 wavefronts per (LDG.128 + LDG.64 + STG.128) where 9 were expected.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1tex_probe scripts/l1tex_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 64
#define THREADS 256

__global__ void k_ldg128(const uint4 *__restrict__ src, uint4 *sink) {
	const uint4 *p = src + (size_t)blockIdx.x * THREADS * ITERS + threadIdx.x;
	uint4 a = make_uint4(0, 0, 0, 0);
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) {
		uint4 v = __ldcg(p + (size_t)i * THREADS);
		a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
	}
	if (a.x == 0x12345678u && a.y == 1) sink[0] = a;
}

__global__ void k_stg128(uint4 *dst) {
	uint4 *p = dst + (size_t)blockIdx.x * THREADS * ITERS + threadIdx.x;
	uint4 a = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) __stcg(p + (size_t)i * THREADS, a);
}

// 256-bit accesses (sm_100+): two row chunks per thread
__global__ void k_ldg256(const uint4 *__restrict__ src, uint4 *sink) {
	const uint4 *p = src + ((size_t)blockIdx.x * THREADS * ITERS + threadIdx.x) * 2;
	unsigned a = 0;
#pragma unroll 8
	for (int i = 0; i < ITERS / 2; i++) {
		unsigned r0, r1, r2, r3, r4, r5, r6, r7;
		asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
		             : "l"(p + (size_t)i * THREADS * 2));
		a ^= r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
	}
	if (a == 0x12345678u) sink[0] = make_uint4(a, 0, 0, 0);
}

__global__ void k_stg256(uint4 *dst) {
	uint4 *p = dst + ((size_t)blockIdx.x * THREADS * ITERS + threadIdx.x) * 2;
	unsigned a = threadIdx.x;
#pragma unroll 8
	for (int i = 0; i < ITERS / 2; i++)
		asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + (size_t)i * THREADS * 2), "r"(a),
		             "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(a + 4), "r"(a + 5), "r"(a + 6), "r"(a + 7)
		             : "memory");
}

// the coefficient load of k_sweep: 4 threads share one 8-byte word
__global__ void k_ldg64_dup4(const unsigned long long *__restrict__ src, unsigned long long *sink) {
	const unsigned long long *p = src + (size_t)blockIdx.x * (THREADS / 4) * ITERS + threadIdx.x / 4;
	unsigned long long a = 0;
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) a ^= __ldg(p + (size_t)i * (THREADS / 4));
	if (a == 0x12345678u) sink[0] = a;
}

// same, one lane in four loads
__global__ void k_ldg64_lane(const unsigned long long *__restrict__ src, unsigned long long *sink) {
	const unsigned long long *p = src + (size_t)blockIdx.x * (THREADS / 4) * ITERS + threadIdx.x / 4;
	unsigned long long a = 0;
#pragma unroll 8
	for (int i = 0; i < ITERS; i++)
		if ((threadIdx.x & 3) == 0) a ^= __ldg(p + (size_t)i * (THREADS / 4));
	if (a == 0x12345678u) sink[0] = a;
}

// fully coalesced 8-byte load (one word per lane)
__global__ void k_ldg64(const unsigned long long *__restrict__ src, unsigned long long *sink) {
	const unsigned long long *p = src + (size_t)blockIdx.x * THREADS * ITERS + threadIdx.x;
	unsigned long long a = 0;
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) a ^= __ldg(p + (size_t)i * THREADS);
	if (a == 0x12345678u) sink[0] = a;
}

__global__ void k_shfl(unsigned *sink) {
	unsigned a = threadIdx.x * 2654435761u;
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) a ^= __shfl_sync(0xffffffffu, a, (threadIdx.x + i) & 28);
	if (a == 0x12345678u) sink[0] = a;
}

__global__ void k_lds128(uint4 *sink) {
	__shared__ uint4 T[THREADS * 2];
	T[threadIdx.x] = make_uint4(threadIdx.x, 1, 2, 3);
	T[threadIdx.x + THREADS] = make_uint4(threadIdx.x, 5, 6, 7);
	__syncthreads();
	uint4 a = make_uint4(0, 0, 0, 0);
	unsigned idx = threadIdx.x;
#pragma unroll 8
	for (int i = 0; i < ITERS; i++) {
		uint4 v = T[idx & (2 * THREADS - 1)];
		a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
		idx += 8 + (a.x & 8);
	}
	if (a.x == 0x12345678u && a.y == 1) sink[0] = a;
}

int main() {
	const int grid = 148 * 8;
	size_t bytes = (size_t)grid * THREADS * ITERS * 16 * 2;
	void *buf, *sink;
	cudaMalloc(&buf, bytes);
	cudaMalloc(&sink, 256);
	cudaMemset(buf, 1, bytes);
	for (int rep = 0; rep < 2; rep++) {
		k_ldg128<<<grid, THREADS>>>((const uint4 *)buf, (uint4 *)sink);
		k_stg128<<<grid, THREADS>>>((uint4 *)buf);
		k_ldg256<<<grid, THREADS>>>((const uint4 *)buf, (uint4 *)sink);
		k_stg256<<<grid, THREADS>>>((uint4 *)buf);
		k_ldg64_dup4<<<grid, THREADS>>>((const unsigned long long *)buf, (unsigned long long *)sink);
		k_ldg64_lane<<<grid, THREADS>>>((const unsigned long long *)buf, (unsigned long long *)sink);
		k_ldg64<<<grid, THREADS>>>((const unsigned long long *)buf, (unsigned long long *)sink);
		k_shfl<<<grid, THREADS>>>((unsigned *)sink);
		k_lds128<<<grid, THREADS>>>((uint4 *)sink);
	}
	cudaError_t e = cudaDeviceSynchronize();
	printf("l1tex_probe: %s; warp-level requests per kernel = %d\n", cudaGetErrorString(e), grid * THREADS / 32 * ITERS);
	return e != cudaSuccess;
}
