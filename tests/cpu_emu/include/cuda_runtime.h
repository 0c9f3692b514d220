/*
 * cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for the CUDA runtime and
 * the device-side language features that gf2bv_b200/csrc uses, so that the SOURCE of
 * the sm_100a kernels can be executed on a machine without a GPU (tests/cpu_emu).
 *
 * It is never part of the product: libgf2b200.so is built by nvcc from the same
 * sources and has no CPU path.  The emulated library is built by
 * tests/cpu_emu/build_emu.py into tests/cpu_emu/_build/ and loaded only by
 * tests/test_emu_kernels.py (in a subprocess).  Purpose: catch logic errors in the
 * kernels (indexing, table layouts, pivot bookkeeping, shard exchange order) before
 * GPU time is spent.  It says nothing about races, memory ordering or speed.
 *
 * Execution model: one OS thread.  A launch runs the CTAs of the grid one after
 * another; the threads of a CTA are fibers (own stack, hand-written context
 * switch) scheduled round-robin and parked at __syncthreads / warp collectives /
 * mbarrier waits until their condition holds.  Warp collectives take all live
 * lanes of the warp (the sources only use full masks).
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define GF2_EMU 1

/* ---- language ------------------------------------------------------------ */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct alignas(16) uint4 {
	unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
	dim3(int x_) : x((unsigned)x_), y(1), z(1) {}
	dim3(long long x_) : x((unsigned)x_), y(1), z(1) {}
};

/* ThreadSanitizer build (build_emu.py --tsan): every GPU thread is a TSan fiber and only
 * the GPU's own synchronisation (barriers, warp collectives, mbarriers, atomics, kernel
 * boundaries) creates happens-before edges, so TSan reports shared/global-memory races
 * between GPU threads -- a CPU-side racecheck. */
#ifdef __SANITIZE_THREAD__
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void *fiber);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
void __tsan_acquire(void *addr);
void __tsan_release(void *addr);
}
#define EMU_ACQUIRE(p) __tsan_acquire((void *)(p))
#define EMU_RELEASE(p) __tsan_release((void *)(p))
#define EMU_INTERNAL __attribute__((no_sanitize("thread"))) /* the emulator's own bookkeeping is not GPU memory */
#else
#define EMU_ACQUIRE(p) ((void)0)
#define EMU_RELEASE(p) ((void)0)
#define EMU_INTERNAL
#endif

namespace emu {

enum WaitKind { W_NONE = 0, W_BARRIER, W_WARP_ENTER, W_WARP_RESULT, W_WORD };

struct Warp {
	unsigned long long vals[32], snap[32];
	unsigned gen;       /* completed collectives (racecheck: orders accesses inside the warp) */
	unsigned live;      /* lanes that have not exited */
	unsigned arrived;   /* lanes waiting in the current collective */
	unsigned departing; /* lanes that still have to read snap */
	unsigned snap_mask; /* lanes that contributed to snap */
};

struct Cta;
struct Fiber {
	void *sp;
	int done;
	int wait;
	unsigned bar_gen;
	const volatile unsigned long long *word; /* W_WORD: wait until (*word & 1) != word_val */
	unsigned long long word_val;
	dim3 tid;
	int lane, warp;
	Cta *cta;
	void *tsan; /* TSan fiber handle (ThreadSanitizer builds) */
	unsigned tma_seen; /* racecheck: bulk copies of this CTA whose mbarrier phase this thread has waited for */
};

struct Cta {
	std::vector<Fiber> f;
	std::vector<Warp> w;
	int alive, bar_arrived;
	unsigned bar_gen;
	unsigned tma_count; /* bulk copies issued in this CTA so far */
	dim3 bid;
	unsigned char *smem; /* this CTA's dynamic shared memory */
};

extern Fiber *cur;
extern Cta *cta_p; /* the CTA of the running fiber */
extern dim3 g_blockDim, g_gridDim;
extern void *sched_sp;
static const size_t DYN_SMEM_CAP = 232 * 1024;

extern "C" void emu_switch(void **save_sp, void *new_sp);
void yield_to_scheduler();
void run_grid(dim3 grid, dim3 block, size_t smem, bool concurrent, void (*thunk)(void *), void *arg);
bool is_dynamic_smem(const void *p);
void warp_complete_if_ready(Warp &W);

EMU_INTERNAL static inline void block_on(int kind) {
	cur->wait = kind;
	yield_to_scheduler();
}

/* every live lane deposits v; returns the snapshot of all lanes (0 for dead lanes) */
EMU_INTERNAL static inline const unsigned long long *warp_exchange(unsigned long long v, unsigned *mask_out = nullptr) {
	Warp &W = cta_p->w[cur->warp];
	const unsigned bit = 1u << cur->lane;
	while (W.departing) block_on(W_WARP_ENTER);
	W.vals[cur->lane] = v;
	W.arrived |= bit;
	EMU_RELEASE(&W);
	warp_complete_if_ready(W);
	while (!(W.departing & bit)) block_on(W_WARP_RESULT);
	EMU_ACQUIRE(&W);
	static thread_local unsigned long long out[32];
	memcpy(out, W.snap, sizeof out);
	if (mask_out) *mask_out = W.snap_mask;
	W.departing &= ~bit;
	return out;
}

EMU_INTERNAL static inline const dim3 &self_tid() { return cur->tid; }
EMU_INTERNAL static inline const dim3 &self_bid() { return cta_p->bid; }
/* `extern __shared__ T name[]` is rewritten by build_emu.py into a pointer to this */
EMU_INTERNAL static inline void *dynamic_smem() { return cta_p->smem; }
EMU_INTERNAL static inline const dim3 &self_bdim() { return g_blockDim; }
EMU_INTERNAL static inline const dim3 &self_gdim() { return g_gridDim; }

} /* namespace emu */

#define threadIdx (emu::self_tid())
#define blockIdx (emu::self_bid())
#define blockDim (emu::self_bdim())
#define gridDim (emu::self_gdim())

/* ---- device intrinsics ---------------------------------------------------- */
EMU_INTERNAL static inline void __syncthreads() {
	emu::Cta &C = *emu::cta_p;
	EMU_RELEASE(&C.bar_gen);
	C.bar_arrived++;
	if (C.bar_arrived == C.alive) {
		C.bar_arrived = 0;
		C.bar_gen++;
		EMU_ACQUIRE(&C.bar_gen);
		return;
	}
	emu::cur->bar_gen = C.bar_gen;
	while (emu::cur->bar_gen == C.bar_gen) emu::block_on(emu::W_BARRIER);
	EMU_ACQUIRE(&C.bar_gen);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
EMU_INTERNAL static inline unsigned __ballot_sync(unsigned, int pred) {
	const unsigned long long *s = emu::warp_exchange(pred ? 1 : 0);
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)(s[i] & 1) << i;
	return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
EMU_INTERNAL static inline int __all_sync(unsigned m, int pred) {
	unsigned live;
	const unsigned long long *s = emu::warp_exchange(pred ? 1 : 0, &live);
	for (int i = 0; i < 32; i++)
		if (((live >> i) & 1) && !s[i]) return 0;
	return 1;
}
EMU_INTERNAL static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)emu::warp_exchange(v)[src & 31]; }
EMU_INTERNAL static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::warp_exchange((unsigned)v)[src & 31]; }
EMU_INTERNAL static inline unsigned __reduce_xor_sync(unsigned, unsigned v) {
	const unsigned long long *s = emu::warp_exchange(v);
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r ^= (unsigned)s[i];
	return r;
}
static inline int __reduce_xor_sync(unsigned m, int v) { return (int)__reduce_xor_sync(m, (unsigned)v); }
EMU_INTERNAL static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
	const unsigned long long *s = emu::warp_exchange(v);
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)s[i];
	return r;
}
static inline int __reduce_or_sync(unsigned m, int v) { return (int)__reduce_or_sync(m, (unsigned)v); }

static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
	return (unsigned)((((unsigned long long)hi << 32) | lo) >> (s & 31));
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
	const unsigned long long src = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		const unsigned s = (sel >> (4 * i)) & 0xF;
		unsigned byte = (unsigned)(src >> (8 * (s & 7))) & 0xFF;
		if (s & 8) byte = (byte & 0x80) ? 0xFF : 0; /* sign replication mode */
		r |= byte << (8 * i);
	}
	return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline void __stcg(T *p, const T &v) { *p = v; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}
/* a spinning thread lets the others run (CTAs of a cooperative launch wait for each other) */
EMU_INTERNAL static inline void __nanosleep(unsigned) { emu::block_on(emu::W_NONE); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }

/* mbarrier + bulk copy stand-ins used by the GF2_EMU branches of the kernels:
 * the barrier word's bit 0 is the phase that completes next */
static inline void emu_mbar_init(void *bar) { __atomic_store_n((unsigned long long *)bar, 0ULL, __ATOMIC_RELEASE); }
static inline void emu_mbar_complete(void *bar) {
	EMU_RELEASE(bar);
	__atomic_fetch_xor((unsigned long long *)bar, 1ULL, __ATOMIC_RELEASE);
}
extern "C" void emu_racecheck_bulk_write(void *dst, size_t bytes, unsigned seq); /* --racecheck builds */
/* the bulk copy (TMA stand-in): done at once, completes the barrier's current phase */
EMU_INTERNAL static inline void emu_bulk_copy(void *dst, const void *src, unsigned bytes, void *bar) {
	memcpy(dst, src, bytes);
	emu::cta_p->tma_count++;
#ifdef EMU_RACECHECK
	emu_racecheck_bulk_write(dst, bytes, emu::cta_p->tma_count);
#endif
	emu_mbar_complete(bar);
}
EMU_INTERNAL static inline void emu_mbar_wait(void *bar, unsigned phase) {
	volatile unsigned long long *b = (volatile unsigned long long *)bar;
	while ((__atomic_load_n((unsigned long long *)bar, __ATOMIC_ACQUIRE) & 1) == phase) {
		emu::cur->word = b;
		emu::cur->word_val = phase;
		emu::block_on(emu::W_WORD);
	}
	EMU_ACQUIRE(bar);
	emu::cur->tma_seen = emu::cta_p->tma_count;
}

/* ---- runtime API ------------------------------------------------------------ */
typedef enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 } cudaError_t;
typedef struct emu_stream_ *cudaStream_t;
struct emu_event_ { double t; };
typedef emu_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2
#define cudaHostAllocPortable 1
#define cudaIpcMemLazyEnablePeerAccess 1
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
struct cudaIpcMemHandle_t { char reserved[64]; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
	const char *s = getenv("GF2_EMU_SMS");
	p->major = 10; p->minor = 0; p->multiProcessorCount = s ? atoi(s) : 3;
	return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
	void *q = nullptr;
	if (posix_memalign(&q, 256, bytes ? bytes : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
	memset(q, 0xA5, bytes); /* device memory is not zeroed: make reliance on it visible */
	*p = (T *)q;
	return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) {
	return posix_memalign(p, 256, bytes) ? cudaErrorMemoryAllocation : cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline double emu_now_ms() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

/* ---- launches: build_emu.py rewrites  k<<<g, b, smem, st>>>(args)  into
 * EMU_LAUNCH(k, g, b, smem, st, args) -------------------------------------- */
namespace emu {
template <typename... P> struct Call {
	void (*k)(P...);
	std::tuple<std::decay_t<P>...> args;
	static void thunk(void *self) {
		Call *c = (Call *)self;
		std::apply(c->k, c->args);
	}
};
template <typename... P, typename... A>
static inline void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, bool concurrent, void (*k)(P...), A &&...a) {
	Call<P...> c{k, std::tuple<std::decay_t<P>...>(std::forward<A>(a)...)};
	run_grid(grid, block, smem, concurrent, &Call<P...>::thunk, &c);
}
} /* namespace emu */

/* cooperative launches: every CTA of the grid is resident at once -- the emulator then
 * interleaves the fibers of ALL CTAs (kernels launched this way must keep their shared
 * memory dynamic: static __shared__ variables are one process-wide instance here) */
enum cudaLaunchAttributeID { cudaLaunchAttributeCooperative = 2 };
struct cudaLaunchAttributeValue { int cooperative; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { /* same member ORDER as CUDA's (callers use positional initialisation:
	                            * gridDim / blockDim are macros in this header) */
	dim3 grid_, block_;
	size_t dynamicSmemBytes;
	cudaStream_t stream;
	cudaLaunchAttribute *attrs;
	unsigned numAttrs;
};
template <typename... P, typename... A>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*k)(P...), A &&...a) {
	bool coop = false;
	for (unsigned i = 0; i < cfg->numAttrs; i++)
		if (cfg->attrs[i].id == cudaLaunchAttributeCooperative && cfg->attrs[i].val.cooperative) coop = true;
	emu::launch(cfg->grid_, cfg->block_, cfg->dynamicSmemBytes, cfg->stream, coop, k, std::forward<A>(a)...);
	return cudaSuccess;
}
#define EMU_LAUNCH(k, g, b, smem, st, ...) emu::launch(dim3(g), dim3(b), (size_t)(smem), (st), false, k, ##__VA_ARGS__)
