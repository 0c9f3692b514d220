/*
 * gf2b200.cu -- host side of libgf2b200.so: the C-ABI declared in
 * include/gf2b200.h on top of the sm_100a kernels in gf2b200_kernels.cuh.
 *
 * Stands where M4RI stands behind gf2bv/_internal.c:429-489 (PLUQ + solve +
 * kernel basis).  No CPU fallback: every entry point needs a CUDA device.
 */
#include "gf2b200_kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gf2b200.h"

using namespace gf2b200;

struct gf2b200_ctx {
	int device;
	int n_sm;
	cudaStream_t own_stream;
	cudaStream_t stream;
	int profile;
	int rank, world;
	void *nccl; /* ncclComm_t (dist contexts) */
	char err[512];
};

struct gf2b200_system {
	gf2b200_ctx *ctx;
	long long m_global, n;
	long long row_begin; /* first global row held here */
	Mat M;
	u64 *d_pc[2];
	SolverState *d_state;
	PanelDesc *d_pd;
	long long *d_hist_r;
	u64 *d_hist_pm;
	uint4 *d_ebuf;
	u64 *d_x; /* particular solution, nw words */
	std::vector<long long> hist_r;
	std::vector<u64> hist_pm;
	long long rank;
	int inconsistent;
	int eliminated;
	gf2b200_stats stats;
	std::vector<cudaEvent_t> ev;
};

static thread_local char g_err[512] = "";

static int fail(gf2b200_ctx *ctx, int code, const char *fmt, const char *a = "", const char *b = "") {
	char buf[512];
	snprintf(buf, sizeof buf, fmt, a, b);
	if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", buf);
	snprintf(g_err, sizeof g_err, "%s", buf);
	return code;
}

#define CK(ctx, call)                                                                      \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                           \
		if (e_ != cudaSuccess)                                                             \
			return fail(ctx, e_ == cudaErrorMemoryAllocation ? GF2B200_ENOMEM : GF2B200_ECUDA, \
			            "%s: %s", #call, cudaGetErrorString(e_));                          \
	} while (0)

extern "C" int gf2b200_abi_version(void) { return GF2B200_ABI_VERSION; }

extern "C" int gf2b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" const char *gf2b200_last_error(const gf2b200_ctx *ctx) { return ctx ? ctx->err : g_err; }

static int ctx_init(gf2b200_ctx **out, int device) {
	if (!out) return fail(nullptr, GF2B200_EINVAL, "out is NULL");
	*out = nullptr;
	int ndev = gf2b200_device_count();
	if (ndev <= 0)
		return fail(nullptr, GF2B200_ENODEV, "no CUDA device: libgf2b200 has no CPU fallback");
	if (device < 0 || device >= ndev) return fail(nullptr, GF2B200_EINVAL, "bad device index");
	gf2b200_ctx *c = (gf2b200_ctx *)calloc(1, sizeof *c);
	if (!c) return fail(nullptr, GF2B200_ENOMEM, "calloc");
	c->device = device;
	c->world = 1;
	cudaError_t e = cudaSetDevice(device);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
	cudaDeviceProp prop;
	if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) {
		int rc = fail(nullptr, GF2B200_ECUDA, "context init: %s", cudaGetErrorString(e));
		free(c);
		return rc;
	}
	if (prop.major < 10) {
		free(c);
		return fail(nullptr, GF2B200_ENODEV, "device is not sm_100-class (built for sm_100a only)");
	}
	c->n_sm = prop.multiProcessorCount;
	c->stream = c->own_stream;
	e = cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(k_backsub, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	if (e != cudaSuccess) {
		int rc = fail(nullptr, GF2B200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
		free(c);
		return rc;
	}
	*out = c;
	return GF2B200_OK;
}

extern "C" int gf2b200_create(gf2b200_ctx **out, int device) { return ctx_init(out, device); }

extern "C" int gf2b200_nccl_unique_id(void *out_id128) {
	(void)out_id128;
	return fail(nullptr, GF2B200_ENCCL, "multi-GPU path not built yet");
}

extern "C" int gf2b200_create_dist(gf2b200_ctx **out, int device, int rank, int world,
                                   const void *nccl_id128) {
	(void)device; (void)rank; (void)nccl_id128;
	if (out) *out = nullptr;
	if (world == 1) return ctx_init(out, device);
	return fail(nullptr, GF2B200_ENCCL, "multi-GPU path not built yet");
}

extern "C" void gf2b200_destroy(gf2b200_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	free(ctx);
}

extern "C" int gf2b200_set_stream(gf2b200_ctx *ctx, void *cuda_stream) {
	if (!ctx) return GF2B200_EINVAL;
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
	return GF2B200_OK;
}

extern "C" int gf2b200_set_profile(gf2b200_ctx *ctx, int profile) {
	if (!ctx) return GF2B200_EINVAL;
	ctx->profile = profile;
	return GF2B200_OK;
}

extern "C" int gf2b200_host_alloc(void **out, size_t bytes) {
	if (!out) return fail(nullptr, GF2B200_EINVAL, "out is NULL");
	*out = nullptr;
	if (gf2b200_device_count() <= 0)
		return fail(nullptr, GF2B200_ENODEV, "no CUDA device: libgf2b200 has no CPU fallback");
	cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		*out = nullptr;
		return fail(nullptr, GF2B200_ENOMEM, "cudaHostAlloc: %s", cudaGetErrorString(e));
	}
	return GF2B200_OK;
}

extern "C" void gf2b200_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

extern "C" void gf2b200_result_free(gf2b200_result *res) {
	if (!res) return;
	free(res->origin);
	free(res->basis);
	free(res->pivcols);
	res->origin = res->basis = nullptr;
	res->pivcols = nullptr;
}

/* ---- systems -------------------------------------------------------------- */

extern "C" void gf2b200_system_destroy(gf2b200_system *sys) {
	if (!sys) return;
	cudaSetDevice(sys->ctx->device);
	cudaFree(sys->M.base);
	cudaFree(sys->d_pc[0]);
	cudaFree(sys->d_pc[1]);
	cudaFree(sys->d_state);
	cudaFree(sys->d_pd);
	cudaFree(sys->d_hist_r);
	cudaFree(sys->d_hist_pm);
	cudaFree(sys->d_ebuf);
	cudaFree(sys->d_x);
	for (cudaEvent_t e : sys->ev) cudaEventDestroy(e);
	delete sys;
}

extern "C" int gf2b200_system_create(gf2b200_ctx *ctx, int64_t m, int64_t n, gf2b200_system **out) {
	if (!ctx || !out) return fail(ctx, GF2B200_EINVAL, "NULL argument");
	*out = nullptr;
	if (m < 1 || n < 1) return fail(ctx, GF2B200_EINVAL, "m and n must be >= 1");
	if (m >= (1LL << 31) - 4096 || (n + 63) / 64 >= (1LL << 27))
		return fail(ctx, GF2B200_EINVAL, "system too large for 32-bit row indices");
	CK(ctx, cudaSetDevice(ctx->device));
	gf2b200_system *s = new gf2b200_system();
	s->ctx = ctx;
	s->m_global = m;
	s->n = n;
	long long r0 = m * ctx->rank / ctx->world, r1 = m * (ctx->rank + 1) / ctx->world;
	s->row_begin = r0;
	Mat &M = s->M;
	M.m = r1 - r0;
	M.n = n;
	M.nw = (int)((n + 63) / 64);
	M.ns = (M.nw + 1 + 7) / 8;
	M.mp = (std::max<long long>(M.m, 1) + 15) / 16 * 16;
	M.base = nullptr;
	s->d_pc[0] = s->d_pc[1] = nullptr;
	s->d_state = nullptr; s->d_pd = nullptr; s->d_hist_r = nullptr; s->d_hist_pm = nullptr;
	s->d_ebuf = nullptr; s->d_x = nullptr;
	s->rank = 0; s->inconsistent = 0; s->eliminated = 0;
	memset(&s->stats, 0, sizeof s->stats);
	size_t mat_bytes = (size_t)M.ns * (size_t)M.mp * 64;
	cudaError_t e = cudaMalloc(&M.base, mat_bytes);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_pc[0], (size_t)M.mp * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_pc[1], (size_t)M.mp * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_state, sizeof(SolverState));
	if (e == cudaSuccess) e = cudaMalloc(&s->d_pd, sizeof(PanelDesc));
	if (e == cudaSuccess) e = cudaMalloc(&s->d_hist_r, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_hist_pm, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_ebuf, (size_t)M.ns * 4096);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_x, (size_t)M.nw * 8);
	if (e != cudaSuccess) {
		int rc = fail(ctx, e == cudaErrorMemoryAllocation ? GF2B200_ENOMEM : GF2B200_ECUDA,
		              "system_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		gf2b200_system_destroy(s);
		return rc;
	}
	*out = s;
	return GF2B200_OK;
}

extern "C" int64_t gf2b200_system_local_rows(const gf2b200_system *sys) { return sys ? sys->M.m : -1; }

static int grid_for(long long items, int threads, int cap) {
	long long g = (items + threads - 1) / threads;
	if (g < 1) g = 1;
	if (g > cap) g = cap;
	return (int)g;
}

extern "C" int gf2b200_system_load_device(gf2b200_system *sys, const uint64_t *dA,
                                          const uint64_t *db, int64_t stride64) {
	if (!sys || !dA) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	if (stride64 < sys->M.nw) return fail(ctx, GF2B200_EINVAL, "stride64 < ceil(n/64)");
	CK(ctx, cudaSetDevice(ctx->device));
	long long total = sys->M.m * sys->M.ns * 8;
	k_layout<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
	    sys->M, (const u64 *)dA, (const u64 *)db, stride64, 0, sys->M.m, 0);
	CK(ctx, cudaGetLastError());
	sys->eliminated = 0;
	return GF2B200_OK;
}

extern "C" int gf2b200_system_load_host(gf2b200_system *sys, const uint64_t *A, const uint64_t *b,
                                        int64_t stride64) {
	if (!sys || !A) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	const Mat &M = sys->M;
	if (stride64 < M.nw) return fail(ctx, GF2B200_EINVAL, "stride64 < ceil(n/64)");
	CK(ctx, cudaSetDevice(ctx->device));
	/* rows travel in chunks through two device staging buffers so the layout
	 * kernel of chunk c overlaps the H2D copy of chunk c+1 */
	const size_t row_bytes = (size_t)stride64 * 8;
	long long chunk_rows = std::max<long long>(1, (long long)((64u << 20) / row_bytes));
	chunk_rows = std::min<long long>(chunk_rows, M.m);
	u64 *stage[2] = {nullptr, nullptr};
	u64 *d_b = nullptr;
	cudaEvent_t done[2] = {nullptr, nullptr};
	int rc = GF2B200_OK;
	cudaError_t e = cudaMalloc(&stage[0], chunk_rows * row_bytes);
	if (e == cudaSuccess && chunk_rows < M.m) e = cudaMalloc(&stage[1], chunk_rows * row_bytes);
	if (e == cudaSuccess && b) e = cudaMalloc(&d_b, (size_t)((M.m + 63) / 64) * 8);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
	if (e == cudaSuccess && b)
		e = cudaMemcpyAsync(d_b, b, (size_t)((M.m + 63) / 64) * 8, cudaMemcpyHostToDevice, ctx->stream);
	int ci = 0;
	for (long long row0 = 0; e == cudaSuccess && row0 < M.m; row0 += chunk_rows, ci ^= 1) {
		long long nr = std::min<long long>(chunk_rows, M.m - row0);
		u64 *st = stage[1] ? stage[ci] : stage[0];
		e = cudaMemcpyAsync(st, A + row0 * stride64, nr * row_bytes, cudaMemcpyHostToDevice, ctx->stream);
		if (e != cudaSuccess) break;
		long long total = nr * M.ns * 8;
		k_layout<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
		    M, st, d_b, stride64, row0, nr, row0);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e != cudaSuccess) rc = fail(ctx, GF2B200_ECUDA, "system_load_host: %s", cudaGetErrorString(e));
	cudaFree(stage[0]);
	cudaFree(stage[1]);
	cudaFree(d_b);
	if (done[0]) cudaEventDestroy(done[0]);
	if (done[1]) cudaEventDestroy(done[1]);
	sys->eliminated = 0;
	return rc;
}

extern "C" int gf2b200_system_generate(gf2b200_system *sys, uint64_t seed) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	const Mat &M = sys->M;
	CK(ctx, cudaSetDevice(ctx->device));
	long long total = M.m * M.ns * 8;
	k_generate<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(M, seed, sys->row_begin);
	/* x* goes through d_x (overwritten later by the solve) */
	k_synth_xstar<<<(M.nw + 255) / 256, 256, 0, ctx->stream>>>(sys->d_x, M.nw, M.n, seed);
	k_synth_dot<<<grid_for(M.m * 32, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
	    M, seed, sys->row_begin, sys->d_x, 0, nullptr);
	CK(ctx, cudaGetLastError());
	sys->eliminated = 0;
	return GF2B200_OK;
}

extern "C" int gf2b200_system_eliminate(gf2b200_system *sys) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	const Mat &M = sys->M;
	cudaStream_t st = ctx->stream;
	CK(ctx, cudaSetDevice(ctx->device));
	const int nw = M.nw;
	const bool prof = ctx->profile != 0;
	size_t need_ev = 4 + (prof ? 2 * (size_t)nw : 0);
	while (sys->ev.size() < need_ev) {
		cudaEvent_t e;
		CK(ctx, cudaEventCreate(&e));
		sys->ev.push_back(e);
	}
	cudaEvent_t ev_begin = sys->ev[0], ev_fwd = sys->ev[1], ev_end = sys->ev[2];
	long long launches = 0;

	CK(ctx, cudaEventRecord(ev_begin, st));
	CK(ctx, cudaMemsetAsync(sys->d_state, 0, sizeof(SolverState), st));
	k_extract_pc<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, 0, sys->d_pc[0], 0);
	launches++;
	const int apply_cap = ctx->n_sm * 4;
	for (int w = 0; w < nw; w++) {
		u64 colmask = ~0ULL;
		if (w == nw - 1 && (M.n & 63)) colmask = (1ULL << (M.n & 63)) - 1;
		u64 *pc_cur = sys->d_pc[w & 1], *pc_next = sys->d_pc[(w + 1) & 1];
		k_select<<<1, SEL_THREADS, 0, st>>>(M, pc_cur, w, colmask, sys->d_state, sys->d_pd,
		                                    sys->d_hist_r, sys->d_hist_pm);
		int s0a = w >> 3;
		k_apply<<<std::min(M.ns - s0a, apply_cap), 256, 0, st>>>(M, sys->d_pd, sys->d_ebuf, s0a);
		if (prof) CK(ctx, cudaEventRecord(sys->ev[4 + 2 * w], st));
		k_sweep<<<ctx->n_sm, SWEEP_THREADS, SWEEP_SMEM, st>>>(M, sys->d_pd, pc_cur, pc_next,
		                                                     sys->d_ebuf, w, (w + 1) >> 3);
		if (prof) CK(ctx, cudaEventRecord(sys->ev[5 + 2 * w], st));
		launches += 3;
	}
	k_check<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, sys->d_state);
	launches++;
	CK(ctx, cudaEventRecord(ev_fwd, st));
	size_t xs_bytes = (size_t)M.ns * 64;
	if (xs_bytes > 200 * 1024) return fail(ctx, GF2B200_EINVAL, "n too large for the back-substitution kernel");
	k_backsub<<<1, 1024, xs_bytes, st>>>(M, sys->d_hist_r, sys->d_hist_pm, nullptr, 1, sys->d_x);
	launches++;
	CK(ctx, cudaEventRecord(ev_end, st));
	CK(ctx, cudaGetLastError());

	sys->hist_r.resize(nw);
	sys->hist_pm.resize(nw);
	SolverState hs;
	CK(ctx, cudaMemcpyAsync(sys->hist_r.data(), sys->d_hist_r, (size_t)nw * 8, cudaMemcpyDeviceToHost, st));
	CK(ctx, cudaMemcpyAsync(sys->hist_pm.data(), sys->d_hist_pm, (size_t)nw * 8, cudaMemcpyDeviceToHost, st));
	CK(ctx, cudaMemcpyAsync(&hs, sys->d_state, sizeof hs, cudaMemcpyDeviceToHost, st));
	CK(ctx, cudaStreamSynchronize(st));
	sys->rank = hs.r;
	sys->inconsistent = hs.inconsistent;
	sys->eliminated = 1;

	gf2b200_stats &S = sys->stats;
	memset(&S, 0, sizeof S);
	float ms = 0;
	CK(ctx, cudaEventElapsedTime(&ms, ev_begin, ev_end));
	S.ms_total = ms;
	CK(ctx, cudaEventElapsedTime(&ms, ev_begin, ev_fwd));
	S.ms_forward = ms;
	CK(ctx, cudaEventElapsedTime(&ms, ev_fwd, ev_end));
	S.ms_backward = ms;
	S.kernel_launches = launches;
	S.panels = nw;
	S.rank = sys->rank;
	S.m_local = M.m;
	for (int w = 0; w < nw; w++) {
		int k = __builtin_popcountll(sys->hist_pm[w]);
		long long r1 = sys->hist_r[w] + k;
		if (k == 0 || r1 >= M.m) continue;
		double bytes = 2.0 * (double)(M.m - r1) * 64.0 * (double)(M.ns - ((w + 1) >> 3));
		S.sweep_bytes += bytes;
		S.sweep_launches++;
		if (prof) {
			CK(ctx, cudaEventElapsedTime(&ms, sys->ev[4 + 2 * w], sys->ev[5 + 2 * w]));
			S.ms_sweep += ms;
			if (ms > S.ms_sweep_max) {
				S.ms_sweep_max = ms;
				S.sweep_bytes_max = bytes;
			}
		}
	}
	return GF2B200_OK;
}

extern "C" int gf2b200_system_stats(const gf2b200_system *sys, gf2b200_stats *out) {
	if (!sys || !out) return GF2B200_EINVAL;
	*out = sys->stats;
	return GF2B200_OK;
}

extern "C" int gf2b200_system_result(gf2b200_system *sys, int mode, gf2b200_result *out) {
	if (!sys || !out) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	memset(out, 0, sizeof *out);
	if (mode != 0 && mode != 1) return fail(ctx, GF2B200_EINVAL, "Invalid mode");
	if (!sys->eliminated) return fail(ctx, GF2B200_EINVAL, "system_result before system_eliminate");
	const Mat &M = sys->M;
	const int nw = M.nw;
	CK(ctx, cudaSetDevice(ctx->device));
	out->rank = sys->rank;
	if (sys->inconsistent) {
		out->status = GF2B200_INCONSISTENT;
		return GF2B200_OK;
	}
	out->origin = (uint64_t *)calloc((size_t)nw, 8);
	out->pivcols = (int64_t *)malloc((size_t)std::max<long long>(sys->rank, 1) * 8);
	if (!out->origin || !out->pivcols) {
		gf2b200_result_free(out);
		return fail(ctx, GF2B200_ENOMEM, "malloc");
	}
	CK(ctx, cudaMemcpyAsync(out->origin, sys->d_x, (size_t)nw * 8, cudaMemcpyDeviceToHost, ctx->stream));
	long long q = 0;
	for (int w = 0; w < nw; w++) {
		u64 pm = sys->hist_pm[w];
		while (pm) {
			out->pivcols[q++] = (int64_t)w * 64 + __builtin_ctzll(pm);
			pm &= pm - 1;
		}
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	if (mode == 1 && sys->rank < M.n) {
		/* free columns in M4RI's sigma order: arrangement after "for i<r: swap(i, p_i)"
		 * (mzd_apply_p_left_trans at _internal.c:348) */
		const long long n = M.n, r = sys->rank, d = n - r;
		std::vector<long long> sigma((size_t)n);
		for (long long i = 0; i < n; i++) sigma[i] = i;
		for (long long i = 0; i < r; i++) std::swap(sigma[i], sigma[out->pivcols[i]]);
		out->basis = (uint64_t *)malloc((size_t)d * (size_t)nw * 8);
		if (!out->basis) {
			gf2b200_result_free(out);
			return fail(ctx, GF2B200_ENOMEM, "malloc basis");
		}
		long long *d_free = nullptr;
		u64 *d_basis = nullptr;
		/* batches bound the device buffer for huge nullities */
		const long long batch = std::min<long long>(d, std::max<long long>(1, (1LL << 28) / ((long long)nw * 8)));
		cudaError_t e = cudaMalloc(&d_free, (size_t)d * 8);
		if (e == cudaSuccess) e = cudaMalloc(&d_basis, (size_t)batch * nw * 8);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(d_free, sigma.data() + r, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream);
		for (long long b0 = 0; e == cudaSuccess && b0 < d; b0 += batch) {
			long long nb = std::min(batch, d - b0);
			k_backsub<<<(unsigned)nb, 1024, (size_t)M.ns * 64, ctx->stream>>>(
			    M, sys->d_hist_r, sys->d_hist_pm, d_free + b0, 0, d_basis);
			e = cudaGetLastError();
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(out->basis + b0 * nw, d_basis, (size_t)nb * nw * 8,
				                    cudaMemcpyDeviceToHost, ctx->stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		}
		cudaFree(d_free);
		cudaFree(d_basis);
		if (e != cudaSuccess) {
			gf2b200_result_free(out);
			return fail(ctx, GF2B200_ECUDA, "kernel basis: %s", cudaGetErrorString(e));
		}
		out->kernel_dim = d;
	}
	out->status = GF2B200_OK;
	return GF2B200_OK;
}

extern "C" int gf2b200_system_check_synthetic(gf2b200_system *sys, uint64_t seed, const uint64_t *x,
                                              int64_t *bad_rows) {
	if (!sys || !x || !bad_rows) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	const Mat &M = sys->M;
	CK(ctx, cudaSetDevice(ctx->device));
	u64 *d_v = nullptr, *d_xs = nullptr;
	unsigned long long *d_cnt = nullptr;
	unsigned long long cnt = 0;
	cudaError_t e = cudaMalloc(&d_v, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&d_xs, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 8);
	if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, 8, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, x, (size_t)M.nw * 8, cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess) {
		/* A (x ^ x*) == 0  <=>  A x == b because b = A x* by construction */
		k_synth_xstar<<<(M.nw + 255) / 256, 256, 0, ctx->stream>>>(d_xs, M.nw, M.n, seed);
		k_xor_vec<<<(M.nw + 255) / 256, 256, 0, ctx->stream>>>(d_v, d_v, d_xs, M.nw);
		k_synth_dot<<<grid_for(M.m * 32, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
		    M, seed, sys->row_begin, d_v, 1, d_cnt);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d_v);
	cudaFree(d_xs);
	cudaFree(d_cnt);
	if (e != cudaSuccess) return fail(ctx, GF2B200_ECUDA, "check_synthetic: %s", cudaGetErrorString(e));
	*bad_rows = (int64_t)cnt;
	return GF2B200_OK;
}

/* ---- one-shot host-buffer solve (what m4ri_solve's body becomes) ---------- */
extern "C" int gf2b200_solve(gf2b200_ctx *ctx, const uint64_t *A, const uint64_t *b, int64_t m,
                             int64_t n, int64_t stride64, int mode, gf2b200_result *out) {
	if (!ctx || !A || !out) return fail(ctx, GF2B200_EINVAL, "NULL argument");
	memset(out, 0, sizeof *out);
	if (mode != 0 && mode != 1) return fail(ctx, GF2B200_EINVAL, "Invalid mode");
	if (ctx->world != 1) return fail(ctx, GF2B200_EINVAL, "gf2b200_solve needs a single-GPU context");
	gf2b200_system *sys = nullptr;
	int rc = gf2b200_system_create(ctx, m, n, &sys);
	if (rc) return rc;
	rc = gf2b200_system_load_host(sys, A, b, stride64);
	if (!rc) rc = gf2b200_system_eliminate(sys);
	if (!rc) rc = gf2b200_system_result(sys, mode, out);
	gf2b200_system_destroy(sys);
	if (rc) out->status = rc;
	return rc;
}
