/*
 * gf2b200.cu -- host side of libgf2b200.so: the C-ABI declared in
 * include/gf2b200.h on top of the sm_100a kernels in gf2b200_kernels.cuh (panel
 * select / apply / sweep) and gf2b200_dist.cuh (row-sharded election, pivot-row
 * exchange, blocked back-substitution).
 *
 * Stands where M4RI stands behind gf2bv/_internal.c:429-489 (PLUQ + solve +
 * kernel basis).  No CPU fallback: every entry point needs a CUDA device.
 *
 * A system is a set of row shards.  Three kinds of context:
 *   single    one shard on one GPU (gf2b200_create);
 *   nccl      one shard per process/GPU (gf2b200_create_dist): NCCL is the rendezvous
 *             (IPC handle exchange, back-substitution slabs), the per-panel pivot
 *             exchange is NVLink loads/stores on peer memory from our own kernels;
 *   loopback  `world` shards on ONE GPU in one process (gf2b200_create_shards): the
 *             peers are local pointers -- same kernels and control flow as the nccl
 *             kind, so the sharded path can be parity-tested on one GPU.
 */
#include "gf2b200_dist.cuh"
#include "gf2b200_persist.cuh"
#include "gf2b200_basis.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/gf2b200.h"

using namespace gf2b200;

/* ---- NCCL, resolved at run time (torch's bundled libnccl.so.2 when the caller
 * already loaded it, else the system one) ---------------------------------- */
struct NcclApi {
	void *handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

struct gf2b200_ctx {
	int device;
	int n_sm;
	cudaStream_t own_stream;
	cudaStream_t stream;
	int profile;
	int rank, world; /* shard index of the first local shard, number of shards */
	int n_local;     /* local shards: 1 (single, nccl) or world (loopback) */
	int coop;        /* a grid of n_sm CTAs of SWEEP_THREADS threads can be launched cooperatively (k_forward, k_sweep_apply) */
	int persist;     /* k_forward: 0 never, 1 for matrices below PERSIST_AUTO_BYTES, 2 always (GF2B200_FORWARD) */
	ncclComm_t nccl; /* nccl contexts */
	struct gf2b200_system *cached; /* device buffers of the last gf2b200_solve, reused for equal shapes */
	/* host-buffer loads: staging buffers, copy stream and events live as long as the context
	 * (an MT19937-class API solve is ~20 ms: allocating these per call showed) */
	u64 *stage[2];
	size_t stage_bytes;
	u64 *d_bstage;
	size_t bstage_bytes;
	cudaStream_t copy_stream;
	cudaEvent_t ev_copied[2], ev_laid[2];
	char err[512];
};

struct Shard {
	Mat M;
	int index;           /* global shard index */
	long long row_begin; /* first global row */
	u64 *d_pc[2];
	SolverState *d_state;
	PanelDesc *d_pd;
	long long *d_hist_r;
	u64 *d_hist_pm;
	uint4 *d_ebuf;
	/* one-GPU systems: second E-tile buffer, barrier block and per-panel timestamps of k_forward */
	uint4 *d_ebuf2;
	void *d_gs;
	unsigned long long *d_tpanel;
	u64 *d_cand; /* two candidate lists of PERSIST_CAND_MAX {row, word} pairs (k_forward's slow path) */
	/* sharded systems only */
	DistPanel *d_dp;
	unsigned char *d_hist_owner;
	XchBlock *xch;    /* behind the matrix, same allocation (one IPC handle maps both) */
	PeerTable *d_pt;  /* peers' matrices / exchange blocks as mapped in this process */
	std::vector<void *> ipc_opened;
	/* back-substitution */
	u64 *d_x;    /* nw + 1 words */
	u64 *d_slab; /* BS_S*64 rows x BS_W */
	u64 *d_slab_all;
	std::vector<long long> hist_r;
	/* kernel basis on a sharded system (alive only inside system_result) */
	void *bk_send, *bk_recv;
};

struct gf2b200_system {
	gf2b200_ctx *ctx;
	long long m_global, n;
	std::vector<Shard> sh;
	std::vector<u64> hist_pm;
	std::vector<unsigned char> hist_owner;
	long long rank;
	int inconsistent;
	int eliminated;
	unsigned epoch_base; /* flag epochs of the peer-memory exchange only ever grow */
	/* a host-buffer load in progress (system_load_begin .. system_load_end) */
	long long ld_stride, ld_chunk_rows;
	int ld_ci, ld_used[2];
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

#define NK(ctx, call)                                                                    \
	do {                                                                                 \
		ncclResult_t r_ = (call);                                                        \
		if (r_ != ncclSuccess)                                                           \
			return fail(ctx, GF2B200_ENCCL, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
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

static int nccl_load(void) {
	if (g_nccl.handle) return 0;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) return fail(nullptr, GF2B200_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
	NcclApi a;
	a.handle = h;
	*(void **)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
	*(void **)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
	*(void **)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
	*(void **)&a.AllGather = dlsym(h, "ncclAllGather");
	*(void **)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
	if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
		return fail(nullptr, GF2B200_ENCCL, "libnccl.so.2 lacks a required symbol");
	g_nccl = a;
	return 0;
}

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
	c->n_local = 1;
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
	if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
	/* Tuning switch (off by default, not yet A/B-ed): GF2B200_CARVEOUT=<percent> asks for the
	 * same shared-memory carve-out in the small per-panel kernels as k_sweep gets (164 of
	 * 228 KiB = 72), so that the SMs are not re-partitioned three times per panel. */
	if (const char *cv = getenv("GF2B200_CARVEOUT")) {
		const int pct = atoi(cv);
		if (e == cudaSuccess && pct > 0 && pct <= 100) {
			e = cudaFuncSetAttribute(k_select, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
			if (e == cudaSuccess) e = cudaFuncSetAttribute(k_apply, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
			if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		}
	}
#if SW == 8
	/* the one-kernel forward elimination needs every CTA of a grid of n_sm resident at once
	 * (cooperative launch); GF2B200_FORWARD=launches keeps the per-panel launch chain */
	if (e == cudaSuccess) e = cudaFuncSetAttribute(k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, PERSIST_SMEM);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
	if (e == cudaSuccess) {
		/* 1: k_forward for systems below PERSIST_AUTO_BYTES (default), 2: always (GF2B200_FORWARD=persist),
		 * 0: never (GF2B200_FORWARD=launches) */
		const char *fw = getenv("GF2B200_FORWARD");
		c->persist = (fw && !strcmp(fw, "launches")) ? 0 : (fw && !strcmp(fw, "persist")) ? 2 : 1;
#ifndef GF2_EMU
		int coop = 0, per_sm = 0;
		if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) coop = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_forward, SWEEP_THREADS, PERSIST_SMEM) != cudaSuccess)
			per_sm = 0;
		cudaGetLastError();
		c->coop = coop && per_sm >= 1;
#else
		c->coop = 1;
#endif
		if (!c->coop) c->persist = 0;
	}
#endif
	if (e != cudaSuccess) {
		int rc = fail(nullptr, GF2B200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
		free(c);
		return rc;
	}
	*out = c;
	return GF2B200_OK;
}

extern "C" int gf2b200_create(gf2b200_ctx **out, int device) { return ctx_init(out, device); }

extern "C" int gf2b200_create_shards(gf2b200_ctx **out, int device, int world) {
	if (out) *out = nullptr;
	if (world < 1 || world > 64) return fail(nullptr, GF2B200_EINVAL, "world must be in 1..64");
	int rc = ctx_init(out, device);
	if (rc) return rc;
	(*out)->world = world;
	(*out)->n_local = world;
	return GF2B200_OK;
}

extern "C" int gf2b200_nccl_unique_id(void *out_id128) {
	if (!out_id128) return fail(nullptr, GF2B200_EINVAL, "out is NULL");
	if (nccl_load()) return GF2B200_ENCCL;
	ncclUniqueId id;
	NK(nullptr, g_nccl.GetUniqueId(&id));
	memcpy(out_id128, &id, sizeof id);
	return GF2B200_OK;
}

extern "C" int gf2b200_create_dist(gf2b200_ctx **out, int device, int rank, int world,
                                   const void *nccl_id128) {
	if (out) *out = nullptr;
	if (world == 1) return ctx_init(out, device);
	if (world < 1 || world > 64 || rank < 0 || rank >= world || !nccl_id128)
		return fail(nullptr, GF2B200_EINVAL, "bad rank/world/id");
	if (nccl_load()) return GF2B200_ENCCL;
	int rc = ctx_init(out, device);
	if (rc) return rc;
	gf2b200_ctx *c = *out;
	c->rank = rank;
	c->world = world;
	ncclUniqueId id;
	memcpy(&id, nccl_id128, sizeof id);
	ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
	if (r != ncclSuccess) {
		rc = fail(nullptr, GF2B200_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
		c->nccl = nullptr;
		gf2b200_destroy(c);
		*out = nullptr;
		return rc;
	}
	return GF2B200_OK;
}

extern "C" void gf2b200_destroy(gf2b200_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->cached) gf2b200_system_destroy(ctx->cached);
	cudaFree(ctx->stage[0]);
	cudaFree(ctx->stage[1]);
	cudaFree(ctx->d_bstage);
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	for (int i = 0; i < 2; i++) {
		if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
		if (ctx->ev_laid[i]) cudaEventDestroy(ctx->ev_laid[i]);
	}
	if (ctx->nccl) g_nccl.CommDestroy(ctx->nccl);
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

static void shard_free(Shard &s) {
	cudaFree(s.M.base);
	cudaFree(s.d_pc[0]);
	cudaFree(s.d_pc[1]);
	cudaFree(s.d_state);
	cudaFree(s.d_pd);
	cudaFree(s.d_hist_r);
	cudaFree(s.d_hist_pm);
	cudaFree(s.d_ebuf);
	cudaFree(s.d_ebuf2);
	cudaFree(s.d_gs);
	cudaFree(s.d_tpanel);
	cudaFree(s.d_cand);
	for (void *p : s.ipc_opened) cudaIpcCloseMemHandle(p);
	s.ipc_opened.clear();
	cudaFree(s.d_dp);
	cudaFree(s.d_hist_owner);
	cudaFree(s.d_pt);
	cudaFree(s.d_x);
	if (s.d_slab_all != s.d_slab) cudaFree(s.d_slab_all);
	cudaFree(s.d_slab);
}

extern "C" void gf2b200_system_destroy(gf2b200_system *sys) {
	if (!sys) return;
	cudaSetDevice(sys->ctx->device);
	for (Shard &s : sys->sh) shard_free(s);
	for (cudaEvent_t e : sys->ev) cudaEventDestroy(e);
	delete sys;
}

static cudaError_t shard_alloc(Shard &s, int world, bool persist) {
	const Mat &M = s.M;
	/* matrix + (sharded) exchange block in ONE allocation */
	const size_t mat_bytes = (size_t)M.ns * (size_t)M.mp * SBYTES;
	cudaError_t e = cudaMalloc(&s.M.base, mat_bytes + (world > 1 ? sizeof(XchBlock) : 0));
	if (e == cudaSuccess) e = cudaMalloc(&s.d_pc[0], (size_t)M.mp * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_pc[1], (size_t)M.mp * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_state, sizeof(SolverState));
	if (e == cudaSuccess) e = cudaMalloc(&s.d_pd, 2 * sizeof(PanelDesc));
	if (e == cudaSuccess) e = cudaMalloc(&s.d_hist_r, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_hist_pm, (size_t)M.nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_ebuf, (size_t)M.ns * EBUF_Q * 16);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_x, (size_t)(M.nw + 1) * 8);
	if (e == cudaSuccess) e = cudaMalloc(&s.d_slab, (size_t)BS_S * 64 * BS_W * 8);
	s.d_slab_all = s.d_slab;
#if SW == 8
	if (persist && world == 1) {
		if (e == cudaSuccess) e = cudaMalloc(&s.d_ebuf2, (size_t)M.ns * EBUF_Q * 16);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_gs, sizeof(GridSync));
		if (e == cudaSuccess) e = cudaMalloc(&s.d_cand, (size_t)2 * PERSIST_CAND_MAX * 2 * 8);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_tpanel, ((size_t)(M.nw + 2) + (PERSIST_TRACE ? (size_t)M.nw * 256 * 8 : 0)) * 8);
	}
#endif
	if (world > 1) {
		s.d_slab_all = nullptr;
		if (e == cudaSuccess) e = cudaMalloc(&s.d_slab_all, (size_t)world * BS_S * 64 * BS_W * 8);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_dp, sizeof(DistPanel));
		if (e == cudaSuccess) e = cudaMalloc(&s.d_hist_owner, (size_t)M.nw * 64);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_pt, sizeof(PeerTable));
		if (e == cudaSuccess) {
			s.xch = reinterpret_cast<XchBlock *>(reinterpret_cast<char *>(s.M.base) + mat_bytes);
			e = cudaMemset(s.xch, 0, sizeof(XchBlock));
		}
	}
	return e;
}

/* Fill every local shard's PeerTable.  Loopback: the peers are the local shards.
 * NCCL context: every rank exports its matrix allocation (matrix + exchange block)
 * as a CUDA IPC handle, the handles travel by ncclAllGather, and each rank maps its
 * peers' allocations -- after this the panel exchange is plain NVLink loads/stores
 * issued by our own kernels. */
static int map_peers(gf2b200_system *sys) {
	gf2b200_ctx *ctx = sys->ctx;
	const int G = ctx->world;
	const long long m = sys->m_global;
	PeerTable pt;
	memset(&pt, 0, sizeof pt);
	for (int g = 0; g < G; g++) {
		long long rows = m * (g + 1) / G - m * g / G;
		pt.mp[g] = (std::max<long long>(rows, 1) + 15) / 16 * 16;
	}
	if (!ctx->nccl) {
		for (Shard &h : sys->sh) {
			pt.base[h.index] = h.M.base;
			pt.xch[h.index] = h.xch;
		}
		for (Shard &h : sys->sh)
			CK(ctx, cudaMemcpy(h.d_pt, &pt, sizeof pt, cudaMemcpyHostToDevice));
		return GF2B200_OK;
	}
	Shard &h = sys->sh[0];
	cudaIpcMemHandle_t mine;
	CK(ctx, cudaIpcGetMemHandle(&mine, h.M.base));
	std::vector<cudaIpcMemHandle_t> all((size_t)G);
	char *d_h = nullptr; /* freed below on every path */
	CK(ctx, cudaMalloc(&d_h, sizeof(mine) * (size_t)(G + 1)));
	cudaError_t e = cudaMemcpyAsync(d_h + sizeof(mine) * (size_t)G, &mine, sizeof mine, cudaMemcpyHostToDevice,
	                                ctx->stream);
	ncclResult_t r = g_nccl.AllGather(d_h + sizeof(mine) * (size_t)G, d_h, sizeof mine, ncclUint8, ctx->nccl,
	                                  ctx->stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * (size_t)G, cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d_h);
	if (r != ncclSuccess) return fail(ctx, GF2B200_ENCCL, "ncclAllGather(ipc handles): %s", g_nccl.GetErrorString(r));
	if (e != cudaSuccess) return fail(ctx, GF2B200_ECUDA, "ipc handle exchange: %s", cudaGetErrorString(e));
	for (int g = 0; g < G; g++) {
		void *p = h.M.base;
		if (g != h.index) {
			e = cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				cudaGetLastError();
				return fail(ctx, GF2B200_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
			}
			h.ipc_opened.push_back(p);
		}
		pt.base[g] = (u64 *)p;
		/* the peer's exchange block sits behind ITS matrix: ns is global, mp is per shard */
		pt.xch[g] = reinterpret_cast<XchBlock *>((char *)p + (size_t)h.M.ns * (size_t)pt.mp[g] * SBYTES);
	}
	CK(ctx, cudaMemcpy(h.d_pt, &pt, sizeof pt, cudaMemcpyHostToDevice));
	return GF2B200_OK;
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
	s->rank = 0;
	s->inconsistent = 0;
	s->eliminated = 0;
	memset(&s->stats, 0, sizeof s->stats);
	s->sh.resize(ctx->n_local);
	cudaError_t e = cudaSuccess;
	for (int l = 0; l < ctx->n_local && e == cudaSuccess; l++) {
		Shard &h = s->sh[l];
		h.index = ctx->rank + l;
		long long r0 = m * h.index / ctx->world, r1 = m * (h.index + 1) / ctx->world;
		h.row_begin = r0;
		Mat &M = h.M;
		M.base = nullptr;
		M.m = r1 - r0;
		M.n = n;
		M.nw = (int)((n + 63) / 64);
		M.ns = (M.nw + 1 + SW - 1) / SW;
		M.mp = (std::max<long long>(M.m, 1) + 15) / 16 * 16;
		h.d_pc[0] = h.d_pc[1] = nullptr;
		h.d_state = nullptr; h.d_pd = nullptr; h.d_hist_r = nullptr; h.d_hist_pm = nullptr;
		h.d_ebuf = nullptr; h.d_dp = nullptr; h.d_hist_owner = nullptr;
		h.d_ebuf2 = nullptr; h.d_gs = nullptr; h.d_tpanel = nullptr; h.d_cand = nullptr;
		h.xch = nullptr; h.d_pt = nullptr;
		h.d_x = h.d_slab = h.d_slab_all = nullptr;
		e = shard_alloc(h, ctx->world, ctx->coop != 0);
	}
	if (e != cudaSuccess) {
		int rc = fail(ctx, e == cudaErrorMemoryAllocation ? GF2B200_ENOMEM : GF2B200_ECUDA,
		              "system_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		gf2b200_system_destroy(s);
		return rc;
	}
	s->epoch_base = 0;
	if (ctx->world > 1) {
		int rc = map_peers(s);
		if (rc) {
			gf2b200_system_destroy(s);
			return rc;
		}
	}
	*out = s;
	return GF2B200_OK;
}

extern "C" int64_t gf2b200_system_local_rows(const gf2b200_system *sys) {
	if (!sys) return -1;
	long long t = 0;
	for (const Shard &s : sys->sh) t += s.M.m;
	return t;
}

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
	if (stride64 < sys->sh[0].M.nw) return fail(ctx, GF2B200_EINVAL, "stride64 < ceil(n/64)");
	CK(ctx, cudaSetDevice(ctx->device));
	const long long base = sys->sh[0].row_begin; /* dA / db start at the first local row */
	for (Shard &h : sys->sh) {
		if (h.M.m == 0) continue;
		long long total = h.M.m * h.M.ns * SW;
		long long off = h.row_begin - base;
		k_layout<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
		    h.M, (const u64 *)dA + off * stride64, (const u64 *)db, stride64, 0, h.M.m, off);
	}
	CK(ctx, cudaGetLastError());
	sys->eliminated = 0;
	return GF2B200_OK;
}

/* ---- host-buffer loads -------------------------------------------------------------
 * Rows travel in chunks through two device staging buffers: the copy of chunk c+1 (copy
 * stream) overlaps the layout kernel of chunk c (solver stream).  Buffers, stream and
 * events belong to the context and are reused by every load.  The load is split in
 * begin / rows / end so that a caller who is still PRODUCING rows (the extension packing
 * PyLongs on worker threads) can hand finished blocks over one by one: H2D and layout of a
 * block overlap the packing of the next (the reference packs everything, then calls
 * M4RI: _internal.c:403-426 followed by :433). */
static int load_error(gf2b200_system *sys, cudaError_t e, const char *where) {
	gf2b200_ctx *ctx = sys->ctx;
	const int rc = fail(ctx, GF2B200_ECUDA, "%s: %s", where, cudaGetErrorString(e));
	if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
	cudaStreamSynchronize(ctx->stream);
	sys->ld_stride = 0;
	return rc;
}

extern "C" int gf2b200_system_load_begin(gf2b200_system *sys, int64_t stride64) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	if (stride64 < sys->sh[0].M.nw) return fail(ctx, GF2B200_EINVAL, "stride64 < ceil(n/64)");
	CK(ctx, cudaSetDevice(ctx->device));
	const long long m_loc = gf2b200_system_local_rows(sys);
	const size_t row_bytes = (size_t)stride64 * 8;
	long long chunk_rows = std::max<long long>(1, (long long)((64u << 20) / row_bytes));
	chunk_rows = std::min<long long>(chunk_rows, std::max<long long>(m_loc, 1));
	const size_t need = (size_t)chunk_rows * row_bytes;
	cudaError_t e = cudaSuccess;
	if (ctx->stage_bytes < need || !ctx->stage[1]) {
		cudaFree(ctx->stage[0]);
		cudaFree(ctx->stage[1]);
		ctx->stage[0] = ctx->stage[1] = nullptr;
		ctx->stage_bytes = 0;
		e = cudaMalloc(&ctx->stage[0], need);
		if (e == cudaSuccess) e = cudaMalloc(&ctx->stage[1], need);
		if (e == cudaSuccess) ctx->stage_bytes = need;
	}
	const size_t b_bytes = (size_t)((m_loc + 63) / 64) * 8;
	if (e == cudaSuccess && ctx->bstage_bytes < b_bytes) {
		cudaFree(ctx->d_bstage);
		ctx->d_bstage = nullptr;
		ctx->bstage_bytes = 0;
		e = cudaMalloc(&ctx->d_bstage, b_bytes);
		if (e == cudaSuccess) ctx->bstage_bytes = b_bytes;
	}
	if (e == cudaSuccess && !ctx->copy_stream) {
		e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
		for (int i = 0; i < 2 && e == cudaSuccess; i++) {
			e = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_laid[i], cudaEventDisableTiming);
		}
	}
	/* the previous load's last layout kernels may still be reading the staging buffers */
	if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_laid[0], ctx->stream);
	if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_laid[0], 0);
	if (e != cudaSuccess) return load_error(sys, e, "system_load_begin");
	sys->ld_stride = stride64;
	sys->ld_chunk_rows = chunk_rows;
	sys->ld_ci = 0;
	sys->ld_used[0] = sys->ld_used[1] = 0;
	sys->eliminated = 0;
	return GF2B200_OK;
}

/* rows [row0, row0 + nrows) of this rank's rows; A_rows points at row row0.  The host memory
 * must stay valid until gf2b200_system_load_end returns.  Any order, no overlaps. */
extern "C" int gf2b200_system_load_rows(gf2b200_system *sys, const uint64_t *A_rows, int64_t row0, int64_t nrows) {
	if (!sys || !A_rows) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	if (!sys->ld_stride) return fail(ctx, GF2B200_EINVAL, "system_load_rows without system_load_begin");
	if (row0 < 0 || nrows < 0 || row0 + nrows > gf2b200_system_local_rows(sys))
		return fail(ctx, GF2B200_EINVAL, "system_load_rows: rows out of range");
	CK(ctx, cudaSetDevice(ctx->device));
	const long long stride64 = sys->ld_stride, chunk_rows = sys->ld_chunk_rows;
	const size_t row_bytes = (size_t)stride64 * 8;
	const long long base = sys->sh[0].row_begin;
	cudaError_t e = cudaSuccess;
	for (Shard &h : sys->sh) {
		/* the part of the block that falls into this shard (loopback contexts hold several) */
		const long long off = h.row_begin - base;
		const long long lo = std::max<long long>(row0, off), hi = std::min<long long>(row0 + nrows, off + h.M.m);
		for (long long r = lo; e == cudaSuccess && r < hi; r += chunk_rows) {
			const long long nr = std::min<long long>(chunk_rows, hi - r);
			const int bi = sys->ld_ci;
			/* the staging buffer is free once the layout kernel that read it is done */
			if (sys->ld_used[bi]) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_laid[bi], 0);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(ctx->stage[bi], A_rows + (r - row0) * stride64, nr * row_bytes,
				                    cudaMemcpyHostToDevice, ctx->copy_stream);
			if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copied[bi], ctx->copy_stream);
			if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[bi], 0);
			if (e != cudaSuccess) break;
			const long long total = nr * h.M.ns * SW;
			k_layout<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(h.M, ctx->stage[bi], nullptr, stride64,
			                                                                      r - off, nr, 0);
			e = cudaGetLastError();
			if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_laid[bi], ctx->stream);
			sys->ld_used[bi] = 1;
			sys->ld_ci ^= 1;
		}
	}
	if (e != cudaSuccess) return load_error(sys, e, "system_load_rows");
	return GF2B200_OK;
}

/* b: the packed bits of the local rows (NULL: homogeneous).  Returns when every copy has left
 * the caller's host buffers; the layout kernels still in flight are ordered before whatever the
 * caller enqueues next on the solver stream (gf2b200_system_eliminate). */
extern "C" int gf2b200_system_load_end(gf2b200_system *sys, const uint64_t *b) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	if (!sys->ld_stride) return fail(ctx, GF2B200_EINVAL, "system_load_end without system_load_begin");
	CK(ctx, cudaSetDevice(ctx->device));
	cudaError_t e = cudaSuccess;
	if (b) {
		const long long m_loc = gf2b200_system_local_rows(sys);
		const long long base = sys->sh[0].row_begin;
		e = cudaMemcpyAsync(ctx->d_bstage, b, (size_t)((m_loc + 63) / 64) * 8, cudaMemcpyHostToDevice, ctx->stream);
		for (Shard &h : sys->sh) {
			if (e != cudaSuccess || h.M.m == 0) continue;
			k_place_b<<<grid_for(h.M.m, 256, ctx->n_sm * 8), 256, 0, ctx->stream>>>(h.M, ctx->d_bstage, h.row_begin - base);
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
	if (e == cudaSuccess && b) e = cudaStreamSynchronize(ctx->stream); /* b travels on the solver stream */
	if (e != cudaSuccess) return load_error(sys, e, "system_load_end");
	sys->ld_stride = 0;
	return GF2B200_OK;
}

extern "C" int gf2b200_system_load_host(gf2b200_system *sys, const uint64_t *A, const uint64_t *b,
                                        int64_t stride64) {
	if (!sys || !A) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	int rc = gf2b200_system_load_begin(sys, stride64);
	if (!rc) rc = gf2b200_system_load_rows(sys, A, 0, gf2b200_system_local_rows(sys));
	if (!rc) rc = gf2b200_system_load_end(sys, b);
	return rc;
}

extern "C" int gf2b200_system_generate(gf2b200_system *sys, uint64_t seed) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	CK(ctx, cudaSetDevice(ctx->device));
	for (Shard &h : sys->sh) {
		const Mat &M = h.M;
		if (M.m == 0) continue;
		long long total = M.m * M.ns * SW;
		k_generate<<<grid_for(total, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(M, seed, h.row_begin);
		/* x* goes through d_x (overwritten later by the solve) */
		k_synth_xstar<<<(M.nw + 255) / 256, 256, 0, ctx->stream>>>(h.d_x, M.nw, M.n, seed);
		k_synth_dot<<<grid_for(M.m * 32, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
		    M, seed, h.row_begin, h.d_x, 0, nullptr);
	}
	CK(ctx, cudaGetLastError());
	sys->eliminated = 0;
	return GF2B200_OK;
}

/* all-gather over the system's shards: `bytes` of every shard's send buffer, in
 * shard order, into every shard's recv buffer (NCCL, or device copies on a
 * loopback context) */
template <typename T, typename U>
static int all_gather(gf2b200_system *sys, T *Shard::*send, U *Shard::*recv, size_t bytes, double *acct) {
	gf2b200_ctx *ctx = sys->ctx;
	if (ctx->nccl) {
		Shard &h = sys->sh[0];
		NK(ctx, g_nccl.AllGather(h.*send, h.*recv, bytes, ncclUint8, ctx->nccl, ctx->stream));
	} else {
		for (Shard &d : sys->sh)
			for (Shard &s : sys->sh)
				CK(ctx, cudaMemcpyAsync((char *)(d.*recv) + (size_t)s.index * bytes, s.*send, bytes,
				                        cudaMemcpyDeviceToDevice, ctx->stream));
	}
	if (acct) *acct += (double)bytes * sys->sh.size();
	return GF2B200_OK;
}

/* forward elimination of a one-shard system (single GPU) */
static int forward_single(gf2b200_system *sys, long long *launches) {
	gf2b200_ctx *ctx = sys->ctx;
	Shard &h = sys->sh[0];
	const Mat &M = h.M;
	cudaStream_t st = ctx->stream;
	const bool prof = ctx->profile != 0;
	const int nw = M.nw;
	k_extract_pc<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, 0, h.d_pc[0], 0);
	(*launches)++;
	const int apply_cap = ctx->n_sm * APPLY_CTAS_PER_SM;
	const bool fused = !getenv("GF2B200_NO_FUSED_SELECT"); /* diagnostic switch */
#if SW == 8
	/* needs the second E-tile buffer and a cooperative launch (both there when k_forward is possible) */
	const bool tail_apply = fused && ctx->coop && h.d_ebuf2 && !getenv("GF2B200_NO_TAIL_APPLY");
#endif
	for (int w = 0; w < nw; w++) {
		u64 colmask = ~0ULL;
		if (w == nw - 1 && (M.n & 63)) colmask = (1ULL << (M.n & 63)) - 1;
		u64 *pc_cur = h.d_pc[w & 1], *pc_next = h.d_pc[(w + 1) & 1];
		/* two descriptions: the sweep of panel w writes panel w+1's while reading its own */
		PanelDesc *pd = h.d_pd + (w & 1), *pdn = (w + 1 < nw && fused) ? h.d_pd + ((w + 1) & 1) : nullptr;
		u64 colmask_next = ~0ULL;
		if (w + 1 == nw - 1 && (M.n & 63)) colmask_next = (1ULL << (M.n & 63)) - 1;
		k_select<<<1, SEL_THREADS, 0, st>>>(M, pc_cur, w, colmask, h.d_state, pd, h.d_hist_r, h.d_hist_pm);
		(*launches)++;
		int s0a = w >> SW_SHIFT;
#if SW == 8
		if (tail_apply) {
			/* k_sweep_apply: the sweep of panel w applies panel w + 1 in its tail (into the other E-tile
			 * buffer); k_apply is then a no-op, as k_select is after a successful look-ahead */
			uint4 *eb = (w & 1) ? h.d_ebuf2 : h.d_ebuf, *eb_next = (w & 1) ? h.d_ebuf : h.d_ebuf2;
			k_apply<<<std::min(M.ns - s0a, apply_cap), APPLY_THREADS, 0, st>>>(M, pd, eb, s0a);
			if (prof) CK(ctx, cudaEventRecord(sys->ev[6 + 2 * w], st));
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeCooperative;
			at[0].val.cooperative = 1;
			cudaLaunchConfig_t cfg = {dim3(ctx->n_sm), dim3(SWEEP_THREADS), SWEEP_SMEM, st, at, 1};
			CK(ctx, cudaLaunchKernelEx(&cfg, k_sweep_apply, M, (const PanelDesc *)pd, (const u64 *)pc_cur, pc_next,
			                           (const uint4 *)eb, eb_next, w, (w + 1) >> SW_SHIFT, pdn, h.d_state, h.d_hist_r,
			                           h.d_hist_pm, colmask_next));
			if (prof) CK(ctx, cudaEventRecord(sys->ev[7 + 2 * w], st));
			*launches += 2;
			continue;
		}
#endif
		k_apply<<<std::min(M.ns - s0a, apply_cap), APPLY_THREADS, 0, st>>>(M, pd, h.d_ebuf, s0a);
		if (prof) CK(ctx, cudaEventRecord(sys->ev[6 + 2 * w], st));
		k_sweep<<<ctx->n_sm, SWEEP_THREADS, SWEEP_SMEM, st>>>(M, pd, pc_cur, pc_next, h.d_ebuf, w,
		                                                     (w + 1) >> SW_SHIFT,
		                                                     pdn, h.d_state, h.d_hist_r, h.d_hist_pm, colmask_next);
		if (prof) CK(ctx, cudaEventRecord(sys->ev[7 + 2 * w], st));
		*launches += 2;
	}
	k_check<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, h.d_state);
	(*launches)++;
	return GF2B200_OK;
}

#if SW == 8
/* systems of at least this many bytes take the launch chain unless GF2B200_FORWARD says otherwise */
#ifndef PERSIST_AUTO_BYTES
#define PERSIST_AUTO_BYTES (256.0 * 1024.0 * 1024.0)
#endif
/* forward elimination of a one-shard system as ONE cooperative kernel (gf2b200_persist.cuh) */
static int forward_single_persist(gf2b200_system *sys, long long *launches) {
	gf2b200_ctx *ctx = sys->ctx;
	Shard &h = sys->sh[0];
	const Mat &M = h.M;
	cudaStream_t st = ctx->stream;
	CK(ctx, cudaMemsetAsync(h.d_gs, 0, sizeof(GridSync), st));
	CK(ctx, cudaMemsetAsync(h.d_tpanel, 0, ((size_t)(M.nw + 2) + (PERSIST_TRACE ? (size_t)M.nw * ctx->n_sm * 8 : 0)) * 8, st));
	k_extract_pc<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, 0, h.d_pc[0], 0);
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeCooperative;
	at[0].val.cooperative = 1;
	cudaLaunchConfig_t cfg = {dim3(ctx->n_sm), dim3(SWEEP_THREADS), PERSIST_SMEM, st, at, 1};
	CK(ctx, cudaEventRecord(sys->ev[3], st));
	CK(ctx, cudaLaunchKernelEx(&cfg, k_forward, M, h.d_pc[0], h.d_pc[1], h.d_ebuf, h.d_ebuf2, h.d_pd, h.d_state,
	                           h.d_hist_r, h.d_hist_pm, (GridSync *)h.d_gs, h.d_tpanel, h.d_cand, 0, M.nw));
	CK(ctx, cudaEventRecord(sys->ev[4], st));
	k_check<<<grid_for(M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(M, h.d_state);
	*launches += 3;
	return GF2B200_OK;
}
#endif

/* forward elimination of a row-sharded system: the per-panel exchange goes through
 * peer memory (gf2b200_dist.cuh); an NCCL context adds the flag waits, a loopback
 * context gets the same ordering from the single stream */
static int forward_sharded(gf2b200_system *sys, long long *launches, double *xbytes) {
	gf2b200_ctx *ctx = sys->ctx;
	cudaStream_t st = ctx->stream;
	const bool prof = ctx->profile != 0;
	const int G = ctx->world;
	const int barriers = ctx->nccl ? 1 : 0;
	const int nw = sys->sh[0].M.nw, ns = sys->sh[0].M.ns;
	const int apply_cap = ctx->n_sm * APPLY_CTAS_PER_SM;
	for (Shard &h : sys->sh) {
		k_extract_pc<<<grid_for(h.M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(h.M, 0, h.d_pc[0], 0);
		(*launches)++;
	}
	if (ctx->nccl) {
		/* Only system_create is collective: a rank may reach this point seconds after its peers
		 * (load_host, first-launch module load).  A small stream-ordered collective makes every
		 * rank's stream arrive here before any kernel starts a timed flag wait. */
		Shard &h0 = sys->sh[0];
		NK(ctx, g_nccl.AllGather(h0.d_slab, h0.d_slab_all, 8, ncclUint8, ctx->nccl, st));
	}
	/* Per panel and shard: [k_select_publish] [k_elect] k_apply_pull k_apply_commit k_sweep_dist.
	 * The first two are no-ops whenever the PREVIOUS sweep's look-ahead already published this
	 * shard's candidates and (one process per GPU) ran the election: the pivot exchange of
	 * panel w+1 then overlaps the sweep of panel w, and what is left between two sweeps is the
	 * pull of the elected rows over NVLink, one flag round (nobody overwrites an elected row
	 * before every peer has pulled it) and the commit. */
	const bool look = !getenv("GF2B200_NO_DIST_LOOKAHEAD"); /* diagnostic switch */
	/* the look-ahead CTA of a sweep runs the local search (~17 us), waits for the peers' candidate
	 * blocks and runs the election (~15-30 us): it is spared that many work units (3.3 us each) */
	int dist_pad = barriers ? 24 : SWEEP_SEL_PAD;
	if (const char *dp = getenv("GF2B200_DIST_PAD")) dist_pad = atoi(dp);
	for (int w = 0; w < nw; w++) {
		u64 colmask = ~0ULL;
		if (w == nw - 1 && (sys->n & 63)) colmask = (1ULL << (sys->n & 63)) - 1;
		u64 colmask_next = ~0ULL;
		if (w + 1 == nw - 1 && (sys->n & 63)) colmask_next = (1ULL << (sys->n & 63)) - 1;
		const int s0a = w >> SW_SHIFT, nsr = ns - s0a;
		const unsigned epoch = sys->epoch_base + (unsigned)w + 1;
		for (Shard &h : sys->sh)
			k_select_publish<<<1, SEL_THREADS, 0, st>>>(h.M, h.d_pc[w & 1], w, colmask, h.d_state, h.d_pt, h.index, G,
			                                            epoch, barriers);
		for (Shard &h : sys->sh)
			k_elect<<<1, 32, 0, st>>>(h.xch, G, h.index, w, colmask, h.d_state, h.d_pd + (w & 1), h.d_dp, h.d_pc[w & 1],
			                          h.d_hist_r, h.d_hist_pm, h.d_hist_owner, epoch, barriers);
		for (Shard &h : sys->sh)
			k_apply_pull<<<std::min(nsr, apply_cap), APPLY_THREADS, 0, st>>>(h.M, h.d_pd + (w & 1), h.d_dp, h.d_pt, h.d_ebuf,
			                                                                s0a, h.d_state, h.index, G, epoch, barriers);
		int li = 0;
		for (Shard &h : sys->sh) {
			k_apply_commit<<<std::min(nsr, apply_cap), APPLY_THREADS, 0, st>>>(h.M, h.d_pd + (w & 1), h.d_dp, h.d_ebuf, s0a,
			                                                                  h.d_state, h.xch, G, epoch, barriers);
			const bool ev = prof && li == 0;
			if (ev) CK(ctx, cudaEventRecord(sys->ev[6 + 2 * w], st));
			DistLook dl;
			dl.pt = h.d_pt;
			dl.xch = h.xch;
			dl.dp = h.d_dp;
			dl.hist_owner = h.d_hist_owner;
			dl.me = h.index;
			dl.G = G;
			dl.epoch_next = epoch + 1;
			dl.barriers = barriers;
			dl.pad = dist_pad;
			PanelDesc *pdn = (look && w + 1 < nw) ? h.d_pd + ((w + 1) & 1) : nullptr;
			k_sweep_dist<<<ctx->n_sm, SWEEP_THREADS, SWEEP_SMEM, st>>>(h.M, h.d_pd + (w & 1), h.d_pc[w & 1],
			                                                          h.d_pc[(w + 1) & 1], h.d_ebuf, w, (w + 1) >> SW_SHIFT,
			                                                          pdn, h.d_state, h.d_hist_r, h.d_hist_pm, colmask_next, dl);
			if (ev) CK(ctx, cudaEventRecord(sys->ev[7 + 2 * w], st));
			li++;
		}
		*launches += 5 * (long long)sys->sh.size();
		/* per shard: candidate blocks stored to G peers + 64 pivot-row pieces pulled per strip */
		*xbytes += (double)sys->sh.size() * ((double)G * CAND_W * 8 + 64.0 * nsr * SBYTES);
	}
	sys->epoch_base += (unsigned)nw;
	for (Shard &h : sys->sh) {
		k_check<<<grid_for(h.M.m, 256, ctx->n_sm * 8), 256, 0, st>>>(h.M, h.d_state);
		(*launches)++;
	}
	CK(ctx, cudaGetLastError());
	return GF2B200_OK;
}

/* blocked back-substitution into every shard's d_x: the particular solution
 * (freecol < 0) or the kernel vector of one free column */
static int backward(gf2b200_system *sys, long long *launches, double *xbytes, long long freecol = -1) {
	gf2b200_ctx *ctx = sys->ctx;
	cudaStream_t st = ctx->stream;
	const int nw = sys->sh[0].M.nw;
	const bool sharded = ctx->world > 1;
	const int nsp = (nw + BS_S - 1) / BS_S;
	for (Shard &h : sys->sh) k_bs_init<<<(nw + 256) / 256, 256, 0, st>>>(h.d_x, nw, freecol);
	for (int P = nsp - 1; P >= 0; --P) {
		for (Shard &h : sys->sh)
			k_bs_outer<<<BS_S * 64 * 32 / 256, 256, 0, st>>>(h.M, h.d_hist_r, h.d_hist_pm,
			                                                sharded ? h.d_hist_owner : nullptr, h.index, h.d_x,
			                                                h.d_slab, P);
		if (sharded) {
			int rc = all_gather(sys, &Shard::d_slab, &Shard::d_slab_all, (size_t)BS_S * 64 * BS_W * 8, xbytes);
			if (rc) return rc;
		}
		for (Shard &h : sys->sh)
			k_bs_inner<<<1, 1024, 0, st>>>(h.d_slab_all, h.d_hist_pm, sharded ? h.d_hist_owner : nullptr, h.d_x,
			                               P, nw);
		*launches += 2 * (long long)sys->sh.size();
	}
	CK(ctx, cudaGetLastError());
	return GF2B200_OK;
}

extern "C" int gf2b200_system_eliminate(gf2b200_system *sys) {
	if (!sys) return fail(nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	cudaStream_t st = ctx->stream;
	CK(ctx, cudaSetDevice(ctx->device));
	const int nw = sys->sh[0].M.nw, ns = sys->sh[0].M.ns;
	const bool prof = ctx->profile != 0;
	const bool sharded = ctx->world > 1;
	size_t need_ev = 6 + (prof ? 2 * (size_t)nw : 0);
	while (sys->ev.size() < need_ev) {
		cudaEvent_t e;
		CK(ctx, cudaEventCreate(&e));
		sys->ev.push_back(e);
	}
	cudaEvent_t ev_begin = sys->ev[0], ev_fwd = sys->ev[1], ev_end = sys->ev[2];
	long long launches = 0;
	double xbytes = 0;

	CK(ctx, cudaEventRecord(ev_begin, st));
	for (Shard &h : sys->sh) {
		CK(ctx, cudaMemsetAsync(h.d_state, 0, sizeof(SolverState), st));
		CK(ctx, cudaMemsetAsync(h.d_pd, 0, 2 * sizeof(PanelDesc), st));
	}
	bool persist = false;
#if SW == 8
	/* Which forward elimination: the persistent kernel saves 2 - 6 us of fixed cost per panel, the
	 * launch chain's k_sweep streams ~3 % faster (same source, but ptxas schedules the loop better
	 * outside the large persistent kernel): per-panel sweeps longer than ~50 us favour the chain.
	 * Measured on one box (profiles/r02_ab.md, calls U - W; k_forward / launch chain): n = 8192
	 * 3.93 / 4.60 ms, n = 32768 21.2 / 22.2 ms, n = 65536 96.3 / 91.6 ms, n = 131072 619 / 601 ms.
	 * (Programmatic dependent launch between the kernels of the chain was measured too: faster at
	 * n = 8192, 4.08 ms, but slower from n = 65536 up, 95.1 and 622 ms -- not kept.) */
	persist = !sharded && ctx->persist && sys->sh[0].d_gs &&
	          (ctx->persist == 2 || (double)sys->sh[0].M.mp * sys->sh[0].M.ns * SBYTES < PERSIST_AUTO_BYTES);
#endif
	int rc;
	if (sharded) rc = forward_sharded(sys, &launches, &xbytes);
#if SW == 8
	else if (persist) rc = forward_single_persist(sys, &launches);
#endif
	else rc = forward_single(sys, &launches);
	if (rc) return rc;
	CK(ctx, cudaEventRecord(ev_fwd, st));
	rc = backward(sys, &launches, &xbytes);
	if (rc) return rc;
	CK(ctx, cudaEventRecord(ev_end, st));
	CK(ctx, cudaGetLastError());

	sys->hist_pm.resize(nw);
	std::vector<SolverState> hs(sys->sh.size());
	CK(ctx, cudaMemcpyAsync(sys->hist_pm.data(), sys->sh[0].d_hist_pm, (size_t)nw * 8, cudaMemcpyDeviceToHost, st));
	if (sharded) {
		sys->hist_owner.resize((size_t)nw * 64);
		CK(ctx, cudaMemcpyAsync(sys->hist_owner.data(), sys->sh[0].d_hist_owner, (size_t)nw * 64,
		                        cudaMemcpyDeviceToHost, st));
	}
	for (size_t l = 0; l < sys->sh.size(); l++) {
		Shard &h = sys->sh[l];
		h.hist_r.resize(nw);
		CK(ctx, cudaMemcpyAsync(h.hist_r.data(), h.d_hist_r, (size_t)nw * 8, cudaMemcpyDeviceToHost, st));
		CK(ctx, cudaMemcpyAsync(&hs[l], h.d_state, sizeof(SolverState), cudaMemcpyDeviceToHost, st));
	}
	std::vector<unsigned long long> tpanel;
#if SW == 8
	GridSync hgs;
	memset(&hgs, 0, sizeof hgs);
	if (persist) {
		tpanel.resize((size_t)nw + 2);
		CK(ctx, cudaMemcpyAsync(tpanel.data(), sys->sh[0].d_tpanel, ((size_t)nw + 2) * 8, cudaMemcpyDeviceToHost, st));
		CK(ctx, cudaMemcpyAsync(&hgs, sys->sh[0].d_gs, sizeof hgs, cudaMemcpyDeviceToHost, st));
	}
#endif
	CK(ctx, cudaStreamSynchronize(st));
#if SW == 8
	if (persist && hgs.fault)
		return fail(ctx, GF2B200_ECUDA, "k_forward: a grid-wide wait timed out (the persistent kernel gave up)");
#if PERSIST_TRACE
	if (persist)
		if (const char *tf = getenv("GF2B200_TRACE_FILE")) {
			const size_t cnt = (size_t)(nw + 2) + (size_t)nw * ctx->n_sm * 8;
			std::vector<unsigned long long> tr(cnt);
			cudaMemcpy(tr.data(), sys->sh[0].d_tpanel, cnt * 8, cudaMemcpyDeviceToHost);
			if (FILE *f = fopen(tf, "wb")) {
				long long hdr[2] = {nw, ctx->n_sm};
				fwrite(hdr, 8, 2, f);
				fwrite(tr.data(), 8, cnt, f);
				fclose(f);
			}
		}
#endif
#endif
	sys->rank = hs[0].r;
	int bad = 0;
	for (const SolverState &s : hs) {
		bad |= s.inconsistent;
		if (s.fault) return fail(ctx, GF2B200_ECUDA, "peer-memory exchange timed out waiting for another rank");
	}
	if (ctx->nccl) {
		/* a shard with an active row "0 = 1" makes the whole system inconsistent */
		int *d_flag = nullptr;
		std::vector<int> flags(ctx->world, 0);
		CK(ctx, cudaMalloc(&d_flag, sizeof(int) * (ctx->world + 1)));
		cudaError_t e = cudaMemcpyAsync(d_flag + ctx->world, &bad, sizeof(int), cudaMemcpyHostToDevice, st);
		ncclResult_t r = g_nccl.AllGather(d_flag + ctx->world, d_flag, sizeof(int), ncclUint8, ctx->nccl, st);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * ctx->world, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		cudaFree(d_flag);
		if (r != ncclSuccess) return fail(ctx, GF2B200_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
		if (e != cudaSuccess) return fail(ctx, GF2B200_ECUDA, "flag exchange: %s", cudaGetErrorString(e));
		for (int f : flags) bad |= f;
	}
	sys->inconsistent = bad;
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
	S.exchange_bytes = xbytes;
	S.panels = nw;
	S.rank = sys->rank;
	S.m_local = gf2b200_system_local_rows(sys);
	/* algorithmic sweep bytes of every local shard; the timed launches (profile
	 * mode) are those of the first local shard */
	for (size_t l = 0; l < sys->sh.size(); l++) {
		const Shard &h = sys->sh[l];
		for (int w = 0; w < nw; w++) {
			const u64 pm = sys->hist_pm[w];
			const int k = __builtin_popcountll(pm);
			int mine = k;
			if (sharded) {
				mine = 0;
				for (int j = 0; j < k; j++) mine += sys->hist_owner[(size_t)w * 64 + j] == h.index;
			}
			const long long r1 = h.hist_r[w] + mine;
			if (k == 0 || r1 >= h.M.m) continue;
			double bytes = 2.0 * (double)(h.M.m - r1) * (double)SBYTES * (double)(ns - ((w + 1) >> SW_SHIFT));
			S.sweep_bytes += bytes;
			S.sweep_launches++;
			if (persist) {
				/* per-panel wall time inside the one kernel (globaltimer at the top of every
				 * panel): sweep + look-ahead search + apply + grid barrier of that panel */
				int wq = w + 1;
				while (wq <= nw && !tpanel[wq]) wq++;
				if (wq <= nw && tpanel[w]) {
					const double pms = (double)(tpanel[wq] - tpanel[w]) * 1e-6;
					S.sweep_bytes_timed += bytes;
					S.sweep_launches_timed++;
					if (pms > S.ms_sweep_max) {
						S.ms_sweep_max = pms;
						S.sweep_bytes_max = bytes;
					}
				}
			} else if (prof && l == 0) {
				CK(ctx, cudaEventElapsedTime(&ms, sys->ev[6 + 2 * w], sys->ev[7 + 2 * w]));
				S.ms_sweep += ms;
				S.sweep_bytes_timed += bytes;
				S.sweep_launches_timed++;
				if (ms > S.ms_sweep_max) {
					S.ms_sweep_max = ms;
					S.sweep_bytes_max = bytes;
				}
			}
		}
	}
	if (persist) {
		/* one launch did all the sweeps: its duration is the denominator of the roofline */
		CK(ctx, cudaEventElapsedTime(&ms, sys->ev[3], sys->ev[4]));
		S.ms_sweep = ms;
		S.forward_kernel_launches = 1;
	}
	return GF2B200_OK;
}

extern "C" int gf2b200_system_stats(const gf2b200_system *sys, gf2b200_stats *out) {
	if (!sys || !out) return GF2B200_EINVAL;
	*out = sys->stats;
	return GF2B200_OK;
}

/* Kernel basis of a row-sharded system: the blocked multi-right-hand-side triangular solve of
 * gf2b200_basis.cuh with F split by rows like the matrix (every shard holds F for its own echelon
 * rows) and ONE small exchange per backward panel.  Every rank of an NCCL context calls this
 * together and gets the whole basis.  Replaces mzd_trsm_upper_left (_internal.c:343). */
struct BasisShard {
	Mat F;
	long long r_loc;
	long long *d_free;
	u64 *d_pc, *d_part, *d_all;
	uint4 *d_ebuf, *d_tile, *d_tiles_all;
	PanelDesc *d_pd;
};

static int basis_sharded(gf2b200_system *sys, const std::vector<long long> &sigma, long long r, long long d,
                         uint64_t *basis_out) {
	gf2b200_ctx *ctx = sys->ctx;
	const int nw = sys->sh[0].M.nw;
	const int G = ctx->world;
	const size_t L = sys->sh.size();
	std::vector<BasisShard> bs(L);
	cudaError_t e = cudaSuccess;
	int rc = GF2B200_OK;
	const int fnw = (int)((d + 63) / 64), fns = (fnw + SW - 1) / SW;
	const size_t tile_bytes = (size_t)fns * EBUF_Q * 16;
	/* batches bound the exchange buffers for huge nullities */
	const long long batch = std::min<long long>(d, std::max<long long>(1, (1LL << 28) / ((long long)nw * 8 * G)));
	for (size_t l = 0; l < L; l++) {
		Shard &h = sys->sh[l];
		BasisShard &b = bs[l];
		memset(&b, 0, sizeof b);
		for (int w = 0; w < nw; w++) {
			const int k = __builtin_popcountll(sys->hist_pm[w]);
			for (int j = 0; j < k; j++) b.r_loc += sys->hist_owner[(size_t)w * 64 + j] == h.index;
		}
		b.F.m = b.r_loc;
		b.F.n = d;
		b.F.nw = fnw;
		b.F.ns = fns;
		b.F.mp = (std::max<long long>(b.r_loc, 1) + 15) / 16 * 16;
		if (e == cudaSuccess) e = cudaMalloc(&b.d_free, (size_t)d * 8);
		if (e == cudaSuccess) e = cudaMalloc(&b.F.base, (size_t)fns * (size_t)b.F.mp * SBYTES);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_pc, (size_t)b.F.mp * 8);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_ebuf, tile_bytes);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_tile, tile_bytes);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_tiles_all, tile_bytes * G);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_pd, sizeof(PanelDesc));
		if (e == cudaSuccess) e = cudaMalloc(&b.d_part, (size_t)batch * nw * 8);
		if (e == cudaSuccess) e = cudaMalloc(&b.d_all, (size_t)batch * nw * 8 * G);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(b.d_free, sigma.data() + r, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream);
		if (e == cudaSuccess && b.r_loc > 0)
			k_basis_gather<<<grid_for(b.r_loc * fns * SW, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
			    h.M, b.F, b.d_free, h.d_hist_r, b.r_loc, d);
	}
	cudaEvent_t tb0 = sys->ev[3], tb1 = sys->ev[4], tb2 = sys->ev[5];
	double bk_bytes = 0;
	long long bk_panels = 0;
	if (e == cudaSuccess) e = cudaEventRecord(tb0, ctx->stream);
	for (int w = nw - 1; w >= 0 && e == cudaSuccess && !rc; --w) {
		const u64 pm = sys->hist_pm[w];
		if (!pm) continue;
		for (size_t l = 0; l < L; l++) {
			Shard &h = sys->sh[l];
			BasisShard &b = bs[l];
			k_basis_tile<<<grid_for((long long)fns * EBUF_Q, 256, ctx->n_sm * 4), 256, 0, ctx->stream>>>(
			    b.F, w, pm, h.hist_r[w], h.d_hist_owner, h.index, b.d_tile);
			h.bk_send = b.d_tile;
			h.bk_recv = b.d_tiles_all;
		}
		rc = all_gather(sys, &Shard::bk_send, &Shard::bk_recv, tile_bytes, nullptr);
		for (size_t l = 0; l < L && !rc; l++) {
			Shard &h = sys->sh[l];
			BasisShard &b = bs[l];
			const long long r_w = h.hist_r[w];
			k_basis_prep_sharded<<<grid_for(std::max<long long>(r_w, (long long)fns * EBUF_Q), 256, ctx->n_sm * 8), 256, 0,
			                       ctx->stream>>>(h.M, b.F, w, pm, r_w, b.d_pc, b.d_ebuf, b.d_pd, b.d_tiles_all, G);
			if (r_w <= 0) continue;
			Mat Fv = b.F;
			Fv.m = r_w; /* this shard's rows above the panel */
			k_sweep<<<ctx->n_sm, SWEEP_THREADS, SWEEP_SMEM, ctx->stream>>>(Fv, b.d_pd, b.d_pc, nullptr, b.d_ebuf, 0, 0, nullptr,
			                                                              nullptr, nullptr, nullptr, 0);
			bk_bytes += 2.0 * (double)r_w * (double)SBYTES * (double)fns;
		}
		bk_panels++;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaEventRecord(tb1, ctx->stream);
	for (long long b0 = 0; e == cudaSuccess && !rc && b0 < d; b0 += batch) {
		const long long nb = std::min(batch, d - b0);
		for (size_t l = 0; l < L; l++) {
			Shard &h = sys->sh[l];
			BasisShard &b = bs[l];
			k_basis_scatter_sharded<<<grid_for(nb * nw, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
			    b.F, h.d_hist_r, h.d_hist_pm, h.d_hist_owner, h.index, b.d_free, nw, b0, nb, b.d_part);
			h.bk_send = b.d_part;
			h.bk_recv = b.d_all;
		}
		rc = all_gather(sys, &Shard::bk_send, &Shard::bk_recv, (size_t)nb * nw * 8, nullptr);
		if (rc) break;
		k_or_parts<<<grid_for(nb * nw, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(bs[0].d_part, bs[0].d_all, nb * nw, G);
		e = cudaGetLastError();
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(basis_out + b0 * nw, bs[0].d_part, (size_t)nb * nw * 8, cudaMemcpyDeviceToHost, ctx->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	}
	if (e == cudaSuccess) e = cudaEventRecord(tb2, ctx->stream);
	if (e == cudaSuccess) e = cudaEventSynchronize(tb2);
	else cudaStreamSynchronize(ctx->stream);
	if (e == cudaSuccess && !rc) {
		float ms1 = 0, ms2 = 0;
		cudaEventElapsedTime(&ms1, tb0, tb1);
		cudaEventElapsedTime(&ms2, tb1, tb2);
		sys->stats.ms_basis_solve = ms1;
		sys->stats.ms_basis_output = ms2;
		sys->stats.basis_sweep_bytes = bk_bytes;
		sys->stats.basis_panels = bk_panels;
	}
	for (BasisShard &b : bs) {
		cudaFree(b.d_free);
		cudaFree(b.F.base);
		cudaFree(b.d_pc);
		cudaFree(b.d_ebuf);
		cudaFree(b.d_tile);
		cudaFree(b.d_tiles_all);
		cudaFree(b.d_pd);
		cudaFree(b.d_part);
		cudaFree(b.d_all);
	}
	for (Shard &h : sys->sh) h.bk_send = h.bk_recv = nullptr;
	if (rc) return rc;
	if (e != cudaSuccess) return fail(ctx, GF2B200_ECUDA, "kernel basis (sharded): %s", cudaGetErrorString(e));
	return GF2B200_OK;
}

extern "C" int gf2b200_system_result(gf2b200_system *sys, int mode, gf2b200_result *out) {
	if (!sys || !out) return fail(sys ? sys->ctx : nullptr, GF2B200_EINVAL, "NULL argument");
	gf2b200_ctx *ctx = sys->ctx;
	memset(out, 0, sizeof *out);
	if (mode != 0 && mode != 1) return fail(ctx, GF2B200_EINVAL, "Invalid mode");
	if (!sys->eliminated) return fail(ctx, GF2B200_EINVAL, "system_result before system_eliminate");
	Shard &h = sys->sh[0];
	const Mat &M = h.M;
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
	cudaError_t e0 = cudaMemcpyAsync(out->origin, h.d_x, (size_t)nw * 8, cudaMemcpyDeviceToHost, ctx->stream);
	long long q = 0;
	for (int w = 0; w < nw; w++) {
		u64 pm = sys->hist_pm[w];
		while (pm) {
			out->pivcols[q++] = (int64_t)w * 64 + __builtin_ctzll(pm);
			pm &= pm - 1;
		}
	}
	if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(ctx->stream);
	if (e0 != cudaSuccess) {
		gf2b200_result_free(out); /* callers do not free the result of a failed call */
		return fail(ctx, GF2B200_ECUDA, "system_result: %s", cudaGetErrorString(e0));
	}
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
		if (ctx->world > 1) {
			/* row-sharded system: the same blocked solve with F split by rows like the matrix and
			 * one small exchange per backward panel (basis_sharded above) */
			const int rc = basis_sharded(sys, sigma, r, d, out->basis);
			if (rc) {
				gf2b200_result_free(out);
				return rc;
			}
			out->kernel_dim = d;
			out->status = GF2B200_OK;
			return GF2B200_OK;
		}
		/* blocked multi-right-hand-side triangular solve on the free-column matrix
		 * (gf2b200_basis.cuh): F = U2, then per panel, last first, the Four-Russians sweep */
		Mat F;
		memset(&F, 0, sizeof F);
		F.m = r;
		F.n = d;
		F.nw = (int)((d + 63) / 64);
		F.ns = (F.nw + SW - 1) / SW;
		F.mp = (std::max<long long>(r, 1) + 15) / 16 * 16;
		long long *d_free = nullptr;
		u64 *d_basis = nullptr, *d_pcF = nullptr;
		uint4 *d_ebufF = nullptr;
		PanelDesc *d_pdF = nullptr;
		/* batches bound the device buffer for huge nullities */
		const long long batch = std::min<long long>(d, std::max<long long>(1, (1LL << 28) / ((long long)nw * 8)));
		cudaError_t e = cudaMalloc(&d_free, (size_t)d * 8);
		if (e == cudaSuccess) e = cudaMalloc(&d_basis, (size_t)batch * nw * 8);
		if (e == cudaSuccess) e = cudaMalloc(&F.base, (size_t)F.ns * (size_t)F.mp * SBYTES);
		if (e == cudaSuccess) e = cudaMalloc(&d_pcF, (size_t)F.mp * 8);
		if (e == cudaSuccess) e = cudaMalloc(&d_ebufF, (size_t)F.ns * EBUF_Q * 16);
		if (e == cudaSuccess) e = cudaMalloc(&d_pdF, sizeof(PanelDesc));
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(d_free, sigma.data() + r, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream);
		cudaEvent_t tb0 = sys->ev[3], tb1 = sys->ev[4], tb2 = sys->ev[5];
		double bk_bytes = 0;
		long long bk_panels = 0;
		if (e == cudaSuccess) e = cudaEventRecord(tb0, ctx->stream);
		if (e == cudaSuccess && r > 0) {
			k_basis_gather<<<grid_for(r * F.ns * SW, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(M, F, d_free, h.d_hist_r, r, d);
			for (int w = nw - 1; w >= 0 && e == cudaSuccess; --w) {
				const u64 pm = sys->hist_pm[w];
				const long long r_w = h.hist_r[w];
				if (!pm || r_w <= 0) continue;
				k_basis_prep<<<grid_for(std::max<long long>(r_w, (long long)F.ns * EBUF_Q), 256, ctx->n_sm * 8), 256, 0,
				               ctx->stream>>>(M, F, w, pm, r_w, d_pcF, d_ebufF, d_pdF);
				Mat Fv = F;
				Fv.m = r_w; /* the rows above the panel */
				k_sweep<<<ctx->n_sm, SWEEP_THREADS, SWEEP_SMEM, ctx->stream>>>(Fv, d_pdF, d_pcF, nullptr, d_ebufF, 0, 0,
				                                                              nullptr, nullptr, nullptr, nullptr, 0);
				bk_bytes += 2.0 * (double)r_w * (double)SBYTES * (double)F.ns;
				bk_panels++;
			}
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaEventRecord(tb1, ctx->stream);
		for (long long b0 = 0; e == cudaSuccess && b0 < d; b0 += batch) {
			long long nb = std::min(batch, d - b0);
			k_basis_scatter<<<grid_for(nb * nw, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
			    F, h.d_hist_r, h.d_hist_pm, d_free, nw, b0, nb, d_basis);
			e = cudaGetLastError();
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(out->basis + b0 * nw, d_basis, (size_t)nb * nw * 8,
				                    cudaMemcpyDeviceToHost, ctx->stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		}
		if (e == cudaSuccess) e = cudaEventRecord(tb2, ctx->stream);
		if (e == cudaSuccess) e = cudaEventSynchronize(tb2);
		if (e == cudaSuccess) {
			float ms1 = 0, ms2 = 0;
			cudaEventElapsedTime(&ms1, tb0, tb1);
			cudaEventElapsedTime(&ms2, tb1, tb2);
			sys->stats.ms_basis_solve = ms1;
			sys->stats.ms_basis_output = ms2;
			sys->stats.basis_sweep_bytes = bk_bytes;
			sys->stats.basis_panels = bk_panels;
		}
		cudaFree(d_free);
		cudaFree(d_basis);
		cudaFree(F.base);
		cudaFree(d_pcF);
		cudaFree(d_ebufF);
		cudaFree(d_pdF);
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
	const int nw = sys->sh[0].M.nw;
	CK(ctx, cudaSetDevice(ctx->device));
	u64 *d_v = nullptr, *d_xs = nullptr;
	unsigned long long *d_cnt = nullptr;
	unsigned long long cnt = 0;
	cudaError_t e = cudaMalloc(&d_v, (size_t)nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&d_xs, (size_t)nw * 8);
	if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 8);
	if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, 8, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, x, (size_t)nw * 8, cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess) {
		/* A (x ^ x*) == 0  <=>  A x == b because b = A x* by construction */
		k_synth_xstar<<<(nw + 255) / 256, 256, 0, ctx->stream>>>(d_xs, nw, sys->n, seed);
		k_xor_vec<<<(nw + 255) / 256, 256, 0, ctx->stream>>>(d_v, d_v, d_xs, nw);
		for (Shard &h : sys->sh)
			if (h.M.m)
				k_synth_dot<<<grid_for(h.M.m * 32, 256, ctx->n_sm * 16), 256, 0, ctx->stream>>>(
				    h.M, seed, h.row_begin, d_v, 1, d_cnt);
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

/* Rows [row0, row0 + nrows) of the dense synthetic system written to
 * HOST memory, row-major, for callers that measure the host-buffer path (bench e2e).
 * A workload generator, not a solver: word(i, w) = mix(seed + PHI*(i*nw + w + 1)),
 * b_i = <A_i, x*>; b is packed relative to row0. */
extern "C" int gf2b200_synth_host(uint64_t *A, uint64_t *b, int64_t row0, int64_t nrows, int64_t n,
                                  uint64_t seed) {
	if (!A || !b || nrows < 0 || n < 1) return fail(nullptr, GF2B200_EINVAL, "bad argument");
	const int64_t nw = (n + 63) / 64;
	std::vector<u64> x((size_t)nw);
	for (int64_t w = 0; w < nw; w++) x[w] = mix64((seed ^ 0xB200ULL) + GF2_PHI * (u64)(w + 1));
	const u64 tail = (n & 63) ? ((1ULL << (n & 63)) - 1) : ~0ULL;
	x[nw - 1] &= tail;
	const int64_t blocks = (nrows + 63) / 64; /* 64 rows = one word of b: no sharing between threads */
	unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 64));
	nt = (unsigned)std::min<int64_t>(nt, std::max<int64_t>(blocks, 1));
	auto work = [&](unsigned t) {
		for (int64_t blk = t; blk < blocks; blk += nt) {
			u64 bw = 0;
			const int64_t i1 = std::min<int64_t>(nrows, (blk + 1) * 64);
			for (int64_t il = blk * 64; il < i1; il++) {
				uint64_t *row = A + il * nw;
				const u64 base = (u64)((row0 + il) * nw);
				u64 acc = 0;
				for (int64_t w = 0; w < nw; w++) {
					u64 v = mix64(seed + GF2_PHI * (base + (u64)w + 1));
					if (w == nw - 1) v &= tail;
					row[w] = v;
					acc ^= v & x[w];
				}
				bw |= (u64)(__builtin_popcountll(acc) & 1) << (il & 63);
			}
			b[blk] = (uint64_t)bw;
		}
	};
	std::vector<std::thread> th;
	for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
	work(0);
	for (auto &t : th) t.join();
	return GF2B200_OK;
}

/* ---- one-shot host-buffer solve (what m4ri_solve's body becomes) ---------- */
/* Repeated solves of one shape (the common case behind LinearSystem) keep their HBM
 * buffers: cudaMalloc/cudaFree of the matrix costs as much as its H2D copy.  open hands
 * out the context's cached system of that shape (or a new one), close solves and puts it
 * back; between the two the caller loads the rows (system_load_host, or begin/rows/end
 * while it is still packing them). */
extern "C" int gf2b200_solve_open(gf2b200_ctx *ctx, int64_t m, int64_t n, gf2b200_system **out) {
	if (!ctx || !out) return fail(ctx, GF2B200_EINVAL, "NULL argument");
	*out = nullptr;
	if (ctx->nccl) return fail(ctx, GF2B200_EINVAL, "gf2b200_solve needs a single-process context");
	gf2b200_system *sys = ctx->cached;
	ctx->cached = nullptr;
	if (sys && (sys->m_global != m || sys->n != n)) {
		gf2b200_system_destroy(sys);
		sys = nullptr;
	}
	if (!sys) {
		const int rc = gf2b200_system_create(ctx, m, n, &sys);
		if (rc) return rc;
	}
	*out = sys;
	return GF2B200_OK;
}

/* mode < 0: give the system up without solving (a load failed). */
extern "C" int gf2b200_solve_close(gf2b200_ctx *ctx, gf2b200_system *sys, int mode, gf2b200_result *out) {
	if (!ctx || !sys) return fail(ctx, GF2B200_EINVAL, "NULL argument");
	if (mode < 0) {
		gf2b200_system_destroy(sys);
		return GF2B200_OK;
	}
	if (!out || (mode != 0 && mode != 1)) {
		gf2b200_system_destroy(sys);
		return fail(ctx, GF2B200_EINVAL, out ? "Invalid mode" : "NULL argument");
	}
	memset(out, 0, sizeof *out);
	int rc = gf2b200_system_eliminate(sys);
	if (!rc) rc = gf2b200_system_result(sys, mode, out);
	if (rc) {
		gf2b200_system_destroy(sys);
		out->status = rc;
	} else {
		if (ctx->cached) gf2b200_system_destroy(ctx->cached);
		ctx->cached = sys;
	}
	return rc;
}

extern "C" int gf2b200_solve(gf2b200_ctx *ctx, const uint64_t *A, const uint64_t *b, int64_t m,
                             int64_t n, int64_t stride64, int mode, gf2b200_result *out) {
	if (!ctx || !A || !out) return fail(ctx, GF2B200_EINVAL, "NULL argument");
	memset(out, 0, sizeof *out);
	if (mode != 0 && mode != 1) return fail(ctx, GF2B200_EINVAL, "Invalid mode");
	gf2b200_system *sys = nullptr;
	int rc = gf2b200_solve_open(ctx, m, n, &sys);
	if (rc) return rc;
	rc = gf2b200_system_load_host(sys, A, b, stride64);
	if (rc) {
		gf2b200_system_destroy(sys);
		out->status = rc;
		return rc;
	}
	return gf2b200_solve_close(ctx, sys, mode, out);
}
