/*
 * gf2b200_dist.cuh -- kernels of the row-sharded (multi-GPU) elimination and of
 * the blocked back-substitution shared by the single- and multi-GPU paths.
 *
 * Sharding: shard g of G holds the contiguous global rows
 * [g*m/G, (g+1)*m/G) in its own strip-major Mat; columns are never split, so all
 * XOR traffic of the sweep stays in local HBM.  The per-panel exchange runs over
 * PEER MEMORY (every rank maps every peer's matrix + exchange block through CUDA
 * IPC; NVLink loads/stores issued by these kernels), not through collectives:
 *
 *   k_select_publish  each shard reduces ITS active rows' panel words to <= 64
 *                     candidate rows (a basis of their span) and STORES the
 *                     candidates' raw panel words + row numbers into every
 *                     peer's exchange block, then signals flag A (release.sys);
 *   k_elect           waits for all peers' flag A (acquire.sys), then every shard
 *                     runs the same deterministic election over the G*64
 *                     candidates (taken round-robin over the shards so rows deplete
 *                     evenly and the pull is spread over all links): the elected
 *                     rows' pivot columns are the panel's
 *                     GLOBAL column rank profile -- what _mzd_pluq reports in Q
 *                     (reference _internal.c:433) -- because the
 *                     union of the local bases spans the global active row space;
 *   k_apply_pull      E = TB * Sel where the elected rows Sel are LOADED straight
 *                     from their owners' matrices over NVLink (the north star's
 *                     "broadcast of each pivot row", pulled by the consumers)
 *                     -> local ebuf for the sweep;
 *   k_peer_barrier    flag B: nobody overwrites an elected row before every
 *                     peer has pulled it;
 *   k_apply_commit    the shard that owned the j-th elected row stores E_j in its
 *                     place; displaced rows move to the vacated positions;
 *   k_sweep           unchanged, on the local active rows.
 *
 * On a loopback context (all shards on one GPU, one stream) the same kernels run
 * phase by phase in stream order and the flag waits are skipped.
 *
 * Back-substitution (k_bs_outer / k_bs_inner) is a blocked triangular solve over
 * super-panels of BS_S panel words; see the comments at those kernels.
 */
#pragma once
#include "gf2b200_kernels.cuh"

namespace gf2b200 {

#define MAX_SHARDS 64
#define CAND_W 104 /* words a shard publishes per panel: count, 64 panel words, 64 row numbers (32 words), pad */

/* Exchange block of one shard; lives right behind its matrix in the same
 * allocation so one IPC handle maps both.  Slot [src] is written by shard src. */
struct XchBlock {
	unsigned flagA[MAX_SHARDS]; /* epoch of the last candidate block published by src */
	unsigned flagB[MAX_SHARDS]; /* epoch up to which src has pulled its pivot rows */
	unsigned flagC[MAX_SHARDS]; /* epoch of the last panel whose sweep src has COMPLETED (its rows may be pulled) */
	u64 cand[MAX_SHARDS][CAND_W];
};

/* Peer mappings of one shard (device resident). */
struct PeerTable {
	u64 *base[MAX_SHARDS];       /* strip-major matrices (self included) */
	XchBlock *xch[MAX_SHARDS];
	long long mp[MAX_SHARDS];    /* padded row counts (strip stride) */
};

/* Per-shard description of its share of the current panel (written by k_elect). */
struct DistPanel {
	int my_cnt;          /* elected rows owned by this shard */
	int my_row[64];      /* local row of my i-th elected row (before the moves) */
	int my_idx_of_j[64]; /* for E row j (pivot-column rank): index i among mine, or -1 */
	int src[64];         /* elected row j: owner shard ... */
	int srow[64];        /* ... and its row number there */
};

#ifndef GF2_EMU
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
#else
/* tests/cpu_emu only: one thread runs everything, plain accesses are ordered */
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ unsigned long long globaltimer_ns() { return (unsigned long long)(emu_now_ms() * 1e6); }
#endif
/* spin until *flag >= epoch (epochs only grow); gives up after DIST_TIMEOUT_NS and reports a
 * fault.  The ranks' streams meet in a small collective before the first wait of an
 * elimination (forward_sharded), so the time-out only has to cover one panel's skew. */
#ifndef DIST_TIMEOUT_NS
#define DIST_TIMEOUT_NS 30000000000ULL
#endif
__device__ __forceinline__ bool wait_flag(const unsigned *flag, unsigned epoch) {
	const unsigned long long t0 = globaltimer_ns();
	while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
		if (globaltimer_ns() - t0 > DIST_TIMEOUT_NS) return false;
		__nanosleep(64);
	}
	return true;
}

/* What the sweep of a sharded system needs for its look-ahead (passed by pointer to a
 * kernel-parameter copy). */
struct DistLook {
	const PeerTable *pt;
	XchBlock *xch;
	DistPanel *dp;
	unsigned char *hist_owner;
	int me, G;
	unsigned epoch_next; /* flag epoch of the NEXT panel's candidate blocks */
	int barriers;        /* 1: one process per GPU (flag waits); 0: loopback shards on one stream */
	int pad;             /* work units the look-ahead CTA is spared */
};
__device__ __forceinline__ int dist_sel_pad(const DistLook *dl) { return dl->pad; }

/* ------------------------------------------------------------------------
 * publish_candidates (all threads of a CTA): the <= 64 rows S.sel[] that the scan selected
 * form a basis of this shard's active panel words; their raw panel words + row numbers go
 * into slot `me` of every peer's exchange block, then flag A.
 * block = [count | 64 raw panel words | 64 local row numbers as ints | pad]
 * ---------------------------------------------------------------------- */
__device__ __forceinline__ void publish_candidates(SelectSmem &S, u64 *blk, const u64 *pc, u64 colmask,
                                                   const PeerTable *__restrict__ pt, int me, int G, unsigned epoch,
                                                   int barriers) {
	const int tid = threadIdx.x;
	if (tid < CAND_W) blk[tid] = 0;
	__syncthreads();
	if (tid < 64) {
		const int n = S.nsel;
		if (tid < n) {
			blk[1 + tid] = __ldcg(pc + S.sel[tid]) & colmask;
			reinterpret_cast<int *>(blk + 65)[tid] = S.sel[tid];
		}
		if (tid == 0) blk[0] = (u64)n;
	}
	__syncthreads();
	for (int t = tid; t < G * CAND_W; t += blockDim.x) {
		const int g = t / CAND_W, i = t - g * CAND_W;
		pt->xch[g]->cand[me][i] = blk[i];
	}
	__threadfence_system();
	__syncthreads();
	if (barriers && tid < G) st_release_sys(&pt->xch[tid]->flagA[me], epoch);
}

/* k_select_publish: the full scan of this shard's active rows + publish.  A no-op when the
 * previous sweep's look-ahead already published this panel's candidates. */
__global__ void __launch_bounds__(SEL_THREADS, 1)
k_select_publish(Mat M, const u64 *__restrict__ pc, int w, u64 colmask, SolverState *st,
                 const PeerTable *__restrict__ pt, int me, int G, unsigned epoch, int barriers) {
	__shared__ SelectSmem S;
	__shared__ u64 blk[CAND_W];
	if (*(const volatile int *)&st->fault) return; /* an earlier wait timed out: the elimination is void */
	if (*(const volatile int *)&st->published == w + 1) return;
	select_init(S);
	__syncthreads();
	select_scan(S, pc, st->r_loc, M.m, colmask);
	publish_candidates(S, blk, pc, colmask, pt, me, G, epoch, barriers);
}

/* ------------------------------------------------------------------------
 * elect_body: global pivot election, ONE WARP, identical on every shard.
 * Candidate code = shard * 64 + slot.  Scratch: S.B, S.TB, S.sel, S.topsel, S.mv_src, S.mv_dst.
 * ---------------------------------------------------------------------- */
__device__ __forceinline__ void elect_body(SelectSmem &S, const XchBlock *__restrict__ xch, int G, int me, int w,
                                           u64 colmask, SolverState *st, PanelDesc *pd, DistPanel *dp,
                                           u64 *__restrict__ pc, long long *hist_r, u64 *hist_pm,
                                           unsigned char *hist_owner, unsigned epoch, int barriers) {
	u64 *B = S.B, *TB = S.TB;
	int *sel = S.sel, *topsel = S.topsel, *mv_src = S.mv_src, *mv_dst = S.mv_dst;
	const int lane = threadIdx.x & 31;
	if (*(volatile int *)&st->fault) return;
	if (barriers) {
		bool ok = true;
		for (int g = lane; g < G; g += 32) ok = ok && wait_flag(&xch->flagA[g], epoch);
		if (!__all_sync(0xffffffffu, ok)) {
			if (lane == 0) st->fault = 1;
			return;
		}
	}
	for (int c = lane; c < 64; c += 32) {
		sel[c] = -1;
		topsel[c] = 0;
	}
	__syncwarp();
	WarpBasis W;
	W.B0 = W.B1 = W.T0 = W.T1 = 0;
	W.pm = 0;
	W.nsel = 0;
	/* Candidates are taken round-robin, ceil(64/G) at a time, starting with shard
	 * w mod G: on a dense system every panel then elects ~64/G rows from EVERY shard,
	 * so the shards deplete evenly and the pivot-row pull of k_apply_pull is spread
	 * over all NVLink ports instead of draining one owner.  Any order gives the same
	 * pivot columns; this one is the same on every shard. */
	const int per = (64 + G - 1) / G;
	for (int q0 = 0; q0 < 64 && W.pm != colmask; q0 += per) {
		for (int jj = 0; jj < G && W.pm != colmask; jj++) {
			const int src = (w + jj) % G;
			const u64 *cb = xch->cand[src];
			const int cnt = (int)__ldcg(cb);
			const int nq = min(per, cnt - q0); /* per <= 32: one candidate per lane */
			if (nq <= 0) continue;
			const u64 mine = (lane < nq) ? __ldcg(cb + 1 + q0 + lane) : 0;
			for (int j = 0; j < nq && W.pm != colmask; j++)
				wb_insert(W, sel, shfl64(mine, j), 0, src * 64 + q0 + j, lane);
		}
	}
	wb_store(W, B, TB, lane);
	__syncwarp();
	const u64 pm = W.pm;
	const int k = W.nsel;
	const long long r_loc = st->r_loc;
	for (int c = lane; c < 64; c += 32) {
		pd->TB[c] = ((pm >> c) & 1) ? TB[c] : 0;
		pd->sel[c] = sel[c];
		const int src = (c < k) ? (sel[c] >> 6) : -1;
		dp->src[c] = src;
		dp->srow[c] = (c < k) ? __ldcg(reinterpret_cast<const int *>(xch->cand[src] + 65) + (sel[c] & 63)) : -1;
		hist_owner[(long long)w * 64 + c] = (c < k) ? (unsigned char)src : 0xFF;
	}
	__syncwarp();
	/* my share, in election order */
	int my_cnt = 0;
	for (int h = 0; h < 2; h++) {
		const int j = lane + 32 * h;
		const bool mine = (j < k) && ((sel[j] >> 6) == me);
		const unsigned bal = __ballot_sync(0xffffffffu, mine);
		const int i = my_cnt + __popc(bal & ((1u << lane) - 1));
		if (mine) {
			dp->my_row[i] = dp->srow[j];
			dp->my_idx_of_j[j] = i;
		} else {
			dp->my_idx_of_j[j] = -1;
		}
		my_cnt += __popc(bal);
	}
	__syncwarp();
	/* local rows r_loc .. r_loc+my_cnt-1 become this shard's echelon rows */
	for (int i = lane; i < my_cnt; i += 32) {
		const int row = dp->my_row[i];
		if (row < r_loc + my_cnt) topsel[row - (int)r_loc] = 1;
	}
	__syncwarp();
	int nvac = 0, ndis = 0;
	for (int h = 0; h < 2; h++) {
		const int i = lane + 32 * h;
		const bool vac = (i < my_cnt) && (dp->my_row[i] >= r_loc + my_cnt);
		const unsigned bv = __ballot_sync(0xffffffffu, vac);
		if (vac) mv_dst[nvac + __popc(bv & ((1u << lane) - 1))] = dp->my_row[i];
		nvac += __popc(bv);
		const bool dis = (i < my_cnt) && !topsel[i];
		const unsigned bd = __ballot_sync(0xffffffffu, dis);
		if (dis) mv_src[ndis + __popc(bd & ((1u << lane) - 1))] = (int)r_loc + i;
		ndis += __popc(bd);
	}
	__syncwarp();
	u64 tmpv[2];
	for (int h = 0; h < 2; h++) {
		const int q = lane + 32 * h;
		tmpv[h] = (q < ndis) ? __ldcg(pc + mv_src[q]) : 0;
	}
	__syncwarp();
	for (int h = 0; h < 2; h++) {
		const int q = lane + 32 * h;
		if (q < ndis) {
			__stcg(pc + mv_dst[q], tmpv[h]);
			pd->mv_src[q] = mv_src[q];
			pd->mv_dst[q] = mv_dst[q];
		}
	}
	if (lane == 0) {
		dp->my_cnt = my_cnt;
		pd->r = r_loc;
		pd->r1 = r_loc + my_cnt;
		pd->k = k;
		pd->nmove = ndis;
		pd->pm = pm;
		st->r += k;
		st->r_loc = r_loc + my_cnt;
		st->elected = w + 1;
		hist_r[w] = r_loc;
		hist_pm[w] = pm;
	}
}

/* k_elect: one warp.  A no-op when the previous sweep's look-ahead already ran this election. */
__global__ void __launch_bounds__(32)
k_elect(const XchBlock *__restrict__ xch, int G, int me, int w, u64 colmask, SolverState *st,
        PanelDesc *pd, DistPanel *dp, u64 *__restrict__ pc, long long *hist_r, u64 *hist_pm,
        unsigned char *hist_owner, unsigned epoch, int barriers) {
	__shared__ SelectSmem S;
	if (*(volatile int *)&st->elected == w + 1) return;
	elect_body(S, xch, G, me, w, colmask, st, pd, dp, pc, hist_r, hist_pm, hist_owner, epoch, barriers);
}

/* The sweep's look-ahead on a sharded system (called by every thread of the CTA that just swept
 * the first active rows; S holds the scan of their new panel words).  If those rows settle this
 * shard's part (full local rank, or no more rows): publish the candidates; and when every rank
 * is its own process, wait for the peers' blocks and run the election here, so that the panel's
 * description is ready before the sweep ends.  Otherwise k_select_publish / k_elect do it. */
__device__ __forceinline__ void dist_lookahead(const DistLook *dl, SelectSmem &S, u64 *pc_next, int wn,
                                               u64 colmask_next, long long r1, long long lim, long long m,
                                               SolverState *st, PanelDesc *pd_next, long long *hist_r, u64 *hist_pm) {
	const int tid = threadIdx.x;
	const bool final_ = (S.pm == colmask_next || lim == m); /* uniform: read after the scan's barrier */
	__syncthreads();
	if (!final_) return;
	/* the candidate block is staged in the scan's (now idle) queue */
	publish_candidates(S, S.qv, pc_next, colmask_next, dl->pt, dl->me, dl->G, dl->epoch_next, dl->barriers);
	if (tid == 0) st->published = wn + 1;
	if (dl->barriers && tid < 32)
		elect_body(S, dl->xch, dl->G, dl->me, wn, colmask_next, st, pd_next, dl->dp, pc_next, hist_r, hist_pm,
		           dl->hist_owner, dl->epoch_next, 1);
}

/* ------------------------------------------------------------------------
 * k_apply_pull: per strip s >= s0: pull the k elected rows' 128-byte pieces from
 * their owners' matrices (peer loads), E_c = XOR_{j in TB[c]} Sel_j -> ebuf[s][c].
 * Nothing is written to any matrix here.
 * ---------------------------------------------------------------------- */
__global__ void __launch_bounds__(APPLY_THREADS)
k_apply_pull(Mat M, const PanelDesc *__restrict__ pd, const DistPanel *__restrict__ dp,
             const PeerTable *__restrict__ pt, uint4 *__restrict__ ebuf, int s0, SolverState *st, int me, int G,
             unsigned epoch, int barriers) {
	__shared__ uint4 Sel[64][SQ];
	__shared__ u64 sTB[64];
	__shared__ int s_last, s_okc;
	if (*(const volatile int *)&st->fault) return;
	const int k = pd->k;
	if (barriers && k > 0) {
		/* every owner has finished sweeping the previous panel (flag C) before its rows are read */
		if (threadIdx.x == 0) s_okc = 1;
		__syncthreads();
		if ((int)threadIdx.x < G && !wait_flag(&pt->xch[me]->flagC[threadIdx.x], epoch - 1)) {
			s_okc = 0;
			st->fault = 1;
		}
		__syncthreads();
		if (!s_okc) return;
	}
	const int tid = threadIdx.x, rr = tid / SQ, ch = tid % SQ;
	if (k > 0) {
	if (tid < 64) sTB[tid] = pd->TB[tid];
	const uint4 *prow = nullptr; /* my elected row's piece of strip 0 on its owner */
	long long pstride = 0;       /* uint4 per strip there */
	if (rr < k) {
		const int src = dp->src[rr];
		pstride = pt->mp[src] * SQ;
		prow = reinterpret_cast<const uint4 *>(pt->base[src]) + (long long)dp->srow[rr] * SQ + ch;
	}
	__syncthreads();
	const uint4 z = make_uint4(0, 0, 0, 0);
	for (int s = s0 + blockIdx.x; s < M.ns; s += gridDim.x) {
		Sel[rr][ch] = (rr < k) ? __ldcg(prow + (long long)s * pstride) : z;
		__syncthreads();
		uint4 acc = z;
		u64 t = sTB[rr];
		while (t) {
			int l = __ffsll((long long)t) - 1;
			t &= t - 1;
			xor4(acc, Sel[l][ch]);
		}
		ebuf[(long long)s * EBUF_Q + rr * SQ + ch] = acc;
		__syncthreads();
	}
	}
	/* flag B ("this shard has pulled its pivot rows of the panel"): the last CTA to finish tells
	 * every peer; k_apply_commit waits for all of them before an elected row is overwritten */
	if (!barriers) return;
	__syncthreads();
	if (tid == 0) {
		__threadfence();
		s_last = (atomicAdd(&st->pull_cnt, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (!s_last) return;
	if (tid == 0) st->pull_cnt = 0;
	__threadfence_system();
	if (tid < G) st_release_sys(&pt->xch[tid]->flagB[me], epoch);
}

/* k_sweep on the local rows of a shard, with the sharded look-ahead (publish + election of the
 * next panel from inside the sweep) */
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
k_sweep_dist(Mat M, const PanelDesc *__restrict__ pd, const u64 *__restrict__ pc_cur, u64 *__restrict__ pc_next,
             const uint4 *__restrict__ ebuf, int w, int s0, PanelDesc *pd_next, SolverState *st, long long *hist_r,
             u64 *hist_pm, u64 colmask_next, DistLook dl) {
	sweep_body(M, pd, pc_cur, pc_next, ebuf, w, s0, pd_next, st, hist_r, hist_pm, colmask_next, &dl);
	/* Flag C.  With the look-ahead a peer learns which of this shard's rows were elected for
	 * panel w+1 while this sweep is still running; it may pull them only once EVERY strip of
	 * them has been swept: the last CTA to finish tells every peer "sweep of panel w done". */
	if (!dl.barriers) return;
	__shared__ int is_last;
	__threadfence_system(); /* this thread's row stores, visible to the peers, before the CTA's ticket */
	__syncthreads();
	if (threadIdx.x == 0) is_last = (atomicAdd(&st->sweep_done, 1u) == gridDim.x - 1);
	__syncthreads();
	if (!is_last) return;
	if (threadIdx.x == 0) st->sweep_done = 0;
	__threadfence_system();
	if ((int)threadIdx.x < dl.G) st_release_sys(&dl.pt->xch[threadIdx.x]->flagC[dl.me], dl.epoch_next - 1);
}

/* k_apply_commit: the owner of elected row j stores E_j (from ebuf) at local row
 * r + (index among its own); displaced rows go to the vacated positions */
__global__ void __launch_bounds__(APPLY_THREADS)
k_apply_commit(Mat M, const PanelDesc *__restrict__ pd, const DistPanel *__restrict__ dp,
               const uint4 *__restrict__ ebuf, int s0, SolverState *st, const XchBlock *xch, int G, unsigned epoch,
               int barriers) {
	__shared__ uint4 Dis[64][SQ];
	__shared__ int ssrc[64], sdst[64], smyidx[64];
	__shared__ int s_ok;
	if (*(const volatile int *)&st->fault) return;
	const int k = pd->k;
	if (k == 0) return;
	if (barriers) {
		/* nobody overwrites an elected row before every peer has pulled it */
		if (threadIdx.x == 0) s_ok = 1;
		__syncthreads();
		if (threadIdx.x < G && !wait_flag(&xch->flagB[threadIdx.x], epoch)) {
			s_ok = 0;
			st->fault = 1;
		}
		__syncthreads();
		if (!s_ok) return;
	}
	const int tid = threadIdx.x, rr = tid / SQ, ch = tid % SQ;
	const long long r = pd->r;
	const int nmove = pd->nmove;
	const u64 pm = pd->pm;
	if (tid < 64) {
		ssrc[tid] = pd->mv_src[tid];
		sdst[tid] = pd->mv_dst[tid];
		smyidx[tid] = dp->my_idx_of_j[tid];
	}
	__syncthreads();
	uint4 *mb = reinterpret_cast<uint4 *>(M.base);
	const bool ispiv = (pm >> rr) & 1;
	const int jrank = __popcll(pm & ((1ULL << rr) - 1));
	const int myi = ispiv ? smyidx[jrank] : -1;
	for (int s = s0 + blockIdx.x; s < M.ns; s += gridDim.x) {
		const long long sb = (long long)s * M.mp;
		if (rr < nmove) Dis[rr][ch] = mb[(sb + ssrc[rr]) * SQ + ch];
		__syncthreads();
		if (myi >= 0) mb[(sb + r + myi) * SQ + ch] = ebuf[(long long)s * EBUF_Q + rr * SQ + ch];
		if (rr < nmove) mb[(sb + sdst[rr]) * SQ + ch] = Dis[rr][ch];
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------
 * Blocked back-substitution of the particular solution (free variables 0;
 * reference _mzd_pluq_solve_left, _internal.c:440-454).
 *
 * The echelon rows of panel w (E_j, RREF inside the panel word) sit at local rows
 * hist_r[w] + i of their owner shard.  Super-panel P covers panel words
 * [P*BS_S, (P+1)*BS_S).  From the last super-panel to the first:
 *   k_bs_outer  (all SMs) for every echelon row of P: y = <row[words beyond P], x>
 *               with x[nw] = 1 standing for the b column, and a copy of the row's
 *               BS_S words inside P  ->  slab[(wp*64 + j)][BS_W]
 *   [exchange]  all-gather of the slabs (multi-GPU only)
 *   k_bs_inner  (one CTA, identical on every shard) solves the BS_S-word
 *               triangular block panel by panel and writes x[P*BS_S ..].
 * ---------------------------------------------------------------------- */
#define BS_S 32 /* panel words per super-panel (2 strips) */
#define BS_W 40 /* slab row: BS_S words, then y at [BS_S], padding */

__global__ void __launch_bounds__(256)
k_bs_outer(Mat M, const long long *__restrict__ hist_r, const u64 *__restrict__ hist_pm,
           const unsigned char *__restrict__ hist_owner, int me, const u64 *__restrict__ x,
           u64 *__restrict__ slab, int P) {
	const int lane = threadIdx.x & 31;
	const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; /* (wp, j) */
	if (gw >= BS_S * 64) return;
	const int wp = gw >> 6, j = gw & 63;
	const int w = P * BS_S + wp;
	u64 *out = slab + (long long)gw * BS_W;
	long long row = -1;
	if (w < M.nw) {
		const u64 pm = hist_pm[w];
		if (j < __popcll(pm)) {
			if (!hist_owner) {
				row = hist_r[w] + j;
			} else if (hist_owner[(long long)w * 64 + j] == me) {
				int i = 0;
				for (int jj = 0; jj < j; jj++) i += (hist_owner[(long long)w * 64 + jj] == me);
				row = hist_r[w] + i;
			}
		}
	}
	if (row < 0) {
		for (int l = lane; l < BS_W; l += 32) out[l] = 0;
		return;
	}
	/* outer part: strips beyond the super-panel, SW lanes per 128-byte piece */
	const int wl = lane & (SW - 1), so = lane >> SW_SHIFT;
	const u64 *rowp = M.base + row * SW + wl;
	u64 acc = 0;
	for (int s = (P + 1) * (BS_S / SW) + so; s < M.ns; s += 32 / SW) {
		const int wd = s * SW + wl;
		if (wd <= M.nw) acc ^= rowp[(long long)s * M.mp * SW] & x[wd];
	}
	int par = __popcll(acc) & 1;
	par = __reduce_xor_sync(0xffffffffu, par);
	/* inner part: the row's words inside the super-panel.  The b word (index nw,
	 * "unknown" fixed to 1) may fall inside the last super-panel: fold it into y. */
	{
		const int wd = P * BS_S + lane;
		const u64 v = (wd <= M.nw) ? M.base[widx(M, row, wd)] : 0;
		const int bpar = (wd == M.nw) ? (int)(v & x[M.nw] & 1) : 0;
		par ^= __reduce_or_sync(0xffffffffu, bpar);
		out[lane] = (wd < M.nw) ? v : 0;
	}
	if (lane == 0) out[BS_S] = (u64)(par & 1);
	if (lane > 0 && lane < BS_W - BS_S) out[BS_S + lane] = 0;
}

__global__ void __launch_bounds__(1024, 1)
k_bs_inner(const u64 *__restrict__ slab_all, const u64 *__restrict__ hist_pm,
           const unsigned char *__restrict__ hist_owner, u64 *__restrict__ x, int P, int nw) {
	__shared__ u64 xs[BS_S];
	__shared__ unsigned long long newbits;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nwp = min(BS_S, nw - P * BS_S);
	/* the free-variable bits of these words (all 0 for the particular solution, one
	 * bit set somewhere for a kernel vector) are part of the right-hand side */
	if (tid < BS_S) xs[tid] = (tid < nwp) ? x[P * BS_S + tid] : 0;
	if (tid == 0) newbits = 0;
	__syncthreads();
	for (int wp = nwp - 1; wp >= 0; --wp) {
		const int w = P * BS_S + wp;
		const u64 pm = hist_pm[w];
		const int k = __popcll(pm);
		for (int j = warp; j < k; j += 32) {
			const int src = hist_owner ? hist_owner[(long long)w * 64 + j] : 0;
			const u64 *row = slab_all + ((long long)src * (BS_S * 64) + wp * 64 + j) * BS_W;
			/* word wp itself: the row is 0 at the panel's other pivot columns (RREF inside
			 * the panel) and xs[wp] holds only free-variable bits until the panel is done */
			u64 a = (lane >= wp) ? (row[lane] & xs[lane]) : 0;
			int par = __popcll(a) & 1;
			par = __reduce_xor_sync(0xffffffffu, par) ^ (int)(row[BS_S] & 1);
			if (lane == 0 && par) {
				u64 t = pm;
				for (int q = 0; q < j; q++) t &= t - 1;
				atomicOr(&newbits, t & (~t + 1));
			}
		}
		__syncthreads();
		if (tid == 0) {
			xs[wp] |= newbits;
			newbits = 0;
		}
		__syncthreads();
	}
	if (tid < nwp) x[P * BS_S + tid] = xs[tid];
}

/* Right-hand side of the back-substitution.  freecol < 0: the particular solution,
 * x[0..nw) = 0 and x[nw] = 1 (the b column's "unknown").  freecol >= 0: the kernel
 * vector that is 1 at that free column and 0 at the other free columns
 * (_internal.c:330-348), x = e_freecol and x[nw] = 0 (homogeneous). */
__global__ void k_bs_init(u64 *x, int nw, long long freecol) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > nw) return;
	u64 v = 0;
	if (freecol < 0) v = (i == nw) ? 1 : 0;
	else if (i == (int)(freecol >> 6)) v = 1ULL << (freecol & 63);
	x[i] = v;
}

} /* namespace gf2b200 */
