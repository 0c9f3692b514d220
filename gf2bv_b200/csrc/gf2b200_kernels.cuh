/*
 * gf2b200_kernels.cuh -- sm_100a kernels of the B200 GF(2) solver.
 *
 * What this replaces in the reference: the M4RI calls behind
 * gf2bv/_internal.c:433 (_mzd_pluq), :440 (_mzd_pluq_solve_left) and :343
 * (mzd_trsm_upper_left inside _mzd_kernel_left_pluq, :309-357).  The algorithm
 * is NOT M4RI's recursive PLE; it is a right-looking blocked elimination with
 * 64-column panels designed around HBM streaming:
 *
 *   HBM layout ("strip-major"): the augmented matrix [A | b] is cut into column
 *   strips of SW = 8 words (64 B; GF2_STRIP_WORDS = 16 keeps the first geometry).
 *   Strip s holds, for every row i, the 64-byte piece words 8s..8s+7 of that row,
 *   rows contiguous:  word(i, w) lives at
 *       base[((w / SW) * mp + i) * SW + (w % SW)].
 *   A sweep work unit (one strip x 1024 rows) is therefore ONE contiguous 64 KiB
 *   region: perfectly coalesced 128-bit loads/stores, no strided DRAM pages, and a
 *   row piece is half a 128-byte shared-memory line of the lookup tables.
 *   b sits alone in word nw (bit 0).
 *
 *   per panel (word column w):
 *     k_select  one CTA: XOR-basis insertion over the dense panel-column array
 *               pc[] (warp ballots compact the rows that survive reduction, one
 *               warp inserts them into a register-resident basis with warp-wide
 *               REDUX.XOR); yields <= 64 pivot rows, the pivot-column mask (= the
 *               panel's column rank profile) and the 64x64 transform TB that
 *               turns the selected rows into RREF.  Usually a no-op: k_sweep of
 *               the previous panel already ran this search on the first 1024
 *               active rows while the other SMs kept sweeping.
 *     k_apply   per strip: E = TB * Sel (reduced pivot rows), stores them at rows
 *               r..r+k-1 (physical swap with the displaced rows) and into the
 *               L2-resident staging tile ebuf[s] (64 x 64 B, indexed by column).
 *     k_sweep   persistent, 1 CTA / SM: TMA bulk-copies ebuf[s] into shared
 *               memory (one strip ahead), builds eight Four-Russians tables stored
 *               as four line pairs (128 KiB), then streams every active row piece:
 *               8 conflict-free table XORs per 16 B.  This is the HBM-bound kernel
 *               (2 * rows * 64 B per strip).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gf2b200 {

typedef unsigned long long u64;

#define GF2_PHI 0x9E3779B97F4A7C15ULL

/* Strip geometry (compile-time): GF2_STRIP_WORDS = 16 -> 128-byte strips, nine
 * lookup tables of whole 128-byte lines; 8 -> 64-byte strips, eight tables stored
 * as four line PAIRS (see k_sweep). */
#ifndef GF2_STRIP_WORDS
#define GF2_STRIP_WORDS 8
#endif
#if GF2_STRIP_WORDS == 16
#define SW 16        /* 64-bit words per strip piece */
#define SW_SHIFT 4
#define SQ 8         /* 16-byte chunks per strip piece */
#define SBYTES 128   /* bytes per strip piece */
#elif GF2_STRIP_WORDS == 8
#define SW 8
#define SW_SHIFT 3
#define SQ 4
#define SBYTES 64
#else
#error "GF2_STRIP_WORDS must be 16 or 8"
#endif

struct Mat {
	u64 *base;      /* strip-major storage, ns * mp * SW words */
	long long mp;   /* padded row count (multiple of 16) */
	long long m;    /* rows held here */
	long long n;    /* unknowns */
	int nw;         /* ceil(n / 64): A words per row */
	int ns;         /* strips: ceil((nw + 1) / SW) */
};

/* Device-resident description of the current panel (written by k_select). */
struct PanelDesc {
	long long r;     /* echelon rows before this panel */
	long long r1;    /* first active (not yet pivot) row after this panel */
	int k;           /* pivots found in this panel */
	int nmove;       /* displaced rows to relocate */
	int valid;       /* panel index + 1 this description belongs to (0: none) */
	int applied;     /* panel index + 1 if the sweep that settled this panel also applies it in its tail
	                  * (k_sweep_apply: k_apply is then a no-op); 0 otherwise */
	u64 pm;          /* pivot column mask within the panel word */
	u64 TB[64];      /* by column c: combination of selected rows giving E_c; 0 if c is free */
	int sel[64];     /* physical row of the l-th selected row */
	int mv_src[64];
	int mv_dst[64];
	unsigned look;   /* panel index + 1 once the look-ahead search of the previous sweep has decided (valid or not) */
	int pad2;
};

struct SolverState {
	long long r;     /* current (global) rank */
	long long r_loc; /* first active row of this shard (== r on a single GPU) */
	int inconsistent;
	int fault;       /* a peer-memory flag wait timed out */
	unsigned sweep_done; /* (unused: the tail-select variant of k_sweep was measured neutral and removed) */
	/* row-sharded systems (gf2b200_dist.cuh): panel + 1 whose candidates this shard has already
	 * published / whose election it has already run from inside the previous sweep, and the
	 * CTA counter of the pivot-row pull */
	int published, elected;
	unsigned pull_cnt;
};

struct DistLook; /* look-ahead of the sharded path inside the sweep (gf2b200_dist.cuh) */
__device__ __forceinline__ int dist_sel_pad(const DistLook *dl);
__device__ __forceinline__ void dist_lookahead(const DistLook *dl, struct SelectSmem &S, u64 *pc_next, int wn,
                                               u64 colmask_next, long long r1, long long lim, long long m,
                                               SolverState *st, PanelDesc *pd_next, long long *hist_r, u64 *hist_pm);

__host__ __device__ __forceinline__ u64 mix64(u64 z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
	z ^= z >> 27; z *= 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return z;
}

__device__ __forceinline__ long long widx(const Mat &M, long long i, int w) {
	return ((long long)(w >> SW_SHIFT) * M.mp + i) * SW + (w & (SW - 1));
}

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
	unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
	unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
	return ((u64)hi << 32) | lo;
}

/* ------------------------------------------------------------------------
 * Loading: row-major words -> strip-major, masking bits >= n (the reference
 * ignores them, _internal.c:45,48) and placing b (bit i of the packed b) in
 * word nw.  One thread per (row, word); 16 consecutive lanes write one 128 B piece.
 * ---------------------------------------------------------------------- */
__global__ void k_layout(Mat M, const u64 *__restrict__ src, const u64 *__restrict__ bsrc,
                         long long stride, long long row0, long long nrows, long long brow0) {
	const int WT = M.ns * SW;
	long long total = nrows * WT;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		long long il = t / WT;
		int w = (int)(t - il * WT);
		u64 v = 0;
		if (w < M.nw) {
			v = src[il * stride + w];
			if (w == M.nw - 1 && (M.n & 63)) v &= (1ULL << (M.n & 63)) - 1;
		} else if (w == M.nw && bsrc) {
			long long bi = brow0 + il;
			v = (bsrc[bi >> 6] >> (bi & 63)) & 1;
		}
		M.base[widx(M, row0 + il, w)] = v;
	}
}

/* right-hand side of a host-buffer load: bit (brow0 + i) of bsrc -> bit 0 of word nw of local row i
 * (k_layout left that word 0; b arrives last, when the caller has finished producing the rows) */
__global__ void k_place_b(Mat M, const u64 *__restrict__ bsrc, long long brow0) {
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M.m; i += (long long)gridDim.x * blockDim.x) {
		const long long bi = brow0 + i;
		M.base[widx(M, i, M.nw)] = (bsrc[bi >> 6] >> (bi & 63)) & 1;
	}
}

/* Synthetic dense system of SURVEY.md 8(d): word(i, w) = mix(seed + PHI*(i*nw + w + 1)),
 * i = GLOBAL row index (grow0 + local). */
__global__ void k_generate(Mat M, u64 seed, long long grow0) {
	const int WT = M.ns * SW;
	long long total = M.m * WT;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		long long il = t / WT;
		int w = (int)(t - il * WT);
		u64 v = 0;
		if (w < M.nw) {
			v = mix64(seed + GF2_PHI * (u64)((grow0 + il) * M.nw + w + 1));
			if (w == M.nw - 1 && (M.n & 63)) v &= (1ULL << (M.n & 63)) - 1;
		}
		M.base[widx(M, il, w)] = v;
	}
}

/* parity(<A_i, vec>) with A regenerated from the seed (no matrix reads).
 * mode 0: write it as b into word nw of row i.  mode 1: count rows where it is 1.
 * One warp per row. */
__global__ void k_synth_dot(Mat M, u64 seed, long long grow0, const u64 *__restrict__ vec,
                            int mode, unsigned long long *count) {
	int lane = threadIdx.x & 31;
	long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
	long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long il = warp; il < M.m; il += nwarps) {
		u64 acc = 0;
		for (int w = lane; w < M.nw; w += 32) {
			u64 v = mix64(seed + GF2_PHI * (u64)((grow0 + il) * M.nw + w + 1));
			if (w == M.nw - 1 && (M.n & 63)) v &= (1ULL << (M.n & 63)) - 1;
			acc ^= v & vec[w];
		}
		int p = __popcll(acc) & 1;
		p = __reduce_xor_sync(0xffffffffu, p);
		if (lane == 0) {
			if (mode == 0) M.base[widx(M, il, M.nw)] = (u64)p;
			else if (p) atomicAdd(count, 1ULL);
		}
	}
}

/* x* of the synthetic system: nw words from the generator with seed ^ 0xB200 */
__global__ void k_synth_xstar(u64 *x, int nw, long long n, u64 seed) {
	int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nw) return;
	u64 v = mix64((seed ^ 0xB200ULL) + GF2_PHI * (u64)(w + 1));
	if (w == nw - 1 && (n & 63)) v &= (1ULL << (n & 63)) - 1;
	x[w] = v;
}

__global__ void k_xor_vec(u64 *dst, const u64 *a, const u64 *b, int nw) {
	int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w < nw) dst[w] = a[w] ^ b[w];
}

/* dense copy of word column w for rows [row0, m) */
__global__ void k_extract_pc(Mat M, int w, u64 *__restrict__ pc, long long row0) {
	for (long long i = row0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M.m;
	     i += (long long)gridDim.x * blockDim.x)
		pc[i] = M.base[widx(M, i, w)];
}

/* ------------------------------------------------------------------------
 * k_select: pivot search for panel word w over the active rows [r, m).
 *
 * Maintains an RREF XOR-basis B[c] (keyed by lowest set bit = leftmost column)
 * of the 64-bit panel slices seen so far; TB[c] records which selected rows were
 * XORed to form B[c].  The set of keys after all active rows are absorbed is the
 * panel's column rank profile (an invariant of the row space), which is what
 * _mzd_pluq's Q reports (_internal.c:433).  Stops early once every
 * valid column of the panel is a pivot.
 *
 * All 32 warps reduce 1024 rows against the basis snapshot and compact the
 * survivors (warp ballot); warp 0 then inserts them one at a time into a
 * register-resident basis (WarpBasis: warp-wide REDUX.XOR reduction).
 * ---------------------------------------------------------------------- */
#define SEL_THREADS 1024

/* XOR basis of <= 64 panel words held in the registers of ONE warp: lane L keeps
 * the basis vectors keyed by columns L and L+32 (0 while the column is not a
 * pivot) and, beside each, the set of selected rows that were XORed to form it.
 * The basis is kept in RREF, so a candidate is reduced against ALL of it at once:
 * every lane offers its vectors where the candidate has the key bit, and two
 * warp-wide REDUX.XOR pairs fold the offers -- no serial walk over set bits. */
struct WarpBasis {
	u64 B0, B1, T0, T1;
	u64 pm;   /* pivot columns so far (uniform) */
	int nsel; /* selected rows so far (uniform) */
};

__device__ __forceinline__ u64 warp_xor64(u64 v) {
	unsigned lo = __reduce_xor_sync(0xffffffffu, (unsigned)v);
	unsigned hi = __reduce_xor_sync(0xffffffffu, (unsigned)(v >> 32));
	return ((u64)hi << 32) | lo;
}

__device__ __forceinline__ void wb_load(WarpBasis &W, const u64 *B, const u64 *TB, u64 pm, int nsel, int lane) {
	W.B0 = B[lane]; W.B1 = B[lane + 32];
	W.T0 = TB[lane]; W.T1 = TB[lane + 32];
	W.pm = pm; W.nsel = nsel;
}

__device__ __forceinline__ void wb_store(const WarpBasis &W, u64 *B, u64 *TB, int lane) {
	B[lane] = W.B0; B[lane + 32] = W.B1;
	TB[lane] = W.T0; TB[lane + 32] = W.T1;
}

/* v, tv, row uniform across the warp.  Returns true if v raised the rank. */
__device__ __forceinline__ bool wb_insert(WarpBasis &W, int *sel, u64 v, u64 tv, int row, int lane) {
	const u64 m0 = 0ULL - ((v >> lane) & 1), m1 = 0ULL - ((v >> (lane + 32)) & 1);
	/* both folds are issued together (four independent REDUX): this loop is a serial
	 * chain on the critical path of every panel, one REDUX latency per candidate */
	const u64 rv = warp_xor64((W.B0 & m0) ^ (W.B1 & m1));
	const u64 rt = warp_xor64((W.T0 & m0) ^ (W.T1 & m1));
	v ^= rv;
	if (!v) return false;
	tv ^= rt;
	const int c = __ffsll((long long)v) - 1;
	tv ^= 1ULL << W.nsel;
	if ((W.B0 >> c) & 1) { W.B0 ^= v; W.T0 ^= tv; }
	if ((W.B1 >> c) & 1) { W.B1 ^= v; W.T1 ^= tv; }
	if (lane == (c & 31)) {
		if (c < 32) { W.B0 = v; W.T0 = tv; }
		else { W.B1 = v; W.T1 = tv; }
	}
	if (lane == 0) sel[W.nsel] = row;
	W.pm |= 1ULL << c;
	W.nsel++;
	return true;
}

/* Shared scan body: absorbs rows [r, m) of pc into the basis.  Returns via smem. */
struct SelectSmem {
	u64 B[64], TB[64];
	u64 qv[SEL_THREADS], qtv[SEL_THREADS];
	int qrow[SEL_THREADS];
	int sel[64];
	int topsel[64];
	int mv_src[64], mv_dst[64];
	int wcount[32];
	u64 pm;
	int nsel;
};

/* list != nullptr: scan a compacted candidate list instead of the dense column -- entry i in
 * [r, m) is the pair {row, panel word} at list[2i], list[2i+1] (k_forward's slow path) */
__device__ __forceinline__ void select_scan(SelectSmem &S, const u64 *__restrict__ pc, long long r,
                                            long long m, u64 colmask, const u64 *list = nullptr) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (long long base = r; base < m; base += SEL_THREADS) {
		u64 pm = S.pm;
		if (pm == colmask) break;
		long long i = base + tid;
		/* ld.cg: inside the persistent kernel another SM wrote these words moments ago */
		u64 v;
		int row = (int)i;
		if (list) {
			v = (i < m) ? (__ldcg(list + 2 * i + 1) & colmask) : 0;
			if (i < m) row = (int)__ldcg(list + 2 * i);
		} else {
			v = (i < m) ? (__ldcg(pc + i) & colmask) : 0;
		}
		u64 tv = 0;
		u64 x = v & pm;
		while (x) {
			int c = __ffsll((long long)x) - 1;
			x &= x - 1;
			v ^= S.B[c];
			tv ^= S.TB[c];
		}
		unsigned bal = __ballot_sync(0xffffffffu, v != 0);
		if (lane == 0) S.wcount[warp] = __popc(bal);
		__syncthreads();
		int off = 0, total = 0;
#pragma unroll
		for (int q = 0; q < 32; q++) {
			int cnt = S.wcount[q];
			if (q < warp) off += cnt;
			total += cnt;
		}
		if (v) {
			int pos = off + __popc(bal & ((1u << lane) - 1));
			S.qv[pos] = v;
			S.qtv[pos] = tv;
			S.qrow[pos] = row;
		}
		__syncthreads();
		if (warp == 0) {
			/* survivors, one at a time, against the register-resident basis */
			WarpBasis W;
			wb_load(W, S.B, S.TB, pm, S.nsel, lane);
			/* the next candidate is read while the current one is inserted */
			u64 cv = S.qv[0], ct = S.qtv[0];
			int cr = S.qrow[0];
			for (int q = 0; q < total && W.pm != colmask; q++) {
				const int qn = min(q + 1, SEL_THREADS - 1);
				const u64 nv = S.qv[qn], nt = S.qtv[qn];
				const int nr = S.qrow[qn];
				wb_insert(W, S.sel, cv, ct, cr, lane);
				cv = nv; ct = nt; cr = nr;
			}
			wb_store(W, S.B, S.TB, lane);
			__syncwarp(); /* every lane has read S.nsel / S.pm before lane 0 replaces them */
			if (lane == 0) {
				S.pm = W.pm;
				S.nsel = W.nsel;
			}
		}
		__syncthreads();
	}
}

__device__ __forceinline__ void select_init(SelectSmem &S) {
	const int tid = threadIdx.x;
	if (tid < 64) {
		S.B[tid] = 0;
		S.TB[tid] = 0;
		S.sel[tid] = -1;
		S.topsel[tid] = 0;
	}
	if (tid == 0) {
		S.pm = 0;
		S.nsel = 0;
	}
}

/* Warp 0: turn the finished scan into the panel description.  Rows r..r+k-1
 * become the echelon rows; selected rows already there stay, the others
 * ("displaced") move to the positions vacated by selected rows. */
__device__ __forceinline__ void select_finalize(SelectSmem &S, u64 *__restrict__ pc, int w, long long r,
                                                SolverState *st, PanelDesc *pd, long long *hist_r,
                                                u64 *hist_pm) {
	const int lane = threadIdx.x & 31;
	const u64 pm = S.pm;
	const int k = S.nsel;
	for (int c = lane; c < 64; c += 32) {
		pd->TB[c] = ((pm >> c) & 1) ? S.TB[c] : 0;
		pd->sel[c] = S.sel[c];
	}
	for (int l = lane; l < k; l += 32) {
		int srow = S.sel[l];
		if (srow < r + k) S.topsel[srow - (int)r] = 1;
	}
	__syncwarp();
	int nvac = 0, ndis = 0;
	for (int h = 0; h < 2; h++) {
		int l = lane + 32 * h;
		bool vac = (l < k) && (S.sel[l] >= r + k);
		unsigned bv = __ballot_sync(0xffffffffu, vac);
		if (vac) S.mv_dst[nvac + __popc(bv & ((1u << lane) - 1))] = S.sel[l];
		nvac += __popc(bv);
		bool dis = (l < k) && !S.topsel[l];
		unsigned bd = __ballot_sync(0xffffffffu, dis);
		if (dis) S.mv_src[ndis + __popc(bd & ((1u << lane) - 1))] = (int)r + l;
		ndis += __popc(bd);
	}
	__syncwarp();
	/* keep the dense panel column consistent with the row moves */
	u64 tmpv[2];
	for (int h = 0; h < 2; h++) {
		int q = lane + 32 * h;
		tmpv[h] = (q < ndis) ? __ldcg(pc + S.mv_src[q]) : 0;
	}
	__syncwarp();
	for (int h = 0; h < 2; h++) {
		int q = lane + 32 * h;
		if (q < ndis) {
			__stcg(pc + S.mv_dst[q], tmpv[h]);
			pd->mv_src[q] = S.mv_src[q];
			pd->mv_dst[q] = S.mv_dst[q];
		}
	}
	if (lane == 0) {
		pd->r = r;
		pd->r1 = r + k;
		pd->k = k;
		pd->nmove = ndis;
		pd->pm = pm;
		pd->valid = w + 1;
		st->r = r + k;
		st->r_loc = r + k;
		hist_r[w] = r;
		hist_pm[w] = pm;
	}
}

/* Pivot search of panel w.  Usually a no-op: the sweep of panel w-1 already ran
 * the search on the first 1024 active rows (see k_sweep) and, when that found a
 * full set of pivots, marked the description valid.  Otherwise scan everything. */
__global__ void __launch_bounds__(SEL_THREADS, 1)
k_select(Mat M, u64 *__restrict__ pc, int w, u64 colmask, SolverState *st, PanelDesc *pd,
         long long *hist_r, u64 *hist_pm) {
	__shared__ SelectSmem S;
	if (pd->valid == w + 1) return;
	const long long r = st->r, m = M.m;
	select_init(S);
	__syncthreads();
	select_scan(S, pc, r, m, colmask);
	if (threadIdx.x >= 32) return;
	select_finalize(S, pc, w, r, st, pd, hist_r, hist_pm);
}

/* ------------------------------------------------------------------------
 * k_apply: per strip s >= s0: E_c = XOR_{l in TB[c]} Sel_l for every pivot column
 * c, stored (a) at matrix row r + rank(c) and (b) in ebuf[s][c] (zero rows for
 * free columns) for the sweep's TMA load; displaced rows go to the vacated
 * positions.  512 threads = 64 rows x 8 x 16 B.
 * ---------------------------------------------------------------------- */
#define APPLY_THREADS (64 * SQ)
#ifndef APPLY_CTAS_PER_SM
#define APPLY_CTAS_PER_SM (2048 / APPLY_THREADS) /* grid cap: a full SM of threads, strips strided over the grid */
#endif

__global__ void __launch_bounds__(APPLY_THREADS)
k_apply(Mat M, const PanelDesc *__restrict__ pd, uint4 *__restrict__ ebuf, int s0) {
	__shared__ uint4 Sel[64][SQ];
	__shared__ uint4 Dis[64][SQ];
	__shared__ u64 sTB[64];
	__shared__ int ssel[64], ssrc[64], sdst[64];
	const int k = pd->k;
	if (k == 0) return;
	if (pd->applied != 0 && pd->applied == pd->valid) return; /* the previous sweep's tail applied this panel */
	const int tid = threadIdx.x, rr = tid / SQ, ch = tid % SQ;
	const long long r = pd->r;
	const int nmove = pd->nmove;
	const u64 pm = pd->pm;
	if (tid < 64) {
		sTB[tid] = pd->TB[tid];
		ssel[tid] = pd->sel[tid];
		ssrc[tid] = pd->mv_src[tid];
		sdst[tid] = pd->mv_dst[tid];
	}
	__syncthreads();
	uint4 *mb = reinterpret_cast<uint4 *>(M.base);
	const int erow = (int)r + __popcll(pm & ((1ULL << rr) - 1));
	const bool ispiv = (pm >> rr) & 1;
	for (int s = s0 + blockIdx.x; s < M.ns; s += gridDim.x) {
		const long long sb = (long long)s * M.mp;
		uint4 z = make_uint4(0, 0, 0, 0);
		Sel[rr][ch] = (rr < k) ? mb[(sb + ssel[rr]) * SQ + ch] : z;
		if (rr < nmove) Dis[rr][ch] = mb[(sb + ssrc[rr]) * SQ + ch];
		__syncthreads();
		uint4 acc = z;
		u64 t = sTB[rr];
		while (t) {
			int l = __ffsll((long long)t) - 1;
			t &= t - 1;
			uint4 v = Sel[l][ch];
			acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
		}
		ebuf[(long long)s * (64 * SQ) + rr * SQ + ch] = acc;
		if (ispiv) mb[(sb + erow) * SQ + ch] = acc;
		if (rr < nmove) mb[(sb + sdst[rr]) * SQ + ch] = Dis[rr][ch];
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------
 * k_sweep: the HBM-bound row-XOR sweep.
 *   rows [r1, m) x strips [s0, ns):  piece ^= XOR_g T_g[field g of (pc_cur[row] & pm)]
 * Persistent grid (1 CTA of 1024 threads per SM).  A work unit is SWEEP_RU rows x one
 * strip (64 KiB contiguous); units are dealt out contiguously in strip-major
 * order so a CTA rebuilds its tables only when it enters a new strip.
 *
 * Four-Russians tables, laid out for conflict-free 128-bit lookups.  A lookup
 * wavefront is a quarter-warp (8 threads x 16 B).  In the first version (64-byte
 * strips, plain 64-byte entries) it held 2 rows x 4 chunks and the two rows
 * collided whenever their indices had equal parity: ncu (profiles/r01a) showed 33%
 * of all shared wavefronts were such replays and the l1tex data pipe, not HBM, was
 * the limiter.  Two layouts avoid that:
 *   SW = 16  a strip piece and a table entry are both one 128-byte line, the 8
 *            threads of a quarter-warp are the 8 chunks of ONE row.  Whole lines
 *            cost capacity: eight 7-bit fields + one 8-bit field, 9 lookups per
 *            16 B, 160 KiB; build scratch and the fused pivot search's scratch
 *            alias table space.
 *   SW = 8   (default) entries are half lines stored in PAIRS of tables, the two
 *            rows of a quarter-warp walk a pair in opposite order and therefore
 *            always sit in opposite bank halves: eight full 8-bit fields, 8
 *            lookups per 16 B, 128 KiB + 24 KiB of scratch behind the tables (E tile
 *            and partial tables during a build, SelectSmem during the fused search).
 * Both stay under the 164 KiB carve-out so the SM keeps its L1 for the loads in flight.
 * The CTA that updates the strip holding word w+1 also emits the dense copy of
 * that word column (pc_next) for the next panel's pivot search.
 * ---------------------------------------------------------------------- */
#define SWEEP_THREADS 1024
#ifndef SWEEP_U
#define SWEEP_U 4 /* row pieces in flight per thread (2, 3, 6 measured slower) */
#endif
#define SWEEP_RU (SWEEP_THREADS / SQ * SWEEP_U) /* rows per unit */
#define EBUF_Q (64 * SQ) /* uint4 per strip in ebuf */
/* SWEEP_EARLY_TILE (64-byte strips only: it needs the E tile outside the tables): request the next
 * strip's E tile (TMA) during the current table build -- hides the TMA round trip of the strip
 * changes, 622.6 -> 608.8 ms at n = 131072 (profiles/r01f_ab.txt).  Measured and removed (DESIGN.md
 * section 3, profiles/r02_ab.md call A): one coefficient load per row + warp shuffle (-2 %), row loads not
 * gated by the coefficient (neutral here; k_forward and the lean unit do it), row loads issued before
 * the table build (+1 %). */
#ifndef SWEEP_EARLY_TILE
#define SWEEP_EARLY_TILE (SW == 8)
#endif
#if SWEEP_EARLY_TILE && SW == 16
#error "SWEEP_EARLY_TILE needs GF2_STRIP_WORDS=8"
#endif
#if SW == 16
#define SWEEP_LINES (8 * 128 + 256)
#define SWEEP_PARTS (8 * 24 + 32)
/* 160 KiB of tables + the mbarrier: the whole CTA stays under the 164 KiB
 * shared-memory carve-out, so the SM keeps 92 KiB of L1 for the streaming loads
 * in flight (a 205 KiB layout measured 20% slower at the 228 KiB carve-out). */
#define SWEEP_SMEM (SWEEP_LINES * 128 + 16)
#else
/* 64-byte strips: eight 256-entry tables of 64-byte entries stored as four line
 * PAIRS -- line e of pair p holds T_{2p}[e] in bytes 0..63 and T_{2p+1}[e] in bytes
 * 64..127.  A lookup wavefront (quarter-warp) is 2 rows x 4 chunks; the two rows
 * visit the two tables of a pair in OPPOSITE order, so at every step one row reads
 * banks 0..15 and the other banks 16..31 whatever their indices are: 8 lookups per
 * 16 B, all single-wavefront, 128 KiB of tables.  The E tile (4 KiB) and the nibble
 * partial tables (16.5 KiB) have their own space: 148.5 KiB, still under the
 * 164 KiB carve-out. */
#define SWEEP_LINES (4 * 256)
#define SWEEP_PSTRIDE 132 /* uint4 per table in P: 32 entries x 4 chunks + 64 B skew (tables 2p / 2p+1 on opposite bank halves) */
#define SWEEP_BUILD_BYTES (EBUF_Q * 16 + 8 * SWEEP_PSTRIDE * 16)
#define SWEEP_SCRATCH_BYTES 24576 /* E | P during a build, the fused pivot search's SelectSmem otherwise */
#define SWEEP_SMEM (SWEEP_LINES * 128 + SWEEP_SCRATCH_BYTES + 16)
#endif

#ifndef GF2_EMU
__device__ __forceinline__ unsigned smem_u32(const void *p) {
	return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
	             "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned phase) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "W_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@!p bra W_%=;\n\t}" ::"r"(smem_u32(bar)),
	    "r"(phase)
	    : "memory");
}
/* 1-D TMA: global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	        smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(smem_u32(bar))
	    : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
#else
/* tests/cpu_emu only (the kernels' source executed on the CPU; never in libgf2b200.so):
 * the copy completes at once and flips the barrier's phase bit */
__device__ __forceinline__ void mbar_init(void *bar, unsigned) { emu_mbar_init(bar); }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_expect_tx(void *, unsigned) {}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned phase) { emu_mbar_wait(bar, phase); }
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
	emu_bulk_copy(dst, src, bytes, bar);
}
#endif

__device__ __forceinline__ void xor4(uint4 &a, const uint4 &b) {
	a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w;
}

#if SW == 16
/* Builds TD from the pivot-row tile E (64 columns x 128 B).  The tile sits in the
 * first lines of field 0 and is dead once the partial tables P exist; P sits
 * inside field 8's area (the last 32 KiB of TD), which is therefore written last,
 * from registers.  All SWEEP_THREADS threads; ends with a __syncthreads. */
__device__ __forceinline__ void sweep_build_tables(uint4 *TD, uint4 *P, const uint4 *E, int tid) {
	/* partial tables: field g < 8 (columns 7g..7g+6): 8 entries over its low 3
	 * columns, 16 over its high 4; field 8 (columns 56..63): 16 + 16 */
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int it = tid + q * SWEEP_THREADS;
		if (it < SWEEP_PARTS * 8) {
			const int id = it >> 3, c8 = it & 7;
			int col0, e;
			if (id < 192) {
				const int g = id / 24, r = id - g * 24;
				if (r < 8) { col0 = 7 * g; e = r; }
				else { col0 = 7 * g + 3; e = r - 8; }
			} else {
				const int r = id - 192;
				col0 = 56 + (r & 16 ? 4 : 0);
				e = r & 15;
			}
			uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
			for (int b = 0; b < 4; b++)
				if ((e >> b) & 1) xor4(acc, E[(col0 + b) * 8 + c8]);
			P[it] = acc;
		}
	}
	__syncthreads();
	/* fields 0..7: item = (line, chunk of the 128-byte line) */
	for (int it = tid; it < 1024 * 8; it += SWEEP_THREADS) {
		const int L = it >> 3, c8 = it & 7;
		const int g = L >> 7, e = L & 127;
		uint4 a = P[(g * 24 + (e & 7)) * 8 + c8];
		xor4(a, P[(g * 24 + 8 + (e >> 3)) * 8 + c8]);
		TD[it] = a;
	}
	/* field 8 overwrites the scratch: values to registers, barrier, then store */
	uint4 f8[2];
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int it = tid + q * SWEEP_THREADS; /* 256 lines x 8 */
		const int e = it >> 3, c8 = it & 7;
		f8[q] = P[(192 + (e & 15)) * 8 + c8];
		xor4(f8[q], P[(208 + (e >> 4)) * 8 + c8]);
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < 2; q++) TD[1024 * 8 + tid + q * SWEEP_THREADS] = f8[q];
	__syncthreads();
}

static_assert(sizeof(SelectSmem) <= 1024 * 128, "pivot-search scratch must fit inside the tables of fields 0..7");
static_assert(SWEEP_PARTS * 8 * 16 <= 256 * 128, "partial tables must fit inside field 8's lines");
static_assert(EBUF_Q * 16 <= 128 * 128, "the E tile must fit inside field 0's lines");
#else
/* Builds the four line pairs from the pivot-row tile E (64 columns x 64 B).
 * Step 1: per table t (columns 8t..8t+7) 16 combinations of its low 4 columns and
 * 16 of its high 4 (one item per thread).  Step 2: entry e = lo[e & 15] ^ hi[e >> 4],
 * written into the half-line of its table.  Ends with a __syncthreads. */
__device__ __forceinline__ void sweep_build_tables(uint4 *TD, uint4 *P, uint4 *E, int tid,
                                                   const uint4 *next_tile, void *bar) {
	{
		const int t = tid >> 7, e5 = (tid >> 2) & 31, c4 = tid & 3;
		const int col0 = 8 * t + (e5 & 16 ? 4 : 0), e = e5 & 15;
		uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int b = 0; b < 4; b++)
			if ((e >> b) & 1) xor4(acc, E[(col0 + b) * 4 + c4]);
		P[t * SWEEP_PSTRIDE + e5 * 4 + c4] = acc;
	}
	__syncthreads();
	/* E is dead from here on: fetch the tile of the strip this CTA enters next */
	if (next_tile && tid == 0) {
		mbar_expect_tx(bar, EBUF_Q * 16);
		tma_bulk_g2s(E, next_tile, EBUF_Q * 16, bar);
	}
	/* thread = (chunk, table of the pair, low nibble, half of the high nibbles, pair):
	 * the low-nibble part is read once, then 8 entries are one LDS + one STS each; the 8
	 * threads of a quarter-warp write one whole 128-byte line */
	{
		const int c4 = tid & 3, tpar = (tid >> 2) & 1, lo = (tid >> 3) & 15, hh = (tid >> 7) & 1, pr = tid >> 8;
		const uint4 *Pt = P + (2 * pr + tpar) * SWEEP_PSTRIDE + c4;
		const uint4 base = Pt[lo * 4];
		uint4 *dst = TD + ((pr * 256 + hh * 128 + lo) * 8 + tpar * 4 + c4);
#pragma unroll
		for (int j = 0; j < 8; j++) {
			uint4 a = Pt[(16 + hh * 8 + j) * 4];
			xor4(a, base);
			dst[j * 16 * 8] = a; /* entry e = (hh*8 + j)*16 + lo */
		}
	}
	__syncthreads();
}

static_assert(SWEEP_THREADS == 8 * 32 * 4, "one partial-table item per thread; (4 chunks, 2, 16, 2, 4 pairs) in the second step");
#endif
#ifndef SWEEP_SEL_PAD
#define SWEEP_SEL_PAD 6 /* units CTA 0 is spared to make room for the fused pivot search (2: 631, 4: 626, 6: 622 ms) */
#endif


/* gpu-scope release / acquire of a flag word (look-ahead verdicts inside a launch, k_forward's barriers) */
#ifndef GF2_EMU
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#else
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#endif

#if SW == 8
/* ---- the lean streaming unit (64-byte strips) -------------------------------------------
 * Units that lie entirely inside the active rows of a strip other than the next panel word's --
 * all but a few hundred of the 32768 units of a large panel -- take this path instead of the
 * general one: no row-range predicates, 32-bit shared-memory addresses computed once per
 * thread, per lookup one byte extract (PRMT, ALU) + one multiply-add (IMAD, FMA pipe) + LDS.128.
 * ncu (profiles/r02z_sweep_ncu.md against r01g): 426 M -> 235 M instructions per launch, issue
 * slots 53 % -> 30 % busy, and the SAME duration -- the loop is bound by the bytes it moves through
 * the l1tex data pipe (88 %), not by what it executes; the lean form is kept for the halved
 * instruction stream.  Measured and dropped on the way (profiles/r02_ab.md
 * calls S - T): the loop behind a call boundary, all eight loads of a unit issued before the first
 * lookup (ptxas' own schedule, which requests row piece q + 2 while piece q is looked up, is the
 * better one), a rolling pipeline over units, a TMA bulk prefetch of the next units into L2. */
#ifndef SWEEP_LEAN_UNITS
#define SWEEP_LEAN_UNITS 1
#endif

#ifndef GF2_EMU
/* a plain (weak, L1-allocating) global load the compiler can neither drop nor turn into ld.global.nc */
__device__ __forceinline__ u64 ld_weak_u64(const u64 *p) {
	u64 v;
	asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
typedef unsigned smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void *p) { return (smem_addr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(smem_addr_t a) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
/* a row load the compiler can neither sink below the test of its coefficient nor drop (the row
 * loads must not wait for the coefficient: two chained L2 latencies per unit, 600 -> 582 ms) */
__device__ __forceinline__ uint4 ldcg128_now(const uint4 *p) {
	uint4 v;
	asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
#else
__device__ __forceinline__ u64 ld_weak_u64(const u64 *p) { return *p; }
__device__ __forceinline__ uint4 ldcg128_now(const uint4 *p) { return *p; }
typedef uintptr_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void *p) { return (smem_addr_t)p; }
__device__ __forceinline__ uint4 lds128(smem_addr_t a) { return *reinterpret_cast<const uint4 *>(a); }
#endif

/* one row piece: v ^= the eight table entries selected by the coefficient bytes (tables of a pair
 * visited in opposite order by the two rows of a quarter-warp: te / to and bsel differ by row parity) */
__device__ __forceinline__ void lean_piece(uint4 &v, u64 cf, smem_addr_t te, smem_addr_t to, unsigned bsel) {
	const unsigned lo = __byte_perm((unsigned)cf, 0, bsel);
	const unsigned hi = __byte_perm((unsigned)(cf >> 32), 0, bsel);
#define LEAN_PAIR(i, cw, j0, j1)                                                       \
	{                                                                                  \
		const uint4 a = lds128(te + (i) * 32768 + __byte_perm(cw, 0, 0x4440 | (j0)) * 128u); \
		const uint4 b = lds128(to + (i) * 32768 + __byte_perm(cw, 0, 0x4440 | (j1)) * 128u); \
		v.x ^= a.x ^ b.x;                                                              \
		v.y ^= a.y ^ b.y;                                                              \
		v.z ^= a.z ^ b.z;                                                              \
		v.w ^= a.w ^ b.w;                                                              \
	}
	LEAN_PAIR(0, lo, 0, 1)
	LEAN_PAIR(1, lo, 2, 3)
	LEAN_PAIR(2, hi, 0, 1)
	LEAN_PAIR(3, hi, 2, 3)
#undef LEAN_PAIR
}

/* a unit whose SWEEP_RU rows are all active and whose strip is not the next panel word's:
 * p = this thread's first row piece, pcp = its first coefficient */
__device__ __forceinline__ void lean_unit(uint4 *__restrict__ p, const u64 *__restrict__ pcp, u64 pm, smem_addr_t te,
                                          smem_addr_t to, unsigned bsel) {
	u64 cf[SWEEP_U];
	uint4 d[SWEEP_U];
#pragma unroll
	for (int q = 0; q < SWEEP_U; q++) cf[q] = ld_weak_u64(pcp + (SWEEP_THREADS / SQ) * q) & pm;
#pragma unroll
	for (int q = 0; q < SWEEP_U; q++) d[q] = ldcg128_now(p + (SWEEP_THREADS / SQ) * q * SQ);
#pragma unroll
	for (int q = 0; q < SWEEP_U; q++) {
		/* branch-free: a zero coefficient looks up the eight zero entries, and the row is not written
		 * (a branch here made ptxas sink the row load below the test of its coefficient) */
		lean_piece(d[q], cf[q], te, to, bsel);
		if (cf[q] != 0) __stcg(p + (SWEEP_THREADS / SQ) * q * SQ, d[q]);
	}
}

#endif /* SW == 8 */

__device__ __forceinline__ void
sweep_body(Mat M, const PanelDesc *__restrict__ pd, const u64 *__restrict__ pc_cur,
           u64 *__restrict__ pc_next, const uint4 *__restrict__ ebuf, int w, int s0,
           PanelDesc *pd_next, SolverState *st, long long *hist_r, u64 *hist_pm, u64 colmask_next,
           const DistLook *dl = nullptr, bool publish_look = false) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint4 *TD = reinterpret_cast<uint4 *>(smem_raw);
#if SW == 16
	uint4 *E = TD;            /* build scratch: the E tile (8 KiB), inside field 0 ... */
	uint4 *P = TD + 1024 * 8; /* ... and the partial tables (28 KiB), inside field 8 */
	u64 *bar = reinterpret_cast<u64 *>(TD + SWEEP_LINES * 8);
#else
	uint4 *E = TD + SWEEP_LINES * 8; /* the E tile and the partial tables follow the tables */
	uint4 *P = E + EBUF_Q;
	u64 *bar = reinterpret_cast<u64 *>(smem_raw + SWEEP_LINES * 128 + SWEEP_SCRATCH_BYTES);
	static_assert(SWEEP_BUILD_BYTES <= SWEEP_SCRATCH_BYTES && sizeof(SelectSmem) <= SWEEP_SCRATCH_BYTES,
	              "build scratch and pivot-search scratch share the region behind the tables");
#endif
	/* scratch of the fused pivot search of panel w+1 (pd_next != nullptr): over the tables */
#if SW == 16
	SelectSmem &S = *reinterpret_cast<SelectSmem *>(smem_raw);
#else
	SelectSmem &S = *reinterpret_cast<SelectSmem *>(smem_raw + SWEEP_LINES * 128); /* tables stay intact */
#endif

	const int tid = threadIdx.x;
	const int k = pd->k;
	const long long r1 = pd->r1;
	const long long m = M.m;
	if (r1 >= m) return;
	const int wn = w + 1; /* next panel word (or the b word): always exists */
	/* pc_next == nullptr: a sweep that feeds no later panel (the kernel-basis solve): no strip is "next" */
	const int snext = pc_next ? (wn >> SW_SHIFT) : -1;
	if (k == 0) {
		/* nothing to eliminate: only hand the next word column to k_select */
		if (!pc_next) return;
		for (long long i = r1 + blockIdx.x * (long long)SWEEP_THREADS + tid; i < m;
		     i += (long long)gridDim.x * SWEEP_THREADS)
			pc_next[i] = M.base[widx(M, i, wn)];
		return;
	}
	const u64 pm = pd->pm;
	const long long rows = m - r1;
	const long long nchunks = (rows + SWEEP_RU - 1) / SWEEP_RU;
	const long long units = (long long)(M.ns - s0) * nchunks;
	/* unit 0 = (strip holding word w+1, first SWEEP_RU active rows) carries the fused
	 * pivot search: its CTA gets SWEEP_SEL_PAD fewer units */
	/* (sharded systems: the search CTA also waits for the peers' candidates and runs the election) */
	const long long vpad = pd_next ? max(0LL, min((long long)(dl ? dist_sel_pad(dl) : SWEEP_SEL_PAD), units / gridDim.x - 1)) : 0;
	const long long vunits = units + vpad;
	const long long u0 = max(0LL, vunits * blockIdx.x / gridDim.x - vpad);
	const long long u1 = vunits * (blockIdx.x + 1) / gridDim.x - vpad;
	if (u0 >= u1) return;
	if (tid == 0) {
		mbar_init(bar, 1);
		mbar_init_fence();
	}
	__syncthreads();
	unsigned phase = 0;
	int cur = -1;
	const int ch = tid % SQ, rl = tid / SQ;  /* chunk of the strip piece, row in the pass */
	const int nch = (wn & (SW - 1)) >> 1;    /* chunk holding word wn inside its strip */
	uint4 *mb = reinterpret_cast<uint4 *>(M.base);
#if SW == 16
	const unsigned char *Tb = reinterpret_cast<const unsigned char *>(TD + ch);
#else
	/* the two rows of a quarter-warp take the tables of a pair in opposite order:
	 * row parity h reads half h at even steps and half 1-h at odd steps, and sees
	 * its coefficient with adjacent bytes swapped when h = 1 */
	const int h = rl & 1;
	const unsigned char *Tbe = reinterpret_cast<const unsigned char *>(TD + 4 * h + ch);
	const unsigned char *Tbo = reinterpret_cast<const unsigned char *>(TD + 4 * (1 - h) + ch);
	const unsigned bsel = h ? 0x2301u : 0x3210u;
#endif

#if SW == 8
	const int s_last = s0 + (int)((u1 - 1) / nchunks); /* last strip this CTA touches */
	int fetched = -1;                                  /* strip whose E tile is already on its way */
#endif
	/* (re)build the tables when unit u lies in another strip than the previous one */
	auto enter_strip = [&](int s, long long u) {
		if (s == cur) return;
		__syncthreads(); /* everyone is done with the previous tables */
#if SW == 16
		if (tid == 0) {
			mbar_expect_tx(bar, EBUF_Q * 16);
			tma_bulk_g2s(E, ebuf + (long long)s * EBUF_Q, EBUF_Q * 16, bar);
		}
		mbar_wait(bar, phase);
		phase ^= 1;
		sweep_build_tables(TD, P, E, tid);
#else
		if (fetched != s && tid == 0) {
			mbar_expect_tx(bar, EBUF_Q * 16);
			tma_bulk_g2s(E, ebuf + (long long)s * EBUF_Q, EBUF_Q * 16, bar);
		}
		mbar_wait(bar, phase);
		phase ^= 1;
		/* not for the unit that carries the fused pivot search: its scratch shares the
		 * E region and wants a quiet barrier */
		const bool more = SWEEP_EARLY_TILE && (s < s_last) && !(pd_next && u == 0);
		sweep_build_tables(TD, P, E, tid, more ? ebuf + (long long)(s + 1) * EBUF_Q : nullptr, bar);
		fetched = more ? s + 1 : -1;
#endif
		cur = s;
	};
	for (long long u = u0; u < u1; ++u) {
		const int s = s0 + (int)(u / nchunks);
		const long long chunk = u % nchunks;
		enter_strip(s, u);
		const long long row0 = r1 + chunk * SWEEP_RU + rl;
		const bool force = (s == snext);
		uint4 *p = mb + ((long long)s * M.mp + row0) * SQ + ch;
#if SW == 8 && SWEEP_LEAN_UNITS
		if (!force && chunk < rows / SWEEP_RU) {
			/* every consecutive unit of this strip whose SWEEP_RU rows are all active */
			const long long nl = min(rows / SWEEP_RU - chunk, u1 - u);
			const smem_addr_t te32 = smem_addr(Tbe), to32 = smem_addr(Tbo);
			const u64 *pcp = pc_cur + row0;
#pragma unroll 1
			for (long long i = 0; i < nl; i++, p += (long long)SWEEP_RU * SQ, pcp += SWEEP_RU)
				lean_unit(p, pcp, pm, te32, to32, bsel);
			u += nl - 1;
			continue;
		}
#endif
		u64 cf[SWEEP_U];
		uint4 d[SWEEP_U];
		bool act[SWEEP_U];
#pragma unroll
		for (int q = 0; q < SWEEP_U; q++) {
			long long row = row0 + (SWEEP_THREADS / SQ) * q;
			cf[q] = (row < m) ? (__ldg(pc_cur + row) & pm) : 0;
		}
#pragma unroll
		for (int q = 0; q < SWEEP_U; q++) {
			long long row = row0 + (SWEEP_THREADS / SQ) * q;
			act[q] = (row < m) && (cf[q] != 0 || force);
			if (act[q]) d[q] = __ldcg(p + (long long)(SWEEP_THREADS / SQ) * q * SQ);
		}
#pragma unroll
		for (int q = 0; q < SWEEP_U; q++) {
			if (!act[q]) continue;
			uint4 v = d[q];
#if SW == 16
			const unsigned lo = (unsigned)cf[q], hi = (unsigned)(cf[q] >> 32);
			const unsigned mid = __funnelshift_r(lo, hi, 28);
			/* byte offset of a line = 128 * (field base + field value) */
#define TLOOK(off) (*reinterpret_cast<const uint4 *>(Tb + (off)))
			xor4(v, TLOOK(0 * 16384 + ((lo << 7) & 0x3F80u)));
			xor4(v, TLOOK(1 * 16384 + (lo & 0x3F80u)));
			xor4(v, TLOOK(2 * 16384 + ((lo >> 7) & 0x3F80u)));
			xor4(v, TLOOK(3 * 16384 + ((lo >> 14) & 0x3F80u)));
			xor4(v, TLOOK(4 * 16384 + ((mid << 7) & 0x3F80u)));
			xor4(v, TLOOK(5 * 16384 + ((hi << 4) & 0x3F80u)));
			xor4(v, TLOOK(6 * 16384 + ((hi >> 3) & 0x3F80u)));
			xor4(v, TLOOK(7 * 16384 + ((hi >> 10) & 0x3F80u)));
			xor4(v, TLOOK(8 * 16384 + ((hi >> 17) & 0x7F80u)));
#undef TLOOK
#else
			const unsigned lo = __byte_perm((unsigned)cf[q], 0, bsel);
			const unsigned hi = __byte_perm((unsigned)(cf[q] >> 32), 0, bsel);
			/* byte offset of a line = 32768 * pair + 128 * (byte of the coefficient) */
#define TLOOK(base, off) (*reinterpret_cast<const uint4 *>((base) + (off)))
			xor4(v, TLOOK(Tbe, 0 * 32768 + ((lo << 7) & 0x7F80u)));
			xor4(v, TLOOK(Tbo, 0 * 32768 + ((lo >> 1) & 0x7F80u)));
			xor4(v, TLOOK(Tbe, 1 * 32768 + ((lo >> 9) & 0x7F80u)));
			xor4(v, TLOOK(Tbo, 1 * 32768 + ((lo >> 17) & 0x7F80u)));
			xor4(v, TLOOK(Tbe, 2 * 32768 + ((hi << 7) & 0x7F80u)));
			xor4(v, TLOOK(Tbo, 2 * 32768 + ((hi >> 1) & 0x7F80u)));
			xor4(v, TLOOK(Tbe, 3 * 32768 + ((hi >> 9) & 0x7F80u)));
			xor4(v, TLOOK(Tbo, 3 * 32768 + ((hi >> 17) & 0x7F80u)));
#undef TLOOK
#endif
			__stcg(p + (long long)(SWEEP_THREADS / SQ) * q * SQ, v);
			if (force && ch == nch) {
				long long row = row0 + (SWEEP_THREADS / SQ) * q;
				pc_next[row] = (wn & 1) ? (((u64)v.w << 32) | v.z) : (((u64)v.y << 32) | v.x);
			}
		}
		if (pd_next && u == 0) {
			/* This CTA just produced panel word w+1 of the first SWEEP_RU active rows.
			 * Search them for the next panel's pivots while the other SMs keep
			 * sweeping; a full set (the usual case on dense systems) or an exhausted
			 * row range makes the description final and turns k_select into a no-op.
			 * The rows' other strips are complete by the time k_apply runs. */
			/* the pc_next words these rows just received are read back by this CTA only:
			 * the barrier orders them; the rest of the grid meets them at the kernel boundary */
			__threadfence_block();
			__syncthreads(); /* also: every lookup of this unit is done (SW = 16: the tables may be clobbered) */
#if SW == 16
			cur = -1;        /* ... and are rebuilt before the next unit */
#endif
			select_init(S);
			__syncthreads();
			const long long lim = min(m, r1 + (long long)SWEEP_RU);
			select_scan(S, pc_next, r1, lim, colmask_next);
			if (dl) {
				/* row-sharded system: this shard's candidates go to every peer now, and (one
				 * process per GPU) the global election runs here too, while the sweep goes on */
				dist_lookahead(dl, S, pc_next, wn, colmask_next, r1, lim, m, st, pd_next, hist_r, hist_pm);
			} else if (tid < 32) {
				const bool final_ = (S.pm == colmask_next || lim == m);
				if (final_)
					select_finalize(S, pc_next, wn, r1, st, pd_next, hist_r, hist_pm);
				else if (tid == 0)
					pd_next->valid = 0;
				if (publish_look) {
					/* k_sweep_apply: the other CTAs wait for this verdict in their tails and, when it is
					 * "valid", apply the next panel themselves */
					__syncwarp();
					if (tid == 0) {
						pd_next->applied = final_ ? wn + 1 : 0;
						__threadfence();
						st_release_gpu(&pd_next->look, (unsigned)wn + 1);
					}
				}
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(SWEEP_THREADS, 1)
k_sweep(Mat M, const PanelDesc *__restrict__ pd, const u64 *__restrict__ pc_cur,
        u64 *__restrict__ pc_next, const uint4 *__restrict__ ebuf, int w, int s0,
        PanelDesc *pd_next, SolverState *st, long long *hist_r, u64 *hist_pm, u64 colmask_next) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	sweep_body(M, pd, pc_cur, pc_next, ebuf, w, s0, pd_next, st, hist_r, hist_pm, colmask_next);
}

/* any active row (i >= rank) with b = 1 makes the system inconsistent
 * (_mzd_pluq_solve_left's check, _internal.c:440-446 -> None) */
__global__ void k_check(Mat M, SolverState *st) {
	const long long r = st->r_loc;
	int bad = 0;
	for (long long i = r + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M.m;
	     i += (long long)gridDim.x * blockDim.x)
		bad |= (int)(M.base[widx(M, i, M.nw)] & 1);
	if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&st->inconsistent, 1);
}

} /* namespace gf2b200 */
