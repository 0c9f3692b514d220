/*
 * gf2b200_persist.cuh -- k_forward: the whole forward elimination of a one-GPU system as
 * ONE persistent cooperative kernel (1 CTA of 1024 threads per SM, every CTA resident).
 *
 * What it replaces: the per-panel launch chain k_select -> k_apply -> k_sweep of
 * gf2b200_kernels.cuh (still used by the row-sharded path and by 128-byte-strip builds).
 * Same arithmetic, same tables, same bit-exact result (M4RI's _mzd_pluq semantics,
 * reference gf2bv/_internal.c:433) -- what changes is WHERE the per-panel
 * fixed cost goes.  Measured in round 1 (profiles/r01c_launches.md):
 * ~30 us per panel outside the streaming part of the sweep (three launches and their
 * gaps, k_apply 7.9 us, a no-op k_select 2.3 us, launch ramp and drain), 10 % of a solve at
 * n = 131072 and two thirds of it at n = 32768.  Here one panel is
 *
 *   sweep(w)      every CTA streams its contiguous range of work units (a unit = one
 *                 64-byte strip x 1024 rows), exactly as k_sweep does;
 *   search(w+1)   the CTA that owns unit 0 (strip of word w+1, first 1024 active rows)
 *                 searches those rows for the next panel's pivots as soon as it has
 *                 swept them (as k_sweep's fused search does) and publishes the result
 *                 with a release store; the other 147 CTAs keep streaming;
 *   apply(w+1)    every CTA, once its own range is done, builds E = TB * Sel for the
 *                 strips whose first-1024-rows unit IT swept (so the rows it reads are
 *                 rows it wrote; no other CTA touches them in this panel), stores the
 *                 pivot rows in place, moves the displaced rows and writes the E tile
 *                 of the strip into the OTHER half of the double-buffered ebuf;
 *   grid barrier  one per panel.
 *
 * When the first 1024 active rows do not settle the next panel (rank-deficient or
 * sparse systems) the panel ends with the slow path: barrier, full scan of the panel
 * column by CTA 0, barrier, apply spread over all CTAs, barrier.
 *
 * Memory model: everything another CTA wrote during this kernel is read with ld.cg
 * (L2) or an acquire load, never through L1 or the non-coherent path; the E tile still
 * arrives by TMA bulk copy (L2 -> shared memory) after a proxy fence.  Every wait has
 * a time-out that raises GridSync::fault and makes all CTAs leave: a logic error ends
 * in an error code, not in a hung GPU.
 */
#pragma once
#include "gf2b200_kernels.cuh"

#if SW == 8
namespace gf2b200 {

struct GridSync {
	unsigned count, gen; /* grid barrier: arrivals, generation */
	unsigned sel_flag;   /* panel index + 1 whose description the look-ahead search made final */
	unsigned need_full;  /* panel index + 1 for which the first 1024 rows were not enough */
	int fault;           /* a wait timed out */
	int done_w;          /* panels finished (diagnostic) */
	unsigned cand_cnt[2]; /* candidate rows collected for panel (slot = panel & 1): see PERSIST_CAND_MAX */
	unsigned list_cnt[2]; /* length of the candidate list the slow path of panel (slot) used, 0 if none */
	unsigned pad2[2];
	/* what every CTA needs at the top of a panel, in ONE 16-byte load: {valid = panel + 1,
	 * r1 = first active row, pm lo, pm hi}; slot = panel & 1 (k = popcount(pm)) */
	uint4 hdr[2];
	/* the same header written by the SLOW path of a panel.  It must not go to hdr[]: every CTA decides
	 * "slow or not" from hdr[panel & 1] at the top of the panel, and a CTA that gets there late -- after
	 * CTA 0 has already published its verdict -- would read a valid header, skip the slow path's apply
	 * and barrier, and leave the grid barriers out of step.  (Never seen at full speed, where CTA 0
	 * needs microseconds and the others nanoseconds; compute-sanitizer's racecheck, which slows and
	 * skews the CTAs, ran into it: profiles/r02_sanitizer.md.) */
	uint4 hdr_slow;
};

/* bit 31 of the first word: "the rows selected for this panel all came from its first 1024 active
 * rows" -- the look-ahead of the next panel is worth trying (set by the look-ahead itself, and by the
 * slow path when a dense-looking panel follows a sparse one) */
__device__ __forceinline__ void hdr_publish(uint4 *slot, int w, long long r1, u64 pm, bool window_ok) {
	__stcg(slot, make_uint4((unsigned)(w + 1) | (window_ok ? 0x80000000u : 0u), (unsigned)r1, (unsigned)pm, (unsigned)(pm >> 32)));
}

/* How the sweep reads a row's coefficient (the panel word pc_cur[row]): a plain ld.global (L1-cached).
 * Safe inside the kernel: the words were written before the last grid barrier, whose acquire
 * invalidates this SM's L1 (PTX memory model: weak loads after an acquire observe what happened
 * before the release); never the non-coherent path (ld.global.nc is outside the model).  The row loads
 * do not wait for the coefficient (rows whose coefficient turns out to be zero are read for nothing
 * and not written): gating them chained two L2 latencies per unit, 600 -> 582 ms.  Measured equal and
 * removed: ld.cg for the coefficient; the first E tile of a panel by plain loads instead of the bulk
 * copy (profiles/r02_ab.md call C). */
/* Developer trace (-DPERSIST_TRACE=1, scripts/trace_forward.py): thread 32 of every CTA (NOT a
 * lane of warp 0: a lone lane stamping there leaves warp 0 divergent and sends the search's
 * warp collectives down their slow BRA.DIV paths -- a 16 us search read 70 us) stamps
 * globaltimer at 8 points of every panel into t_panel + (nw + 2) + (w * gridDim + cta) * 8. */
#ifndef PERSIST_TRACE
#define PERSIST_TRACE 0
#endif
#if PERSIST_TRACE
#define TRACE(slot) do { if (tid == 32) t_panel[(size_t)(nw + 2) + ((size_t)w * G + blockIdx.x) * 8 + (slot)] = gtimer_ns(); } while (0)
/* inside warp 0: all 32 lanes store (warp-uniform, no divergence); slots 8.. of CTA 0 live in the trace row of CTA 1 + */
#define TRACE_W0(slot) do { t_panel[(size_t)(nw + 2) + ((size_t)w * G + blockIdx.x) * 8 + (slot)] = gtimer_ns(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_W0(slot) do { } while (0)
#endif
/* Sparse / rank-deficient systems: when the first 1024 active rows do not settle a panel,
 * the pivots must be looked for among ALL active rows.  While a panel that was itself
 * settled the slow way is swept, the CTAs that update the strip of the next panel word
 * append every row whose new panel word is non-zero to a candidate list (warp-aggregated
 * atomics); CTA 0 then scans that list (a few % of the rows on MT19937-class systems)
 * instead of the whole column.  More than PERSIST_CAND_MAX candidates: full scan. */
#ifndef PERSIST_GJ_SEARCH
#define PERSIST_GJ_SEARCH 1
#endif
#ifndef PERSIST_SEL_PAD
#define PERSIST_SEL_PAD 14
#endif
#ifndef PERSIST_CAND_MAX
#define PERSIST_CAND_MAX 8192
#endif
#ifndef PERSIST_TIMEOUT_NS
#define PERSIST_TIMEOUT_NS 20000000000ULL
#endif

#ifndef GF2_EMU
__device__ __forceinline__ unsigned long long gtimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
/* generic-proxy writes (other CTAs' st.global, made visible by the acquire before this)
 * before the async-proxy read of the bulk copy that follows */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
__device__ __forceinline__ unsigned long long gtimer_ns() { return (unsigned long long)(emu_now_ms() * 1e6); }
__device__ __forceinline__ void fence_proxy_async() {}
#endif

/* one thread: wait until *p - target >= 0 (flags only grow) or the other flag does; returns 1 / 2 for
 * the flag that did (both loads are acquires), 0 after a fault or a time-out */
__device__ __forceinline__ int persist_wait(const unsigned *p, unsigned target, const unsigned *alt, GridSync *gs) {
	const unsigned long long t0 = gtimer_ns();
	unsigned it = 0;
	for (;;) {
		if ((int)(ld_acquire_gpu(p) - target) >= 0) return 1;
		if (alt && (int)(ld_acquire_gpu(alt) - target) >= 0) return 2;
		if ((++it & 31) == 0) {
			if (*(volatile int *)&gs->fault) return 0;
			if (gtimer_ns() - t0 > PERSIST_TIMEOUT_NS) {
				atomicOr(&gs->fault, 1);
				return 0;
			}
		}
		__nanosleep(40);
	}
}

/* all threads of all CTAs; false after a fault (uniform over the CTA) */
__device__ __forceinline__ bool grid_barrier(GridSync *gs, int *s_ok) {
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned g = *(volatile unsigned *)&gs->gen; /* cannot advance before this CTA arrives */
		__threadfence();                                   /* this CTA's writes before its arrival */
		bool ok = true;
		if (atomicAdd(&gs->count, 1u) == gridDim.x - 1) {
			atomicExch(&gs->count, 0u);
			st_release_gpu(&gs->gen, g + 1); /* cumulative: orders everything this thread observed */
		} else {
			ok = persist_wait(&gs->gen, g + 1, nullptr, gs) != 0; /* acquire loads */
		}
		*s_ok = (ok && !*(volatile int *)&gs->fault) ? 1 : 0;
	}
	__syncthreads();
	return *s_ok != 0;
}

/* Look-ahead pivot search, warp 0 only: the 1024 panel words of the first active rows sit in
 * S.qv (slot = row - base8, 0 for rows outside the active range) where the sweep of unit 0
 * left them -- no reload from global memory, no compaction pass.  32 slots at a time; zero
 * words are skipped by ballot; stops as soon as every column of the panel is a pivot.
 *
 * The RREF basis is held COLUMN-wise: lane L owns panel columns L and L + 32 as bit masks over
 * the pivots (bit c of Clo = entry of basis vector c in column L) and rows L and L + 32 of the
 * transform (bit c of Tlo = "selected row L takes part in E_c").  Reducing a candidate v
 * against the whole basis is then lane-local -- its bit in column j is
 * v_j ^ parity(C_j & v & pm) -- and two ballots hand every lane the reduced vector; the only
 * other cross-lane traffic of an insertion is the broadcast of column c of the basis (which
 * vectors must absorb the new one to stay reduced).  The REDUX.XOR formulation of
 * wb_insert measured 210 ns per insertion in this position (profiles/r02_trace.md): with
 * ~66 insertions per panel on the critical path of EVERY panel it was the longest link of
 * the chain for n <= 32768 and for the last quarter of the panels of any n. */
__device__ __forceinline__ void window_search(SelectSmem &S, long long base8, u64 colmask, int lane) {
	for (int c = lane; c < 64; c += 32) S.topsel[c] = 0;
	/* S.sel[] is initialised by the lane that fills it below (no write after another lane's write:
	 * compute-sanitizer racecheck flagged the lane-parallel form) */
	if (lane == 0)
		for (int c = 0; c < 64; c++) S.sel[c] = -1;
	u64 Clo = 0, Chi = 0, Tlo = 0, Thi = 0, pm = 0;
	int nsel = 0;
	for (int g = 0; g < SWEEP_RU / 32 && pm != colmask; g++) {
		const u64 mine = S.qv[32 * g + lane] & colmask;
		unsigned bal = __ballot_sync(0xffffffffu, mine != 0);
		while (bal && pm != colmask) {
			const int src = __ffs((int)bal) - 1;
			bal &= bal - 1;
			const u64 v = shfl64(mine, src);
			const u64 mm = v & pm; /* the basis vectors v picks up */
			const unsigned vlo = ((unsigned)(v >> lane) & 1u) ^ ((unsigned)__popcll(Clo & mm) & 1u);
			const unsigned vhi = ((unsigned)(v >> (lane + 32)) & 1u) ^ ((unsigned)__popcll(Chi & mm) & 1u);
			const u64 vr = (u64)__ballot_sync(0xffffffffu, vlo) | ((u64)__ballot_sync(0xffffffffu, vhi) << 32);
			if (!vr) continue; /* dependent on the rows selected so far */
			const int c = __ffsll((long long)vr) - 1;
			/* transform bits of the new vector: what it picked up, plus itself (selected row nsel) */
			unsigned tlo = (unsigned)__popcll(Tlo & mm) & 1u, thi = (unsigned)__popcll(Thi & mm) & 1u;
			if (lane == (nsel & 31)) {
				if (nsel < 32) tlo ^= 1u;
				else thi ^= 1u;
			}
			/* basis vectors with a 1 in column c absorb the new vector (keeps the basis reduced) */
			const u64 colc = shfl64(c < 32 ? Clo : Chi, c & 31);
			const u64 bit = 1ULL << c;
			Clo = (Clo ^ (vlo ? colc : 0)) | (vlo ? bit : 0);
			Chi = (Chi ^ (vhi ? colc : 0)) | (vhi ? bit : 0);
			Tlo = (Tlo ^ (tlo ? colc : 0)) | (tlo ? bit : 0);
			Thi = (Thi ^ (thi ? colc : 0)) | (thi ? bit : 0);
			if (lane == 0) S.sel[nsel] = (int)(base8 + 32 * g + src);
			pm |= bit;
			nsel++;
		}
	}
	/* hand the transform over in the by-pivot-column form select_finalize / the apply read:
	 * TB[c] bit l = "selected row l takes part in E_c" (a 64 x 64 bit transpose by ballots) */
	u64 mine_lo = 0, mine_hi = 0;
#pragma unroll 8
	for (int c = 0; c < 64; c++) {
		const u64 wv = (u64)__ballot_sync(0xffffffffu, (unsigned)(Tlo >> c) & 1u) |
		               ((u64)__ballot_sync(0xffffffffu, (unsigned)(Thi >> c) & 1u) << 32);
		if (lane == (c & 31)) {
			if (c < 32) mine_lo = wv;
			else mine_hi = wv;
		}
	}
	S.TB[lane] = mine_lo;
	S.TB[lane + 32] = mine_hi;
	if (lane == 0) {
		S.pm = pm;
		S.nsel = nsel;
	}
	__syncwarp();
}

/* The same search as a Gauss-Jordan elimination, for the usual case that the window holds a full
 * set of pivots and the panel has all 64 columns.  window_search inserts ~66 rows one after the
 * other, ~100 dependent instructions each: 14.5 us on the critical path of every small panel
 * (profiles/r02_trace.md).  Here lane L holds window slots L and L + 32 as ROWS (value v, transform t
 * over the 64 slots):
 *   phase A  the first 64 slots are eliminated column by column: two ballots find a row that is
 *            not a pivot yet and has a 1 in column c, two 64-bit shuffles broadcast it, every other
 *            row with a 1 there absorbs it (~25 instructions per column);
 *   phase B  a random 64 x 64 block has full rank with probability 0.29 only (expected defect 0.85):
 *            the missing columns come from the following slots, each candidate reduced against the
 *            basis with one lane-local selection + a warp XOR, and a row that survives moves into
 *            a slot phase A left without a pivot (its index there is its "selected row" number).
 * Which rows are selected does not matter (the pivot COLUMNS are the invariant);
 * selected row l = slot l, TB[c] = transform of the row whose pivot is column c.  Returns with
 * S.pm != all-ones when the window does not hold a full set: the caller then takes the all-rows
 * search, as before. */
__device__ __forceinline__ void window_search_gj(SelectSmem &S, long long base8, int lane) {
	for (int c = lane; c < 64; c += 32) {
		S.sel[c] = -1;
		S.topsel[c] = 0;
	}
	u64 v0 = S.qv[lane], v1 = S.qv[lane + 32];
	u64 t0 = 1ULL << lane, t1 = 1ULL << (lane + 32);
	int c0 = -1, c1 = -1; /* pivot column of my two rows, -1: none (yet) */
	u64 pm = 0;
#pragma unroll 1
	for (int c = 0; c < 64; c++) {
		const bool b0 = (v0 >> c) & 1, b1 = (v1 >> c) & 1;
		const unsigned f0 = __ballot_sync(0xffffffffu, b0 && c0 < 0), f1 = __ballot_sync(0xffffffffu, b1 && c1 < 0);
		if (!(f0 | f1)) continue; /* no row of this block has a pivot in column c */
		const bool low = f0 != 0;
		const int src = __ffs((int)(low ? f0 : f1)) - 1;
		const u64 pv = shfl64(low ? v0 : v1, src), pt = shfl64(low ? t0 : t1, src);
		const bool me0 = low && lane == src, me1 = !low && lane == src;
		if (b0 && !me0) {
			v0 ^= pv;
			t0 ^= pt;
		}
		if (b1 && !me1) {
			v1 ^= pv;
			t1 ^= pt;
		}
		if (me0) c0 = c;
		if (me1) c1 = c;
		pm |= 1ULL << c;
	}
	/* phase B: the columns still missing, from the slots that follow */
	for (int g = 2; g < SWEEP_RU / 32 && pm != ~0ULL; g++) {
		const u64 mine = S.qv[32 * g + lane];
		unsigned bal = __ballot_sync(0xffffffffu, mine != 0);
		while (bal && pm != ~0ULL) {
			const int src = __ffs((int)bal) - 1;
			bal &= bal - 1;
			const u64 x = shfl64(mine, src);
			/* the basis is reduced: x loses exactly the rows whose pivot column it has a 1 in */
			const bool u0 = c0 >= 0 && ((x >> c0) & 1), u1 = c1 >= 0 && ((x >> c1) & 1);
			const u64 xr = x ^ warp_xor64((u0 ? v0 : 0) ^ (u1 ? v1 : 0));
			if (!xr) continue; /* dependent on the rows selected so far */
			const int c = __ffsll((long long)xr) - 1;
			const unsigned h0 = __ballot_sync(0xffffffffu, c0 < 0), h1 = __ballot_sync(0xffffffffu, c1 < 0);
			const bool low = h0 != 0; /* a slot without a pivot exists: fewer than 64 pivots so far */
			const int hl = __ffs((int)(low ? h0 : h1)) - 1;
			const int slot = hl + (low ? 0 : 32);
			const u64 tn = warp_xor64((u0 ? t0 : 0) ^ (u1 ? t1 : 0)) ^ (1ULL << slot);
			if (c0 >= 0 && ((v0 >> c) & 1)) {
				v0 ^= xr;
				t0 ^= tn;
			}
			if (c1 >= 0 && ((v1 >> c) & 1)) {
				v1 ^= xr;
				t1 ^= tn;
			}
			if (lane == hl) {
				if (low) {
					v0 = xr;
					t0 = tn;
					c0 = c;
				} else {
					v1 = xr;
					t1 = tn;
					c1 = c;
				}
				S.sel[slot] = (int)(base8 + 32 * g + src);
			}
			pm |= 1ULL << c;
		}
	}
	/* phase A pivots sit in their own slots; rows that moved in during phase B wrote S.sel themselves */
	if (c0 >= 0) {
		S.TB[c0] = t0;
		if (S.sel[lane] < 0) S.sel[lane] = (int)(base8 + lane);
	}
	if (c1 >= 0) {
		S.TB[c1] = t1;
		if (S.sel[lane + 32] < 0) S.sel[lane + 32] = (int)(base8 + lane + 32);
	}
	if (lane == 0) {
		S.pm = pm;
		S.nsel = __popcll(pm);
	}
	__syncwarp();
}

/* Shared-memory image of the panel description an apply needs. */
struct ApplySmem {
	uint4 Sel[4][64][SQ]; /* four strips at a time, 256 threads each */
	uint4 Dis[4][64][SQ];
	u64 TB[64];
	int sel[64], src[64], dst[64];
	int k, nmove;
	long long r;
	u64 pm;
};

/* E = TB * Sel for the strips list(i), i in [0, count): rows r..r+k-1 <- E (pivot rows),
 * displaced rows -> vacated positions, E tile -> ebuf_dst.  All SWEEP_THREADS threads.
 * The description is read with ld.cg: another CTA may have written it in this kernel. */
template <typename StripOf>
__device__ __forceinline__ void persist_apply(const Mat &M, const PanelDesc *pdn, uint4 *__restrict__ ebuf_dst,
                                              ApplySmem &A, int count, StripOf strip_of) {
	const int tid = threadIdx.x;
	if (tid < 64) {
		A.TB[tid] = __ldcg(&pdn->TB[tid]);
		A.sel[tid] = __ldcg(&pdn->sel[tid]);
		A.src[tid] = __ldcg(&pdn->mv_src[tid]);
		A.dst[tid] = __ldcg(&pdn->mv_dst[tid]);
	}
	if (tid == 0) {
		A.k = __ldcg(&pdn->k);
		A.nmove = __ldcg(&pdn->nmove);
		A.r = __ldcg(&pdn->r);
		A.pm = __ldcg(&pdn->pm);
	}
	__syncthreads();
	const int k = A.k, nmove = A.nmove;
	if (k == 0) return;
	const int g = tid >> 8, t8 = tid & 255, rr = t8 / SQ, ch = t8 % SQ;
	const u64 pm = A.pm;
	const long long erow = A.r + __popcll(pm & ((1ULL << rr) - 1));
	const bool ispiv = (pm >> rr) & 1;
	uint4 *mb = reinterpret_cast<uint4 *>(M.base);
	const uint4 z = make_uint4(0, 0, 0, 0);
	for (int i0 = 0; i0 < count; i0 += 4) {
		const int i = i0 + g;
		const bool on = i < count;
		const int s = on ? strip_of(i) : 0;
		const long long sb = (long long)s * M.mp;
		if (on) {
			A.Sel[g][rr][ch] = (rr < k) ? __ldcg(mb + (sb + A.sel[rr]) * SQ + ch) : z;
			if (rr < nmove) A.Dis[g][rr][ch] = __ldcg(mb + (sb + A.src[rr]) * SQ + ch);
		}
		__syncthreads();
		if (on) {
			uint4 acc = z;
			u64 t = A.TB[rr];
			while (t) {
				const int l = __ffsll((long long)t) - 1;
				t &= t - 1;
				xor4(acc, A.Sel[g][l][ch]);
			}
			__stcg(ebuf_dst + (long long)s * EBUF_Q + rr * SQ + ch, acc);
			if (ispiv) __stcg(mb + (sb + erow) * SQ + ch, acc);
			if (rr < nmove) __stcg(mb + (sb + A.dst[rr]) * SQ + ch, A.Dis[g][rr][ch]);
		}
		__syncthreads();
	}
}

/* The list units of a sparse sweep (units [u0, u1), numbered from the first list unit): strip
 * s0 + 1 + u / ochunks, 1024 entries of the candidate list instead of 1024 consecutive rows, one row
 * piece at a time per thread, same tables and the same eight lookups as the streaming loop.  Kept
 * out of line: inlined, its registers pushed spills into the streaming loop of k_forward.  Returns
 * the phase of the tile mbarrier. */
__device__ __noinline__ unsigned sparse_sweep(uint4 *mb, long long mp, uint4 *TD, uint4 *P, uint4 *E, u64 *bar, unsigned phase,
                                              const uint4 *ebuf, const u64 *list, unsigned lcnt, long long u0, long long u1,
                                              long long ochunks, int s_first, long long r1, const u64 *pc_cur, u64 pm) {
	const int tid = threadIdx.x;
	const int ch = tid % SQ, rl = tid / SQ;
	const int h = rl & 1;
	const unsigned char *Tbe = reinterpret_cast<const unsigned char *>(TD + 4 * h + ch);
	const unsigned char *Tbo = reinterpret_cast<const unsigned char *>(TD + 4 * (1 - h) + ch);
	const unsigned bsel = h ? 0x2301u : 0x3210u;
	int cur = -1;
	int s = s_first + (int)(u0 / ochunks);
	long long chunk = u0 % ochunks;
#pragma unroll 1
	for (long long u = u0; u < u1; ++u, ++chunk) {
		if (chunk == ochunks) {
			chunk = 0;
			++s;
		}
		if (s != cur) {
			__syncthreads(); /* everyone is done with the previous tables */
			if (tid == 0) {
				fence_proxy_async();
				mbar_expect_tx(bar, EBUF_Q * 16);
				tma_bulk_g2s(E, ebuf + (long long)s * EBUF_Q, EBUF_Q * 16, bar);
			}
			mbar_wait(bar, phase);
			phase ^= 1;
			sweep_build_tables(TD, P, E, tid, nullptr, bar);
			cur = s;
		}
		uint4 *strip = mb + (long long)s * mp * SQ;
#pragma unroll 1
		for (int q = 0; q < SWEEP_U; q++) {
			const long long idx = chunk * SWEEP_RU + rl + (SWEEP_THREADS / SQ) * q;
			if (idx >= (long long)lcnt) continue;
			const long long row = (long long)__ldcg(list + 2 * idx);
			if (row < r1) continue; /* became a pivot row of this panel */
			const u64 cfq = ld_weak_u64(pc_cur + row) & pm;
			if (!cfq) continue;
			uint4 *pq = strip + row * SQ + ch;
			uint4 v = __ldcg(pq);
			const unsigned lo = __byte_perm((unsigned)cfq, 0, bsel);
			const unsigned hi = __byte_perm((unsigned)(cfq >> 32), 0, bsel);
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
			__stcg(pq, v);
		}
	}
	return phase;
}

/* What a CTA derives from the header of a panel: its share [u0, u1) of the panel's work units and
 * the geometry behind it.  Computed before the streaming loop and AGAIN after it, from words
 * re-read from global memory: written once and used on both sides of the loop, these values
 * stayed in registers across it, and ptxas paid for them inside the loop (it re-derived the
 * per-thread constants for every row piece: 275 instructions per lean unit, 233 without; an
 * intermediate build with more such state ran the whole solve 6 % slower). */
struct PanelGeo {
	long long r1, base8, nchunks, u0, u1;
	long long sparse_chunks; /* list chunks per strip in sparse mode, else 0 */
	u64 pm;
	int s0;
	bool try_window;
};

/* lc: length of the candidate list the slow path of this panel worked from (0: none / fast path) */
__device__ __forceinline__ PanelGeo panel_geo(const Mat &M, uint4 hd, unsigned lc, int w, int G, int b) {
	PanelGeo g;
	g.r1 = (long long)hd.y;
	g.pm = ((u64)hd.w << 32) | hd.z;
	/* sparse / rank-deficient mode: a panel whose pivots did NOT all come from its first 1024 rows
	 * is followed by one that most likely needs all rows too -- the look-ahead search (one CTA busy,
	 * and everybody's wait for its verdict) is skipped and the panel goes straight to the
	 * candidate list.  A panel settled from its first rows switches the look-ahead back on. */
	g.try_window = (hd.x >> 31) != 0;
	/* Sparse sweep: when the slow path worked from a candidate list (the rows whose word of THIS
	 * panel is non-zero) and that list is short, only the listed rows can have a non-zero
	 * coefficient -- every strip but the one of the next panel word is swept over the list
	 * instead of over all active rows (a unit of 1024 rows costs 8 KiB of coefficient reads even
	 * when none of them has work: on MT19937 that was 20 us of a 44 us panel).  Row moves of the
	 * panel are harmless: moved rows land on listed positions, and the coefficient is re-read. */
	g.sparse_chunks = 0;
	if (lc > 0 && (long long)lc * 4 <= M.m - g.r1) {
		g.sparse_chunks = ((long long)lc + SWEEP_RU - 1) / SWEEP_RU;
		g.try_window = false;
	}
	const int wn = w + 1;
	g.s0 = wn >> SW_SHIFT;
	g.base8 = g.r1 & ~7LL; /* chunks start on 512-byte boundaries of the strip */
	g.nchunks = (M.m - g.base8 + SWEEP_RU - 1) / SWEEP_RU;
	/* chunks per strip: the strip of the next panel word always takes every active row (it hands
	 * pc_next and the next candidate list over); the others take the list in sparse mode */
	const long long units = g.nchunks + (long long)(M.ns - g.s0 - 1) * (g.sparse_chunks ? g.sparse_chunks : g.nchunks);
	/* the CTA that owns unit 0 runs the look-ahead search with every other warp of its SM idle and
	 * restarts without a prefetched tile: it is dealt PERSIST_SEL_PAD fewer units */
	const long long vpad = (wn < M.nw && g.try_window) ? max(0LL, min((long long)PERSIST_SEL_PAD, units / G - 1)) : 0;
	const long long vunits = units + vpad;
	g.u0 = max(0LL, vunits * b / G - vpad);
	/* sparse mode: units [0, nchunks) are the rows of strip s0 (streaming loop); the list units that
	 * follow are dealt out separately */
	g.u1 = g.sparse_chunks ? min(vunits * (b + 1) / G, g.nchunks) : vunits * (b + 1) / G - vpad;
	return g;
}

#ifndef GF2_EMU
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
	uint4 v;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}
#else
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
	const volatile unsigned *q = reinterpret_cast<const volatile unsigned *>(p);
	return make_uint4(q[0], q[1], q[2], q[3]);
}
#endif

#define PERSIST_CTRL_BYTES 64
#define PERSIST_SMEM (SWEEP_LINES * 128 + SWEEP_SCRATCH_BYTES + 16 + PERSIST_CTRL_BYTES)
static_assert(sizeof(ApplySmem) <= SWEEP_LINES * 128, "the apply scratch aliases the (dead) tables");

__global__ void __launch_bounds__(SWEEP_THREADS, 1)
k_forward(Mat M, u64 *pc0, u64 *pc1, uint4 *ebuf0, uint4 *ebuf1, PanelDesc *pd2, SolverState *st,
          long long *hist_r, u64 *hist_pm, GridSync *gs, unsigned long long *t_panel, u64 *cand, int w_begin,
          int w_end) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint4 *TD = reinterpret_cast<uint4 *>(smem_raw);
	uint4 *E = TD + SWEEP_LINES * 8;
	uint4 *P = E + EBUF_Q;
	u64 *bar = reinterpret_cast<u64 *>(smem_raw + SWEEP_LINES * 128 + SWEEP_SCRATCH_BYTES);
	int *s_ok = reinterpret_cast<int *>(smem_raw + SWEEP_LINES * 128 + SWEEP_SCRATCH_BYTES + 16);
	int *s_state = s_ok + 1; /* 1: description of the next panel is final, 2: needs the full scan, 0: fault */
	SelectSmem &S = *reinterpret_cast<SelectSmem *>(smem_raw + SWEEP_LINES * 128);
	ApplySmem &AP = *reinterpret_cast<ApplySmem *>(smem_raw);

	const int tid = threadIdx.x;
	const int G = gridDim.x;
	const long long m = M.m;
	const int nw = M.nw;
	uint4 *mb = reinterpret_cast<uint4 *>(M.base);
	if (tid == 0) {
		mbar_init(bar, 1);
		mbar_init_fence();
	}
	__syncthreads();
	unsigned phase = 0;
	bool list_ready = false; /* the candidate list of the panel about to start was collected */

	for (int w = w_begin; w < w_end; ++w) {
		u64 *pc_cur = (w & 1) ? pc1 : pc0, *pc_next = (w & 1) ? pc0 : pc1;
		const uint4 *ebuf = (w & 1) ? ebuf1 : ebuf0;
		PanelDesc *pd = pd2 + (w & 1), *pdn = pd2 + ((w + 1) & 1);
		if (blockIdx.x == 0 && tid == 0) t_panel[w] = gtimer_ns();
		TRACE(0);

		/* ---- slow path: nobody settled this panel ahead of time ---------------- */
		uint4 hd = __ldcg(&gs->hdr[w & 1]);
		const bool slow = (int)(hd.x & 0x7fffffffu) != w + 1;
		if (slow) {
			u64 colmask = ~0ULL;
			if (w == nw - 1 && (M.n & 63)) colmask = (1ULL << (M.n & 63)) - 1;
			if (blockIdx.x == 0) {
				const long long r = *(volatile long long *)&st->r;
				const unsigned cnt = list_ready ? *(volatile unsigned *)&gs->cand_cnt[w & 1] : 0xffffffffu;
				select_init(S);
				__syncthreads();
				if (cnt <= PERSIST_CAND_MAX)
					select_scan(S, pc_cur, 0, (long long)cnt, colmask, cand + (size_t)(w & 1) * PERSIST_CAND_MAX * 2);
				else
					select_scan(S, pc_cur, r, m, colmask);
				if (tid < 32) {
					/* did every selected row come from the first 1024 active rows?  (S.sel is read before
					 * select_finalize, which leaves it untouched) */
					int far = 0;
					for (int l = tid; l < S.nsel; l += 32) far |= (S.sel[l] >= (r & ~7LL) + SWEEP_RU);
					const bool window_ok = !__any_sync(0xffffffffu, far);
					select_finalize(S, pc_cur, w, r, st, pd, hist_r, hist_pm);
					__syncwarp();
					if (tid == 0) {
						hdr_publish(&gs->hdr_slow, w, r + S.nsel, S.pm, window_ok);
						gs->list_cnt[w & 1] = (cnt <= PERSIST_CAND_MAX) ? cnt : 0u;
						__threadfence();
						st_release_gpu(&gs->sel_flag, (unsigned)w + 1);
					}
				}
			}
			/* every CTA: wait for the description, apply its share of the strips, ONE barrier */
			__syncthreads();
			if (tid == 0) {
				const bool ok = persist_wait(&gs->sel_flag, (unsigned)w + 1, nullptr, gs) != 0;
				*s_state = ok ? 1 : 0;
			}
			__syncthreads();
			if (*s_state == 0) return;
			const int s0a = w >> SW_SHIFT;
			const int mine = (M.ns - s0a - (int)blockIdx.x + G - 1) / G; /* strips s0a + b + i*G */
			persist_apply(M, pd, const_cast<uint4 *>(ebuf), AP, mine > 0 ? mine : 0,
			              [&](int i) { return s0a + (int)blockIdx.x + i * G; });
			if (!grid_barrier(gs, s_ok)) return;
			hd = __ldcg(&gs->hdr_slow);
		}
		/* the list of this panel is consumed (or was not needed); its slot is refilled two panels on */
		if (blockIdx.x == 0 && tid == 0) gs->cand_cnt[w & 1] = 0;
		/* collect candidates for the next panel while sweeping a panel that needed the slow path */
		const bool collect = slow;
		list_ready = false;
		const PanelGeo geo = panel_geo(M, hd, slow ? __ldcg(&gs->list_cnt[w & 1]) : 0u, w, G, (int)blockIdx.x);
		const long long r1 = geo.r1;
		const bool try_window = geo.try_window;

		const u64 pm = ((u64)hd.w << 32) | hd.z;
		const int k = __popcll(pm);
		const int wn = w + 1;
		const bool has_next = wn < nw;
		if (r1 >= m) {
			/* every row is an echelon row: the remaining panels have no pivots */
			if (blockIdx.x == 0)
				for (int x = wn + tid; x < nw; x += SWEEP_THREADS) {
					hist_r[x] = r1;
					hist_pm[x] = 0;
				}
			break;
		}
		if (k == 0) {
			/* nothing to eliminate: hand the next word column to the (slow-path) search */
			for (long long i = r1 + blockIdx.x * (long long)SWEEP_THREADS + tid; i < m; i += (long long)G * SWEEP_THREADS)
				pc_next[i] = __ldcg(M.base + widx(M, i, wn));
			if (!grid_barrier(gs, s_ok)) return;
			continue;
		}

		/* ---- sweep(w): rows [r1, m) x strips [s0, ns) ---------------------------- */
		const int s0 = geo.s0;
		/* 32-bit counters inside the streaming loop (units of a panel = strips x chunks < 2^31 for any
		 * matrix that fits a GPU): every 64-bit value kept across the loop costs it two registers */
		const long long base8 = geo.base8;
		const int nchunks = (int)geo.nchunks, u0 = (int)geo.u0, u1 = (int)geo.u1;
		const int nch = (wn & (SW - 1)) >> 1;
		u64 colmask_next = ~0ULL;
		if (wn == nw - 1 && (M.n & 63)) colmask_next = (1ULL << (M.n & 63)) - 1;
		if (u0 < u1) {
			/* per-thread lookup constants are (re)derived here, not kept live across the search /
			 * apply / barrier code of the panel loop (64 registers per thread, no spills) */
			const int ch = tid % SQ, rl = tid / SQ;
			const int h = rl & 1;
			const unsigned char *Tbe = reinterpret_cast<const unsigned char *>(TD + 4 * h + ch);
			const unsigned char *Tbo = reinterpret_cast<const unsigned char *>(TD + 4 * (1 - h) + ch);
			const unsigned bsel = h ? 0x2301u : 0x3210u;
#if SWEEP_LEAN_UNITS
			const smem_addr_t te32 = smem_addr(Tbe), to32 = smem_addr(Tbo);
			/* chunks [lean_lo, lean_hi) of a strip other than s0 lie entirely inside the active rows */
			const int lean_lo = (base8 >= r1) ? 0 : 1, lean_hi = (int)((m - base8) / SWEEP_RU);
#endif
			int cur = -1, fetched = -1;
			const int s_last = s0 + (int)((u1 - 1) / nchunks);
			int s = s0 + (int)(u0 / nchunks);
			int chunk = u0 % nchunks; /* (strip, row chunk) of unit u, advanced without dividing */
			for (int u = u0; u < u1; ++u, ++chunk) {
				if (chunk == nchunks) {
					chunk = 0;
					++s;
				}
				if (s != cur) {
					__syncthreads(); /* everyone is done with the previous tables */
					{
						if (fetched != s && tid == 0) {
							fence_proxy_async();
							mbar_expect_tx(bar, EBUF_Q * 16);
							tma_bulk_g2s(E, ebuf + (long long)s * EBUF_Q, EBUF_Q * 16, bar);
						}
						mbar_wait(bar, phase);
						phase ^= 1;
					}
					const bool more = SWEEP_EARLY_TILE && (s < s_last) && !(has_next && u == 0);
					if (more && tid == 0) fence_proxy_async();
					sweep_build_tables(TD, P, E, tid, more ? ebuf + (long long)(s + 1) * EBUF_Q : nullptr, bar);
					fetched = more ? s + 1 : -1;
					if (cur < 0) TRACE(1); /* first tables of the panel built */
					cur = s;
				}
				const long long row0 = base8 + (long long)chunk * SWEEP_RU + rl;
				const bool force = (s == s0);
				uint4 *p = mb + ((long long)s * M.mp + row0) * SQ + ch;
#if SWEEP_LEAN_UNITS
				if (!force && chunk >= lean_lo && chunk < lean_hi) {
					/* every consecutive lean unit of this strip in one tight loop (few live values: the
					 * per-thread constants stay in registers instead of being re-derived per row piece) */
					const int nl = min(lean_hi - chunk, u1 - u);
					const u64 *pcp = pc_cur + row0;
#pragma unroll 1
					for (int i = 0; i < nl; i++, p += (long long)SWEEP_RU * SQ, pcp += SWEEP_RU)
						lean_unit(p, pcp, pm, te32, to32, bsel);
					u += nl - 1;
					chunk += nl - 1;
					continue;
				}
#endif
				u64 cf[SWEEP_U];
				uint4 d[SWEEP_U];
				bool act[SWEEP_U];
#pragma unroll
				for (int q = 0; q < SWEEP_U; q++) {
					const long long row = row0 + (SWEEP_THREADS / SQ) * q;
					cf[q] = (row >= r1 && row < m) ? (ld_weak_u64(pc_cur + row) & pm) : 0;
				}
#pragma unroll
				for (int q = 0; q < SWEEP_U; q++) {
					const long long row = row0 + (SWEEP_THREADS / SQ) * q;
					if (row >= r1 && row < m) d[q] = __ldcg(p + (long long)(SWEEP_THREADS / SQ) * q * SQ);
					act[q] = (row >= r1 && row < m) && (cf[q] != 0 || force);
				}
#pragma unroll
				for (int q = 0; q < SWEEP_U; q++) {
					if (!act[q]) {
						if (has_next && try_window && u == 0 && ch == nch) S.qv[rl + (SWEEP_THREADS / SQ) * q] = 0;
						continue;
					}
					uint4 v = d[q];
					const unsigned lo = __byte_perm((unsigned)cf[q], 0, bsel);
					const unsigned hi = __byte_perm((unsigned)(cf[q] >> 32), 0, bsel);
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
					__stcg(p + (long long)(SWEEP_THREADS / SQ) * q * SQ, v);
					if (force && ch == nch) {
						const long long row = row0 + (SWEEP_THREADS / SQ) * q;
						const u64 nv = (wn & 1) ? (((u64)v.w << 32) | v.z) : (((u64)v.y << 32) | v.x);
						__stcg(pc_next + row, nv);
						if (has_next && try_window && u == 0) S.qv[rl + (SWEEP_THREADS / SQ) * q] = nv; /* for the look-ahead search */
					}
				}
				if (force && collect && has_next) {
					/* (uniform branch, sparse / rank-deficient systems only) warp-aggregated append of
					 * {row, new panel word} to the next panel's candidate list; the words are read back
					 * from pc_next, where this thread stored them a moment ago (same thread: ordered) */
#pragma unroll
					for (int q = 0; q < SWEEP_U; q++) {
						const long long row = row0 + (SWEEP_THREADS / SQ) * q;
						const u64 nvq = (ch == nch && row >= r1 && row < m) ? (__ldcg(pc_next + row) & colmask_next) : 0;
						const unsigned bal = __ballot_sync(0xffffffffu, nvq != 0);
						if (bal) {
							const int lane = tid & 31, leader = __ffs((int)bal) - 1;
							unsigned base_i = 0;
							if (lane == leader) base_i = atomicAdd(&gs->cand_cnt[wn & 1], (unsigned)__popc(bal));
							base_i = __shfl_sync(0xffffffffu, base_i, leader);
							const unsigned idx = base_i + __popc(bal & ((1u << lane) - 1));
							if (nvq && idx < PERSIST_CAND_MAX) {
								u64 *ce = cand + ((size_t)(wn & 1) * PERSIST_CAND_MAX + idx) * 2;
								__stcg(ce, (u64)row);
								__stcg(ce + 1, nvq);
							}
						}
					}
				}
				if (has_next && try_window && u == 0) {
					/* look-ahead: this CTA just produced word w+1 of the first active rows; search
					 * them for the next panel's pivots while the other SMs keep streaming */
					__threadfence_block();
					__syncthreads(); /* S.qv complete; this unit's pc_next words are written */
					TRACE(6);
					if (tid < 32) {
						const long long lim = min(m, base8 + (long long)SWEEP_RU);
						/* the Gauss-Jordan form needs all 64 columns and is only good for a full set of
						 * pivots: the last panel of an n that is no multiple of 64 and the last window of
						 * the matrix (whose partial result is final) keep the row-by-row search */
#if PERSIST_GJ_SEARCH
						if (colmask_next == ~0ULL && lim < m) window_search_gj(S, base8, tid);
						else
#endif
							window_search(S, base8, colmask_next, tid);
						TRACE_W0(1); /* (the search CTA's slot 1 is re-used: window search done) */
						const bool final_ = (S.pm == colmask_next || lim == m);
						if (final_) select_finalize(S, pc_next, wn, r1, st, pdn, hist_r, hist_pm);
						__syncwarp();
						TRACE_W0(0); /* (slot 0 re-used by the search CTA: finalize done) */
						if (tid == 0) {
							if (final_) hdr_publish(&gs->hdr[wn & 1], wn, r1 + S.nsel, S.pm, true);
							__threadfence();
							st_release_gpu(final_ ? &gs->sel_flag : &gs->need_full, (unsigned)wn + 1);
						}
					}
					__syncthreads();
					TRACE(7);
				}
			}
		}

		/* ---- after the streaming loop: the panel's geometry is derived AGAIN, from re-read words
		 * (see PanelGeo), so that the loop above shares no registers with what follows -------- */
		const uint4 hd2 = ld_volatile_u4(slow ? &gs->hdr_slow : &gs->hdr[w & 1]);
		const unsigned lc2 = slow ? *(volatile unsigned *)&gs->list_cnt[w & 1] : 0u;
		const PanelGeo g2 = panel_geo(M, hd2, lc2, w, G, (int)blockIdx.x);
		const int wn2 = w + 1;
		const bool has_next2 = wn2 < M.nw;
		if (g2.sparse_chunks) {
			/* sparse mode: this CTA's share of the list units */
			const long long all = g2.nchunks + (long long)(M.ns - g2.s0 - 1) * g2.sparse_chunks;
			const long long a0 = max(all * blockIdx.x / G, g2.nchunks), a1 = all * (blockIdx.x + 1) / G;
			if (a0 < a1)
				phase = sparse_sweep(mb, M.mp, TD, P, E, bar, phase, ebuf, cand + (size_t)(w & 1) * PERSIST_CAND_MAX * 2, lc2,
				                     a0 - g2.nchunks, a1 - g2.nchunks, g2.sparse_chunks, g2.s0 + 1, g2.r1, pc_cur, g2.pm);
		}

		list_ready = slow && has_next2; /* candidates for the next panel were collected while this one was swept */
		/* ---- apply(w+1) for the strips whose first-rows unit this CTA swept ------- */
		TRACE(2); /* my units are done */
		if (has_next2) {
			__syncthreads(); /* the tables are dead: their space is the apply scratch */
			if (tid == 0) {
				/* 1: the look-ahead settled the next panel, 2: it could not (or did not run: slow path), 0: fault.
				 * The acquire load that saw the flag orders this thread; the CTA barrier below hands that
				 * order on to the threads that read the description (ld.cg: never a stale L1 line) */
				*s_state = g2.try_window ? persist_wait(&gs->sel_flag, (unsigned)wn2 + 1, &gs->need_full, gs) : 2;
			}
			__syncthreads();
			const int state = *s_state;
			if (state == 0) return;
			TRACE(3); /* the next panel's description is there */
			if (state == 1 && g2.u0 < g2.u1) {
				/* strips (relative) whose chunk 0 is mine */
				const long long sa = (g2.u0 + g2.nchunks - 1) / g2.nchunks, sb = (g2.u1 - 1) / g2.nchunks;
				const int cnt = (int)(sb - sa + 1);
				const int sfirst = g2.s0 + (int)sa;
				persist_apply(M, pd2 + (wn2 & 1), (w & 1) ? ebuf0 : ebuf1, AP, cnt > 0 ? cnt : 0, [&](int i) { return sfirst + i; });
			}
		}
		if (blockIdx.x == 0 && tid == 0) gs->done_w = wn2;
		TRACE(4); /* apply done */
		if (!grid_barrier(gs, s_ok)) return;
		TRACE(5);
	}
	if (blockIdx.x == 0 && tid == 0) t_panel[w_end] = gtimer_ns();
}


/* ---- k_sweep_apply: the launch chain's sweep with the NEXT panel's apply in its tail ------------
 * The launch chain (k_select -> k_apply -> k_sweep per panel) streams faster than k_forward on large
 * matrices but pays ~16 us per panel outside the sweep (k_apply 7.8 us, a no-op k_select 2.7 us,
 * gaps: profiles/r02y_launches.md).  This kernel is k_sweep plus what k_forward does after its
 * streaming loop: a CTA that has swept its units waits for the verdict of the look-ahead search
 * (PanelDesc::look, released by the CTA that ran it inside sweep_body) and, when the next panel is
 * settled, builds E = TB * Sel for the strips whose first-rows unit IT swept -- rows nobody else
 * touches in this launch -- into the OTHER E-tile buffer.  k_apply of the next panel then returns at
 * once (PanelDesc::applied).  Launched cooperatively: the waits need every CTA resident. */
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
k_sweep_apply(Mat M, const PanelDesc *__restrict__ pd, const u64 *__restrict__ pc_cur, u64 *__restrict__ pc_next,
              const uint4 *__restrict__ ebuf, uint4 *__restrict__ ebuf_next, int w, int s0, PanelDesc *pd_next,
              SolverState *st, long long *hist_r, u64 *hist_pm, u64 colmask_next) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	sweep_body(M, pd, pc_cur, pc_next, ebuf, w, s0, pd_next, st, hist_r, hist_pm, colmask_next, nullptr, true);
	if (!pd_next) return;
	/* this CTA's share of the units, as sweep_body dealt them */
	const int k = pd->k;
	const long long r1 = pd->r1, m = M.m;
	if (r1 >= m || k == 0) return; /* no sweep, no look-ahead: k_select / k_apply settle the next panel */
	const long long nchunks = (m - r1 + SWEEP_RU - 1) / SWEEP_RU;
	const long long units = (long long)(M.ns - s0) * nchunks;
	const long long vpad = max(0LL, min((long long)SWEEP_SEL_PAD, units / gridDim.x - 1));
	const long long vunits = units + vpad;
	const long long u0 = max(0LL, vunits * blockIdx.x / gridDim.x - vpad);
	const long long u1 = vunits * (blockIdx.x + 1) / gridDim.x - vpad;
	if (u0 >= u1) return;
	const long long sa = (u0 + nchunks - 1) / nchunks, sb = (u1 - 1) / nchunks; /* strips (relative) whose chunk 0 is mine */
	if (sa > sb) return;
	__shared__ int s_verdict;
	__syncthreads(); /* the tables are dead: their space is the apply scratch */
	if (threadIdx.x == 0) {
		const unsigned long long t0 = gtimer_ns();
		int v = -1;
		unsigned it = 0;
		while (ld_acquire_gpu(&pd_next->look) != (unsigned)w + 2) {
			if ((++it & 63) == 0 && gtimer_ns() - t0 > PERSIST_TIMEOUT_NS) {
				v = 0; /* (cannot happen with every CTA resident; k_apply would then redo the panel -- see applied) */
				break;
			}
			__nanosleep(40);
		}
		if (v < 0) v = (*(volatile int *)&pd_next->valid == w + 2) ? 1 : 0;
		s_verdict = v;
	}
	__syncthreads();
	if (!s_verdict) return;
	ApplySmem &AP = *reinterpret_cast<ApplySmem *>(smem_raw);
	const int sfirst = s0 + (int)sa;
	persist_apply(M, pd_next, ebuf_next, AP, (int)(sb - sa + 1), [&](int i) { return sfirst + i; });
}

} /* namespace gf2b200 */
#endif /* SW == 8 */
