/*
 * gf2b200_basis.cuh -- the kernel basis as ONE blocked multi-right-hand-side triangular
 * solve (what the reference does with mzd_trsm_upper_left over all n - r right-hand sides,
 * gf2bv/_internal.c:330-348, call at :343), instead of one back-substitution per free
 * column.
 *
 * After the forward elimination the echelon rows are U = [U1 | U2] in pivot / free column
 * order, each row fully reduced inside its own 64-column panel.  The kernel vectors are the
 * columns of  Q^T [U1^-1 U2 ; I]:
 *
 *   k_basis_gather   F = U2: the d = n - r free columns (in M4RI's sigma order)
 *                    of the r echelon rows, gathered into a dense strip-major r x d bit matrix;
 *   per panel w, last first (rows hist_r[w] .. are final: every later panel is already out):
 *     k_basis_prep   coefficient column = word w of the echelon rows ABOVE the panel, and the
 *                    tiles E[s][c] = F row of pivot column c (zero for free columns);
 *     k_sweep        the same Four-Russians row-XOR sweep as the forward elimination, on F,
 *                    rows [0, hist_r[w]):  F_i ^= XOR_c U[i, c] F_row(c);
 *   k_basis_scatter  vector i = e_{f_i} + sum_j F[j, i] e_{piv_j}, written as rows of
 *                    ceil(n/64) words for the host.
 *
 * Coefficients are read from the echelon form itself: a pivot row is zero left of its own
 * panel (pivot AND free columns -- a column is free because no active row had a 1 there), so
 * eliminating a later panel never changes an earlier panel's columns.
 *
 * Algorithmic bytes: gather r * d / 8 written (+ r * d scattered bit reads), sweeps
 * sum_w 2 * hist_r[w] * d / 8  ~  r^2 d / 512 * (64 / k),  scatter d * n / 8 written.
 */
#pragma once
#include "gf2b200_kernels.cuh"

namespace gf2b200 {

/* F[i][jj] = U[i][freecols[jj]] for i < r.  One thread per (row, 64-bit word of F);
 * consecutive threads take consecutive rows.  An echelon row of panel w_i is ZERO left of
 * word w_i by construction, but the elimination never rewrites those words (the sweeps start
 * at the strip of the next panel word), so what memory holds there is stale: columns in
 * words < w_i read as 0 (the rule the per-column back-substitution applied as `wd >= p`). */
__global__ void k_basis_gather(Mat M, Mat F, const long long *__restrict__ freecols,
                               const long long *__restrict__ hist_r, long long r, long long d) {
	const int WT = F.ns * SW;
	const long long total = r * WT;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		const int jw = (int)(t / r);
		const long long i = t - (long long)jw * r;
		/* panel of row i: the last w with hist_r[w] <= i (hist_r is non-decreasing) */
		int lo = 0, hi = M.nw - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (hist_r[mid] <= i) lo = mid;
			else hi = mid - 1;
		}
		const int wi = lo;
		u64 v = 0;
		const long long j0 = (long long)jw * 64;
		const int nb = (int)max(0LL, min(64LL, d - j0));
		for (int b = 0; b < nb; b++) {
			const long long f = freecols[j0 + b];
			const int fw = (int)(f >> 6);
			if (fw >= wi) v |= ((M.base[widx(M, i, fw)] >> (f & 63)) & 1ULL) << b;
		}
		F.base[widx(F, i, jw)] = v;
	}
}

/* Backward panel w: pc[i] = word w of echelon row i (i < r_w); E tiles of F for every strip;
 * the three fields of the panel description k_sweep reads. */
__global__ void k_basis_prep(Mat M, Mat F, int w, u64 pm, long long r_w, u64 *__restrict__ pc,
                             uint4 *__restrict__ ebufF, PanelDesc *pd) {
	const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	const long long gs = (long long)gridDim.x * blockDim.x;
	if (gt == 0) {
		pd->k = __popcll(pm);
		pd->r = 0;
		pd->r1 = 0;
		pd->pm = pm;
		pd->nmove = 0;
	}
	for (long long i = gt; i < r_w; i += gs) pc[i] = M.base[widx(M, i, w)];
	const uint4 *fb = reinterpret_cast<const uint4 *>(F.base);
	const long long tiles = (long long)F.ns * EBUF_Q;
	for (long long t = gt; t < tiles; t += gs) {
		const int s = (int)(t / EBUF_Q), q = (int)(t % EBUF_Q);
		const int c = q / SQ, ch = q % SQ;
		uint4 v = make_uint4(0, 0, 0, 0);
		if ((pm >> c) & 1) {
			const long long row = r_w + __popcll(pm & ((1ULL << c) - 1));
			v = fb[((long long)s * F.mp + row) * SQ + ch];
		}
		ebufF[t] = v;
	}
}

/* Rows [i0, i0 + cnt) of the basis: out[ii][wo] for every word wo of the unknowns. */
__global__ void k_basis_scatter(Mat F, const long long *__restrict__ hist_r, const u64 *__restrict__ hist_pm,
                                const long long *__restrict__ freecols, int nw, long long i0, long long cnt,
                                u64 *__restrict__ out) {
	const long long total = cnt * nw;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		const int wo = (int)(t / cnt);
		const long long ii = t - (long long)wo * cnt;
		const long long i = i0 + ii;
		u64 pm = hist_pm[wo];
		long long j = hist_r[wo];
		u64 v = 0;
		const int fw = (int)(i >> 6), fb = (int)(i & 63);
		while (pm) {
			const int c = __ffsll((long long)pm) - 1;
			pm &= pm - 1;
			v |= ((F.base[widx(F, j, fw)] >> fb) & 1ULL) << c;
			j++;
		}
		const long long f = freecols[i];
		if ((int)(f >> 6) == wo) v |= 1ULL << (f & 63);
		out[ii * nw + wo] = v;
	}
}

/* ---- row-sharded systems --------------------------------------------------------------
 * Every shard keeps F for ITS echelon rows (local rows 0 .. r_local-1, in panel order).  The k
 * pivot rows of a panel are spread over the shards (hist_owner[w * 64 + j] = owner of the j-th
 * pivot of panel w); per backward panel each shard contributes the F rows it owns to the tile,
 * the contributions are all-gathered (peers' HBM / NCCL -- 64 x d / 8 bytes per shard and panel,
 * where the per-column scheme ran a whole back-substitution with its own exchanges per free
 * column), and every shard sweeps its own rows above the panel. */
__device__ __forceinline__ long long owned_before(const unsigned char *__restrict__ hist_owner, int w, int j, int me) {
	long long c = 0;
	for (int jj = 0; jj < j; jj++) c += hist_owner[(long long)w * 64 + jj] == me;
	return c;
}

/* this shard's contribution to the tile of backward panel w (zero where another shard owns the pivot) */
__global__ void k_basis_tile(Mat F, int w, u64 pm, long long r_w, const unsigned char *__restrict__ hist_owner, int me,
                             uint4 *__restrict__ tile) {
	const uint4 *fb = reinterpret_cast<const uint4 *>(F.base);
	const long long tiles = (long long)F.ns * EBUF_Q;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tiles; t += (long long)gridDim.x * blockDim.x) {
		const int s = (int)(t / EBUF_Q), q = (int)(t % EBUF_Q);
		const int c = q / SQ, ch = q % SQ;
		uint4 v = make_uint4(0, 0, 0, 0);
		if ((pm >> c) & 1) {
			const int j = __popcll(pm & ((1ULL << c) - 1));
			if (hist_owner[(long long)w * 64 + j] == me)
				v = fb[((long long)s * F.mp + r_w + owned_before(hist_owner, w, j, me)) * SQ + ch];
		}
		tile[t] = v;
	}
}

/* k_basis_prep for a shard: the tile is the OR of the G gathered contributions */
__global__ void k_basis_prep_sharded(Mat M, Mat F, int w, u64 pm, long long r_w, u64 *__restrict__ pc,
                                     uint4 *__restrict__ ebufF, PanelDesc *pd, const uint4 *__restrict__ tiles_all, int G) {
	const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	const long long gs = (long long)gridDim.x * blockDim.x;
	if (gt == 0) {
		pd->k = __popcll(pm);
		pd->r = 0;
		pd->r1 = 0;
		pd->pm = pm;
		pd->nmove = 0;
	}
	for (long long i = gt; i < r_w; i += gs) pc[i] = M.base[widx(M, i, w)];
	const long long tiles = (long long)F.ns * EBUF_Q;
	for (long long t = gt; t < tiles; t += gs) {
		uint4 v = tiles_all[t];
		for (int g = 1; g < G; g++) {
			const uint4 o = tiles_all[(long long)g * tiles + t];
			v.x |= o.x;
			v.y |= o.y;
			v.z |= o.z;
			v.w |= o.w;
		}
		ebufF[t] = v;
	}
}

/* this shard's part of rows [i0, i0 + cnt) of the basis: the bits at the pivot columns it owns
 * (shard 0 adds the 1 at the free column itself); the parts are OR-ed after an all-gather */
__global__ void k_basis_scatter_sharded(Mat F, const long long *__restrict__ hist_r, const u64 *__restrict__ hist_pm,
                                        const unsigned char *__restrict__ hist_owner, int me,
                                        const long long *__restrict__ freecols, int nw, long long i0, long long cnt,
                                        u64 *__restrict__ out) {
	const long long total = cnt * nw;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		const int wo = (int)(t / cnt);
		const long long ii = t - (long long)wo * cnt;
		const long long i = i0 + ii;
		u64 pm = hist_pm[wo];
		long long j = hist_r[wo];
		u64 v = 0;
		const int fw = (int)(i >> 6), fb = (int)(i & 63);
		for (int jj = 0; pm; jj++) {
			const int c = __ffsll((long long)pm) - 1;
			pm &= pm - 1;
			if (hist_owner[(long long)wo * 64 + jj] == me) {
				v |= ((F.base[widx(F, j, fw)] >> fb) & 1ULL) << c;
				j++;
			}
		}
		const long long f = freecols[i];
		if (me == 0 && (int)(f >> 6) == wo) v |= 1ULL << (f & 63);
		out[ii * nw + wo] = v;
	}
}

/* dst[t] = OR over the G gathered parts */
__global__ void k_or_parts(u64 *__restrict__ dst, const u64 *__restrict__ all, long long words, int G) {
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < words; t += (long long)gridDim.x * blockDim.x) {
		u64 v = all[t];
		for (int g = 1; g < G; g++) v |= all[(long long)g * words + t];
		dst[t] = v;
	}
}

} /* namespace gf2b200 */
