/*
 * gf2_oracle.c -- CPU ORACLE for the gf2bv hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of what the reference computes on the path
 *   LinearSystem.solve_one / solve_all  ->  _internal.m4ri_solve  ->  M4RI
 * (reference: gf2bv/_internal.c:359-502 control flow, :309-357 kernel basis,
 *  :32-59 bit order in/out, :63-122 + :181-204 + :242-273 enumeration / get).
 *
 * The arithmetic itself lives in M4RI (malb/m4ri; setup.py:14-17 pins release
 * 20260122, sha256 68196ed4...d4cf), a third-party library that is NOT vendored
 * in /root/reference and is absent from this image, so it cannot be compiled
 * here.  What is restated is M4RI's *published semantics* at the reference's
 * call sites:
 *   - _mzd_pluq (_internal.c:433): pivot columns = left-to-right column rank
 *     profile of A; an invariant of A, independent of row pivoting.
 *   - _mzd_pluq_solve_left(..., inconsistency_check=1) (_internal.c:440): the
 *     unique x with A x = b and x_f = 0 on every free column f; -1 (-> None)
 *     when inconsistent.
 *   - _mzd_kernel_left_pluq (_internal.c:309-357): basis vector i has a 1 at free
 *     column sigma(r+i), 0 at the other free columns, where sigma is the column
 *     arrangement produced by Q's transposition sequence (swap i <-> p_i).
 *
 * PARITY PINNING.  Unique-solution systems are pinned by the reference's own
 * example asserts (examples/mt.py:38 seed 3142, lfsr.py:20, xoshiro.py:16,
 * README.md:73) -- tests/golden holds those vectors.  For UNDERDETERMINED
 * systems the reference holds no golden vector (which particular solution, the
 * kernel-basis order, solve_all order): PARITY UNPINNED there; this oracle pins
 * them to M4RI's documented semantics above and says so.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this.  The product (gf2bv_b200) never does.
 *
 * Layout: row-major, `stride` 64-bit words per row, bit j of row i is
 *   (A[i*stride + j/64] >> (j%64)) & 1          (M4RI's mzd_t convention).
 * b: m bits, LSB-first packed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GF2O_OK 0
#define GF2O_INCONSISTENT 1
#define GF2O_ENOMEM -1

typedef struct {
	int32_t status;      /* 0 ok, 1 inconsistent, <0 error */
	int64_t rank;
	int64_t kernel_dim;  /* n - rank when mode==1, else 0 */
	uint64_t *origin;    /* ceil(n/64) words */
	uint64_t *basis;     /* kernel_dim x ceil(n/64) words, sigma order */
	int64_t *pivcols;    /* rank entries, ascending */
	double t_forward;    /* seconds (tier 2 only) */
	double t_backward;
} gf2o_result;

static inline int getbit(const uint64_t *row, int64_t j) {
	return (int)((row[j >> 6] >> (j & 63)) & 1);
}

void gf2o_result_free(gf2o_result *r) {
	if (!r) return;
	free(r->origin);
	free(r->basis);
	free(r->pivcols);
	r->origin = r->basis = NULL;
	r->pivcols = NULL;
}

int gf2o_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads(); /* after gf2o_set_threads(n): n */
#else
	return 1;
#endif
}

static double now_s(void) {
#ifdef _OPENMP
	return omp_get_wtime();
#else
	return 0.0;
#endif
}

/* sigma = arrangement of column indices after "for i<r: swap(i, p_i)"
 * (M4RI mzp_t = LAPACK-style transposition sequence, applied by
 * mzd_apply_p_left_trans at _internal.c:348).  The free column of basis vector
 * i is sigma[r+i]. */
void gf2o_sigma_order(int64_t n, int64_t r, const int64_t *piv, int64_t *sigma) {
	for (int64_t i = 0; i < n; i++) sigma[i] = i;
	for (int64_t i = 0; i < r; i++) {
		int64_t t = sigma[i];
		sigma[i] = sigma[piv[i]];
		sigma[piv[i]] = t;
	}
}

/* ---------------------------------------------------------------------------
 * Tier 1: schoolbook Gauss-Jordan.  This is THE SPEC: column by column, left to
 * right; a column is a pivot iff some not-yet-used row has a 1 there after
 * reduction by the earlier pivots (= column rank profile, _internal.c:433).
 * ------------------------------------------------------------------------- */
int gf2o_solve_schoolbook(const uint64_t *A, const uint64_t *b, int64_t m, int64_t n,
                          int64_t stride, int mode, gf2o_result *out) {
	memset(out, 0, sizeof(*out));
	int64_t W = (n + 1 + 63) / 64; /* augmented [A | b], b at bit n */
	int64_t nw = (n + 63) / 64;
	uint64_t *M = (uint64_t *)calloc((size_t)(m ? m : 1) * (size_t)W, 8);
	int64_t *piv = (int64_t *)malloc((size_t)(n ? n : 1) * 8);
	if (!M || !piv) { free(M); free(piv); return out->status = GF2O_ENOMEM; }
	for (int64_t i = 0; i < m; i++) {
		uint64_t *row = M + i * W;
		memcpy(row, A + i * stride, (size_t)nw * 8);
		/* bits above `cols` are ignored (_internal.c:45,48) */
		if (n & 63) row[nw - 1] &= (1ULL << (n & 63)) - 1;
		if (b && ((b[i >> 6] >> (i & 63)) & 1)) row[n >> 6] |= 1ULL << (n & 63);
	}
	int64_t r = 0;
	for (int64_t c = 0; c < n && r < m; c++) {
		int64_t p = -1;
		for (int64_t i = r; i < m; i++)
			if (getbit(M + i * W, c)) { p = i; break; }
		if (p < 0) continue;
		if (p != r) {
			for (int64_t w = 0; w < W; w++) {
				uint64_t t = M[r * W + w];
				M[r * W + w] = M[p * W + w];
				M[p * W + w] = t;
			}
		}
		const uint64_t *pr = M + r * W;
		for (int64_t i = 0; i < m; i++) {
			if (i == r || !getbit(M + i * W, c)) continue;
			uint64_t *row = M + i * W;
			for (int64_t w = c >> 6; w < W; w++) row[w] ^= pr[w];
		}
		piv[r++] = c;
	}
	/* consistency: a zero A-row with b = 1 (_internal.c:440-446 -> None) */
	int bad = 0;
	for (int64_t i = r; i < m && !bad; i++) bad = getbit(M + i * W, n);
	out->rank = r;
	int rc = GF2O_OK;
	if (bad) {
		rc = GF2O_INCONSISTENT;
	} else {
		out->origin = (uint64_t *)calloc((size_t)(nw ? nw : 1), 8);
		out->pivcols = (int64_t *)malloc((size_t)(r ? r : 1) * 8);
		memcpy(out->pivcols, piv, (size_t)r * 8);
		/* free variables = 0; pivot variables = reduced b (_internal.c:449-454) */
		for (int64_t i = 0; i < r; i++)
			if (getbit(M + i * W, n)) out->origin[piv[i] >> 6] |= 1ULL << (piv[i] & 63);
		if (mode == 1 && r < n) {
			/* R = Q^T [U1^-1 U2 ; I] (_internal.c:330-348): in RREF, U1^-1 U2 is
			 * just the free columns of the pivot rows. */
			int64_t d = n - r;
			int64_t *sigma = (int64_t *)malloc((size_t)n * 8);
			out->basis = (uint64_t *)calloc((size_t)d * (size_t)nw, 8);
			gf2o_sigma_order(n, r, piv, sigma);
			for (int64_t i = 0; i < d; i++) {
				int64_t f = sigma[r + i];
				uint64_t *v = out->basis + i * nw;
				v[f >> 6] |= 1ULL << (f & 63);
				for (int64_t j = 0; j < r; j++)
					if (getbit(M + j * W, f)) v[piv[j] >> 6] |= 1ULL << (piv[j] & 63);
			}
			free(sigma);
			out->kernel_dim = d;
		}
	}
	free(M);
	free(piv);
	return out->status = rc;
}

/* ---------------------------------------------------------------------------
 * Tier 2: blocked Method-of-Four-Russians port (the timed CPU baseline).
 * Same semantics as tier 1; 64-column panels, eight 256-entry XOR tables per
 * panel and column chunk, OpenMP over column chunks.  This is "a CPU port of the
 * M4RI path (PLUQ-style elimination + triangular solves)", NOT M4RI itself.
 *
 * Augmented layout inside: W = nw + 1 words per row, b alone in word nw bit 0.
 * Forward elimination leaves rows 0..r-1 as an echelon form whose rows are
 * fully reduced inside their own 64-column panel; back-substitution then gives
 * the particular solution and (mode 1) one kernel vector per free column.
 * ------------------------------------------------------------------------- */
/* Column-chunk width of the sweep (64-bit words).  Chosen per solve so that there are
 * at least ~3 chunks per thread at any n (the first version used a fixed 64 words: at
 * n = 16384 that left 5 chunks for 16 cores) and so that a thread's eight 256-entry
 * tables (8 x 256 x cw x 8 B) stay inside its L2. */
#define CHUNK_MAX 32

static int g_threads_override = 0;

/* bench.py sets the thread count explicitly (torchrun exports OMP_NUM_THREADS=1) */
void gf2o_set_threads(int n) {
	g_threads_override = n > 0 ? n : 0;
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#endif
}

static int64_t chunk_words(int64_t W, int nthr) {
	int64_t cw = (W + 3 * (int64_t)nthr - 1) / (3 * (int64_t)nthr);
	cw = (cw + 7) / 8 * 8;
	if (cw < 8) cw = 8;
	if (cw > CHUNK_MAX) cw = CHUNK_MAX;
	return cw;
}

static int forward_m4rm(uint64_t *M, int64_t m, int64_t n, int64_t W, int64_t *piv,
                        int64_t *rank_out) {
	int64_t nw = (n + 63) / 64;
	int64_t r = 0;
	int nthr = gf2o_threads();
	const int64_t CH = chunk_words(W, nthr);
	uint64_t *E = (uint64_t *)malloc((size_t)64 * (size_t)W * 8); /* reduced pivot rows */
	uint64_t *tmp = (uint64_t *)malloc((size_t)W * 8);
	uint64_t *coef = (uint64_t *)malloc((size_t)(m ? m : 1) * 8);
	uint64_t *T = (uint64_t *)aligned_alloc(64, (size_t)nthr * 8 * 256 * CH * 8);
	if (!E || !tmp || !coef || !T) { free(E); free(tmp); free(coef); free(T); return GF2O_ENOMEM; }
	for (int64_t w = 0; w < nw && r < m; w++) {
		uint64_t colmask = ~0ULL;
		if (w == nw - 1 && (n & 63)) colmask = (1ULL << (n & 63)) - 1;
		/* 1. select original rows whose 64-bit slices form a basis of the slice
		 *    row space (XOR-basis insertion keyed by lowest set bit = leftmost
		 *    column); the set of keys is the panel's column rank profile */
		int64_t sel[64];
		uint64_t basis[64];
		uint64_t pm = 0;
		int k = 0;
		memset(basis, 0, sizeof(basis));
		for (int64_t i = r; i < m && pm != colmask; i++) {
			uint64_t v = M[i * W + w] & colmask;
			while (v) {
				int c = __builtin_ctzll(v);
				if (!basis[c]) { basis[c] = v; pm |= 1ULL << c; sel[k++] = i; break; }
				v ^= basis[c];
			}
		}
		if (k == 0) continue;
		/* 2. RREF of the k selected slices, tracking the row transform t[] */
		uint64_t v[64], t[64];
		int used[64], pcol[64], order[64], slot_of_bit[64];
		for (int l = 0; l < k; l++) { v[l] = M[sel[l] * W + w] & colmask; t[l] = 1ULL << l; used[l] = 0; pcol[l] = -1; }
		for (int c = 0; c < 64; c++) {
			if (!((pm >> c) & 1)) continue;
			int p = -1;
			for (int l = 0; l < k; l++) if (!used[l] && ((v[l] >> c) & 1)) { p = l; break; }
			for (int l = 0; l < k; l++) if (l != p && ((v[l] >> c) & 1)) { v[l] ^= v[p]; t[l] ^= t[p]; }
			used[p] = 1; pcol[p] = c;
		}
		for (int l = 0; l < k; l++) order[__builtin_popcountll(pm & ((1ULL << pcol[l]) - 1))] = l;
		for (int c = 0; c < 64; c++) slot_of_bit[c] = -1;
		/* 3. E_j = XOR of the originals named by t[] (words w..W-1) */
#pragma omp parallel for schedule(static) if (W - w > 512)
		for (int j = 0; j < k; j++) {
			uint64_t *e = E + (int64_t)j * W;
			memset(e + w, 0, (size_t)(W - w) * 8);
			uint64_t tm = t[order[j]];
			while (tm) {
				int l = __builtin_ctzll(tm); tm &= tm - 1;
				const uint64_t *src = M + sel[l] * W;
				for (int64_t x = w; x < W; x++) e[x] ^= src[x];
			}
		}
		for (int j = 0; j < k; j++) {
			slot_of_bit[pcol[order[j]]] = j;
			piv[r + j] = w * 64 + pcol[order[j]];
		}
		/* 4. pivot rows -> rows r..r+k-1; displaced rows -> vacated positions */
		{
			int target_is_piv[64];
			int64_t vac[64];
			int nv = 0, vi = 0;
			for (int j = 0; j < k; j++) target_is_piv[j] = 0;
			for (int l = 0; l < k; l++) {
				if (sel[l] < r + k) target_is_piv[sel[l] - r] = 1; else vac[nv++] = sel[l];
			}
			for (int j = 0; j < k; j++) {
				if (target_is_piv[j]) continue;
				memcpy(tmp, M + (r + j) * W, (size_t)W * 8);
				memcpy(M + vac[vi++] * W, tmp, (size_t)W * 8);
			}
			for (int j = 0; j < k; j++)
				memcpy(M + (r + j) * W + w, E + (int64_t)j * W + w, (size_t)(W - w) * 8);
		}
		int64_t r1 = r + k;
		/* 5. sweep rows r1..m-1: row ^= sum_j coef_j E_j, coef = raw panel word
		 *    (E is RREF on the pivot columns so no sequential dependency).
		 *    Work items = (column chunk) x (row block), dealt out dynamically, so every
		 *    core has work at any n; a thread rebuilds its tables only when it moves
		 *    to another column chunk. */
#pragma omp parallel for schedule(static) if (m - r1 > 4096)
		for (int64_t i = r1; i < m; i++) coef[i] = M[i * W + w] & pm;
		const int64_t nchunk = (W - w + CH - 1) / CH;
		const int64_t rows = m - r1;
		int64_t rblocks = 1;
		if (nchunk < 3 * (int64_t)nthr) rblocks = (3 * (int64_t)nthr + nchunk - 1) / nchunk;
		if (rblocks > (rows + 255) / 256) rblocks = (rows + 255) / 256;
		if (rblocks < 1) rblocks = 1;
		const int64_t rb = (rows + rblocks - 1) / rblocks;
#pragma omp parallel
		{
#ifdef _OPENMP
			int tid = omp_get_thread_num();
#else
			int tid = 0;
#endif
			uint64_t *Tt = T + (size_t)tid * 8 * 256 * CH;
			int64_t built = -1;
#pragma omp for schedule(dynamic, 1)
			for (int64_t item = 0; item < nchunk * rblocks; item++) {
				const int64_t ci = item / rblocks, bi = item - ci * rblocks;
				const int64_t c0 = w + ci * CH;
				const int64_t cw = (W - c0 < CH) ? (W - c0) : CH;
				if (built != ci) {
					for (int g = 0; g < 8; g++) {
						uint64_t *Tg = Tt + (size_t)g * 256 * CH;
						memset(Tg, 0, (size_t)cw * 8);
						for (int idx = 1; idx < 256; idx++) {
							int s = slot_of_bit[g * 8 + __builtin_ctz(idx)];
							uint64_t *dst = Tg + (size_t)idx * CH;
							const uint64_t *pv = Tg + (size_t)(idx & (idx - 1)) * CH;
							if (s < 0) {
								memcpy(dst, pv, (size_t)cw * 8);
							} else {
								const uint64_t *e = E + (int64_t)s * W + c0;
								for (int64_t x = 0; x < cw; x++) dst[x] = pv[x] ^ e[x];
							}
						}
					}
					built = ci;
				}
				const int64_t i0 = r1 + bi * rb, i1 = (i0 + rb < m) ? i0 + rb : m;
				for (int64_t i = i0; i < i1; i++) {
					uint64_t cf = coef[i];
					if (!cf) continue;
					uint64_t *restrict row = M + i * W + c0;
					const uint64_t *restrict t0 = Tt + ((size_t)0 * 256 + (cf & 255)) * CH;
					const uint64_t *restrict t1 = Tt + ((size_t)1 * 256 + ((cf >> 8) & 255)) * CH;
					const uint64_t *restrict t2 = Tt + ((size_t)2 * 256 + ((cf >> 16) & 255)) * CH;
					const uint64_t *restrict t3 = Tt + ((size_t)3 * 256 + ((cf >> 24) & 255)) * CH;
					const uint64_t *restrict t4 = Tt + ((size_t)4 * 256 + ((cf >> 32) & 255)) * CH;
					const uint64_t *restrict t5 = Tt + ((size_t)5 * 256 + ((cf >> 40) & 255)) * CH;
					const uint64_t *restrict t6 = Tt + ((size_t)6 * 256 + ((cf >> 48) & 255)) * CH;
					const uint64_t *restrict t7 = Tt + ((size_t)7 * 256 + ((cf >> 56) & 255)) * CH;
					for (int64_t x = 0; x < cw; x++)
						row[x] ^= t0[x] ^ t1[x] ^ t2[x] ^ t3[x] ^ t4[x] ^ t5[x] ^ t6[x] ^ t7[x];
				}
			}
		}
		r = r1;
	}
	free(E);
	free(tmp);
	free(coef);
	free(T);
	*rank_out = r;
	return 0;
}

/* x[piv[j]] = rhs_j ^ <U_j[words >= own], x>, j = r-1 .. 0; x preloaded with the
 * free-variable assignment */
static void back_substitute(const uint64_t *M, int64_t W, int64_t nw, int64_t r,
                            const int64_t *piv, int use_b, uint64_t *x) {
	for (int64_t j = r - 1; j >= 0; j--) {
		const uint64_t *row = M + j * W;
		int64_t p = piv[j];
		uint64_t acc = 0;
		for (int64_t w = p >> 6; w < nw; w++) acc ^= row[w] & x[w];
		int bit = __builtin_parityll(acc);
		if (use_b) bit ^= (int)(row[nw] & 1);
		if (bit) x[p >> 6] |= 1ULL << (p & 63);
	}
}

int gf2o_solve_m4rm(const uint64_t *A, const uint64_t *b, int64_t m, int64_t n,
                    int64_t stride, int mode, gf2o_result *out) {
	memset(out, 0, sizeof(*out));
	int64_t nw = (n + 63) / 64;
	int64_t W = nw + 1;
	uint64_t *M = (uint64_t *)malloc((size_t)(m ? m : 1) * (size_t)W * 8);
	int64_t *piv = (int64_t *)malloc((size_t)(n ? n : 1) * 8);
	if (!M || !piv) { free(M); free(piv); return out->status = GF2O_ENOMEM; }
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < m; i++) {
		uint64_t *row = M + i * W;
		memcpy(row, A + i * stride, (size_t)nw * 8);
		if (n & 63) row[nw - 1] &= (1ULL << (n & 63)) - 1;
		row[nw] = b ? ((b[i >> 6] >> (i & 63)) & 1) : 0;
	}
	double t0 = now_s();
	int64_t r = 0;
	int rc = forward_m4rm(M, m, n, W, piv, &r);
	if (rc) { free(M); free(piv); return out->status = rc; }
	double t1 = now_s();
	out->t_forward = t1 - t0;
	out->rank = r;
	int bad = 0;
	for (int64_t i = r; i < m && !bad; i++) bad = (int)(M[i * W + nw] & 1);
	if (bad) {
		rc = GF2O_INCONSISTENT;
	} else {
		out->origin = (uint64_t *)calloc((size_t)(nw ? nw : 1), 8);
		out->pivcols = (int64_t *)malloc((size_t)(r ? r : 1) * 8);
		memcpy(out->pivcols, piv, (size_t)r * 8);
		back_substitute(M, W, nw, r, piv, 1, out->origin);
		if (mode == 1 && r < n) {
			int64_t d = n - r;
			int64_t *sigma = (int64_t *)malloc((size_t)n * 8);
			out->basis = (uint64_t *)calloc((size_t)d * (size_t)nw, 8);
			gf2o_sigma_order(n, r, piv, sigma);
#pragma omp parallel for schedule(dynamic, 1)
			for (int64_t i = 0; i < d; i++) {
				int64_t f = sigma[r + i];
				uint64_t *v = out->basis + i * nw;
				v[f >> 6] |= 1ULL << (f & 63);
				back_substitute(M, W, nw, r, piv, 0, v);
			}
			free(sigma);
			out->kernel_dim = d;
		}
	}
	out->t_backward = now_s() - t1;
	free(M);
	free(piv);
	return out->status = rc;
}

/* ---------------------------------------------------------------------------
 * Synthetic dense inputs: stateless, so the GPU, every shard and
 * the CPU regenerate identical rows.
 *   word(i, w) = mix(seed + PHI * (i*nw + w + 1)),  mix = splitmix64 finaliser
 *   x*         = the same generator with seed ^ 0xB200 (nw words, tail masked)
 *   b          = A x*
 * ------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
	z ^= z >> 27; z *= 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return z;
}
#define PHI 0x9E3779B97F4A7C15ULL

void gf2o_synth_xstar(int64_t n, uint64_t seed, uint64_t *x) {
	int64_t nw = (n + 63) / 64;
	for (int64_t w = 0; w < nw; w++) x[w] = mix64((seed ^ 0xB200ULL) + PHI * (uint64_t)(w + 1));
	if (n & 63) x[nw - 1] &= (1ULL << (n & 63)) - 1;
}

/* A: m x nw words (stride nw), tail bits masked; b: ceil(m/64) words */
void gf2o_synth(int64_t m, int64_t n, uint64_t seed, uint64_t *A, uint64_t *b) {
	int64_t nw = (n + 63) / 64;
	uint64_t *x = (uint64_t *)malloc((size_t)nw * 8);
	gf2o_synth_xstar(n, seed, x);
	memset(b, 0, (size_t)((m + 63) / 64) * 8);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < m; i++) {
		uint64_t *row = A + i * nw;
		uint64_t acc = 0;
		for (int64_t w = 0; w < nw; w++) {
			uint64_t v = mix64(seed + PHI * (uint64_t)(i * nw + w + 1));
			if (w == nw - 1 && (n & 63)) v &= (1ULL << (n & 63)) - 1;
			row[w] = v;
			acc ^= v & x[w];
		}
		if (__builtin_parityll(acc)) {
#pragma omp atomic
			b[i >> 6] |= 1ULL << (i & 63);
		}
	}
	free(x);
}

/* residual check: returns number of rows with A x != b */
int64_t gf2o_residual(const uint64_t *A, const uint64_t *b, int64_t m, int64_t n,
                      int64_t stride, const uint64_t *x) {
	int64_t nw = (n + 63) / 64, bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
	for (int64_t i = 0; i < m; i++) {
		uint64_t acc = 0;
		for (int64_t w = 0; w < nw; w++) {
			uint64_t v = A[i * stride + w];
			if (w == nw - 1 && (n & 63)) v &= (1ULL << (n & 63)) - 1;
			acc ^= v & x[w];
		}
		int bi = b ? (int)((b[i >> 6] >> (i & 63)) & 1) : 0;
		bad += (__builtin_parityll(acc) != bi);
	}
	return bad;
}
