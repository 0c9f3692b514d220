/*
 * gf2b200.h -- C-ABI of libgf2b200.so, the B200 (sm_100a) GF(2) solver that sits
 * where M4RI sits behind gf2bv's `_internal.m4ri_solve`.
 *
 * Plain C types only: no Python, no torch, no exceptions, no abort() across the
 * boundary.  Every call returns 0 on success or a negative GF2B200_E* code; the
 * message is available from gf2b200_last_error().  The library has NO CPU
 * fallback: without a CUDA device gf2b200_create() fails with GF2B200_ENODEV.
 *
 * Bit conventions (identical to the reference's mzd_t usage, _internal.c:411-425
 * and :32-39): A is row-major, `stride64` 64-bit words per row, bit j of row i is
 * (A[i*stride64 + j/64] >> (j%64)) & 1; b is m bits packed LSB-first; solution
 * word w bit k is x_{64w+k}.
 *
 * What each entry point replaces in the reference (gf2bv/_internal.c):
 *   gf2b200_solve            :429-489  _mzd_pluq (:433) + _mzd_pluq_solve_left
 *                                      (:440) + transpose (:449-454) and, mode 1,
 *                                      _mzd_kernel_left_pluq (:309-357, :474-489)
 *   gf2b200_result_free      :467-470 / :285-290  (mzd_free of sol0 / tker)
 *   gf2b200_system_*         the same steps split so a caller can keep the
 *                            matrix resident in HBM (bench `value`, multi-GPU)
 */
#ifndef GF2B200_H
#define GF2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GF2B200_ABI_VERSION 3

#define GF2B200_OK 0
#define GF2B200_INCONSISTENT 1 /* result.status only: system has no solution (-> None) */
#define GF2B200_EINVAL (-1)
#define GF2B200_ENODEV (-2)  /* no CUDA device / driver: there is no CPU fallback */
#define GF2B200_ECUDA (-3)
#define GF2B200_ENOMEM (-4)
#define GF2B200_ENCCL (-5)

typedef struct gf2b200_ctx gf2b200_ctx;
typedef struct gf2b200_system gf2b200_system;

/* Result of one solve.  All pointers are malloc'd by the library and released by
 * gf2b200_result_free().  Replaces the (sol0, tker) pair of _internal.c:436-489. */
typedef struct {
	int32_t status;      /* GF2B200_OK, GF2B200_INCONSISTENT or <0 */
	int64_t rank;
	int64_t kernel_dim;  /* n - rank when mode == 1, else 0 */
	uint64_t *origin;    /* ceil(n/64) words: x with every free variable 0 (NULL if inconsistent) */
	uint64_t *basis;     /* kernel_dim x ceil(n/64) words, M4RI's sigma order (mode 1) */
	int64_t *pivcols;    /* rank ascending pivot columns (column rank profile) */
} gf2b200_result;

/* Timing / work counters of the last gf2b200_system_eliminate (on-device times
 * from CUDA events on the solver's stream). */
typedef struct {
	double ms_total;        /* forward + consistency + back-substitution */
	double ms_forward;
	double ms_backward;
	double ms_sweep;        /* sum of k_sweep durations (profile mode only, else 0) */
	double sweep_bytes;     /* algorithmic bytes of all sweeps: sum 2 * rows * strip bytes * strips */
	double exchange_bytes;  /* multi-GPU: bytes this rank contributed to collectives */
	int64_t sweep_launches; /* sweeps that had work (k > 0 and active rows) */
	int64_t kernel_launches; /* all kernels launched by the elimination */
	int64_t panels;
	int64_t rank;
	int64_t m_local;        /* rows held by this rank */
	double ms_sweep_max;    /* profile mode: the longest single sweep and its bytes */
	double sweep_bytes_max;
	double sweep_bytes_timed;     /* profile mode: algorithmic bytes of the launches in ms_sweep */
	int64_t sweep_launches_timed; /* (a loopback context times its first shard only) */
	int64_t forward_kernel_launches; /* 1: the forward elimination ran as ONE persistent kernel (k_forward);
	                                  * ms_sweep is then that kernel's duration (always measured),
	                                  * sweep_launches counts its panels with work and ms_sweep_max /
	                                  * sweep_bytes_max describe its longest panel.  0: per-panel launches */
	/* last gf2b200_system_result(mode 1) on a one-GPU system: the blocked multi-right-hand-side
	 * triangular solve of the kernel basis (replaces mzd_trsm_upper_left, _internal.c:343) */
	double ms_basis_solve;    /* gather of the free columns + every backward sweep */
	double ms_basis_output;   /* scatter into basis vectors + D2H */
	double basis_sweep_bytes; /* algorithmic bytes of the backward sweeps: sum 2 * rows above * row bytes of F */
	int64_t basis_panels;
} gf2b200_stats;

int gf2b200_abi_version(void);
int gf2b200_device_count(void);

/* One context per caller thread and device (stream, workspaces). */
int gf2b200_create(gf2b200_ctx **out, int device);
/* Multi-GPU context: one process per GPU of one NVLink-connected node; `nccl_id128`
 * is the 128-byte ncclUniqueId from gf2b200_nccl_unique_id() on rank 0, distributed
 * by the caller.  NCCL is the rendezvous only: systems of such a context map each
 * other's HBM with CUDA IPC and exchange pivot rows with NVLink loads/stores, so
 * gf2b200_system_create / _destroy are collective (same order on every rank). */
int gf2b200_nccl_unique_id(void *out_id128);
int gf2b200_create_dist(gf2b200_ctx **out, int device, int rank, int world,
                        const void *nccl_id128);
/* Loopback context: `world` row shards on ONE device in this process; the
 * exchanges of the sharded algorithm become device copies.  Same kernels and
 * control flow as the multi-process kind -- used to parity-test the sharded path on one
 * GPU.  Systems of such a context take the WHOLE matrix in system_load_*. */
int gf2b200_create_shards(gf2b200_ctx **out, int device, int world);
void gf2b200_destroy(gf2b200_ctx *ctx);
const char *gf2b200_last_error(const gf2b200_ctx *ctx);

/* Page-locked host memory for the caller's packed A / b (full-speed, truly
 * asynchronous H2D).  The extension packs PyLongs straight into such a buffer
 * (replaces mzd_init for A and B, _internal.c:398-399). */
int gf2b200_host_alloc(void **out, size_t bytes);
void gf2b200_host_free(void *p);

/* Run on the caller's CUDA stream (a cudaStream_t passed as void*); NULL = the
 * context's own stream.  profile != 0 brackets every sweep launch with events. */
int gf2b200_set_stream(gf2b200_ctx *ctx, void *cuda_stream);
int gf2b200_set_profile(gf2b200_ctx *ctx, int profile);

/* Host-buffer solve: H2D, eliminate, back-substitute, (mode 1) kernel basis, D2H.
 * mode 0 = particular solution only, 1 = + kernel basis (SOLVE_MODE_* of
 * _internal.h:25-26).  b may be NULL (homogeneous).  Requires m >= 1, n >= 1 (the
 * reference additionally requires m >= n at :390-395; that check stays in the
 * extension).  Returns 0 and sets out->status (OK / INCONSISTENT). */
int gf2b200_solve(gf2b200_ctx *ctx, const uint64_t *A, const uint64_t *b, int64_t m,
                  int64_t n, int64_t stride64, int mode, gf2b200_result *out);
void gf2b200_result_free(gf2b200_result *res);
/* gf2b200_solve in two halves, for a caller that streams the rows in between
 * (gf2b200_system_load_begin / _rows / _end): open hands out the context's cached system of
 * that shape (or a new one); close eliminates, fetches the result and returns the system to the
 * cache.  close with mode < 0 gives the system up without solving (a load failed). */
int gf2b200_solve_open(gf2b200_ctx *ctx, int64_t m, int64_t n, gf2b200_system **out);
int gf2b200_solve_close(gf2b200_ctx *ctx, gf2b200_system *sys, int mode, gf2b200_result *out);

/* ---- device-resident systems -------------------------------------------- */
/* m, n are GLOBAL sizes; with a dist context each rank holds rows
 * [rank*m/world, (rank+1)*m/world) of the global system.  On a sharded system
 * (dist or loopback) system_result(mode 1) runs the same blocked multi-right-hand-side
 * triangular solve as on one GPU, every shard on its own echelon rows, with one small
 * exchange per backward panel; on a dist context every rank must call it and gets the
 * whole basis. */
int gf2b200_system_create(gf2b200_ctx *ctx, int64_t m, int64_t n, gf2b200_system **out);
void gf2b200_system_destroy(gf2b200_system *sys);
int64_t gf2b200_system_local_rows(const gf2b200_system *sys);
/* Load this rank's rows from host memory (pinned preferred) / device memory.
 * A points at the first local row; b is the packed bits of the LOCAL rows. */
int gf2b200_system_load_host(gf2b200_system *sys, const uint64_t *A, const uint64_t *b,
                             int64_t stride64);
int gf2b200_system_load_device(gf2b200_system *sys, const uint64_t *dA, const uint64_t *db,
                               int64_t stride64);
/* The same load handed over in blocks while the caller is still PRODUCING the rows: the
 * extension packs PyLongs on worker threads and passes each finished block on, so the H2D copy
 * and the layout kernel of a block overlap the packing of the next (the reference packs
 * everything, then calls M4RI: _internal.c:403-426, then :433).  begin; any number of
 * load_rows (local rows [row0, row0 + nrows), A_rows points at row row0, any order, no overlap;
 * the host memory must stay valid until load_end returns); end takes b (the packed bits of the
 * local rows, NULL = homogeneous) and returns when every copy has left the host buffers. */
int gf2b200_system_load_begin(gf2b200_system *sys, int64_t stride64);
int gf2b200_system_load_rows(gf2b200_system *sys, const uint64_t *A_rows, int64_t row0, int64_t nrows);
int gf2b200_system_load_end(gf2b200_system *sys, const uint64_t *b);
/* Dense synthetic system of SURVEY.md 8(d) generated in HBM: word(i,w) =
 * splitmix64-mix(seed + PHI*(i*ceil(n/64) + w + 1)), b = A x*, x* from seed^0xB200. */
int gf2b200_system_generate(gf2b200_system *sys, uint64_t seed);
/* Forward elimination (column rank profile), consistency check and
 * back-substitution of the particular solution; everything stays in HBM.
 * Asynchronous work is complete when this returns. */
int gf2b200_system_eliminate(gf2b200_system *sys);
/* Fetch the result (mode 1 additionally back-substitutes the kernel basis). */
int gf2b200_system_result(gf2b200_system *sys, int mode, gf2b200_result *out);
int gf2b200_system_stats(const gf2b200_system *sys, gf2b200_stats *out);
/* Number of rows i of the synthetic system (regenerated from `seed`) with
 * A_i x != b_i, summed over this rank's rows; x = ceil(n/64) host words. */
int gf2b200_system_check_synthetic(gf2b200_system *sys, uint64_t seed, const uint64_t *x,
                                   int64_t *bad_rows);
/* Rows [row0, row0+nrows) of the same synthetic system written row-major to HOST
 * memory (A: nrows x ceil(n/64) words, b: nrows bits packed from bit 0): the input
 * of the host-buffer (e2e) measurement.  A workload generator, not a solver. */
int gf2b200_synth_host(uint64_t *A, uint64_t *b, int64_t row0, int64_t nrows, int64_t n,
                       uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* GF2B200_H */
