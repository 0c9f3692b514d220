/*
 * internal_ext.c -- the CPython extension `gf2bv_b200._internal`.
 *
 * Drop-in for the reference's `gf2bv._internal` (gf2bv/_internal.c method table
 * :767-803, types :829-831): same exported names, argument meaning, return values
 * and error behaviour, but `m4ri_solve` hands the packed system to libgf2b200.so
 * (include/gf2b200.h, sm_100a CUDA) instead of M4RI.  There is NO CPU solver in
 * this file: without the library or a CUDA device m4ri_solve raises RuntimeError.
 *
 * Differences from the reference, all deliberate:
 *   - equations are packed a PyLong digit (30 bits) at a time straight into a
 *     pinned staging buffer, not bit by bit (reference :41-59, :403-426);
 *   - solutions become Python ints through the byte-array constructor, not an
 *     ASCII '0'/'1' string (reference :32-39);
 *   - AffineSpace owns plain uint64_t buffers instead of mzd_t (reference
 *     _internal.h:7-10);
 *   - AffineSpace.get() with the wrong arity returns the TypeError instead of
 *     reading args[0] anyway (reference :246-249 forgets the return);
 *   - eqs_to_sage_mat_helper (Sage/libgd bridge, reference :687-765) is outside the
 *     solve path and raises RuntimeError, as the reference does when libgd is absent.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <dlfcn.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gf2b200.h"

/* ---- PyLong internals (no public digit API; same situation as reference :5-16) */
#if PY_VERSION_HEX >= 0x030C0000
#define LONG_NDIGITS(o) ((Py_ssize_t)(((PyLongObject *)(o))->long_value.lv_tag >> 3))
#define LONG_DIGITS(o) (((PyLongObject *)(o))->long_value.ob_digit)
#else
#define LONG_NDIGITS(o) (Py_ABS(Py_SIZE(o)))
#define LONG_DIGITS(o) (((PyLongObject *)(o))->ob_digit)
#endif

/* ------------------------------------------------------------------------
 * libgf2b200.so, resolved lazily on the first solve (the reference resolves
 * libgd the same way, :690-712).
 * ---------------------------------------------------------------------- */
typedef struct {
	void *handle;
	int (*abi_version)(void);
	int (*create)(gf2b200_ctx **, int);
	void (*destroy)(gf2b200_ctx *);
	const char *(*last_error)(const gf2b200_ctx *);
	int (*solve)(gf2b200_ctx *, const uint64_t *, const uint64_t *, int64_t, int64_t, int64_t, int,
	             gf2b200_result *);
	void (*result_free)(gf2b200_result *);
	int (*host_alloc)(void **, size_t);
	void (*host_free)(void *);
	int (*solve_open)(gf2b200_ctx *, int64_t, int64_t, gf2b200_system **);
	int (*solve_close)(gf2b200_ctx *, gf2b200_system *, int, gf2b200_result *);
	int (*load_begin)(gf2b200_system *, int64_t);
	int (*load_rows)(gf2b200_system *, const uint64_t *, int64_t, int64_t);
	int (*load_end)(gf2b200_system *, const uint64_t *);
} shim_t;

static shim_t g_shim;
/* The reference's m4ri_solve can run concurrently from several Python threads (the GIL is
 * released around M4RI, _internal.c:429).  Here every concurrent call takes its own solver
 * slot -- a context (stream, HBM buffers of the last shape) plus a pinned staging buffer --
 * out of a small pool; the lock only guards the pool bookkeeping, never a solve. */
#define MAX_SLOTS 4
typedef struct {
	gf2b200_ctx *ctx;
	uint64_t *stage; /* pinned staging, grow-only */
	size_t stage_bytes;
	int busy;
} slot_t;
static slot_t g_slots[MAX_SLOTS];
static int g_one_slot; /* the kernel tests' CPU emulation build is loaded: never two solves at once */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER; /* guards g_slots[].busy and shim_load */
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;

/* GIL released.  Prefers a slot that already has a context. */
static slot_t *slot_acquire(void) {
	pthread_mutex_lock(&g_lock);
	for (;;) {
		slot_t *pick = NULL;
		const int nslots = g_one_slot ? 1 : MAX_SLOTS;
		for (int i = 0; i < nslots; i++)
			if (!g_slots[i].busy && g_slots[i].ctx) { pick = &g_slots[i]; break; }
		for (int i = 0; !pick && i < nslots; i++)
			if (!g_slots[i].busy) pick = &g_slots[i];
		if (pick) {
			pick->busy = 1;
			pthread_mutex_unlock(&g_lock);
			return pick;
		}
		pthread_cond_wait(&g_cv, &g_lock);
	}
}

static void slot_release(slot_t *sl) {
	pthread_mutex_lock(&g_lock);
	sl->busy = 0;
	pthread_cond_signal(&g_cv);
	pthread_mutex_unlock(&g_lock);
}

static int shim_load(void) {
	if (g_shim.handle) return 0;
	char path[4096];
	const char *env = getenv("GF2B200_LIB");
	if (env && *env) {
		snprintf(path, sizeof path, "%s", env);
	} else {
		Dl_info info;
		if (!dladdr((void *)&shim_load, &info) || !info.dli_fname) {
			PyErr_SetString(PyExc_RuntimeError, "gf2b200: cannot locate the extension on disk");
			return -1;
		}
		snprintf(path, sizeof path, "%s", info.dli_fname);
		char *slash = strrchr(path, '/');
		size_t dirlen = slash ? (size_t)(slash - path) + 1 : 0;
		snprintf(path + dirlen, sizeof path - dirlen, "libgf2b200.so");
	}
	void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!h) {
		PyErr_Format(PyExc_RuntimeError,
		             "gf2b200: cannot load %s (%s); build it with __graft_entry__.build() -- there is no CPU fallback",
		             path, dlerror());
		return -1;
	}
	/* tests/cpu_emu builds a CPU emulation of the kernels for kernel tests; such a build is
	 * never a way to run this extension without a GPU unless a test says so explicitly */
	if (dlsym(h, "gf2b200_emulated_build")) {
		const char *ok = getenv("GF2B200_TEST_EMULATION");
		if (!ok || strcmp(ok, "1") != 0) {
			PyErr_Format(PyExc_RuntimeError,
			             "gf2b200: %s is the CPU emulation build of tests/cpu_emu (test infrastructure); "
			             "refusing to use it -- there is no CPU fallback", path);
			dlclose(h);
			return -1;
		}
		g_one_slot = 1; /* the emulator is one OS thread by design: one solve at a time */
	}
	shim_t s;
	memset(&s, 0, sizeof s);
	s.handle = h;
#define RESOLVE(field, name)                                                             \
	do {                                                                                 \
		*(void **)(&s.field) = dlsym(h, name);                                           \
		if (!s.field) {                                                                  \
			PyErr_Format(PyExc_RuntimeError, "gf2b200: %s lacks symbol %s", path, name); \
			dlclose(h);                                                                  \
			return -1;                                                                   \
		}                                                                                \
	} while (0)
	RESOLVE(abi_version, "gf2b200_abi_version");
	RESOLVE(create, "gf2b200_create");
	RESOLVE(destroy, "gf2b200_destroy");
	RESOLVE(last_error, "gf2b200_last_error");
	RESOLVE(solve, "gf2b200_solve");
	RESOLVE(result_free, "gf2b200_result_free");
	RESOLVE(host_alloc, "gf2b200_host_alloc");
	RESOLVE(host_free, "gf2b200_host_free");
	RESOLVE(solve_open, "gf2b200_solve_open");
	RESOLVE(solve_close, "gf2b200_solve_close");
	RESOLVE(load_begin, "gf2b200_system_load_begin");
	RESOLVE(load_rows, "gf2b200_system_load_rows");
	RESOLVE(load_end, "gf2b200_system_load_end");
#undef RESOLVE
	if (s.abi_version() != GF2B200_ABI_VERSION) {
		PyErr_Format(PyExc_RuntimeError, "gf2b200: ABI version %d, extension built for %d", s.abi_version(),
		             GF2B200_ABI_VERSION);
		dlclose(h);
		return -1;
	}
	g_shim = s;
	return 0;
}

/* slot owned, GIL held */
static int ctx_ready(slot_t *sl) {
	if (sl->ctx) return 0;
	const char *dev = getenv("GF2B200_DEVICE");
	int rc = g_shim.create(&sl->ctx, dev ? atoi(dev) : 0);
	if (rc) {
		PyErr_Format(PyExc_RuntimeError, "gf2b200_create failed (%d): %s", rc, g_shim.last_error(NULL));
		sl->ctx = NULL;
		return -1;
	}
	return 0;
}

static uint64_t *stage_reserve(slot_t *sl, size_t bytes) {
	if (bytes <= sl->stage_bytes) return sl->stage;
	if (sl->stage) g_shim.host_free(sl->stage);
	sl->stage = NULL;
	sl->stage_bytes = 0;
	void *p = NULL;
	size_t want = bytes + bytes / 8 + 4096;
	if (g_shim.host_alloc(&p, want) != 0 || !p) {
		PyErr_Format(PyExc_MemoryError, "gf2b200: cannot pin %zu bytes of host memory: %s", want,
		             g_shim.last_error(NULL));
		return NULL;
	}
	sl->stage = (uint64_t *)p;
	sl->stage_bytes = want;
	return sl->stage;
}

/* ------------------------------------------------------------------------
 * bit codec
 * ---------------------------------------------------------------------- */

/* Equation int -> one matrix row.  Bit 0 is the constant term (returned), bit k
 * (1..cols) is the coefficient of unknown k-1 and lands at row bit k-1; higher
 * bits are dropped and the sign is ignored (digits are the magnitude), exactly
 * what the reference's bit walk does (:41-59, :411-425). */
#if PyLong_SHIFT == 30
/* 32 digits of 30 bits are exactly 15 words: with the loop fully unrolled every shift is a
 * constant (about three operations per digit where a running 128-bit accumulator needed ~25
 * cycles), and a block of zero digits -- almost all of them on MT19937-class systems -- is
 * recognised with one OR-reduction. */
static inline void digits32_to_words15(const digit *d, uint64_t *V) {
	uint64_t any = 0;
#pragma GCC unroll 32
	for (int i = 0; i < 32; i++) any |= d[i];
	if (!any) {
		memset(V, 0, 15 * 8);
		return;
	}
	uint64_t w[16] = {0};
#pragma GCC unroll 32
	for (int i = 0; i < 32; i++) {
		const int bit = 30 * i, j = bit >> 6, sh = bit & 63;
		const uint64_t x = d[i];
		w[j] |= x << sh;
		if (sh > 34) w[j + 1] |= x >> (64 - sh);
	}
	memcpy(V, w, 15 * 8);
}

static inline int pack_equation(PyObject *eq, uint64_t *row, int64_t nw, int64_t cols) {
	Py_ssize_t nd = LONG_NDIGITS(eq);
	if (nd == 0) {
		memset(row, 0, (size_t)nw * 8);
		return 0;
	}
	const digit *d = LONG_DIGITS(eq);
	const int cbit = (int)(d[0] & 1);
	/* only the digits that reach bits 0..cols of the int */
	const Py_ssize_t need = ((Py_ssize_t)cols + 1 + 29) / 30;
	if (nd > need) nd = need;
	/* V = the int's bits as 64-bit words: words 0..nw-1 go to row[], word nw (bit `cols` can sit there) to top */
	uint64_t top = 0;
	Py_ssize_t i = 0;
	int64_t w = 0;
	for (; i + 32 <= nd && w + 15 <= nw; i += 32, w += 15) digits32_to_words15(d + i, row + w);
	if (w < nw) memset(row + w, 0, (size_t)(nw - w) * 8);
	for (; i < nd; i++) {
		const int64_t bit = 30 * (int64_t)i, j = bit >> 6;
		const int sh = (int)(bit & 63);
		const uint64_t x = d[i];
		if (j < nw) row[j] |= x << sh;
		else if (j == nw) top |= x << sh;
		if (sh > 34) {
			if (j + 1 < nw) row[j + 1] |= x >> (64 - sh);
			else if (j + 1 == nw) top |= x >> (64 - sh);
		}
	}
	/* row = V >> 1 (bit 0 was the constant term) */
	for (int64_t j = 0; j + 1 < nw; j++) row[j] = (row[j] >> 1) | (row[j + 1] << 63);
	row[nw - 1] = (row[nw - 1] >> 1) | (top << 63);
	if (cols & 63) row[nw - 1] &= (1ULL << (cols & 63)) - 1;
	return cbit;
}
#else
static inline int pack_equation(PyObject *eq, uint64_t *row, int64_t nw, int64_t cols) {
	const Py_ssize_t nd = LONG_NDIGITS(eq);
	if (nd == 0) {
		memset(row, 0, (size_t)nw * 8);
		return 0;
	}
	const digit *d = LONG_DIGITS(eq);
	const int cbit = (int)(d[0] & 1);
	unsigned __int128 acc = (unsigned __int128)(d[0] >> 1);
	int nbits = PyLong_SHIFT - 1; /* valid bits waiting in acc */
	Py_ssize_t i = 1;
	int64_t w = 0;
	for (; w < nw; w++) {
		while (nbits < 64 && i < nd) {
			acc |= (unsigned __int128)d[i++] << nbits;
			nbits += PyLong_SHIFT;
		}
		row[w] = (uint64_t)acc;
		acc >>= 64;
		nbits = nbits > 64 ? nbits - 64 : 0;
		if (i >= nd && nbits == 0) {
			w++;
			break;
		}
	}
	if (w < nw) memset(row + w, 0, (size_t)(nw - w) * 8);
	if (cols & 63) row[nw - 1] &= (1ULL << (cols & 63)) - 1;
	return cbit;
}
#endif

/* All equations -> A (rows x nw words) and b (rows bits).  Large systems are packed
 * by several threads WITHOUT the GIL: the caller's list items are pinned by strong
 * references for the duration (ints are immutable, so their digits can be read from
 * any thread).  The workers take blocks of rows (multiples of 64, so no two threads share
 * a word of b) off a shared counter; when `sys` is given, the calling thread hands every
 * finished block to gf2b200_system_load_rows as soon as it is packed, so the H2D copy and
 * the layout kernel of a block overlap the packing of the next ones.
 * The reference packs one bit at a time under the GIL, then solves (_internal.c:403-426). */
typedef struct {
	PyObject **items;
	uint64_t *A, *b;
	int64_t nw, cols;
	Py_ssize_t rows, block_rows, nblocks;
	Py_ssize_t next; /* next block to pack (atomic) */
	unsigned char *done;
	pthread_mutex_t mu;
	pthread_cond_t cv;
	int any_b;
} pack_pool;

static int pack_rows(PyObject **items, uint64_t *A, uint64_t *b, int64_t nw, int64_t cols, Py_ssize_t r0,
                     Py_ssize_t r1) {
	int any = 0;
	for (Py_ssize_t r = r0; r < r1; r++) {
		if ((r & 63) == 0 || r == r0) b[r >> 6] = 0;
		if (pack_equation(items[r], A + (size_t)r * nw, nw, cols)) {
			b[r >> 6] |= 1ULL << (r & 63);
			any = 1;
		}
	}
	return any;
}

static void *pack_worker(void *arg) {
	pack_pool *p = (pack_pool *)arg;
	for (;;) {
		const Py_ssize_t blk = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
		if (blk >= p->nblocks) break;
		const Py_ssize_t r0 = blk * p->block_rows, r1 = r0 + p->block_rows < p->rows ? r0 + p->block_rows : p->rows;
		const int any = pack_rows(p->items, p->A, p->b, p->nw, p->cols, r0, r1);
		pthread_mutex_lock(&p->mu);
		p->done[blk] = 1;
		p->any_b |= any;
		pthread_cond_signal(&p->cv);
		pthread_mutex_unlock(&p->mu);
	}
	return NULL;
}

#define PACK_MAX_THREADS 16
#define PACK_MIN_WORDS_PER_THREAD (1 << 19)
#define PACK_BLOCK_BYTES (2u << 20)

/* does a system of this size take the multi-threaded, streaming path? */
static int pack_threads(Py_ssize_t rows, int64_t nw) {
	int64_t minw = PACK_MIN_WORDS_PER_THREAD;
	const char *env = getenv("GF2B200_PACK_MIN_WORDS"); /* tests: take the streaming path on small systems */
	if (env && atoll(env) > 0) minw = atoll(env);
	int64_t n = ((int64_t)rows * nw) / minw;
	/* never more workers than cores (one core stays with the thread that feeds the GPU) */
	const long cores = sysconf(_SC_NPROCESSORS_ONLN);
	if (cores > 1 && n > cores - 1) n = cores - 1;
	return n > PACK_MAX_THREADS ? PACK_MAX_THREADS : (int)n;
}

/* GIL held on entry and exit; items are already type-checked.  Returns any_b, -1 on a Python
 * error (set), or -2 when a gf2b200_system_load_rows call failed (*load_rc holds its code). */
static int pack_all(PyObject *list, Py_ssize_t rows, uint64_t *A, uint64_t *b, int64_t nw, int64_t cols,
                    gf2b200_system *sys, int *load_rc) {
	PyObject **items = ((PyListObject *)list)->ob_item;
	const int nthreads = pack_threads(rows, nw);
	if (nthreads < 2) {
		const int any = pack_rows(items, A, b, nw, cols, 0, rows);
		if (sys && (*load_rc = g_shim.load_rows(sys, A, 0, rows)) != 0) return -2;
		return any;
	}
	/* the list may be mutated by other Python threads once the GIL is released:
	 * work on a private, reference-holding copy of the item pointers */
	pack_pool p;
	memset(&p, 0, sizeof p);
	const char *benv = getenv("GF2B200_PACK_BLOCK_BYTES");
	const size_t block_bytes = (benv && atoll(benv) > 0) ? (size_t)atoll(benv) : PACK_BLOCK_BYTES;
	p.block_rows = (Py_ssize_t)((block_bytes / ((size_t)nw * 8)) & ~(size_t)63);
	if (p.block_rows < 64) p.block_rows = 64;
	p.nblocks = (rows + p.block_rows - 1) / p.block_rows;
	PyObject **held = (PyObject **)malloc((size_t)rows * sizeof(PyObject *));
	p.done = (unsigned char *)calloc((size_t)p.nblocks, 1);
	if (!held || !p.done) {
		free(held);
		free(p.done);
		PyErr_NoMemory();
		return -1;
	}
	for (Py_ssize_t r = 0; r < rows; r++) held[r] = Py_NewRef(items[r]);
	p.items = held;
	p.A = A;
	p.b = b;
	p.nw = nw;
	p.cols = cols;
	p.rows = rows;
	pthread_mutex_init(&p.mu, NULL);
	pthread_cond_init(&p.cv, NULL);
	pthread_t tids[PACK_MAX_THREADS];
	int started = 0, rc = 0;
	Py_BEGIN_ALLOW_THREADS
	for (int t = 0; t < nthreads; t++)
		if (pthread_create(&tids[started], NULL, pack_worker, &p) == 0) started++;
	if (!started) pack_worker(&p); /* thread creation failed: pack here */
	/* blocks are handed over in index order: the workers take them in that order too */
	for (Py_ssize_t blk = 0; blk < p.nblocks; blk++) {
		pthread_mutex_lock(&p.mu);
		while (!p.done[blk]) pthread_cond_wait(&p.cv, &p.mu);
		pthread_mutex_unlock(&p.mu);
		if (sys && !rc) {
			const Py_ssize_t r0 = blk * p.block_rows, nr = r0 + p.block_rows < rows ? p.block_rows : rows - r0;
			rc = g_shim.load_rows(sys, A + (size_t)r0 * nw, r0, nr);
		}
	}
	for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
	Py_END_ALLOW_THREADS
	pthread_mutex_destroy(&p.mu);
	pthread_cond_destroy(&p.cv);
	for (Py_ssize_t r = 0; r < rows; r++) Py_DECREF(held[r]);
	free(held);
	free(p.done);
	if (rc) {
		*load_rc = rc;
		return -2;
	}
	return p.any_b;
}

/* packed little-endian words -> Python int (bit c of the int = bit c of the row;
 * replaces mzd_vector_to_pylong, reference :32-39) */
static PyObject *words_to_pylong(const uint64_t *w, int64_t nw) {
#if PY_VERSION_HEX >= 0x030D0000
	return PyLong_FromUnsignedNativeBytes(w, (size_t)nw * 8, Py_ASNATIVEBYTES_LITTLE_ENDIAN);
#else
	return _PyLong_FromByteArray((const unsigned char *)w, (size_t)nw * 8, 1, 0);
#endif
}

/* ------------------------------------------------------------------------
 * AffineSpace and its two iterators (reference _internal.h:7-23, _internal.c:61-304)
 * ---------------------------------------------------------------------- */
typedef struct {
	PyObject_HEAD
	int64_t cols, nw, dim;
	uint64_t *origin; /* nw words */
	uint64_t *basis;  /* dim x nw words, M4RI's sigma order */
} SpaceObject;

typedef struct {
	PyObject_HEAD
	SpaceObject *space;
	uint64_t *cur;  /* Gray walk: current vector, NULL once exhausted */
	uint64_t idx;   /* Gray walk: step counter */
	uint8_t *state; /* counter walk: dim digits + end sentinel */
} SpaceIterObject;

static PyTypeObject Space_Type, SpaceIterGray_Type, SpaceIterSlow_Type;

static inline void xor_words(uint64_t *dst, const uint64_t *src, int64_t nw) {
	for (int64_t i = 0; i < nw; i++) dst[i] ^= src[i];
}

static void spaceiter_dealloc(SpaceIterObject *self) {
	Py_XDECREF(self->space);
	free(self->cur);
	free(self->state);
	Py_TYPE(self)->tp_free((PyObject *)self);
}

/* reflected Gray code over the basis: 2^dim vectors, one row XOR per step
 * (order observable through solve_all; reference :101-122) */
static PyObject *spaceiter_next_gray(SpaceIterObject *self) {
	if (!self->cur) return NULL;
	SpaceObject *sp = self->space;
	PyObject *ret = words_to_pylong(self->cur, sp->nw);
	const uint64_t before = self->idx ^ (self->idx >> 1);
	self->idx++;
	const uint64_t after = self->idx ^ (self->idx >> 1);
	const uint64_t flipped = before ^ after;
	const int j = flipped ? __builtin_ctzll(flipped) : 64;
	if (j >= sp->dim || (sp->dim == 64 && self->idx == 0)) {
		free(self->cur);
		self->cur = NULL;
	} else {
		xor_words(self->cur, sp->basis + (int64_t)j * sp->nw, sp->nw);
	}
	return ret;
}

/* dimension > 64: little-endian binary counter, vector rebuilt every step
 * (reference :63-91) */
static PyObject *spaceiter_next_slow(SpaceIterObject *self) {
	SpaceObject *sp = self->space;
	const int64_t d = sp->dim;
	if (self->state[d]) return NULL;
	uint64_t *v = (uint64_t *)malloc((size_t)sp->nw * 8);
	if (!v) return PyErr_NoMemory();
	memcpy(v, sp->origin, (size_t)sp->nw * 8);
	for (int64_t r = 0; r < d; r++)
		if (self->state[r]) xor_words(v, sp->basis + r * sp->nw, sp->nw);
	int64_t r = 0;
	while (r < d && self->state[r]) self->state[r++] = 0; /* carry */
	if (r < d) self->state[r] = 1;
	else self->state[d] = 1;
	PyObject *ret = words_to_pylong(v, sp->nw);
	free(v);
	return ret;
}

static PyObject *space_iter(PyObject *obj) {
	SpaceObject *sp = (SpaceObject *)obj;
	const int gray = sp->dim <= 64;
	SpaceIterObject *it = PyObject_New(SpaceIterObject, gray ? &SpaceIterGray_Type : &SpaceIterSlow_Type);
	if (!it) return NULL;
	it->space = NULL;
	it->cur = NULL;
	it->state = NULL;
	it->idx = 0;
	if (gray) {
		it->cur = (uint64_t *)malloc((size_t)sp->nw * 8);
		if (it->cur) memcpy(it->cur, sp->origin, (size_t)sp->nw * 8);
	} else {
		it->state = (uint8_t *)calloc((size_t)sp->dim + 1, 1);
	}
	if (!it->cur && !it->state) {
		Py_DECREF(it);
		return PyErr_NoMemory();
	}
	Py_INCREF(sp);
	it->space = sp;
	return (PyObject *)it;
}

static PyObject *space_get_dimension(SpaceObject *self, void *c) { return PyLong_FromLongLong(self->dim); }
static PyObject *space_get_origin(SpaceObject *self, void *c) { return words_to_pylong(self->origin, self->nw); }
static PyObject *space_get_basis(SpaceObject *self, void *c) {
	PyObject *t = PyTuple_New((Py_ssize_t)self->dim);
	if (!t) return NULL;
	for (int64_t r = 0; r < self->dim; r++) {
		PyObject *v = words_to_pylong(self->basis + r * self->nw, self->nw);
		if (!v) {
			Py_DECREF(t);
			return NULL;
		}
		PyTuple_SET_ITEM(t, (Py_ssize_t)r, v);
	}
	return t;
}

/* get(i): origin ^ XOR of basis[j] over the set bits j < dimension of |i|
 * (plain binary, not Gray; reference :242-273) */
static PyObject *space_get(SpaceObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 1) {
		PyErr_SetString(PyExc_TypeError, "get requires 1 argument");
		return NULL;
	}
	if (!PyLong_Check(args[0])) {
		PyErr_SetString(PyExc_TypeError, "Index must be an integer");
		return NULL;
	}
	uint64_t *v = (uint64_t *)malloc((size_t)self->nw * 8);
	if (!v) return PyErr_NoMemory();
	memcpy(v, self->origin, (size_t)self->nw * 8);
	const Py_ssize_t nd = LONG_NDIGITS(args[0]);
	const digit *d = LONG_DIGITS(args[0]);
	for (Py_ssize_t i = 0; i < nd; i++) {
		digit x = d[i];
		while (x) {
			const int64_t j = (int64_t)i * PyLong_SHIFT + __builtin_ctz(x);
			x &= x - 1;
			if (j < self->dim) xor_words(v, self->basis + j * self->nw, self->nw);
		}
	}
	PyObject *ret = words_to_pylong(v, self->nw);
	free(v);
	return ret;
}

static void space_dealloc(SpaceObject *self) {
	free(self->origin);
	free(self->basis);
	Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyGetSetDef space_getset[] = {
    {"dimension", (getter)space_get_dimension, NULL, "Dimension of the affine space", NULL},
    {"origin", (getter)space_get_origin, NULL, "Origin of the affine space", NULL},
    {"basis", (getter)space_get_basis, NULL, "Basis of the affine space (tuple of ints)", NULL},
    {NULL}};

static PyMethodDef space_methods[] = {
    {"get", _PyCFunction_CAST(space_get), METH_FASTCALL,
     "get(n)\n--\n\nn-th element of the affine space (binary digits of n select basis vectors); "
     "check 0 <= n < 2**dimension first."},
    {NULL}};

static PyTypeObject Space_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_internal.AffineSpace",
    .tp_basicsize = sizeof(SpaceObject),
    .tp_dealloc = (destructor)space_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_iter = space_iter,
    .tp_methods = space_methods,
    .tp_getset = space_getset,
};

static PyTypeObject SpaceIterGray_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_internal.AffineSpaceIterator",
    .tp_basicsize = sizeof(SpaceIterObject),
    .tp_dealloc = (destructor)spaceiter_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_iter = PyObject_SelfIter,
    .tp_iternext = (iternextfunc)spaceiter_next_gray,
};

static PyTypeObject SpaceIterSlow_Type = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_internal.AffineSpaceIteratorSlow",
    .tp_basicsize = sizeof(SpaceIterObject),
    .tp_dealloc = (destructor)spaceiter_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_iter = PyObject_SelfIter,
    .tp_iternext = (iternextfunc)spaceiter_next_slow,
};

/* takes ownership of origin / basis (malloc'd) */
static PyObject *space_new(int64_t cols, int64_t dim, uint64_t *origin, uint64_t *basis) {
	SpaceObject *sp = PyObject_New(SpaceObject, &Space_Type);
	if (!sp) {
		free(origin);
		free(basis);
		return NULL;
	}
	sp->cols = cols;
	sp->nw = (cols + 63) / 64;
	sp->dim = dim;
	sp->origin = origin;
	sp->basis = basis;
	return (PyObject *)sp;
}

/* _make_affine_space(origin:int, basis:sequence of ints, cols) -- builds an
 * AffineSpace from Python ints.  Host-only helper used by the CPU-side tests of
 * the iterators / get(); not part of the reference surface. */
static PyObject *make_affine_space(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 3 || !PyLong_Check(args[0]) || !PyLong_Check(args[2])) {
		PyErr_SetString(PyExc_TypeError, "_make_affine_space(origin, basis, cols)");
		return NULL;
	}
	const int64_t cols = PyLong_AsLongLong(args[2]);
	if (cols <= 0) {
		if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "Number of columns must be positive");
		return NULL;
	}
	PyObject *seq = PySequence_Fast(args[1], "basis must be a sequence of ints");
	if (!seq) return NULL;
	const int64_t dim = PySequence_Fast_GET_SIZE(seq), nw = (cols + 63) / 64;
	uint64_t *origin = (uint64_t *)calloc((size_t)nw + 1, 8);
	uint64_t *basis = (uint64_t *)calloc((size_t)(dim ? dim : 1) * nw + 1, 8);
	if (!origin || !basis) {
		free(origin);
		free(basis);
		Py_DECREF(seq);
		return PyErr_NoMemory();
	}
	/* reuse the equation packer: shift left by one so bit c lands at row bit c */
	PyObject *one = PyLong_FromLong(1);
	for (int64_t r = -1; one && r < dim; r++) {
		PyObject *v = r < 0 ? args[0] : PySequence_Fast_GET_ITEM(seq, r);
		PyObject *sh = PyLong_Check(v) ? PyNumber_Lshift(v, one) : NULL;
		if (!sh) {
			if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "basis items must be integers");
			free(origin);
			free(basis);
			Py_DECREF(seq);
			Py_DECREF(one);
			return NULL;
		}
		pack_equation(sh, r < 0 ? origin : basis + r * nw, nw, cols);
		Py_DECREF(sh);
	}
	Py_XDECREF(one);
	Py_DECREF(seq);
	return space_new(cols, dim, origin, basis);
}

/* ------------------------------------------------------------------------
 * m4ri_solve(equations, cols, mode)   (reference :359-502)
 * ---------------------------------------------------------------------- */
static PyObject *m4ri_solve(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 3) {
		PyErr_SetString(PyExc_TypeError, "m4ri_solve requires 3 arguments");
		return NULL;
	}
	PyObject *eqs = args[0];
	if (!PyList_Check(eqs)) {
		PyErr_SetString(PyExc_TypeError, "The first argument equations must be a list");
		return NULL;
	}
	const Py_ssize_t cols = PyLong_AsSsize_t(args[1]);
	if (cols <= 0) {
		if (cols == -1 && PyErr_Occurred()) return NULL;
		PyErr_SetString(PyExc_ValueError, "Number of columns must be positive");
		return NULL;
	}
	const long mode = PyLong_AsLong(args[2]);
	if (mode == -1 && PyErr_Occurred()) return NULL;
	if (mode != 0 && mode != 1) {
		PyErr_SetString(PyExc_ValueError, "Invalid mode");
		return NULL;
	}
	const Py_ssize_t rows = PyList_GET_SIZE(eqs);
	if (rows < cols) {
		PyErr_SetString(PyExc_ValueError,
		                "Number of rows must be greater than or equal to number of columns, try pad with zeros.");
		return NULL;
	}
	for (Py_ssize_t r = 0; r < rows; r++) {
		if (!PyLong_Check(PyList_GET_ITEM(eqs, r))) {
			PyErr_SetString(PyExc_TypeError, "List items must be integers");
			return NULL;
		}
	}

	const int64_t nw = ((int64_t)cols + 63) / 64;
	const int64_t bw = ((int64_t)rows + 63) / 64;

	/* other threads may be inside solves with the GIL released: take a free solver slot
	 * (waiting, if all MAX_SLOTS are busy, without holding the GIL) */
	pthread_mutex_lock(&g_lock);
	const int lrc = shim_load(); /* GIL held: errors become Python exceptions */
	pthread_mutex_unlock(&g_lock);
	if (lrc) return NULL;
	slot_t *sl;
	Py_BEGIN_ALLOW_THREADS
	sl = slot_acquire();
	Py_END_ALLOW_THREADS

	PyObject *ret = NULL;
	if (ctx_ready(sl)) goto out;
	uint64_t *A = stage_reserve(sl, ((size_t)rows * nw + bw) * 8);
	if (!A) goto out;
	uint64_t *b = A + (size_t)rows * nw; /* the packer writes every word of A and b itself */
	/* large systems: the rows stream to the GPU block by block while the packer threads are
	 * still producing the rest (gf2b200_system_load_begin / _rows / _end) */
	gf2b200_system *sys = NULL;
	int rc = 0;
	if (pack_threads(rows, nw) >= 2) {
		Py_BEGIN_ALLOW_THREADS
		rc = g_shim.solve_open(sl->ctx, rows, cols, &sys);
		if (!rc && (rc = g_shim.load_begin(sys, nw)) != 0) {
			g_shim.solve_close(sl->ctx, sys, -1, NULL);
			sys = NULL;
		}
		Py_END_ALLOW_THREADS
		if (rc) {
			PyErr_Format(PyExc_RuntimeError, "gf2b200_solve failed (%d): %s", rc, g_shim.last_error(sl->ctx));
			goto out;
		}
	}
	const int any_b = pack_all(eqs, rows, A, b, nw, cols, sys, &rc);
	if (any_b < 0) {
		if (sys) g_shim.solve_close(sl->ctx, sys, -1, NULL);
		if (any_b == -2)
			PyErr_Format(PyExc_RuntimeError, "gf2b200_solve failed (%d): %s", rc, g_shim.last_error(sl->ctx));
		goto out;
	}

	gf2b200_result res;
	Py_BEGIN_ALLOW_THREADS /* reference releases the GIL around M4RI too (:429) */
	if (sys) {
		rc = g_shim.load_end(sys, any_b ? b : NULL);
		if (rc) g_shim.solve_close(sl->ctx, sys, -1, NULL);
		else rc = g_shim.solve_close(sl->ctx, sys, (int)mode, &res);
	} else {
		rc = g_shim.solve(sl->ctx, A, any_b ? b : NULL, rows, cols, nw, (int)mode, &res);
	}
	Py_END_ALLOW_THREADS
	if (rc) {
		PyErr_Format(PyExc_RuntimeError, "gf2b200_solve failed (%d): %s", rc, g_shim.last_error(sl->ctx));
		goto out;
	}
	if (res.status == GF2B200_INCONSISTENT) {
		g_shim.result_free(&res);
		ret = Py_NewRef(Py_None);
		goto out;
	}
	if (mode == 0) {
		ret = words_to_pylong(res.origin, nw);
		g_shim.result_free(&res);
		goto out;
	}
	{
		const int64_t dim = res.kernel_dim;
		uint64_t *origin = (uint64_t *)malloc((size_t)nw * 8);
		uint64_t *basis = (uint64_t *)malloc((size_t)(dim ? dim : 1) * nw * 8);
		if (!origin || !basis) {
			free(origin);
			free(basis);
			g_shim.result_free(&res);
			PyErr_NoMemory();
			goto out;
		}
		memcpy(origin, res.origin, (size_t)nw * 8);
		if (dim) memcpy(basis, res.basis, (size_t)dim * nw * 8);
		g_shim.result_free(&res);
		ret = space_new(cols, dim, origin, basis);
	}
out:
	slot_release(sl);
	return ret;
}

/* ------------------------------------------------------------------------
 * tuple helpers used by the symbolic layer (reference :504-676); host-only
 * ---------------------------------------------------------------------- */

/* to_bits(n, a) -> tuple of n bools, LSB first, of |a| */
static PyObject *to_bits(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 2) {
		PyErr_SetString(PyExc_TypeError, "to_bits requires 2 arguments");
		return NULL;
	}
	const Py_ssize_t n = PyLong_AsSsize_t(args[0]);
	if (n < 0) {
		if (n == -1 && PyErr_Occurred()) return NULL;
		PyErr_SetString(PyExc_ValueError, "n must be non-negative");
		return NULL;
	}
	if (!PyLong_Check(args[1])) {
		PyErr_SetString(PyExc_TypeError, "a must be an integer");
		return NULL;
	}
	PyObject *t = PyTuple_New(n);
	if (!t) return NULL;
	const Py_ssize_t nd = LONG_NDIGITS(args[1]);
	const digit *d = LONG_DIGITS(args[1]);
	for (Py_ssize_t k = 0; k < n; k++) {
		const Py_ssize_t i = k / PyLong_SHIFT;
		const int set = i < nd && ((d[i] >> (k % PyLong_SHIFT)) & 1);
		PyTuple_SET_ITEM(t, k, Py_NewRef(set ? Py_True : Py_False));
	}
	return t;
}

static void magnitude_bits(uint8_t *out, Py_ssize_t n, PyObject *v) {
	const Py_ssize_t nd = LONG_NDIGITS(v);
	const digit *d = LONG_DIGITS(v);
	for (Py_ssize_t k = 0; k < n; k++) {
		const Py_ssize_t i = k / PyLong_SHIFT;
		out[k] = (uint8_t)(i < nd && ((d[i] >> (k % PyLong_SHIFT)) & 1));
	}
}

/* mul_bit_quad(n, a, b, v, basis): OR into v the monomial x_i x_j (j < i) basis
 * element whenever a_i b_j ^ a_j b_i = 1; monomials are numbered row by row
 * starting at 1 + n (reference :538-604) */
static PyObject *mul_bit_quad(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 5) {
		PyErr_SetString(PyExc_TypeError, "mul_bit_quad requires 5 arguments");
		return NULL;
	}
	const Py_ssize_t n = PyLong_AsSsize_t(args[0]);
	if (n <= 0) {
		if (n == -1 && PyErr_Occurred()) return NULL;
		PyErr_SetString(PyExc_ValueError, "n must be positive");
		return NULL;
	}
	if (!PyLong_Check(args[1]) || !PyLong_Check(args[2]) || !PyLong_Check(args[3])) {
		PyErr_SetString(PyExc_TypeError, "a and b and v must be integers");
		return NULL;
	}
	PyObject *basis = args[4];
	if (!PyList_Check(basis)) {
		PyErr_SetString(PyExc_TypeError, "basis must be a list");
		return NULL;
	}
	if (PyList_GET_SIZE(basis) != 1 + n + n * (n - 1) / 2) {
		PyErr_SetString(PyExc_ValueError, "The length of basis is not correct");
		return NULL;
	}
	uint8_t *ab = (uint8_t *)malloc((size_t)n * 2);
	if (!ab) return PyErr_NoMemory();
	uint8_t *bb = ab + n;
	magnitude_bits(ab, n, args[1]);
	magnitude_bits(bb, n, args[2]);
	PyObject *acc = Py_NewRef(args[3]);
	Py_ssize_t mono = 1 + n;
	for (Py_ssize_t i = 0; i < n && acc; i++) {
		if (!ab[i] && !bb[i]) { /* whole row of monomials contributes nothing */
			mono += i;
			continue;
		}
		for (Py_ssize_t j = 0; j < i; j++, mono++) {
			if ((ab[i] & bb[j]) ^ (ab[j] & bb[i])) {
				PyObject *nv = PyNumber_Or(acc, PyList_GET_ITEM(basis, mono));
				Py_DECREF(acc);
				acc = nv;
				if (!acc) {
					PyErr_SetString(PyExc_TypeError, "Failed to compute or, list items must be integers");
					break;
				}
			}
		}
	}
	free(ab);
	return acc;
}

static PyObject *xor_tuple(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 2) {
		PyErr_SetString(PyExc_TypeError, "xor_tuple requires 2 arguments");
		return NULL;
	}
	PyObject *a = args[0], *b = args[1];
	if (!PyTuple_Check(a) || !PyTuple_Check(b)) {
		PyErr_SetString(PyExc_TypeError, "a and b must be tuples");
		return NULL;
	}
	const Py_ssize_t n = PyTuple_GET_SIZE(a);
	if (PyTuple_GET_SIZE(b) != n) {
		PyErr_SetString(PyExc_ValueError, "The length of a and b is not equal");
		return NULL;
	}
	PyObject *out = PyTuple_New(n);
	if (!out) return NULL;
	for (Py_ssize_t i = 0; i < n; i++) {
		PyObject *x = PyNumber_Xor(PyTuple_GET_ITEM(a, i), PyTuple_GET_ITEM(b, i));
		if (!x) {
			Py_DECREF(out);
			PyErr_SetString(PyExc_TypeError, "Failed to compute xor, list items must be integers");
			return NULL;
		}
		PyTuple_SET_ITEM(out, i, x);
	}
	return out;
}

/* tuple_where(cond, a, b): like np.where, but -- as in the reference (:641-676) --
 * the selection is written INTO `cond`, which is also the return value. */
static PyObject *tuple_where(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 3) {
		PyErr_SetString(PyExc_TypeError, "tuple_where requires 3 arguments");
		return NULL;
	}
	PyObject *cond = args[0], *a = args[1], *b = args[2];
	if (!PyTuple_Check(cond)) {
		PyErr_SetString(PyExc_TypeError, "cond must be a list");
		return NULL;
	}
	const Py_ssize_t n = PyTuple_GET_SIZE(cond);
	const int a_seq = PyTuple_Check(a), b_seq = PyTuple_Check(b);
	if (a_seq && PyTuple_GET_SIZE(a) != n) {
		PyErr_SetString(PyExc_ValueError, "The length of a and cond is not equal");
		return NULL;
	}
	if (b_seq && PyTuple_GET_SIZE(b) != n) {
		PyErr_SetString(PyExc_ValueError, "The length of b and cond is not equal");
		return NULL;
	}
	for (Py_ssize_t i = 0; i < n; i++) {
		PyObject *c = PyTuple_GET_ITEM(cond, i);
		const int truth = PyObject_IsTrue(c);
		if (truth < 0) return NULL;
		PyObject *pick = truth ? (a_seq ? PyTuple_GET_ITEM(a, i) : a) : (b_seq ? PyTuple_GET_ITEM(b, i) : b);
		Py_INCREF(pick);
		PyTuple_SET_ITEM(cond, i, pick);
		Py_DECREF(c);
	}
	return Py_NewRef(cond);
}

static PyObject *eqs_to_sage_mat_helper(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	PyErr_SetString(PyExc_RuntimeError,
	                "eqs_to_sage_mat_helper: the Sage/libgd bridge is outside the gf2b200 solve path "
	                "(use LinearSystem.get_sage_mat_slow)");
	return NULL;
}

/* _pack_probe(equations, cols) -> (bytes of the packed A rows, bytes of packed b):
 * exposes the bit codec to the CPU-side tests; no device involved. */
static PyObject *pack_probe(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
	if (nargs != 2 || !PyList_Check(args[0])) {
		PyErr_SetString(PyExc_TypeError, "_pack_probe(equations:list, cols)");
		return NULL;
	}
	const Py_ssize_t cols = PyLong_AsSsize_t(args[1]);
	if (cols <= 0) {
		if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "Number of columns must be positive");
		return NULL;
	}
	const Py_ssize_t rows = PyList_GET_SIZE(args[0]);
	const int64_t nw = ((int64_t)cols + 63) / 64, bw = ((int64_t)rows + 63) / 64;
	PyObject *pa = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(rows * nw * 8));
	PyObject *pb = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(bw * 8));
	if (!pa || !pb) {
		Py_XDECREF(pa);
		Py_XDECREF(pb);
		return NULL;
	}
	uint64_t *A = (uint64_t *)PyBytes_AS_STRING(pa), *b = (uint64_t *)PyBytes_AS_STRING(pb);
	memset(b, 0, (size_t)bw * 8);
	for (Py_ssize_t r = 0; r < rows; r++) {
		if (!PyLong_Check(PyList_GET_ITEM(args[0], r))) {
			Py_DECREF(pa);
			Py_DECREF(pb);
			PyErr_SetString(PyExc_TypeError, "List items must be integers");
			return NULL;
		}
	}
	if (pack_all(args[0], rows, A, b, nw, cols, NULL, NULL) < 0) {
		Py_DECREF(pa);
		Py_DECREF(pb);
		return NULL;
	}
	PyObject *ret = PyTuple_Pack(2, pa, pb);
	Py_DECREF(pa);
	Py_DECREF(pb);
	return ret;
}

static PyMethodDef module_methods[] = {
    {"m4ri_solve", _PyCFunction_CAST(m4ri_solve), METH_FASTCALL,
     "m4ri_solve(equations, cols, mode)\n--\n\nSolve a linear system over GF(2) on the B200 "
     "(same contract as gf2bv._internal.m4ri_solve)"},
    {"to_bits", _PyCFunction_CAST(to_bits), METH_FASTCALL,
     "to_bits(n, number)\n--\n\nConvert an integer to a tuple of bits (bool values)"},
    {"mul_bit_quad", _PyCFunction_CAST(mul_bit_quad), METH_FASTCALL,
     "mul_bit_quad(n, a, b, v, basis)\n--\n\nMultiply two linear symbolic bits into a linearized quadratic bit"},
    {"xor_tuple", _PyCFunction_CAST(xor_tuple), METH_FASTCALL, "xor_tuple(a, b)\n--\n\nXOR two tuples of integers"},
    {"tuple_where", _PyCFunction_CAST(tuple_where), METH_FASTCALL,
     "tuple_where(cond, a, b)\n--\n\nSelect from a or b by cond like np.where, writing into cond"},
    {"eqs_to_sage_mat_helper", _PyCFunction_CAST(eqs_to_sage_mat_helper), METH_FASTCALL,
     "eqs_to_sage_mat_helper(equations, cols)\n--\n\nNot available (Sage bridge is out of scope)"},
    {"_make_affine_space", _PyCFunction_CAST(make_affine_space), METH_FASTCALL,
     "_make_affine_space(origin, basis, cols)\n--\n\nBuild an AffineSpace from ints (host-only test helper)"},
    {"_pack_probe", _PyCFunction_CAST(pack_probe), METH_FASTCALL,
     "_pack_probe(equations, cols)\n--\n\nPacked (A, b) bytes of the equation codec (host-only test helper)"},
    {NULL}};

static struct PyModuleDef module_def = {PyModuleDef_HEAD_INIT, "_internal", NULL, -1, module_methods};

PyMODINIT_FUNC PyInit__internal(void) {
	if (PyType_Ready(&Space_Type) < 0 || PyType_Ready(&SpaceIterGray_Type) < 0 ||
	    PyType_Ready(&SpaceIterSlow_Type) < 0)
		return NULL;
	PyObject *mod = PyModule_Create(&module_def);
	if (!mod) return NULL;
	if (PyModule_AddType(mod, &Space_Type) < 0 || PyModule_AddType(mod, &SpaceIterGray_Type) < 0 ||
	    PyModule_AddType(mod, &SpaceIterSlow_Type) < 0) {
		Py_DECREF(mod);
		return NULL;
	}
	return mod;
}
