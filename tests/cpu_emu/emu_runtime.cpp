/*
 * emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): the
 * fiber scheduler.  An ordinary launch runs the grid's CTAs one after another on the
 * calling thread; a cooperative launch keeps every CTA alive and interleaves the
 * fibers of all of them, so CTAs can wait for each other (grid barriers).
 */
#include <cuda_runtime.h>
#include <sys/mman.h>
#ifdef __SANITIZE_ADDRESS__
#include <sanitizer/asan_interface.h>
#endif

#ifdef EMU_RACECHECK
extern "C" void emu_racecheck_launch(void);
#endif

/* marker: the package refuses a library that exports this unless GF2B200_TEST_EMULATION=1 */
extern "C" int gf2b200_emulated_build(void) { return 1; }

namespace emu {

Fiber *cur = nullptr;
Cta *cta_p = nullptr;
dim3 g_blockDim, g_gridDim;
void *sched_sp = nullptr;
static std::vector<Cta *> ctas; /* CTA objects (and their dynamic shared memory), reused across launches */

static const size_t STACK = 64 * 1024;
static char *stacks = nullptr;
static size_t n_stacks = 0;
/* for the hazard checker (emu_racecheck.cpp) */
char *stacks_base = nullptr;
size_t stacks_bytes = 0;
unsigned launch_seq = 0;
static void (*g_thunk)(void *) = nullptr;
static void *g_arg = nullptr;

/* x86-64 System V: callee-saved registers on the old stack, swap rsp, restore */
__asm__(
    ".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size emu_switch,.-emu_switch\n");

/* EMU_TSAN (build_emu.py --tsan): this file itself is compiled WITHOUT -fsanitize=thread --
 * the scheduler's bookkeeping is not GPU memory -- and only drives TSan's fiber API */
#ifdef EMU_TSAN
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
void __tsan_acquire(void *addr);
void __tsan_release(void *addr);
}
static void *sched_tsan = nullptr;
static std::vector<void *> tsan_fibers;
static int launch_obj; /* kernel boundary: host -> every GPU thread -> host */
#define TSAN_TO(f) __tsan_switch_to_fiber((f), 1 /* no implied synchronisation */)
#undef EMU_ACQUIRE
#undef EMU_RELEASE
#define EMU_ACQUIRE(p) __tsan_acquire((void *)(p))
#define EMU_RELEASE(p) __tsan_release((void *)(p))
#else
#define TSAN_TO(f) ((void)0)
#endif

void yield_to_scheduler() {
	TSAN_TO(sched_tsan);
	emu_switch(&cur->sp, sched_sp);
}

void warp_complete_if_ready(Warp &W) {
	if (W.departing || !W.arrived || W.arrived != W.live) return;
	W.gen++;
	for (int i = 0; i < 32; i++) W.snap[i] = ((W.arrived >> i) & 1) ? W.vals[i] : 0;
	W.snap_mask = W.arrived;
	W.departing = W.arrived;
	W.arrived = 0;
}

bool is_dynamic_smem(const void *p) {
	for (const Cta *c : ctas)
		if (c->smem && (const unsigned char *)p >= c->smem && (const unsigned char *)p < c->smem + DYN_SMEM_CAP) return true;
	return false;
}

static void fiber_main() {
#ifdef EMU_TSAN
	EMU_ACQUIRE(&launch_obj);
#endif
	g_thunk(g_arg);
#ifdef EMU_TSAN
	EMU_RELEASE(&launch_obj);
#endif
	Fiber *f = cur;
	Cta &C = *f->cta;
	f->done = 1;
	C.alive--;
	Warp &W = C.w[f->warp];
	W.live &= ~(1u << f->lane);
	warp_complete_if_ready(W);
	if (C.alive > 0 && C.bar_arrived == C.alive) { /* the rest were waiting for this one */
		C.bar_arrived = 0;
		C.bar_gen++;
	}
	TSAN_TO(sched_tsan);
	emu_switch(&f->sp, sched_sp);
	abort(); /* a finished fiber is never resumed */
}

static bool runnable(const Fiber &f) {
	const Cta &C = *f.cta;
	switch (f.wait) {
	case W_NONE: return true;
	case W_BARRIER: return f.bar_gen != C.bar_gen;
	case W_WARP_ENTER: return C.w[f.warp].departing == 0;
	case W_WARP_RESULT: return (C.w[f.warp].departing >> f.lane) & 1;
	case W_WORD: return (__atomic_load_n((const unsigned long long *)f.word, __ATOMIC_ACQUIRE) & 1) != f.word_val;
	}
	return true;
}

/* (re)initialise CTA object `ci` as block `bx` of the grid; its fibers use stacks [stack0, stack0 + nt) */
static void cta_start(size_t ci, unsigned bx, dim3 block, size_t nt, size_t stack0, size_t smem) {
	while (ctas.size() <= ci) ctas.push_back(new Cta());
	Cta &C = *ctas[ci];
	if (!C.smem && posix_memalign((void **)&C.smem, 1024, DYN_SMEM_CAP)) abort();
#ifdef __SANITIZE_ADDRESS__
	/* a launch may touch only the dynamic shared memory it asked for */
	ASAN_UNPOISON_MEMORY_REGION(C.smem, DYN_SMEM_CAP);
	const size_t keep = (smem + 7) & ~(size_t)7;
	if (keep < DYN_SMEM_CAP) ASAN_POISON_MEMORY_REGION(C.smem + keep, DYN_SMEM_CAP - keep);
#else
	(void)smem;
#endif
	C.f.resize(nt);
	C.w.resize((nt + 31) / 32);
	C.bid = dim3(bx);
	C.alive = (int)nt;
	C.bar_arrived = 0;
	C.bar_gen = 0;
	C.tma_count = 0;
	for (Warp &W : C.w) memset(&W, 0, sizeof W);
	for (size_t t = 0; t < nt; t++) {
		Fiber &f = C.f[t];
		f.done = 0;
		f.wait = W_NONE;
		f.tma_seen = 0;
		f.cta = &C;
		f.tid = dim3((unsigned)(t % block.x), (unsigned)(t / block.x));
		f.lane = (int)(t & 31);
		f.warp = (int)(t >> 5);
		C.w[f.warp].live |= 1u << f.lane;
		/* initial frame: six callee-saved slots, then the entry point as return address;
		 * after the `ret` rsp is 8 mod 16 as at any function entry */
		uintptr_t top = ((uintptr_t)(stacks + (stack0 + t + 1) * STACK)) & ~(uintptr_t)15;
		void **sp = (void **)(top - 8);
		*--sp = (void *)fiber_main;
		for (int i = 0; i < 6; i++) *--sp = nullptr;
		f.sp = sp;
#ifdef EMU_TSAN
		if (stack0 + t >= tsan_fibers.size()) tsan_fibers.resize(stack0 + t + 1, nullptr);
		if (!tsan_fibers[stack0 + t]) tsan_fibers[stack0 + t] = __tsan_create_fiber(0);
		f.tsan = tsan_fibers[stack0 + t];
#endif
	}
}

/* run the fibers of CTA objects [0, n) until all are done */
static void run_ctas(size_t n, size_t nt) {
	size_t remaining = n * nt;
	/* GF2_EMU_STARVE=<seed>: in a cooperative launch one CTA at a time is held back for a burst of 8 - 71
	 * scheduler rounds -- CTAs drift apart by many rounds, as they do on a GPU under a tool or a debugger.
	 * The plain round-robin never lets a CTA be "late", which hid an ordering bug of k_forward's slow path
	 * (profiles/r02_sanitizer.md). */
	static unsigned long long starve = 0;
	static bool starve_init = false;
	if (!starve_init) {
		const char *e = getenv("GF2_EMU_STARVE");
		starve = e ? strtoull(e, nullptr, 10) * 2 + 1 : 0;
		starve_init = true;
	}
	size_t victim = 0, burst = 0;
	while (remaining > 0) {
		bool progress = false;
		bool skipped = false;
		if (starve && n > 1) {
			if (burst) {
				burst--;
			} else {
				starve = starve * 6364136223846793005ULL + 1442695040888963407ULL;
				if (((starve >> 40) & 3) == 0) {
					victim = (size_t)((starve >> 44) % n);
					burst = 8 + (size_t)((starve >> 52) & 63);
				}
			}
		}
		for (size_t ci = 0; ci < n; ci++) {
			Cta &C = *ctas[ci];
			if (!C.alive) continue;
			if (burst && ci == victim) {
				skipped = true;
				continue;
			}
			for (size_t t = 0; t < nt; t++) {
				Fiber &f = C.f[t];
				if (f.done || !runnable(f)) continue;
				/* a thread that merely yielded (spin wait) counts as progress only if
				 * something else moves too; a grid where everybody spins is a deadlock
				 * that the caller's own time-out has to break */
				f.wait = W_NONE;
				cur = &f;
				cta_p = &C;
				TSAN_TO(f.tsan);
				emu_switch(&sched_sp, f.sp);
				progress = true;
				if (f.done) remaining--;
			}
		}
		if (!progress && !skipped) {
			fprintf(stderr, "emu: deadlock (%zu threads parked)\n", remaining);
			abort();
		}
	}
}

void run_grid(dim3 grid, dim3 block, size_t smem, bool concurrent, void (*thunk)(void *), void *arg) {
	const size_t nt = (size_t)block.x * block.y * block.z;
	const size_t need = concurrent ? nt * grid.x : nt;
	if (need > n_stacks) {
		if (stacks) munmap(stacks, n_stacks * STACK);
		stacks = (char *)mmap(nullptr, need * STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (stacks == MAP_FAILED) abort();
		n_stacks = need;
		stacks_base = stacks;
		stacks_bytes = need * STACK;
	}
	launch_seq++;
#ifdef EMU_RACECHECK
	emu_racecheck_launch();
#endif
	g_thunk = thunk;
	g_arg = arg;
	g_gridDim = grid;
	g_blockDim = block;
#ifdef EMU_TSAN
	sched_tsan = __tsan_get_current_fiber();
	EMU_RELEASE(&launch_obj);
#endif
	if (concurrent) {
		for (unsigned bx = 0; bx < grid.x; bx++) cta_start(bx, bx, block, nt, (size_t)bx * nt, smem);
		run_ctas(grid.x, nt);
	} else {
		for (unsigned bx = 0; bx < grid.x; bx++) {
			cta_start(0, bx, block, nt, 0, smem);
			run_ctas(1, nt);
		}
	}
#ifdef EMU_TSAN
	EMU_ACQUIRE(&launch_obj);
#endif
	cur = nullptr;
	cta_p = nullptr;
}

} /* namespace emu */
