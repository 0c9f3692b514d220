/*
 * emu_racecheck.cpp -- TEST INFRASTRUCTURE ONLY (tests/cpu_emu, build_emu.py --racecheck).
 *
 * A hazard checker for the emulated kernels in the spirit of `compute-sanitizer
 * --tool racecheck`, for a container without a GPU.  The kernels' translation unit is
 * compiled with -fsanitize=thread ONLY to get the compiler's memory-access hooks; this
 * file implements those hooks itself (libtsan is not linked: its 256 thread slots
 * cannot tell apart the 1024 threads of a sweep CTA).
 *
 * Model.  Every instrumented access by a GPU thread (a fiber of the emulator) to
 * shared memory (the library's own static storage) or device memory (heap blocks) is
 * recorded per 8-byte cell with a byte mask.  Two accesses to overlapping bytes, at
 * least one a write, by different threads are a hazard unless the GPU's own
 * synchronisation orders them:
 *   - a kernel boundary (different launch),
 *   - a __syncthreads between them (CTA barrier generation),
 *   - a warp collective between them when both threads sit in the same warp,
 *   - for data written by the bulk copy (TMA stand-in): the reader has waited on the
 *     mbarrier phase that the copy completed,
 *   - both are atomics.
 * Different CTAs of one launch are ordered only through the ticket pattern
 * "__threadfence; __syncthreads; one thread does an atomic RMW on X" in CTA c1 and a LATER
 * atomic RMW on the same X in CTA c2: what c1 did before that barrier is then ordered before
 * what c2 does after the atomic (same thread) or after c2's next barrier (other threads).
 * (The last-CTA scan of SWEEP_TAIL_SELECT and grid barriers are built from exactly this.)
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <link.h>

#include <unordered_map>
#include <vector>

namespace emu {
extern char *stacks_base;
extern size_t stacks_bytes;
extern unsigned launch_seq;
}

namespace {

struct Acc {
	unsigned launch;
	int cta, tid;
	unsigned bar_gen, warp_gen;
	unsigned tma_seq; /* != 0: written by the bulk copy number tma_seq of this CTA */
	unsigned char mask;
	unsigned char atomic;
	const void *pc;
};
struct Cell {
	Acc w;
	Acc r[2];
};

std::unordered_map<uintptr_t, Cell> shadow;

/* inter-CTA ordering through atomics: per atomic address, what each CTA has released there;
 * per CTA, what it has acquired from every other CTA */
struct Rel { int cta; unsigned bar_gen; };                  /* accesses of `cta` with generation < bar_gen */
struct Acq { unsigned upto; unsigned since_gen; int by_tid; }; /* ... are visible to accesses after since_gen (or to by_tid at once) */
std::unordered_map<uintptr_t, std::vector<Rel>> released;   /* keyed by the atomic's address */
std::unordered_map<long long, Acq> acquired;                /* key = (acquiring cta << 32) | releasing cta */
inline long long akey(int to, int from) { return ((long long)to << 32) | (unsigned)from; }
uintptr_t img_lo = 0, img_hi = 0;
unsigned long n_hazards = 0, n_reported = 0;
bool busy = false;

int find_image(struct dl_phdr_info *info, size_t, void *) {
	const uintptr_t me = (uintptr_t)&shadow;
	uintptr_t lo = ~(uintptr_t)0, hi = 0;
	bool mine = false;
	for (int i = 0; i < info->dlpi_phnum; i++) {
		const ElfW(Phdr) &ph = info->dlpi_phdr[i];
		if (ph.p_type != PT_LOAD) continue;
		const uintptr_t a = info->dlpi_addr + ph.p_vaddr, b = a + ph.p_memsz;
		if (me >= a && me < b) mine = true;
		if (a < lo) lo = a;
		if (b > hi) hi = b;
	}
	if (mine) {
		img_lo = lo;
		img_hi = hi;
		return 1;
	}
	return 0;
}

inline bool on_fiber_stack(uintptr_t a) {
	return a >= (uintptr_t)emu::stacks_base && a < (uintptr_t)emu::stacks_base + emu::stacks_bytes;
}

/* is the earlier access `a` ordered before the current access of thread f? */
inline bool ordered(const Acc &a, bool a_shared, const emu::Fiber *f, int cta, unsigned tma_seen) {
	if (a.launch != emu::launch_seq) return true;
	if (a.cta != cta) {
		if (a_shared) return true; /* shared memory of an earlier CTA is another memory */
		auto it = acquired.find(akey(cta, a.cta));
		/* an atomic on the ticket word itself is ordered by the atomics' own order */
		if (it == acquired.end() || (a.atomic ? a.bar_gen > it->second.upto : a.bar_gen >= it->second.upto)) return false;
		return f->cta->bar_gen > it->second.since_gen || (int)f->tid.x == it->second.by_tid;
	}
	if (a.tma_seq) return tma_seen >= a.tma_seq;
	if (a.tid == (int)f->tid.x) return true;
	if (a.bar_gen != f->cta->bar_gen) return true;
	if ((a.tid >> 5) == f->warp && a.warp_gen != f->cta->w[f->warp].gen) return true;
	return false;
}

void report(const char *what, uintptr_t addr, bool shared, const Acc &prev, const emu::Fiber *f, const void *pc) {
	n_hazards++;
	if (n_reported >= 40) return;
	n_reported++;
	Dl_info d0, d1;
	const char *s0 = (dladdr(pc, &d0) && d0.dli_sname) ? d0.dli_sname : "?";
	const char *s1 = (dladdr(prev.pc, &d1) && d1.dli_sname) ? d1.dli_sname : "?";
	fprintf(stderr,
	        "EMU-RACECHECK hazard %s on %s memory %p (launch %u, cta %d): thread %d at +0x%lx [%s] vs thread %d%s at +0x%lx [%s]\n",
	        what, shared ? "shared" : "device", (void *)addr, emu::launch_seq, (int)f->cta->bid.x, (int)f->tid.x,
	        (unsigned long)((uintptr_t)pc - img_lo), s0, prev.tid, prev.tma_seq ? " (bulk copy)" : "",
	        (unsigned long)((uintptr_t)prev.pc - img_lo), s1);
}

void access(uintptr_t addr, size_t size, bool is_write, bool is_atomic, const void *pc, unsigned tma_seq = 0) {
	const emu::Fiber *f = emu::cur;
	if (!f || busy || !size) return; /* host code, or the checker's own allocations */
	if (on_fiber_stack(addr)) return;
	if (!img_lo) dl_iterate_phdr(find_image, nullptr);
	const bool shared = (addr >= img_lo && addr < img_hi) || emu::is_dynamic_smem((const void *)addr);
	busy = true;
	const int cta = (int)f->cta->bid.x;
	const unsigned tma_seen = f->tma_seen;
	for (uintptr_t a = addr; a < addr + size;) {
		const uintptr_t cell = a >> 3;
		const unsigned off = (unsigned)(a & 7);
		const unsigned n = (unsigned)((addr + size - a < 8 - off) ? addr + size - a : 8 - off);
		const unsigned char mask = (unsigned char)(((1u << n) - 1) << off);
		Cell &c = shadow[cell];
		Acc me{emu::launch_seq, cta, (int)f->tid.x, f->cta->bar_gen, f->cta->w[f->warp].gen, tma_seq, mask,
		       (unsigned char)is_atomic, pc};
		if (c.w.pc && (c.w.mask & mask) && !(c.w.atomic && is_atomic) && !ordered(c.w, shared, f, cta, tma_seen))
			report(is_write ? "write-after-write" : "read-after-write", a, shared, c.w, f, pc);
		if (is_write) {
			for (Acc &r : c.r)
				if (r.pc && (r.mask & mask) && !(r.atomic && is_atomic) && !ordered(r, shared, f, cta, tma_seen))
					report("write-after-read", a, shared, r, f, pc);
			/* bytes not covered by this write keep their history only approximately: a
			 * wider earlier write stays recorded when the new one is narrower */
			if (!c.w.pc || (c.w.mask & ~mask) == 0 || ordered(c.w, shared, f, cta, tma_seen)) c.w = me;
			c.r[0].pc = c.r[1].pc = nullptr;
		} else {
			/* two reader slots: the latest reader and one other thread */
			if (c.r[0].pc && c.r[0].launch == me.launch && c.r[0].cta == me.cta && c.r[0].tid != me.tid) c.r[1] = c.r[0];
			c.r[0] = me;
		}
		a += n;
	}
	busy = false;
}

/* an atomic read-modify-write on device memory: release what this CTA did before its last
 * barrier, acquire what other CTAs released here earlier */
void atomic_rmw(uintptr_t addr) {
	const emu::Fiber *f = emu::cur;
	if (!f || busy) return;
	busy = true;
	const int me = (int)f->cta->bid.x;
	std::vector<Rel> &rl = released[addr];
	for (const Rel &r : rl)
		if (r.cta != me) {
			Acq &q = acquired[akey(me, r.cta)];
			if (r.bar_gen > q.upto || q.upto == 0) q = Acq{r.bar_gen, f->cta->bar_gen, (int)f->tid.x};
		}
	bool found = false;
	for (Rel &r : rl)
		if (r.cta == me) {
			r.bar_gen = f->cta->bar_gen;
			found = true;
		}
	if (!found) rl.push_back(Rel{me, f->cta->bar_gen});
	busy = false;
}

#define PC __builtin_return_address(0)

} /* namespace */

/* called by the emulator's bulk-copy stand-in: the copy's writes, issued by the current thread */
extern "C" void emu_racecheck_bulk_write(void *dst, size_t bytes, unsigned seq) {
	access((uintptr_t)dst, bytes, true, false, PC, seq);
}
extern "C" void emu_racecheck_summary(void) {
	fprintf(stderr, "EMU-RACECHECK: %lu hazard(s)\n", n_hazards);
}
extern "C" unsigned long emu_racecheck_count(void) { return n_hazards; }
/* a launch starts from a clean history: everything before it is ordered by the kernel boundary */
extern "C" void emu_racecheck_launch(void) {
	static bool registered = false;
	if (!registered) {
		registered = true;
		atexit(emu_racecheck_summary);
	}
	busy = true;
	shadow.clear();
	released.clear();
	acquired.clear();
	busy = false;
}

extern "C" {
void __tsan_init(void) {}
void __tsan_func_entry(void *) {}
void __tsan_func_exit(void) {}
void __tsan_vptr_update(void **, void *) {}
void __tsan_acquire(void *) {}
void __tsan_release(void *) {}
void __tsan_read1(void *p) { access((uintptr_t)p, 1, false, false, PC); }
void __tsan_read2(void *p) { access((uintptr_t)p, 2, false, false, PC); }
void __tsan_read4(void *p) { access((uintptr_t)p, 4, false, false, PC); }
void __tsan_read8(void *p) { access((uintptr_t)p, 8, false, false, PC); }
void __tsan_read16(void *p) { access((uintptr_t)p, 16, false, false, PC); }
void __tsan_write1(void *p) { access((uintptr_t)p, 1, true, false, PC); }
void __tsan_write2(void *p) { access((uintptr_t)p, 2, true, false, PC); }
void __tsan_write4(void *p) { access((uintptr_t)p, 4, true, false, PC); }
void __tsan_write8(void *p) { access((uintptr_t)p, 8, true, false, PC); }
void __tsan_write16(void *p) { access((uintptr_t)p, 16, true, false, PC); }
void __tsan_unaligned_read2(void *p) { access((uintptr_t)p, 2, false, false, PC); }
void __tsan_unaligned_read4(void *p) { access((uintptr_t)p, 4, false, false, PC); }
void __tsan_unaligned_read8(void *p) { access((uintptr_t)p, 8, false, false, PC); }
void __tsan_unaligned_read16(void *p) { access((uintptr_t)p, 16, false, false, PC); }
void __tsan_unaligned_write2(void *p) { access((uintptr_t)p, 2, true, false, PC); }
void __tsan_unaligned_write4(void *p) { access((uintptr_t)p, 4, true, false, PC); }
void __tsan_unaligned_write8(void *p) { access((uintptr_t)p, 8, true, false, PC); }
void __tsan_unaligned_write16(void *p) { access((uintptr_t)p, 16, true, false, PC); }
void __tsan_read_range(void *p, unsigned long n) { access((uintptr_t)p, n, false, false, PC); }
void __tsan_write_range(void *p, unsigned long n) { access((uintptr_t)p, n, true, false, PC); }

/* atomics: performed for real; recorded as atomic accesses */
int __tsan_atomic32_load(const volatile int *p, int) {
	access((uintptr_t)p, 4, false, true, PC);
	return __atomic_load_n(p, __ATOMIC_SEQ_CST);
}
void __tsan_atomic32_store(volatile int *p, int v, int) {
	access((uintptr_t)p, 4, true, true, PC);
	__atomic_store_n(p, v, __ATOMIC_SEQ_CST);
}
int __tsan_atomic32_fetch_or(volatile int *p, int v, int) {
	access((uintptr_t)p, 4, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}
int __tsan_atomic32_fetch_add(volatile int *p, int v, int) {
	access((uintptr_t)p, 4, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
int __tsan_atomic32_exchange(volatile int *p, int v, int) {
	access((uintptr_t)p, 4, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
long long __tsan_atomic64_load(const volatile long long *p, int) {
	access((uintptr_t)p, 8, false, true, PC);
	return __atomic_load_n(p, __ATOMIC_SEQ_CST);
}
void __tsan_atomic64_store(volatile long long *p, long long v, int) {
	access((uintptr_t)p, 8, true, true, PC);
	__atomic_store_n(p, v, __ATOMIC_SEQ_CST);
}
long long __tsan_atomic64_fetch_add(volatile long long *p, long long v, int) {
	access((uintptr_t)p, 8, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
long long __tsan_atomic64_fetch_or(volatile long long *p, long long v, int) {
	access((uintptr_t)p, 8, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}
long long __tsan_atomic64_fetch_xor(volatile long long *p, long long v, int) {
	access((uintptr_t)p, 8, true, true, PC);
	atomic_rmw((uintptr_t)p);
	return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST);
}
}
