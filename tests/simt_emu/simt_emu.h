// TEST INFRASTRUCTURE ONLY.  A tiny single-threaded SIMT emulator: runs the *same* device source
// (gama_tts_b200/csrc/tube_kernel.cuh, compiled with -DGTTS_EMU by g++) one CTA at a time, each CUDA
// thread as a ucontext fiber, with warp-level (__syncwarp / __shfl_sync) and CTA-level
// (__syncthreads, named barriers) rendezvous.  It exists so that the kernel's indexing, blocking and
// lane-role logic can be checked against the oracle in this GPU-less container; it is never linked
// into the product library and the product has no CPU path.
#ifndef SIMT_EMU_H_
#define SIMT_EMU_H_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>
#include <ucontext.h>

struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct float2 { float x, y; };

namespace simt {

struct Barrier { uint32_t arrived = 0; uint64_t gen = 0; };

struct Fiber {
	ucontext_t ctx;
	std::vector<char> stack;
	int state = 0;              // 0 runnable, 1 blocked, 2 done
	bool spun = false;          // yielded from a polling loop (spin_yield): runnable, but made no progress
};

struct Cta {
	int nthreads = 0;
	std::vector<Fiber> fibers;
	std::vector<std::map<uint32_t, Barrier>> warpBarriers;     // per warp, keyed by participant mask
	std::vector<std::vector<uint64_t>> xchg;                   // per warp, 32 slots
	std::map<int, std::pair<int, uint64_t>> ctaBarriers;       // id -> (arrived count, generation)
	std::map<int, std::vector<int>> ctaWaiters;
	ucontext_t sched;
	int cur = -1;
	std::function<void(int)> body;
};

extern Cta* g_cta;

inline void yield_to_scheduler() { swapcontext(&g_cta->fibers[g_cta->cur].ctx, &g_cta->sched); }

// Called inside a polling loop on shared memory (the kernels' role_wait): lets every other fiber run once before the
// poll is repeated.
inline void spin_yield()
{
	g_cta->fibers[g_cta->cur].spun = true;
	yield_to_scheduler();
}

inline void syncwarp(uint32_t mask)
{
	Cta& c = *g_cta;
	const int tid = c.cur, warp = tid >> 5, lane = tid & 31;
	Barrier& b = c.warpBarriers[warp][mask];
	b.arrived |= 1u << lane;
	if ((b.arrived & mask) == mask) {
		b.arrived = 0;
		b.gen++;
		for (int l = 0; l < 32; ++l) {
			if ((mask >> l) & 1) {
				const int t = warp * 32 + l;
				if (t < c.nthreads && c.fibers[t].state == 1) c.fibers[t].state = 0;
			}
		}
		return;
	}
	const uint64_t gen = b.gen;
	while (c.warpBarriers[warp][mask].gen == gen) {
		c.fibers[tid].state = 1;
		yield_to_scheduler();
	}
}

inline void cta_barrier(int id, int count)
{
	Cta& c = *g_cta;
	const int tid = c.cur;
	auto& b = c.ctaBarriers[id];
	b.first++;
	if (b.first >= count) {
		b.first = 0;
		b.second++;
		for (int t : c.ctaWaiters[id]) c.fibers[t].state = 0;
		c.ctaWaiters[id].clear();
		return;
	}
	const uint64_t gen = b.second;
	c.ctaWaiters[id].push_back(tid);
	while (c.ctaBarriers[id].second == gen) {
		c.fibers[tid].state = 1;
		yield_to_scheduler();
	}
}

// bar.arrive: counts towards the barrier without waiting for it
inline void cta_barrier_arrive(int id, int count)
{
	Cta& c = *g_cta;
	auto& b = c.ctaBarriers[id];
	b.first++;
	if (b.first >= count) {
		b.first = 0;
		b.second++;
		for (int t : c.ctaWaiters[id]) c.fibers[t].state = 0;
		c.ctaWaiters[id].clear();
	}
}

template<class T>
inline T shfl(uint32_t mask, T v, int src)
{
	static_assert(sizeof(T) <= 8, "shfl: at most 64 bits");
	Cta& c = *g_cta;
	const int warp = c.cur >> 5, lane = c.cur & 31;
	uint64_t bits = 0;
	std::memcpy(&bits, &v, sizeof(T));
	c.xchg[warp][lane] = bits;
	syncwarp(mask);
	bits = c.xchg[warp][src & 31];
	syncwarp(mask);
	T r;
	std::memcpy(&r, &bits, sizeof(T));
	return r;
}

void fiber_entry(int tid);

// Runs one CTA of `nthreads` threads; body(tid) is the kernel body.
inline void run_cta(int nthreads, std::function<void(int)> body, size_t stackBytes = 256 * 1024)
{
	Cta cta;
	cta.nthreads = nthreads;
	cta.fibers.resize(nthreads);
	cta.warpBarriers.resize((nthreads + 31) / 32);
	cta.xchg.assign((nthreads + 31) / 32, std::vector<uint64_t>(32, 0));
	cta.body = body;
	g_cta = &cta;
	for (int t = 0; t < nthreads; ++t) {
		Fiber& f = cta.fibers[t];
		f.stack.resize(stackBytes);
		getcontext(&f.ctx);
		f.ctx.uc_stack.ss_sp = f.stack.data();
		f.ctx.uc_stack.ss_size = f.stack.size();
		f.ctx.uc_link = &cta.sched;
		makecontext(&f.ctx, (void (*)()) fiber_entry, 1, t);
	}
	// Scheduler: drain one warp as far as it goes, then the next; stop when all fibers are done.  Fibers that poll
	// (spin_yield) are resumed once per round; a round in which nothing but polling happened counts towards the
	// deadlock limit.
	long idleRounds = 0;
	for (;;) {
		bool anyAlive = false, progressed = false, anySpun = false;
		for (int w = 0; w * 32 < nthreads; ++w) {
			bool ranInWarp = true;
			std::vector<char> spunThisRound(32, 0);
			while (ranInWarp) {
				ranInWarp = false;
				for (int l = 0; l < 32; ++l) {
					const int t = w * 32 + l;
					if (t >= nthreads) break;
					Fiber& f = cta.fibers[t];
					if (f.state == 2) continue;
					anyAlive = true;
					if (f.state == 0 && !spunThisRound[l]) {
						cta.cur = t;
						f.spun = false;
						swapcontext(&cta.sched, &f.ctx);
						if (f.spun) {
							spunThisRound[l] = 1;
							anySpun = true;
						} else {
							ranInWarp = true;
							progressed = true;
						}
					}
				}
			}
		}
		if (!anyAlive) break;
		if (progressed) {
			idleRounds = 0;
		} else if (!anySpun || ++idleRounds > 100000) {
			std::fprintf(stderr, "simt_emu: deadlock (all live threads blocked%s)\n", anySpun ? " or polling" : "");
			std::abort();
		}
	}
	g_cta = nullptr;
}

} // namespace simt

// ---- the CUDA surface the kernel source uses ---------------------------------------------------------
#define GTTS_DEV static inline
#define GTTS_DEV_NOINLINE static inline
#define GTTS_CONST
#define __syncwarp(...) simt::syncwarp(0xffffffffu)
#define __syncthreads() simt::cta_barrier(0, simt::g_cta->nthreads)
template<class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	const int lane = simt::g_cta->cur & 31;
	const int s = (lane & ~(width - 1)) | (src & (width - 1));
	return simt::shfl<T>(mask, v, s);
}
template<class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	const int lane = simt::g_cta->cur & 31;
	const int src = ((lane & (width - 1)) >= (int) delta) ? lane - (int) delta : lane;
	return simt::shfl<T>(mask, v, src);
}
inline int atomicAdd(int* p, int v) { const int old = *p; *p += v; return old; }
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline int __reduce_max_sync(unsigned mask, int v)
{
	int m = v;
	for (int l = 0; l < 32; ++l) if ((mask >> l) & 1) { const int o = __shfl_sync(mask, v, l, 32); if (o > m) m = o; }
	return m;
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) if ((mask >> l) & 1) { if (__shfl_sync(mask, pred, l, 32)) r |= 1u << l; }
	return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) if ((mask >> l) & 1) r |= __shfl_sync(mask, v, l, 32);
	return r;
}
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline unsigned __double2uint_rz(double x) { return x <= 0.0 ? 0u : (unsigned) x; }
// The emulator build uses pow() so that its results can be compared bit-for-bit with the oracle.
#define gtts_exp2(x) std::pow(2.0, (x))
#define gtts_exp10(x) std::pow(10.0, (x))
using std::rint; using std::tan; using std::cos; using std::sin;

#endif
