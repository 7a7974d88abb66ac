// Decoupled, warp-specialised tube kernel (sm_100a), second generation of tube_kernel_v1.cuh.
//
// One persistent CTA per SM (23 warps, 80 registers) steps kSlots = 7 utterances, 32 internal samples ("block") per
// iteration, through a software pipeline of warp roles.  Block b of an utterance is handled at slot iteration
// it = b + stage:
//
//   stage 0   coef worker   float32 walk of parameter 0 (pitch)                              lane = parameter
//   stage 1   slot helper   f0 -> oscillator increments                                      lane = sample
//   stage 2   chain A       oscillator phase recurrence (serial)                             lane = slot
//             coef worker   float32 walk of parameters 1..6
//   stage 3   slot helper   amplitudes, frication taps, bandpass coefficients, noise (LCG jump-ahead), wavetable
//                           lookup, 49-tap FIR, mixing                                        lane = sample
//             coef worker   float32 walk of parameters 7..15 (radii, velum)
//   stage 4   chain A2      frication bandpass biquad + tap signals (serial)                 lane = slot
//             coef worker   junction coefficients (9 divisions per sample)                   lane = sample
//   stage 5   tube warps    the waveguide: one cell per lane, 16 lanes per utterance, three shuffles per sample
//   stage 6   chain B       radiation filters + throat low-pass (serial), output sum         lane = slot x filter
//   stage 7   SRC worker    windowed-sinc sample-rate conversion, two consecutive outputs per lane from one
//                           register-held window, whole 256-byte row pairs                    lane = output pair
//
// What changed against v1 (measurements in DESIGN.md section 4):
//  * No CTA-wide barrier.  Before iteration i a role waits only for the roles it shares a buffer with -- its
//    producers AND its consumers -- to have completed iteration i - 1: one named hardware barrier per role type and
//    iteration parity, neighbours arrive without blocking, the type's warps wait without issuing.  That is the
//    guarantee the CTA barrier gave, restricted to the pairs that need it; roles that are not neighbours drift apart
//    by up to three iterations (bounded by the four-deep ring of slot control blocks).
//  * The float32 parameter walks (Controller.cpp:297-311) run on the slots' coefficient workers (lanes 0..15 = the
//    sixteen parameters, walk_slot), off the helpers.
//  * SRC: every lane forms two consecutive outputs from ONE 27-sample window held in registers (half the window
//    loads), coefficients still shared by all aligned slots.
//  * The six junction coefficients that are per-voice constants stay in the tube lanes' registers; alpha_u of the
//    3-way junction is derived from alpha (alpha_l + alpha_r + alpha_u = 2); serial chains load two samples per
//    16-byte shared-memory access.
//
// The per-sample arithmetic follows tube_kernel.cuh (v0, the general kernel); the arithmetic contract is the one
// stated there.  The file is also compiled for the host under tests/simt_emu (GTTS_EMU).
#ifndef GTTS_TUBE_KERNEL_V2_CUH_
#define GTTS_TUBE_KERNEL_V2_CUH_

#include "tube_kernel.cuh"

#ifndef GTTS_ROLE_PRESET
#define GTTS_ROLE_PRESET 0
#endif

namespace gtts {
namespace v2 {

enum {
	kSlots = 7,
	kWarps = 23,
	kThreads = kWarps * 32,
	kTubeWarps = 4,
	kParamRow = 36,               // floats per walked-parameter row: 16-byte aligned, rows 4 banks apart
	kRow = 34,                    // doubles per per-sample row: rows start on 16 bytes (two samples per access), 4 banks apart
	kXr = kSrcRing + 32,          // tube-output ring with its first 32 entries mirrored behind the end: any 27-sample window is contiguous
	kSrcPad = 256,                // zero entries behind the SRC table (taps one past either wing read 0)
	// pipeline stages (slot iteration - block)
	kStWalk0 = 0, kStF0 = 1, kStPhase = 2, kStWalk1 = 2, kStSrc = 3, kStWalk7 = 3, kStCoef = 4, kStTube = 5, kStRad = 6, kStOut = 7,
	kStages = 7,
	// roles (one warp each)
	kRoleTube0 = 0,
	kRoleChainA = 4,
	kRoleChainB = 5,
	kRoleChainA2 = 6,
	kRoleHelper0 = 7,             // + slot
	kRoleCoef0 = 14,              // + slot: the slot's float32 walks, junction coefficients, per-slot SRC
	kRoleSrcB = 21,               // shared SRC tasks (slots in lockstep), slots 4..6; per-slot SRC runs on the slot's coefficient worker
	kRoleSrcA = 22,               // shared SRC tasks, slots 0..3
	kCtrSched = 24,               // counter: control blocks published by the scheduler
	kCounters = 32,
	kCtlRing = 4,
	kFar = 2,                     // the scheduler may run this many iterations ahead of any role (ring of 4 control blocks)
	kNever = 1 << 28,
};

struct SlotSm {
	double osc[2][kRow];
	double pos[2][2][kRow];
	double sig[2][kRow];
	double bp[2][3][kRow];        // b0, a1, a2
	double tapa[2][kRow], tapb[2][kRow];
	double thr[4][kRow];
	double in[3][kRow];
	double2 pab[2][kRow];         // {tapA * fric, tapB * fric}
	double2 kab[2][5][kRow];      // the ten per-sample junction coefficients (times damping), see kVarCell
	double onepk7[3][kRow];
	double endm[2][kRow], endn[2][kRow];
	double rad[3][kRow];          // mouth radiation, nose radiation, throat outputs of one block
	double2 vw[kVRing];           // 2x oscillator stream {even phase, odd phase}: [0, 24) history, [24, 56) current block
	double xring[kXr];
	alignas(16) float cur[2][7][kParamRow];     // walked parameters 0..6, [block parity][parameter][sample]
	alignas(16) float pscr[2][9][kParamRow];    // walked parameters 7..15
	signed char ip[3][kBlock];    // integer part of the frication position (-50: none)
	VoiceDev V[2];                // voice constants of the slot's utterances, alternating (Ctl::vbuf)
	int    fric[2];               // block has frication (some tap * bandpassed noise != 0), per pab buffer
	float  ckey[9];               // parameters 7..15 at the last sample of the previous block (NaN: none)
	// control blocks: the scheduler writes ctl[(i + 1) % 4] during iteration i, the roles read ctl[i % 4]
	struct Ctl {
		UttDesc U;
		int it;                   // iteration counter of the current utterance; -1: idle
		int nblocks;
		int voice;
		int last_len;             // samples in the last block
		int vbuf;                 // which V[] holds this utterance's voice
		int gran;                 // output rows are completed in units of this many samples (64 or 32)
		long long src_k0, src_k1; // outputs of the block in the SRC stage this iteration (see src_rows)
		long long eq;
		unsigned erem, eQ, eR, inc;
	} ctl[kCtlRing];
	int     pad_[14];             // slot stride = 16 mod 128 bytes: the 7 slots start in different banks
};
static_assert(sizeof(SlotSm) % 128 == 16, "slot stride should be 16 mod 128 bytes (the chains read one 16-byte word per slot and lane); adjust pad_");

struct CtaSm {
	double2 tab[kSrcFilterLen + kSrcPad];
	SlotSm slot[kSlots];
	struct Sched {
		int live;                 // some slot has work
		int src_shared;           // 1: every slot in the SRC stage is at the same block of an equally long utterance of the same rate
		int src_mask;             // slots with a block in the SRC stage
		int src_tasks;            // shared SRC tasks (64 outputs each) this iteration
		long long src_k0, src_k1; // shared SRC tasks: first output of the first row pair, end of the range
		long long src_k0w;        // first output to write (>= src_k0)
	} sched[kCtlRing];
	int done[kCounters];          // iterations completed per role; [kCtrSched]: control blocks published
};

struct KernelParamsV2 {
	const VoiceDev* voices;
	const double* tables;         // per-voice glottal wavetables, 512 doubles each (host-built)
	const UttDesc* utts;
	const int32_t* order;
	const float* frames;
	float* out;
	const double2* src_tab;
	int32_t* queue;
	int32_t n_utt;
	UttStateV2* states;           // streaming: per-utterance state between chunks (UttDesc::state_index), else nullptr
	int32_t debug_skip;           // development builds only (-DGTTS_EXPERIMENTS)
	long long* prof;              // development builds only (-DGTTS_ROLE_PROFILE): [grid][2 * kWarps + 1]
};

#ifndef GTTS_SRC1_UNROLL
#define GTTS_SRC1_UNROLL 1
#endif

#ifdef GTTS_EXPERIMENTS
#define GTTS_SKIP_BITS2(P) const int skip = (P).debug_skip
#else
#define GTTS_SKIP_BITS2(P) constexpr int skip = 0
#endif

// ---- role synchronisation --------------------------------------------------------------------------------
#ifndef GTTS_EMU
GTTS_DEV int ld_acquire_shared(const int* p)
{
	int v;
	asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned) __cvta_generic_to_shared(p)) : "memory");
	return v;
}
GTTS_DEV void st_release_shared(int* p, int v)
{
	asm volatile("st.release.cta.shared.b32 [%0], %1;" :: "r"((unsigned) __cvta_generic_to_shared(p)), "r"(v) : "memory");
}
GTTS_DEV void backoff(unsigned ns) { __nanosleep(ns); }
// Named hardware barriers (ids 1..15; 0 is __syncthreads): `count` threads take part, some arriving (not blocking),
// some waiting; prior shared-memory writes of the arriving threads are visible to the waiting ones when it completes.
GTTS_DEV void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }
GTTS_DEV void bar_wait(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
#define GTTS_OPAQUE_INT(x) asm volatile("" : "+r"(x))
#else
GTTS_DEV int ld_acquire_shared(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
GTTS_DEV void st_release_shared(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
GTTS_DEV void backoff(unsigned) { simt::spin_yield(); }
GTTS_DEV void bar_arrive(int id, int count) { simt::cta_barrier_arrive(id, count); }
GTTS_DEV void bar_wait(int id, int count) { simt::cta_barrier(id, count); }
#define GTTS_OPAQUE_INT(x) ((void) 0)
#endif

// ---- who waits for whom ----------------------------------------------------------------------------------------
// Two roles are neighbours when they share a buffer, in either direction (producer -> consumer: the data must be
// there; consumer -> producer: the buffer must be free again; every buffer is as deep as its stage distance + 1).
// A role starts iteration i when every neighbour has completed iteration i - 1: the guarantee the CTA barrier of v1
// gave, restricted to the pairs that need it.  Neighbourhood is kept per role TYPE:
//
//   type      warps  neighbours (shared buffers)
//   tube        4    helper (in, ip), coef (kab), chain A2 (pab, fric), chain B (endm, endn)
//   chain A     1    helper (osc, pos)
//   chain A2    1    helper (sig, bp, taps), tube
//   chain B     1    tube, helper (thr), coef (onepk7; xring for the per-slot SRC), SRC (xring)
//   helper      7    chain A, chain A2, tube, chain B, walker (cur)
//   coef        7    walker (pscr), tube, chain B
//   SRC         2    chain B
//   walker      1    helper, coef
//
// and implemented with one named hardware barrier per type and iteration parity (14 of the 15 ids): before
// iteration i the warps of a type wait on barrier (type, i & 1); every neighbour warp arrives on it -- without
// blocking -- when it has completed its iteration i - 1.  A waiting warp issues nothing.  (A neighbour cannot arrive
// for iteration i + 2 before the type has passed iteration i: it would have to complete i + 1 first, which needs the
// type's iteration i.)  The walker, the only type nobody produces for, polls its neighbours' completion counters
// instead; so does the scheduler for its loose bound (the ring of control blocks), and every role checks that its
// control block is published -- it is, up to kFar iterations ahead.
enum { kTypeTube = 0, kTypeA, kTypeA2, kTypeB, kTypeHelper, kTypeCoef, kTypeSrc, kTypes };

GTTS_DEV constexpr int type_of_role(int role)
{
	if (role < kTubeWarps) return kTypeTube;
	if (role == kRoleChainA) return kTypeA;
	if (role == kRoleChainB) return kTypeB;
	if (role == kRoleChainA2) return kTypeA2;
	if (role < kRoleCoef0) return kTypeHelper;
	if (role < kRoleSrcB) return kTypeCoef;
	return kTypeSrc;
}

// bit t set: type t is a neighbour
GTTS_DEV constexpr unsigned neighbour_types(int type)
{
	switch (type) {
	case kTypeTube:   return (1u << kTypeHelper) | (1u << kTypeCoef) | (1u << kTypeA2) | (1u << kTypeB);
	case kTypeA:      return (1u << kTypeHelper);
	case kTypeA2:     return (1u << kTypeHelper) | (1u << kTypeTube);
	case kTypeB:      return (1u << kTypeTube) | (1u << kTypeHelper) | (1u << kTypeCoef) | (1u << kTypeSrc);
	case kTypeHelper: return (1u << kTypeA) | (1u << kTypeA2) | (1u << kTypeTube) | (1u << kTypeB) | (1u << kTypeCoef);
	case kTypeCoef:   return (1u << kTypeHelper) | (1u << kTypeTube) | (1u << kTypeB);
	default:          return (1u << kTypeB);
	}
}

GTTS_DEV constexpr int warps_of_type(int type)
{
	return type == kTypeTube ? kTubeWarps : ((type == kTypeHelper || type == kTypeCoef) ? kSlots : (type == kTypeSrc ? 2 : 1));
}

// threads taking part in the barrier of `type`: its own warps (waiting) and its neighbours' (arriving)
GTTS_DEV constexpr int barrier_threads(int type)
{
	int warps = warps_of_type(type);
	const unsigned nb = neighbour_types(type);
	for (int t = 0; t < kTypes; ++t) if ((nb >> t) & 1u) warps += warps_of_type(t);
	return 32 * warps;
}

GTTS_DEV constexpr int barrier_id(int type, int it) { return 1 + 2 * type + (it & 1); }

// Polls until done[q] >= need for every counter q (lane q reads counter q; lanes with takesPart == false are left
// out).  Between polls the warp sleeps, 100 ns at first and twice as long every time up to 0.8 us.  Used where a
// wait is rare or the waiting warp is a single light one (see above).
GTTS_DEV void poll_counters(CtaSm* C, int lane, int need, bool takesPart)
{
	unsigned ns = 100;
	for (;;) {
		const int v = ld_acquire_shared(&C->done[lane]);
		const bool ok = !takesPart || v >= need;
		if (__all_sync(0xffffffffu, ok)) break;
		backoff(ns);
		ns = ns < 800u ? 2u * ns : ns;
	}
}

// Before iteration `it`: the neighbours have completed it - 1, control block `it` is published.
template<int TYPE>
GTTS_DEV void role_wait(CtaSm* C, int lane, int it)
{
	if (it > 0) {
		constexpr int threads = barrier_threads(TYPE);
		bar_wait(barrier_id(TYPE, it), threads);
	}
	poll_counters(C, lane, it + 1, lane == kCtrSched);
}

// iteration `it` of `role` is complete: publish the counter, arrive on the neighbour types' barriers of iteration it + 1
template<int TYPE, int T>
GTTS_DEV void arrive_on_neighbours(int it)
{
	if (T < kTypes) {
		constexpr unsigned nb = neighbour_types(TYPE);
		if ((nb >> T) & 1u) {
			constexpr int threads = barrier_threads(T);
			bar_arrive(barrier_id(T, it + 1), threads);
		}
		arrive_on_neighbours<TYPE, (T < kTypes ? T + 1 : T)>(it);
	}
}

template<int TYPE>
GTTS_DEV void role_signal(CtaSm* C, int role, int lane, int it)
{
	__syncwarp();
	if (lane == 0) st_release_shared(&C->done[role], it + 1);
	arrive_on_neighbours<TYPE, 0>(it);
}

// control blocks 0 .. n - 1 are published
GTTS_DEV void sched_signal(CtaSm* C, int lane, int n)
{
	__syncwarp();
	if (lane == 0) st_release_shared(&C->done[kCtrSched], n);
}

// Streaming (gtts_stream_*): an utterance arrives in chunks of whole control periods.  A chunk runs like a short
// utterance -- every pipeline stage sees every block of it, so all roles end at the same sample -- except that each
// role starts from the state it saved at the end of the previous chunk (UttDesc::flags bit 0) and, unless it is the
// last one (bit 1), the SRC is not flushed.  Sample and output indices are absolute (UttDesc::n_in_base, out_begin =
// destination - outputs produced so far).
// ST: the kernel is instantiated twice; the batch kernel (ST = false) carries none of the state handling.
template<bool ST>
GTTS_DEV UttStateV2* chunk_state(const KernelParamsV2& P, const UttDesc& U)
{
	if (!ST) return nullptr;
	return (P.states != nullptr && U.state_index >= 0) ? &P.states[U.state_index] : nullptr;
}
GTTS_DEV bool chunk_resumes(const UttStateV2* st, const UttDesc& U) { return st != nullptr && (U.flags & 1) != 0 && st->started != 0; }

GTTS_DEV int block_len(const SlotSm::Ctl& k, int b)
{
	return b == k.nblocks - 1 ? k.last_len : kBlock;
}

// ---- float32 walk (Controller.cpp:297-311), lane = (slot, parameter) -----------------------------------------
// one parameter of nb consecutive frames -> the lane's row (zero-filled to the block), 16 bytes per store
GTTS_DEV_NOINLINE void copy_frames(const float* src, int nb, float* out)
{
#pragma unroll 2
	for (int j = 0; j < kBlock; j += 4) {
		float v[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) v[q] = (j + q < nb) ? src[(long long) (j + q) * kNumParams] : 0.f;
		reinterpret_cast<float4*>(out)[j >> 2] = make_float4(v[0], v[1], v[2], v[3]);
	}
}

struct WalkRegs { float cur, delta, fn1, fn2; int off, frame; };

GTTS_DEV void cursor_init(const float* frames, long long nFrames, float invSteps, int param, WalkRegs& w)
{
	const float a = frames[param];
	const float b = (nFrames > 1) ? frames[kNumParams + param] : a;
	const float c = (nFrames > 2) ? frames[2 * kNumParams + param] : b;
	w.cur = a;
	w.delta = __fmul_rn(__fsub_rn(b, a), invSteps);
	w.fn1 = b;
	w.fn2 = c;
	w.off = 0;
	w.frame = 0;
}

// Walks the lane's parameter over one block: out[j] receives the value used for sample j.  Control periods are at
// least one block long, so at most one frame boundary falls inside the block; at it the walk restarts from the next
// frame value, not from the accumulated one (Controller.cpp:297-300).  The cursor carries the next two frame values
// (fn1 = frame[f+1], fn2 = frame[f+2], clamped to the last frame): they are fetched from global memory one control
// period ahead, and fn2 is read only in a block that contains a boundary, so that a fetch still in flight (pinned
// host memory: microseconds over PCIe) never stalls the walk.  One compact loop (restart by select): the code is
// shared by the seven coefficient workers, which matters more than its instruction count (instruction cache).
GTTS_DEV void walk_block(const float* frames, long long nFrames, int steps, float invSteps, int param,
			int nb, WalkRegs& w, float* out, bool active)
{
	const bool perSample = steps == 1;
	const int first = (steps - w.off) < nb ? (steps - w.off) : nb;
	const bool reaches = active && !perSample && (w.off + first == steps);
	const int restart = reaches ? first : -1;
	float c = w.cur, d = w.delta, c1 = 0.f, d1 = 0.f;
	if (__any_sync(0xffffffffu, reaches)) {
		c1 = w.fn1;
		d1 = __fmul_rn(__fsub_rn(w.fn2, w.fn1), invSteps);
	}
	// the groups of four samples in which some lane's control period ends (all 32 lanes vote here: the lanes of a
	// per-sample utterance leave below)
	const unsigned slowGroups = __reduce_or_sync(0xffffffffu, (restart >= 0 && restart < kBlock) ? (1u << (restart >> 2)) : 0u);
	if (perSample) {
		// one frame per internal sample (the plugin shim records the reference's per-sample parameters): the values
		// ARE the frames, no walk; `frame` counts the samples consumed so far
		if (active) {
			copy_frames(frames + (long long) w.frame * kNumParams + param, nb, out);
			w.frame += nb;
		}
		return;
	}
	float4* o = reinterpret_cast<float4*>(out);
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock; j0 += 4) {
		float v[4];
		// the control period ends inside these four samples on some lane (the three groups of parameters of a slot are
		// at different blocks): once per period; everywhere else the walk is four additions and one 16-byte store
		if ((slowGroups >> (j0 >> 2)) & 1u) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const bool r = (j0 + q) == restart;
				c = r ? c1 : c;
				d = r ? d1 : d;
				v[q] = c;
				c = __fadd_rn(c, d);
			}
		} else {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				v[q] = c;
				c = __fadd_rn(c, d);
			}
		}
		if (active) o[j0 >> 2] = make_float4(v[0], v[1], v[2], v[3]);
	}
	if (restart == kBlock) { c = c1; d = d1; }
	if (active) {
		w.cur = c;
		w.delta = d;
		if (reaches) {
			w.frame += 1; w.off = nb - first; w.fn1 = w.fn2;
			// refill for the boundary after the next one: frame[f+3], needed a whole control period from now
			if ((long long) w.frame + 2 < nFrames) w.fn2 = frames[((long long) w.frame + 2) * kNumParams + param];
		} else {
			w.off += nb;
		}
	}
}

// The slot's sixteen walks (lanes 0..15 = parameters): parameter 0 of block it, 1..6 of block it - 2, 7..15 of block it - 3.
template<bool ST>
GTTS_DEV void walk_slot(SlotSm* S, const KernelParamsV2& P, int lane, int p, WalkRegs& w)
{
	const int param = lane & 15;
	const int stage = param == 0 ? kStWalk0 : (param < 7 ? kStWalk1 : kStWalk7);
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - stage;
	const bool active = lane < 16 && K.it >= 0 && b >= 0 && b < K.nblocks;
	const float* frames = P.frames + K.U.frame_begin * kNumParams;
	UttStateV2* st = chunk_state<ST>(P, K.U);
	if (active && b == 0) {
		cursor_init(frames, K.U.n_frames, K.U.inv_steps, param, w);
		if (chunk_resumes(st, K.U)) {
			// a chunk of a stream starts inside the control period of its first frame: the accumulated value comes
			// from the previous chunk, the position in the period from the sample count
			w.cur = st->walk_cur[param];
			w.off = (int) (K.U.n_in_base % K.U.steps);
		}
	}
	const int nb = active ? block_len(K, b) : kBlock;
	float* row = param < 7 ? S->cur[b & 1][param] : S->pscr[b & 1][param - 7];
	walk_block(frames, K.U.n_frames, active ? K.U.steps : kBlock, K.U.inv_steps, param, nb, w, row, active);
	if (active && st != nullptr && b == K.nblocks - 1) st->walk_cur[param] = w.cur;
}

// ---- slot helper: stage 1 (f0) and stage 3 (source) ------------------------------------------------------------
struct HelperRegs {
	unsigned long long mult;      // 377^(lane+1) mod 2^44, this lane's jump-ahead multiplier (loaded once)
	unsigned long long lcg;       // noise generator state on the 2^-44 grid: next samples are lcg * 377^(j+1) mod 2^44
	double noise_x1;
	int low;                      // lowest wavetable closure point below div1 seen so far (kNoLowMark: none), see rise_segment_scan
	// conversions of the previous block, re-used while the parameter does not change (the reference caches the same
	// way: BandpassFilter.h:93, WavetableGlottalSource.h:164)
	float c_p1, c_p2, c_p3, c_p5, c_p6;
	double c_ax, c_ah1, c_fa, c_a2, c_a1, c_b0;
};

// Rare path of the wavetable lookup.  Glottal volume above 60 dB (ax > 1) with tn_min != tn_max: the reference's
// table rewrite (setup(), WavetableGlottalSource.h:162-184) then zeroes [newDiv2, div2) with newDiv2 < div1, i.e.
// part of the RISE segment, which nothing ever rewrites: from that sample on the entries at or above the lowest
// closure point seen so far read 0 until reset().  `mine` is the lane's closure point if it fell below div1
// (kNoLowMark otherwise), `carried` the minimum of the earlier blocks; returns the running minimum up to and including
// the lane's sample (by value: a reference into the role's register struct would move the struct to local memory).
GTTS_DEV_NOINLINE int rise_segment_scan(int mine, int lane, int carried)
{
#pragma unroll 1
	for (int dlt = 1; dlt < 32; dlt <<= 1) {
		const int o = __shfl_up_sync(0xffffffffu, mine, dlt, 32);
		if (lane >= dlt) mine = mine < o ? mine : o;
	}
	return mine < carried ? mine : carried;
}

template<bool ST>
GTTS_DEV void helper_iteration(SlotSm* S, const KernelParamsV2& P, int lane, HelperRegs& h, int p)
{
	const SlotSm::Ctl& K = S->ctl[p];
	const int it = K.it;
	if (it < 0) return;
	const VoiceDev& V = S->V[K.vbuf];
	const int b1 = it - kStF0, b3 = it - kStSrc;
	if (b1 >= 0 && b1 < K.nblocks) {
		// Util::frequency (VTMUtil.h:76-84) and the oscillator increment of one half sample
		const int nb = block_len(K, b1);
		if (lane < nb) {
			const double f0 = 220.0 * gtts_exp2(((double) S->cur[b1 & 1][0][lane] + 3.0) * (1.0 / 12.0));
			S->osc[b1 & 1][lane] = (f0 / 2.0) * V.basic_inc;
		}
	}
	if (b3 < 0 || b3 >= K.nblocks) return;
	UttStateV2* st = chunk_state<ST>(P, K.U);
	if (b3 == 0) {
		// new utterance: clear the oscillator history, reset the noise generator and the caches (a chunk of a
		// stream: take them over from the previous chunk)
		if (chunk_resumes(st, K.U)) {
			if (lane < 24) S->vw[lane] = make_double2(st->vw[lane][0], st->vw[lane][1]);
			h.lcg = st->lcg; h.noise_x1 = st->noise_x1;
			h.low = st->table_low;
		} else {
			if (lane < 24) S->vw[lane] = make_double2(0.0, 0.0);
			h.lcg = c_lcg_init; h.noise_x1 = 0.0;
			h.low = kNoLowMark;
		}
		h.c_p1 = h.c_p2 = h.c_p3 = h.c_p5 = h.c_p6 = __int_as_float(0x7fc00000);     // NaN: nothing cached
		__syncwarp();
	}
	const int nb = block_len(K, b3);
	const int buf = b3 & 1;
	const float (*cur)[kParamRow] = S->cur[buf];
	double ax = 0.0, ah1 = 0.0;
	{
		const bool live = lane < nb;
		const int col = lane < nb ? lane : 0;
		const float p1 = cur[1][col], p2 = cur[2][col], p3 = cur[3][col], p5 = cur[5][col], p6 = cur[6][col];
		// glottal, aspiration and frication amplitudes (VTMUtil.h:50-67).  A value is re-used while its parameter is
		// uniform over the block and unchanged since the previous block; all three are converted in ONE straight-line
		// region when any of them has to be (three independent exp10 chains overlap).
		double fa = 0.0;
		{
			const float q1 = __shfl_sync(0xffffffffu, p1, 0, 32), q2 = __shfl_sync(0xffffffffu, p2, 0, 32);
			const float q3 = __shfl_sync(0xffffffffu, p3, 0, 32);
			const float e1 = live ? p1 : q1, e2 = live ? p2 : q2, e3 = live ? p3 : q3;
			const bool u1 = __all_sync(0xffffffffu, e1 == q1), u2 = __all_sync(0xffffffffu, e2 == q2);
			const bool u3 = __all_sync(0xffffffffu, e3 == q3);
			if (u1 && u2 && u3 && q1 == h.c_p1 && q2 == h.c_p2 && q3 == h.c_p3) {
				ax = h.c_ax; ah1 = h.c_ah1; fa = h.c_fa;
			} else {
				ax = amp60((double) e1); ah1 = amp60((double) e2); fa = amp60((double) e3);
			}
			const float nan = __int_as_float(0x7fc00000);
			h.c_p1 = u1 ? q1 : nan; h.c_p2 = u2 ? q2 : nan; h.c_p3 = u3 ? q3 : nan;
			h.c_ax = ax; h.c_ah1 = ah1; h.c_fa = fa;
		}
		// frication taps (VocalTractModel0.h:524-552)
		const double fpos = (double) cur[4][col];
		int ip = (int) fpos;
		const double comp = fpos - ip;
		double ta = (1.0 - comp) * fa, tb = comp * fa;
		if (ip < 0 || ip > 7) { ta = 0.0; tb = 0.0; ip = -50; }
		// bandpass coefficients (BandpassFilter.h:91-110)
		double a2, a1, b0;
		{
			const float p50 = __shfl_sync(0xffffffffu, p5, 0, 32), p60 = __shfl_sync(0xffffffffu, p6, 0, 32);
			const float e5 = live ? p5 : p50, e6 = live ? p6 : p60;
			const bool uniform = __all_sync(0xffffffffu, e5 == p50 && e6 == p60);
			if (uniform && p50 == h.c_p5 && p60 == h.c_p6) {
				a2 = h.c_a2; a1 = h.c_a1; b0 = h.c_b0;
			} else {
				const double pi = 3.14159265358979323846;
#ifndef GTTS_EMU
				// a2 = (1 - tan x) / (1 + tan x) = (cos x - sin x) / (cos x + sin x): one division
				double sx, cx, sy, cv;
				gtts_sincos(pi * (double) e6 * V.Ts, sx, cx);
				gtts_sincos(2.0 * pi * (double) e5 * V.Ts, sy, cv);
				a2 = div_fast(cx - sx, cx + sx);
#else
				const double tv = tan(pi * (double) e6 * V.Ts);
				const double cv = cos(2.0 * pi * (double) e5 * V.Ts);
				a2 = (1.0 - tv) / (1.0 + tv);
#endif
				a1 = -(1.0 + a2) * cv;
				b0 = 0.5 - 0.5 * a2;
			}
			h.c_p5 = uniform ? p50 : __int_as_float(0x7fc00000);
			h.c_p6 = uniform ? p60 : __int_as_float(0x7fc00000);
			h.c_a2 = a2; h.c_a1 = a1; h.c_b0 = b0;
		}
		if (live) {
			S->tapa[buf][lane] = ta;
			S->tapb[buf][lane] = tb;
			S->ip[b3 % 3][lane] = (signed char) ip;
			S->bp[buf][2][lane] = a2;
			S->bp[buf][1][lane] = a1;
			S->bp[buf][0][lane] = b0;
		}
	}
	// noise (NoiseSource.h:40-44 as the integer LCG it is, NoiseFilter.h:63-68)
	double lp;
	{
		const unsigned long long sj = (h.lcg * h.mult) & ((1ull << 44) - 1);
		const double n = (double) (long long) sj * (1.0 / 17592186044416.0) - 0.5;
		double prev = shfl_d(n, (lane + 31) & 31, 32);
		if (lane == 0) prev = h.noise_x1;
		h.noise_x1 = shfl_d(n, nb - 1, 32);
		h.lcg = __shfl_sync(0xffffffffu, sj, nb - 1, 32);
		lp = n + prev;
	}
	// wavetable lookup of both half samples (WavetableGlottalSource.h:212-228).  With tn_min != tn_max the fall
	// segment [div1, div2) is a function of the current amplitude (setup(), :162-184) and is evaluated analytically;
	// entries at or above `low` read 0 (see rise_segment_scan; the table's own entries from div2 on are 0 anyway).
	{
		const bool dynamic = (V.waveform == 0) && (V.tn_delta != 0.0);
		double nd2 = 0.0, inv = 0.0;
		int low = h.low;
		if (dynamic) {
			nd2 = (double) V.div2 - rint(ax * V.tn_delta);
			nd2 = nd2 > 0.0 ? nd2 : 0.0;
			inv = 1.0 / (nd2 - (double) V.div1);
			const int mine = (lane < nb && nd2 < (double) V.div1) ? (int) nd2 : kNoLowMark;
			if (__any_sync(0xffffffffu, (mine != kNoLowMark) || (h.low != kNoLowMark))) {
				low = rise_segment_scan(mine, lane, h.low);
				h.low = __shfl_sync(0xffffffffu, low, nb - 1, 32);
			}
		}
		if (lane < nb) {
			const double* table = P.tables + (size_t) K.voice * kTableLen;
			double v[2];
#pragma unroll
			for (int s = 0; s < 2; ++s) {
				const double pos = S->pos[buf][s][lane];
				const unsigned lo = __double2uint_rz(pos);
				const unsigned up = (lo + 1 > 511u) ? lo + 1 - 512u : lo + 1;
				double tl, tu;
				if (dynamic && lo >= (unsigned) V.div1 && lo < (unsigned) V.div2) {
					const double x = (double) (int) (lo - V.div1) * inv;
					tl = (lo >= (unsigned) nd2) ? 0.0 : 1.0 - (x * x);
				} else {
					tl = ((int) lo >= low) ? 0.0 : table[lo & (kTableLen - 1)];   // the mask only acts on an absurd pitch (position beyond the table): no stray read
				}
				if (dynamic && up >= (unsigned) V.div1 && up < (unsigned) V.div2) {
					const double x = (double) (int) (up - V.div1) * inv;
					tu = (up >= (unsigned) nd2) ? 0.0 : 1.0 - (x * x);
				} else {
					tu = ((int) up >= low) ? 0.0 : table[up & (kTableLen - 1)];
				}
				v[s] = tl + ((pos - (double) lo) * (tu - tl));
			}
			// linear window: [0, 24) = the last 24 samples of the previous blocks, [24, 56) = this block
			S->vw[24 + lane] = make_double2(v[0], v[1]);
		}
	}
	__syncwarp();
	double firOut = 0.0;
	{
		// y = sum_i c[i] x2[2n+1-i], i ascending: taps 2m and 2m+1 read the odd and the even phase of sample n - m
		// (m = 0..23), tap 48 the odd phase of n - 24: one 16-byte load per pair of taps.
		const double2* pw = S->vw + 24 + lane;
#ifndef GTTS_EMU
		// two accumulator chains (odd-phase taps | even-phase taps): 25 dependent multiply-adds instead of 49
		double accO = 0.0, accE = 0.0;
#pragma unroll 4
		for (int m = 0; m < 24; ++m) {
			const double2 x = pw[-m];
			accO += x.y * c_fir[2 * m];
			accE += x.x * c_fir[2 * m + 1];
		}
		accO += pw[-24].y * c_fir[48];
		firOut = accO + accE;
#else
		// CPU emulation (tests/simt_emu): the reference's single ascending sum, so that the emulated kernel stays
		// bit-identical to the oracle
		double acc = 0.0;
		for (int m = 0; m < 24; ++m) {
			acc += pw[-m].y * c_fir[2 * m];
			acc += pw[-m].x * c_fir[2 * m + 1];
		}
		acc += pw[-24].y * c_fir[48];
		firOut = acc;
#endif
	}
	__syncwarp();
	if (lane < 24) S->vw[lane] = S->vw[32 + lane];      // slide the window: source [32, 56) and destination [0, 24) do not overlap
	if (lane < nb) {
		double pulse = firOut;
		const double pn = lp * pulse;
		pulse = ax * ((pulse * V.one_minus_breath) + (pn * V.breath));
		double sig;
		if (V.modulation) {
			double cm = ax * V.crossmix;
			cm = (cm < 1.0) ? cm : 1.0;
			sig = (pn * cm) + (lp * (1.0 - cm));
		} else {
			sig = lp;
		}
		S->sig[buf][lane] = sig;
		S->in[b3 % 3][lane] = (pulse + (ah1 * sig)) * 0.125;
		S->thr[b3 & 3][lane] = pulse * 0.125;
	}
	if (st != nullptr && b3 == K.nblocks - 1) {
		// end of a chunk (whole blocks): what the next chunk starts from
		__syncwarp();
		if (lane < 24) { st->vw[lane][0] = S->vw[lane].x; st->vw[lane][1] = S->vw[lane].y; }
		if (lane == 0) { st->lcg = h.lcg; st->noise_x1 = h.noise_x1; st->table_low = h.low; st->started = 1; }
	}
}

// ---- coefficient worker: junction coefficients of block it - 4 (VocalTractModel0.h:484-512, 698-716) ----------
// The ten per-sample coefficients are stored times the damping factor in kab[buf][v >> 1][sample].{x, y}, v being
// the index of the tube cell among the cells with a per-sample coefficient (kVarCell in tube_iteration):
//   v = 0..2 junctions S1|S2, S2|S3, S3|S4;  v = 3 the 3-way junction, (alpha - 1) d with alpha = alpha_left = alpha_right;
//   v = 4 S5|S6;  v = 5..7 S7|S8, S8|S9, S9|S10;  v = 8 mouth end, k8 b0 d;  v = 9 velum | N1.
GTTS_DEV void coef_task(SlotSm* S, int lane, int p)
{
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - kStCoef;
	if (K.it < 0 || b < 0 || b >= K.nblocks) return;
	const VoiceDev& V = S->V[K.vbuf];
	const int nb = block_len(K, b);
	const float (*scr)[kParamRow] = S->pscr[b & 1];
	// Radii and velum unchanged since the last sample of the previous block (a held posture): the coefficients are
	// the previous block's last row, copied instead of recomputed (9 divisions saved).
	const int buf = b & 1;
	bool same = lane >= nb;
	if (lane < nb) {
		same = true;
#pragma unroll
		for (int i = 0; i < 9; ++i) same = same && (scr[i][lane] == S->ckey[i]);
	}
	const bool reuse = b > 0 && __all_sync(0xffffffffu, same);
	__syncwarp();
	if (lane == kBlock - 1) {
#pragma unroll
		for (int i = 0; i < 9; ++i) S->ckey[i] = (nb == kBlock) ? scr[i][lane] : __int_as_float(0x7fc00000);
	}
	if (reuse) {
		if (lane < nb) {
#pragma unroll
			for (int r = 0; r < 5; ++r) S->kab[buf][r][lane] = S->kab[buf ^ 1][r][kBlock - 1];
			S->onepk7[b % 3][lane] = S->onepk7[(b + 2) % 3][kBlock - 1];
		}
	} else if (lane < nb) {
		// Nine scattering coefficients k = (a - b) / (a + b) in ONE rolled loop (instruction footprint): i = 0..6 oral
		// junctions r_i | r_{i+1}, i = 7 mouth r_8 | aperture, i = 8 velum | first nasal section.  Radii as in
		// setAllParameters: max(r * coef, 0.01).
		double* kabFlat = &S->kab[buf][0][0].x;         // [row][sample]{x, y} -> (row * kRow + sample) * 2 + component
		const double vel = (double) scr[8][lane];
		const double v2 = vel * vel;
		const double dmp = V.damping;
		double a2, r2_3 = 0.0, k7 = 0.0;
		{
			double r = (double) scr[0][lane] * V.radius_coef[0];
			r = r > 0.01 ? r : 0.01;
			a2 = r * r;
		}
#pragma unroll 1
		for (int i = 0; i < 9; ++i) {
			double b2;
			if (i < 7) {
				double r = (double) scr[i + 1][lane] * V.radius_coef[i + 1];
				r = r > 0.01 ? r : 0.01;
				b2 = r * r;
			} else if (i == 7) {
				b2 = V.ap2;
			} else {
				a2 = v2;
				b2 = V.nr1_2;
			}
			if (i == 3) r2_3 = a2;
			const double k = div_fast(a2 - b2, a2 + b2);
			if (i == 7) k7 = k;
			const int dst = i <= 2 ? i : i + 1;
			kabFlat[((dst >> 1) * kRow + lane) * 2 + (dst & 1)] = ((i == 7) ? k * V.refl_b0_m : k) * dmp;   // damping (and the mouth end's b0) folded in
			a2 = b2;
		}
		const double sum = div_fast(2.0, r2_3 + r2_3 + v2);
		S->kab[buf][1][lane].y = ((sum * r2_3) - 1.0) * dmp;   // alpha left == alpha right, stored as (alpha - 1) d
		S->onepk7[b % 3][lane] = 1.0 + k7;
	}
	__syncwarp();
}

// ---- output rows of the SRC stage ------------------------------------------------------------------------------
// Block b completes the outputs [e(b), e(b + 1)), e(b) = ceil(32 b 65536 / inc) (closed form of
// SampleRateConverter.h:295-361).  The SRC tasks write whole aligned rows instead: every block but the last two
// stops at the last row boundary (an absolute multiple of `gran` samples in the output buffer: 64 when the ratio is
// at least 2, else 32) and leaves the remainder to the next block, whose window still holds the inputs (the ring
// keeps 128: at most 27 + 26 + 32 back, 32 being written).  Aligned rows are what lets the kernel write straight into
// pinned host memory at PCIe rate.  The boundary between the last two blocks stays exact: the 26 flush zeros that
// chain B appends after the last block would otherwise reach, around the ring, the oldest inputs of deferred outputs.
GTTS_DEV void src_rows(SlotSm::Ctl& N, int b, long long e0, long long e1)
{
	const long long m = N.gran - 1;
	const long long a = N.U.out_begin & m;
	long long k0 = (b == 0 || b == N.nblocks - 2) ? e0 : ((e0 + a) & ~m) - a;     // e0 = 0 for a whole utterance, the outputs of the earlier chunks for a stream
	long long k1;
	if (b == N.nblocks - 1) {
		k1 = N.U.n_out;                             // flush: chain B appended the 26 zeros
	} else {
		k1 = (b == N.nblocks - 3) ? e1 : ((e1 + a) & ~m) - a;
		if (k1 > N.U.n_out) k1 = N.U.n_out;
	}
	if (k0 < 0) k0 = 0;
	if (k1 < k0) k1 = k0;
	N.src_k0 = k0;
	N.src_k1 = k1;
}

// ---- SRC: 64 outputs of NS slots (SampleRateConverter.h:295-361) -----------------------------------------------
// Lane l forms outputs k = kb + 2 l and k + 1.  Output k is centred on input e = (k inc) >> 16 with phase
// f = (k inc) & 0xFFFF: left wing taps h[L + 256 j] (+ deltaH interpolation) on x[e - 13 - j], right wing taps (from
// ~f) on x[e - 12 + j], j = 0..12.  The window x[e - 25 .. e + 1] of the first output (27 samples) also holds the
// 26 samples of the second one, whose centre is e or e + 1 (ratio >= 1): the window is loaded ONCE per slot, position
// by position (12 down to 0, then 13 up to 26: the reference's order of summation for the first output and, up to
// the place of one term, for the second), and every sample feeds both accumulators.  The coefficient a position has
// for an output comes from the table with a per-lane pointer; the tap one past either wing reads the zero padding
// behind the table.  With NS = 7 all slots are in lockstep (same rate, same block): the 26 + 26 interpolated
// coefficients are formed once for the seven windows.
template<int NS>
GTTS_DEV void src_task(CtaSm* C, const KernelParamsV2& P, int lane, int slot0, int mask, unsigned inc,
			long long kb, long long k0w, long long k1, int p)
{
	const long long ka = kb + 2 * lane;
	// outputs before the utterance's first one (partial first row of an unaligned utterance) are computed as output 0
	// and not stored
	const unsigned long long ta = (unsigned long long) (ka < 0 ? 0 : ka) * inc;
	const unsigned long long tb = (unsigned long long) (ka + 1 < 0 ? 0 : ka + 1) * inc;
	const int ea = (int) (ta >> 16);
	const int dl = (int) (tb >> 16) - ea;                       // 0 or 1
	const unsigned fa = (unsigned) (ta & 0xFFFFu), fb = (unsigned) (tb & 0xFFFFu);
	const unsigned ga = (~fa) & 0xFFFFu, gb = (~fb) & 0xFFFFu;
	const double iLa = (double) (fa & 0xFFu) / 256, iRa = (double) (ga & 0xFFu) / 256;
	const double iLb = (double) (fb & 0xFFu) / 256, iRb = (double) (gb & 0xFFu) / 256;
	const double2* pLa = C->tab + (fa >> 8);
	const double2* pRa = C->tab + (ga >> 8);
	const double2* pLb = C->tab + (fb >> 8) + 256 * dl;         // position 12 - t holds tap t + dl of the second output
	const double2* pRb = C->tab + (gb >> 8);
	const double* xw = C->slot[slot0].xring + ((ea - 25) & (kSrcRing - 1));
	constexpr int kStride = (int) (sizeof(SlotSm) / sizeof(double));
	double accA[NS], accB[NS];
#pragma unroll
	for (int q = 0; q < NS; ++q) { accA[q] = 0.0; accB[q] = 0.0; }
	// one slot: the loop is a chain of load -> coefficient -> accumulate; unrolled, the loads of several positions are
	// in flight together (with several slots the accumulators of one position are independent work enough)
	constexpr int kUnroll = NS == 1 ? GTTS_SRC1_UNROLL : 1;
	// left wings: window positions 12 .. 0
#pragma unroll kUnroll
	for (int t = 0; t < kSrcZeroCrossings; ++t) {
		const double2 ca = pLa[256 * t], cb = pLb[256 * t];
		const double cca = ca.x + (ca.y * iLa), ccb = cb.x + (cb.y * iLb);
#pragma unroll
		for (int q = 0; q < NS; ++q) {
			const double x = xw[q * kStride + 12 - t];
			accA[q] += (x * cca);
			accB[q] += (x * ccb);
		}
	}
	// position 13: first tap of the right wing of the first output; of the second one too unless its centre moved on
	{
		const double2 ca = pRa[0], cb = dl ? C->tab[fb >> 8] : pRb[0];
		const double cca = ca.x + (ca.y * iRa), ccb = cb.x + (cb.y * (dl ? iLb : iRb));
#pragma unroll
		for (int q = 0; q < NS; ++q) {
			const double x = xw[q * kStride + 13];
			accA[q] += (x * cca);
			accB[q] += (x * ccb);
		}
	}
	// right wings: window positions 14 .. 26
	pRb -= 256 * dl;
#pragma unroll kUnroll
	for (int t = 1; t <= kSrcZeroCrossings; ++t) {
		const double2 ca = pRa[256 * t], cb = pRb[256 * t];
		const double cca = ca.x + (ca.y * iRa), ccb = cb.x + (cb.y * iRb);
#pragma unroll
		for (int q = 0; q < NS; ++q) {
			const double x = xw[q * kStride + 13 + t];
			accA[q] += (x * cca);
			accB[q] += (x * ccb);
		}
	}
	const bool okA = ka >= k0w && ka < k1, okB = ka + 1 >= k0w && ka + 1 < k1;
	if (okA || okB) {
#pragma unroll
		for (int q = 0; q < NS; ++q) {
			if ((mask >> q) & 1) {
				float* o = P.out + C->slot[slot0 + q].ctl[p].U.out_begin + ka;
				if (okA && okB && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
					*reinterpret_cast<float2*>(o) = make_float2((float) accA[q], (float) accB[q]);
				} else {
					if (okA) o[0] = (float) accA[q];
					if (okB) o[1] = (float) accB[q];
				}
			}
		}
	}
}

// Shared SRC tasks (all slots in the SRC stage in lockstep): the slots are split into two groups, one per SRC warp,
// each of which forms the coefficients for its NS windows (`keep`: the slots of the group this warp stores).
template<int NS>
GTTS_DEV void src_shared_group(CtaSm* C, const KernelParamsV2& P, int lane, int slot0, int keep, int p)
{
	const CtaSm::Sched& D = C->sched[p];
	if (!D.src_shared) return;
	const int mask = (D.src_mask >> slot0) & keep;
	if (mask == 0) return;
	int ref = 0;
	while (!((D.src_mask >> ref) & 1)) ++ref;
	const unsigned inc = C->slot[ref].ctl[p].inc;
#pragma unroll 1
	for (int task = 0; task < D.src_tasks; ++task) src_task<NS>(C, P, lane, slot0, mask, inc, D.src_k0 + 64ll * task, D.src_k0w, D.src_k1, p);
}

// Down-sampling SRC (internal rate above the output rate: vocal tracts shorter than 7.3 cm at 48 kHz), lane = output
// (SampleRateConverter.h:362-415).  Output k is centred on ring position e = (k inc) >> 16 -- input i sits at ring
// position pad + i -- with phase rint(f ratio), f = (k inc) & 0xFFFF; the left wing walks down from position e, the
// right wing (phase from ~f) up from e + 1, each while its filter index phase >> 8 is inside the table, the phase
// advancing by phaseIncrement = rint(ratio 65536) per tap: at most pad taps a wing, in the reference's order.
GTTS_DEV_NOINLINE void src_down_task(CtaSm* C, const KernelParamsV2& P, int lane, int slot, int p)
{
	SlotSm* S = &C->slot[slot];
	const SlotSm::Ctl& K = S->ctl[p];
	const VoiceDev& V = S->V[K.vbuf];
	const unsigned inc = K.inc, pinc = V.src_phase_inc;
	const int pad = V.src_pad;
	const double ratio = V.src_ratio;
	float* out = P.out + K.U.out_begin;
	// Inputs at and after nEnd are the flush zeros (SampleRateConverter.h:462-471).  They are not written into the
	// ring here: 2 pad of them (up to 64) behind the last block would wrap onto inputs the SRC warp may still be
	// reading for the block before (window 2 pad + 32, chain B one block ahead: 128 entries hold that and no more).
	const long long nEnd = K.U.n_in_base + K.U.n_internal;
#pragma unroll 1
	for (long long k = K.src_k0 + lane; k < K.src_k1; k += 32) {
		const unsigned long long t = (unsigned long long) k * inc;
		const long long e = (long long) (t >> 16);
		const unsigned f = (unsigned) (t & 0xFFFFu);
		double acc = 0.0;
		unsigned ph = (unsigned) rint((double) f * ratio);
		long long pos = e - pad;                                   // input index of ring position e
		unsigned ii;
#pragma unroll 1
		while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
			const double2 c = C->tab[ii];
			const double imp = c.x + (c.y * ((double) (ph & 0xFFu) / 256));
			acc += (pos < nEnd ? S->xring[(int) (pos & (kSrcRing - 1))] : 0.0) * imp;
			pos -= 1;
			ph += pinc;
		}
		ph = (unsigned) rint((double) ((~f) & 0xFFFFu) * ratio);
		pos = e - pad + 1;
#pragma unroll 1
		while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
			const double2 c = C->tab[ii];
			const double imp = c.x + (c.y * ((double) (ph & 0xFFu) / 256));
			acc += (pos < nEnd ? S->xring[(int) (pos & (kSrcRing - 1))] : 0.0) * imp;
			pos += 1;
			ph += pinc;
		}
		out[k] = (float) acc;
	}
}

// One output per lane, 32 consecutive outputs = one 128-byte row per call: the unit in which the per-slot SRC (slots
// not in lockstep: mixed voices, ragged lengths -- BASELINE config 3) is shared out between three warps.  Same taps in
// the same order as the reference (left wing j = 0..12 on x[e - 13 - j], then the right wing on x[e - 12 + j]).
GTTS_DEV void src_task_row(CtaSm* C, const KernelParamsV2& P, int lane, int slot, unsigned inc, long long kb, long long k0w, long long k1, int p)
{
	const long long ka = kb + lane;
	const unsigned long long ta = (unsigned long long) (ka < 0 ? 0 : ka) * inc;
	const int ea = (int) (ta >> 16);
	const unsigned fa = (unsigned) (ta & 0xFFFFu), ga = (~fa) & 0xFFFFu;
	const double iL = (double) (fa & 0xFFu) / 256, iR = (double) (ga & 0xFFu) / 256;
	const double2* pL = C->tab + (fa >> 8);
	const double2* pR = C->tab + (ga >> 8);
	const double* xw = C->slot[slot].xring + ((ea - 25) & (kSrcRing - 1));
	double acc = 0.0;
	constexpr int kUnroll = GTTS_SRC1_UNROLL;
#pragma unroll kUnroll
	for (int t = 0; t < kSrcZeroCrossings; ++t) {
		const double2 c = pL[256 * t];
		acc += (xw[12 - t] * (c.x + (c.y * iL)));
	}
#pragma unroll kUnroll
	for (int t = 0; t < kSrcZeroCrossings; ++t) {
		const double2 c = pR[256 * t];
		acc += (xw[13 + t] * (c.x + (c.y * iR)));
	}
	if (ka >= k0w && ka < k1) P.out[C->slot[slot].ctl[p].U.out_begin + ka] = (float) acc;
}

// Per-slot SRC (slots not in lockstep): the slot's outputs of this block in rows of 32, dealt round-robin -- rotating
// with the block, so that the shares even out -- to the slot's three workers: 0 its coefficient worker, 1 its helper,
// 2 one of the SRC warps.  (Measured on the config-3 shape, where all of it ran on five warps: those five were the last
// to arrive in every iteration, 9.3-9.8 k busy cycles against 5.3-5.8 k of the other coefficient workers and helpers.)
// The down-sampling converter (variable tap count) stays whole on worker 0.
GTTS_DEV void src_slot_task(CtaSm* C, const KernelParamsV2& P, int lane, int slot, int p, int part)
{
	const SlotSm::Ctl& K = C->slot[slot].ctl[p];
	const int b = K.it - kStOut;
	if (K.it < 0 || b < 0 || b >= K.nblocks) return;
	if (K.inc > 65536u) { if (part == 0) src_down_task(C, P, lane, slot, p); return; }
	const long long k0 = K.src_k0, k1 = K.src_k1;
	const long long a = K.U.out_begin & 31;
	const long long kStart = ((k0 + a) & ~31ll) - a;           // rows start on multiples of 32 samples of the output buffer
	int t = (part + 3 - b % 3) % 3;                            // this worker's first row: (row + b) % 3 == part
#pragma unroll 1
	for (long long kb = kStart + 32ll * t; kb < k1; kb += 96) src_task_row(C, P, lane, slot, K.inc, kb, k0, k1, p);
}

// ---- chain A: oscillator phase of block it - 2, lane = slot (WavetableGlottalSource.h:196-199, 265-272) --------
// Two half-sample increments per sample, wrap above 511; 42 dependent cycles per sample.  Lanes whose slot has no
// such block run on dummy data (their results are never read), which keeps the loop free of divergent branches.
template<bool ST>
GTTS_DEV void chain_a_iteration(CtaSm* C, const KernelParamsV2& P, int lane, double& posReg, int p)
{
	SlotSm* S = &C->slot[lane < kSlots ? lane : 0];
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - kStPhase;
	if (ST && !__any_sync(0xffffffffu, lane < kSlots && K.it >= 0 && b >= 0 && b < K.nblocks)) return;
	if (lane >= kSlots) return;
	UttStateV2* st = chunk_state<ST>(P, K.U);
	if (b == 0) posReg = chunk_resumes(st, K.U) ? st->pos : 0.0;
	const int buf = b & 1;
	const double2* osc = reinterpret_cast<const double2*>(S->osc[buf]);
	double2* p0 = reinterpret_cast<double2*>(S->pos[buf][0]);
	double2* p1 = reinterpret_cast<double2*>(S->pos[buf][1]);
	double pos = posReg;
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock / 2; j0 += 2) {
		const double2 ia = osc[j0], ib = osc[j0 + 1];
		const double inc[4] = {ia.x, ia.y, ib.x, ib.y};
		double o0[4], o1[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			double s = pos + inc[q];
			pos = (s > 511.0) ? s - 512.0 : s;
			o0[q] = pos;
			s = pos + inc[q];
			pos = (s > 511.0) ? s - 512.0 : s;
			o1[q] = pos;
		}
		p0[j0] = make_double2(o0[0], o0[1]); p0[j0 + 1] = make_double2(o0[2], o0[3]);
		p1[j0] = make_double2(o1[0], o1[1]); p1[j0 + 1] = make_double2(o1[2], o1[3]);
	}
	posReg = pos;
	if (st != nullptr && K.it >= 0 && b == K.nblocks - 1) st->pos = pos;
}

// ---- chain A2: frication bandpass of block it - 4 and the two tap signals, lane = slot (BandpassFilter.h:114-122)
template<bool ST>
GTTS_DEV void chain_a2_iteration(CtaSm* C, const KernelParamsV2& P, int lane, BandpassState& st, int p)
{
	SlotSm* S = &C->slot[lane < kSlots ? lane : 0];
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - kStCoef;
	if (ST && !__any_sync(0xffffffffu, lane < kSlots && K.it >= 0 && b >= 0 && b < K.nblocks)) return;
	if (lane >= kSlots) return;
	UttStateV2* cs = chunk_state<ST>(P, K.U);
	if (b == 0) {
		if (chunk_resumes(cs, K.U)) { st.x1 = cs->bp[0]; st.x2 = cs->bp[1]; st.y1 = cs->bp[2]; st.y2 = cs->bp[3]; }
		else { st.x1 = st.x2 = st.y1 = st.y2 = 0.0; }
	}
	const int buf = b & 1;
	const double2* sig = reinterpret_cast<const double2*>(S->sig[buf]);
	const double2* c0 = reinterpret_cast<const double2*>(S->bp[buf][0]);
	const double2* c1 = reinterpret_cast<const double2*>(S->bp[buf][1]);
	const double2* c2 = reinterpret_cast<const double2*>(S->bp[buf][2]);
	const double2* ta = reinterpret_cast<const double2*>(S->tapa[buf]);
	const double2* tb = reinterpret_cast<const double2*>(S->tapb[buf]);
	double2* pab = S->pab[buf];
	double x1 = st.x1, x2 = st.x2, y1 = st.y1, y2 = st.y2;
	bool anyFric = false;
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock / 2; ++j0) {
		const double2 xv = sig[j0], b0v = c0[j0], a1v = c1[j0], a2v = c2[j0], tav = ta[j0], tbv = tb[j0];
		const double x[2] = {xv.x, xv.y}, b0[2] = {b0v.x, b0v.y}, a1[2] = {a1v.x, a1v.y}, a2[2] = {a2v.x, a2v.y};
		const double tA[2] = {tav.x, tav.y}, tB[2] = {tbv.x, tbv.y};
		double oa[2], ob[2];
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const double y = b0[q] * (x[q] - x2) - a1[q] * y1 - a2[q] * y2;
			x2 = x1; x1 = x[q]; y2 = y1; y1 = y;
			oa[q] = tA[q] * y;
			ob[q] = tB[q] * y;
			anyFric = anyFric || oa[q] != 0.0 || ob[q] != 0.0;
		}
		pab[2 * j0] = make_double2(oa[0], ob[0]);
		pab[2 * j0 + 1] = make_double2(oa[1], ob[1]);
	}
	S->fric[buf] = anyFric ? 1 : 0;
	st.x1 = x1; st.x2 = x2; st.y1 = y1; st.y2 = y2;
	if (cs != nullptr && K.it >= 0 && b == K.nblocks - 1) { cs->bp[0] = x1; cs->bp[1] = x2; cs->bp[2] = y1; cs->bp[3] = y2; }
}

// ---- chain B: radiation filters + throat low-pass of block it - 6, lane = slot * 4 + filter ----------------------
// One code path for the three one-pole filters: y = b0 x + b1 x1 - a1 y1, out = y * gain
//   f = 0 mouth radiation (b0 = A, b1 = a1 = -A, gain 1) on (1 + k7) T[S10]     (RadiationFilter.h:73-79)
//   f = 1 nose radiation on (1 + nk5) NT[N6]
//   f = 2 throat low-pass (b0, b1 = 0, a1, gain = throat gain) on pulse * 0.125  (Throat.h:80-85)
// then the output sum (lane = sample) into the SRC ring, which this role also clears for a new utterance.
struct ChainBRegs { double x1, y1; };

template<bool ST>
GTTS_DEV void chain_b_iteration(CtaSm* C, const KernelParamsV2& P, int lane, ChainBRegs& r, int p)
{
	const int s = lane >> 2, f = lane & 3;
	{
		const SlotSm::Ctl& K0 = C->slot[s < kSlots ? s : 0].ctl[p];
		const int b0 = K0.it - kStRad;
		if (ST && !__any_sync(0xffffffffu, s < kSlots && K0.it >= 0 && b0 >= 0 && b0 < K0.nblocks)) return;
	}
	if (s < kSlots && f < 3) {
		SlotSm* S = &C->slot[s];
		const SlotSm::Ctl& K = S->ctl[p];
		const int b = K.it - kStRad;
		const VoiceDev& V = S->V[K.vbuf];
		UttStateV2* st = chunk_state<ST>(P, K.U);
		if (b == 0) {
			if (chunk_resumes(st, K.U)) { r.x1 = st->rad[f][0]; r.y1 = st->rad[f][1]; }
			else { r.x1 = 0.0; r.y1 = 0.0; }
		}
		const double b0 = f == 0 ? V.rad_m : (f == 1 ? V.rad_n : V.throat_b0);
		const double b1 = f == 0 ? -V.rad_m : (f == 1 ? -V.rad_n : 0.0);
		const double a1 = f == 0 ? -V.rad_m : (f == 1 ? -V.rad_n : V.throat_a1);
		const double gain = f == 2 ? V.throat_gain : 1.0;
		const double onePlusN = 1.0 + V.nasal_k[5];
		const double2* in = reinterpret_cast<const double2*>(f == 0 ? S->endm[b & 1] : (f == 1 ? S->endn[b & 1] : S->thr[b & 3]));
		const double2* scale = reinterpret_cast<const double2*>(S->onepk7[(b % 3 + 3) % 3]);
		double2* out = reinterpret_cast<double2*>(S->rad[f]);
		double x1 = r.x1, y1 = r.y1;
#pragma unroll 1
		for (int j0 = 0; j0 < kBlock / 2; j0 += 2) {
			const double2 ra = in[j0], rb = in[j0 + 1], sa = scale[j0], sb = scale[j0 + 1];
			const double raw[4] = {ra.x, ra.y, rb.x, rb.y}, sc[4] = {sa.x, sa.y, sb.x, sb.y};
			double o[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const double m = f == 0 ? sc[q] : onePlusN;
				const double x = f == 2 ? raw[q] : m * raw[q];
				const double y = b0 * x + b1 * x1 - a1 * y1;
				x1 = x;
				y1 = y;
				o[q] = y * gain;
			}
			out[j0] = make_double2(o[0], o[1]);
			out[j0 + 1] = make_double2(o[2], o[3]);
		}
		r.x1 = x1; r.y1 = y1;
		if (st != nullptr && K.it >= 0 && b == K.nblocks - 1) { st->rad[f][0] = x1; st->rad[f][1] = y1; }
	}
	__syncwarp();
	// output sum, lane = sample: (mouth + nose) + throat (VocalTractModel0.h:657-660, 441)
#pragma unroll 1
	for (int q = 0; q < kSlots; ++q) {
		SlotSm* Q = &C->slot[q];
		const SlotSm::Ctl& QK = Q->ctl[p];
		const int qb = QK.it - kStRad;
		if (QK.it < 0 || qb < 0 || qb >= QK.nblocks) continue;
		UttStateV2* qs = chunk_state<ST>(P, QK.U);
		if (qb == 0) {
			// new utterance: inputs before the first one are zero (SampleRateConverter.h:98-115); a chunk of a stream:
			// the ring as the previous chunk left it
			const bool resumes = chunk_resumes(qs, QK.U);
			for (int i = lane; i < kXr; i += 32) Q->xring[i] = resumes ? qs->xring[i & (kSrcRing - 1)] : 0.0;
			__syncwarp();
		}
		const int qn = block_len(QK, qb);
		const long long n0 = (ST ? QK.U.n_in_base : 0) + (long long) qb * kBlock;
		if (lane < qn) {
			const int idx = (int) ((n0 + lane) & (kSrcRing - 1));
			const double v = (Q->rad[0][lane] + Q->rad[1][lane]) + Q->rad[2][lane];
			Q->xring[idx] = v;
			if (idx < kXr - kSrcRing) Q->xring[idx + kSrcRing] = v;
		}
		if (qb == QK.nblocks - 1 && (QK.U.flags & 2) == 0) {
			// flushBuffer(): 2 pad zeros after the last input (26 when up-sampling; SampleRateConverter.h:462-471);
			// not between the chunks of a stream
			// (up-sampling only: the down-sampling SRC takes inputs past the end as zero, see src_down_task)
			const int nz = Q->V[QK.vbuf].src_upsample ? 2 * Q->V[QK.vbuf].src_pad : 0;
			for (int i = lane; i < nz; i += 32) {
				const int idx = (int) (((ST ? QK.U.n_in_base : 0) + QK.U.n_internal + i) & (kSrcRing - 1));
				Q->xring[idx] = 0.0;
				if (idx < kXr - kSrcRing) Q->xring[idx + kSrcRing] = 0.0;
			}
		}
		if (qs != nullptr && qb == QK.nblocks - 1) {
			__syncwarp();
			for (int i = lane; i < kSrcRing; i += 32) qs->xring[i] = Q->xring[i];
		}
	}
}

// ---- tube warps: block it - 5, one cell per lane, 16 lanes per utterance, two utterances per warp --------------
// Cells: u = 0..9 oral S1..S10 (u = 3: the 3-way junction after S4, u = 9: mouth end), u = 10..15 nasal N1..N6
// (u = 15: nose end).  The forward wave goes u -> u + 1, the backward wave u + 1 -> u, the velum branch links
// u = 3 <-> u = 10: three 64-bit shuffles per sample.  A single warp issues in order and every FP64 instruction costs
// it two cycles whether the result is needed soon or not (tools/microbench/tube_chain.cu), so the step is written for
// the fewest FP64 operations (12) and the shortest chain behind the incoming wave T (two operations).  The three
// kinds of cell are ONE formula with per-lane constants; kd = k d is what the coefficient worker stores:
//        ts = T + sigma Bn + tau nb       Tout = kd ts + (d T + tap)       Bout = kd ts + cB B*       Lout = kd ts + (m3 d T + E)
//      2-port junction  sigma = -1, tau = 0, cB = d, B* = Bn: dl = k (T - Bn), Tout = (T + dl) d + tap,
//                       Bout = (Bn + dl) d; Lout == Bout (m3 = 0, E = cB B*)                   (VocalTractModel0.h:575-600)
//      3-way junction   k = alpha - 1 (alpha = alpha_left = alpha_right), alpha_u = 2 - 2 alpha = -2 k, hence
//                       sigma = +1, tau = -2: kd ts = d (jp - T - Bn) with jp = alpha (T + Bn) + alpha_u nb;
//                       Tout = (jp - Bn) d + tap, Bout = (jp - T) d, and the wave into the nose
//                       Lout = (jp - nb) d = kd ts + d T + d (Bn - nb)  (m3 = 1, E = d (Bn - nb))              (:602-617)
//      open end         sigma = tau = 0, kd = b0 k_end d, cB = -a1, B* = the lane's own Bout of the previous sample
//                       (it shuffles from itself): Bout / d = b0 (k T) - a1 y1 is the reflection low-pass, y1 = B* / d (:619-630)
// The incoming wave of the next sample: T = from the previous cell | the link (N1) | d B[S1] + glottal input (S1).  Cells whose coefficient is a per-voice constant (S6|S7: 0, the nasal
// junctions N1|N2 .. N5|N6 and the nose end) keep it in a register; the other ten read it per sample (kab row v >> 1,
// component v & 1).  Lanes of slots without a block at this stage run on dummy data: their state is reset when
// their block 0 arrives.
struct TubeCell { double T, Bn, nb, last; };

template<bool ST>
GTTS_DEV void tube_iteration(CtaSm* C, const KernelParamsV2& P, int warp, int lane, TubeCell& t, int p)
{
	const int u = lane >> 1;
	const int sbit = lane & 1;
	const int slot = warp + kTubeWarps * sbit;         // slots w and w + 4: their rows are 64 bytes (16 banks) apart modulo 128
	SlotSm* S = &C->slot[slot < kSlots ? slot : 0];
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = (slot < kSlots) ? K.it - kStTube : -1;
	const bool hasBlock = slot < kSlots && K.it >= 0 && b >= 0 && b < K.nblocks;
	const int buf = b & 1, b3 = (b % 3 + 3) % 3;
	if (ST && !__any_sync(0xffffffffu, hasBlock)) return;    // pipeline filling or draining (every iteration of a short stream chunk but a few)
	const bool fricBlock = __any_sync(0xffffffffu, hasBlock && S->fric[buf] != 0);
	const VoiceDev& V = S->V[K.vbuf];
	UttStateV2* st = hasBlock ? chunk_state<ST>(P, K.U) : nullptr;
	if (b == 0) {
		if (hasBlock && chunk_resumes(st, K.U)) { t.T = st->tube[u][0]; t.Bn = st->tube[u][1]; t.nb = st->tube[u][2]; t.last = st->tube[u][3]; }
		else { t.T = t.Bn = t.nb = t.last = 0.0; }
	}
	const double d = V.damping;
	const bool is3 = u == 3, isEnd = (u == 9) || (u == 15), isGlot = u == 0, isN1 = u == 10;
	const bool storesEnd = isEnd && slot < kSlots;     // the second group of the last warp is a dummy: it must not store
	// per-lane constants of the one formula (see above)
	const double sigma = is3 ? 1.0 : (isEnd ? 0.0 : -1.0);
	const double tau = is3 ? -2.0 : 0.0;               // alpha_u d = (2 - 2 alpha) d = -2 k: folded into ts
	const double cB = isEnd ? -(u == 9 ? V.refl_a1_m : V.refl_a1_n) : d;
	const double m3d = is3 ? d : 0.0;
	const double mG = isGlot ? d : 0.0;
	const double mP = (isGlot || isN1) ? 0.0 : 1.0;
	const int tap = (u >= 1 && u <= 8) ? u - 1 : -100;
	// index among the cells with a per-sample coefficient: u = 0..4 -> 0..4, 6..10 -> 5..9; -1: constant
	const int var = u <= 4 ? u : (u >= 6 && u <= 10 ? u - 1 : -1);
	const bool isVar = var >= 0;
	const double kConst = (u == 5) ? 0.0 : ((u == 15 ? V.nasal_k[5] * V.refl_b0_n : V.nasal_k[u >= 11 ? u - 10 : 1]) * d);
	const double* kRowp = &S->kab[buf][isVar ? var >> 1 : 0][0].x + (var & 1);     // stride 2 doubles per sample
	const double2* pabRow = S->pab[buf];
	const double* inRow = S->in[b3];                   // read by lane u = 0 only
	const signed char* ipRow = S->ip[b3];
	double* endRow = (u == 15) ? S->endn[buf] : S->endm[buf];
	// an end cell has no next cell: it takes its own backward output instead, so that Bn is its filter state y1 d
	const int srcPrev = 2 * ((u + 15) & 15) + sbit, srcNext = isEnd ? lane : 2 * ((u + 1) & 15) + sbit, srcLink = 2 * (is3 ? 10 : 3) + sbit;
	// T = forward wave into the cell, Bn = backward wave from the next cell (end cells: own previous backward output),
	// nb = wave on the velum link, last = the cell's own backward output of the previous sample (glottis reflection)
	double T = t.T, Bn = t.Bn, nb = t.nb, last = t.last;
	double inv[4] = {0.0, 0.0, 0.0, 0.0};              // glottal input (lane u = 0): zero on every other lane
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock; j0 += 4) {
		double kv[4], tf[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			kv[q] = kConst;
			if (isVar) kv[q] = kRowp[2 * (j0 + q)];
			if (isGlot) inv[q] = inRow[j0 + q];
			tf[q] = 0.0;
		}
		if (fricBlock) {
			// frication injected at taps ip, ip + 1 (the reference adds tap * 0 = 0 everywhere else)
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const double2 pab = pabRow[j0 + q];
				const int ip = ipRow[j0 + q];
				tf[q] = (tap == ip) ? pab.x : ((tap == ip + 1) ? pab.y : 0.0);
			}
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			// before T is known: everything that depends only on the neighbours' waves and the lane's own state
			const double w = (tau * nb) + (sigma * Bn);
			const double cBB = cB * Bn;                                  // d Bn | -a1 (y1 d) of the end filter
			const double el0 = cBB - (m3d * nb);                         // 3-way junction: d (Bn - nb)
			const double pre = (mG * last) + inv[q];                     // glottis: T[S1] = B[S1] d + input, B[S1] of the previous sample
			// behind T: two dependent operations per output
			const double ts = T + w;
			const double dTf = (d * T) + tf[q];
			const double el = (m3d * T) + el0;
			if (storesEnd) endRow[j0 + q] = T;
			const double Tout = (kv[q] * ts) + dTf;
			const double Bout = (kv[q] * ts) + cBB;
			const double Lout = (kv[q] * ts) + el;                       // 3-way junction: the wave into the nose; else == Bout
			const double fromPrev = shfl_d(Tout, srcPrev, 32);
			const double fromNext = shfl_d(Bout, srcNext, 32);
			const double link = shfl_d(Lout, srcLink, 32);
			last = Bout;
			T = (mP * fromPrev) + (isN1 ? link : pre);
			Bn = fromNext;
			nb = link;
		}
	}
	t.T = T; t.Bn = Bn; t.nb = nb; t.last = last;
	if (st != nullptr && b == K.nblocks - 1) { st->tube[u][0] = T; st->tube[u][1] = Bn; st->tube[u][2] = nb; st->tube[u][3] = last; }
}

// ---- slot bookkeeping for iteration i + 1 (chain A warp, lane = slot) -------------------------------------------
// Reads ctl[i % 4], writes ctl[(i + 1) % 4] and sched[(i + 1) % 4], copies the voice constants of new utterances.
template<bool ST>
GTTS_DEV void schedule_slots(CtaSm* C, const KernelParamsV2& P, int lane, int i, bool first)
{
	const int p = i & (kCtlRing - 1), q = (i + 1) & (kCtlRing - 1);
	int alive = 0, valid = 0, it = -1, fresh = 0, voice = 0, vbuf = 0;
	long long nInternal = 0;
	unsigned inc = 0;
	int phase = 0;
	if (lane < kSlots) {
		SlotSm* S = &C->slot[lane];
		const SlotSm::Ctl& K = S->ctl[p];
		SlotSm::Ctl& N = S->ctl[q];
		it = first ? -1 : K.it;
		if (it >= 0) {
			it += 1;
			if (it > K.nblocks - 1 + kStages) it = -1;          // every stage has seen every block
		}
		if (it >= 0) {
			N.U = K.U; N.nblocks = K.nblocks; N.voice = K.voice; N.it = it; N.last_len = K.last_len;
			N.vbuf = K.vbuf; N.gran = K.gran;
			N.eq = K.eq; N.erem = K.erem; N.eQ = K.eQ; N.eR = K.eR; N.inc = K.inc;
		} else {
			N.it = -1; N.nblocks = 0; N.voice = first ? 0 : K.voice; N.U = K.U;
			N.vbuf = first ? 0 : K.vbuf; N.gran = 32; N.inc = 0;
			for (;;) {
				const int u = atomicAdd(P.queue, 1);
				if (u >= P.n_utt) break;
				const UttDesc U = P.utts[P.order[u]];
				if (U.n_internal == 0) {
					// no input at all: finishSynthesis() alone converts the 26 flush zeros into zeros
#pragma unroll 1
					for (long long k = 0; k < U.n_out; ++k) P.out[U.out_begin + k] = 0.0f;
					continue;
				}
				N.U = U;
				N.voice = U.voice;
				N.vbuf = N.vbuf ^ 1;
				N.nblocks = (int) ((U.n_internal + kBlock - 1) / kBlock);
				N.last_len = (int) (U.n_internal - (long long) (N.nblocks - 1) * kBlock);
				N.it = 0;
				// output counter: e(0) = 0 = (0 + inc - 1) / inc, remainder inc - 1; one block adds 32 * 65536
				const unsigned uinc = P.voices[U.voice].src_inc;
				N.inc = uinc;
				N.gran = uinc <= 32768u ? 64 : (uinc <= 65536u ? 32 : 1);     // down-sampling: no row deferral (the window is up to 64 inputs long)
				N.eQ = (unsigned) (kBlock << 16) / uinc;
				N.eR = (unsigned) (kBlock << 16) % uinc;
				// e(0) = ceil(n_in_base 65536 / inc) as quotient / remainder of (n_in_base 65536 + inc - 1) / inc
				// (0 and inc - 1 for a whole utterance)
				if (ST) {
					const unsigned long long num = ((unsigned long long) U.n_in_base << 16) + uinc - 1;
					N.eq = (long long) (num / uinc);
					N.erem = (unsigned) (num % uinc);
				} else {
					N.eq = 0;
					N.erem = uinc - 1;
				}
				it = 0;
				fresh = 1; voice = U.voice; vbuf = N.vbuf;
				break;
			}
		}
		alive = it >= 0;
		valid = it >= 0 && it - kStOut >= 0 && it - kStOut < N.nblocks;
		if (valid) {
			nInternal = N.U.n_internal; inc = N.inc; phase = (int) (N.U.out_begin & (N.gran - 1));
			// the block entering the SRC stage: advance the output counter by one block, derive its rows
			const long long e0 = N.eq;
			unsigned rem = N.erem + N.eR;
			long long e1 = e0 + N.eQ;
			if (rem >= inc) { rem -= inc; e1 += 1; }
			N.eq = e1; N.erem = rem;
			src_rows(N, it - kStOut, e0, e1);
		}
	}
	// voice constants of the new utterances into the slots' other V buffer (the previous utterance of the slot may
	// still be in its last stages)
	unsigned freshMask = __ballot_sync(0xffffffffu, fresh);
	while (freshMask) {
		const int s = __ffs((int) freshMask) - 1;
		freshMask &= freshMask - 1;
		const int vs = __shfl_sync(0xffffffffu, voice, s, 32), bs = __shfl_sync(0xffffffffu, vbuf, s, 32);
		const double* src = reinterpret_cast<const double*>(&P.voices[vs]);
		double* dst = reinterpret_cast<double*>(&C->slot[s].V[bs]);
		for (int k = lane; k < (int) (sizeof(VoiceDev) / sizeof(double)); k += 32) dst[k] = src[k];
	}
	const unsigned any = __ballot_sync(0xffffffffu, alive);
	// SRC stage alignment: all slots that have a block at this stage are at the same block of equally long utterances
	// with the same SRC increment and row phase -> shared SRC tasks
	const unsigned vmask = __ballot_sync(0xffffffffu, valid);
	const int refLane = vmask ? __ffs((int) vmask) - 1 : 0;
	const int itRef = __shfl_sync(0xffffffffu, it, refLane, 32);
	const long long nRef = __shfl_sync(0xffffffffu, nInternal, refLane, 32);
	const unsigned incRef = __shfl_sync(0xffffffffu, inc, refLane, 32);
	const int phaseRef = __shfl_sync(0xffffffffu, phase, refLane, 32);
	const int same = !valid || (it == itRef && nInternal == nRef && inc == incRef && phase == phaseRef && inc <= 65536u);
	const bool allSame = __ballot_sync(0xffffffffu, same) == 0xffffffffu;
	__syncwarp();                                  // lane 0 reads what the reference slot's lane wrote into ctl[q]
	if (lane == 0) {
		CtaSm::Sched& D = C->sched[q];
		D.live = any != 0;
		D.src_mask = (int) vmask;
		D.src_shared = (allSame && __popc(vmask) >= 2) ? 1 : 0;
		D.src_tasks = 0;
		if (D.src_shared) {
			const SlotSm::Ctl& RK = C->slot[refLane].ctl[q];
			long long k0 = RK.src_k0;
			const long long k1 = RK.src_k1;
			const long long m = RK.gran - 1;
			const long long a = RK.U.out_begin & m;
			D.src_k0w = k0;
			k0 = ((k0 + a) & ~m) - a;
			D.src_k0 = k0;
			D.src_k1 = k1;
			D.src_tasks = (int) ((k1 - k0 + 63) / 64);
		}
	}
}

#ifdef GTTS_ROLE_PROFILE
#define GTTS_PROF_DECL long long profBusy = 0, profWait = 0, profIters = 0, profT0 = 0, profT1 = 0
#define GTTS_PROF_T0() (profT0 = clock64())
#define GTTS_PROF_T1() (profT1 = clock64(), profWait += profT1 - profT0)
#define GTTS_PROF_T2() (profBusy += clock64() - profT1, profIters += 1)
#define GTTS_PROF_STORE(P, role, lane) \
	do { if ((P).prof != nullptr && (lane) == 0) { long long* row = (P).prof + (size_t) blockIdx.x * (2 * kWarps + 1); \
		row[role] = profBusy; row[kWarps + 1 + (role)] = profWait; if ((role) == 0) row[kWarps] = profIters; } } while (0)
#else
#define GTTS_PROF_DECL
#define GTTS_PROF_T0() ((void) 0)
#define GTTS_PROF_T1() ((void) 0)
#define GTTS_PROF_T2() ((void) 0)
#define GTTS_PROF_STORE(P, role, lane) ((void) 0)
#endif

// A role's loop: before iteration i wait for the neighbours' iteration i - 1 (and for control block i), run the body
// on control block p = i % 4, publish i + 1.
#define GTTS_ROLE_LOOP2(TYPE, BODY)                                              \
	{                                                                            \
		GTTS_PROF_DECL;                                                          \
		for (int it = 0;; ++it) {                                                \
			GTTS_PROF_T0();                                                      \
			role_wait<TYPE>(C, lane, it);                                        \
			GTTS_PROF_T1();                                                      \
			const int p = it & (kCtlRing - 1);                                   \
			if (!C->sched[p].live) break;                                        \
			BODY                                                                 \
			role_signal<TYPE>(C, role, lane, it);                                \
			GTTS_PROF_T2();                                                      \
		}                                                                        \
		GTTS_PROF_STORE(P, role, lane);                                          \
	}

template<bool ST>
GTTS_DEV void tube_v2_cta_body(const KernelParamsV2& P, unsigned char* smem, int tid)
{
	CtaSm* C = reinterpret_cast<CtaSm*>(smem);
	const int lane = tid & 31;
	// Hardware warp w runs on SM sub-partition w % 4; which role runs where matters by a few percent (v1 measurements):
	// tube warps one per sub-partition, the chains and the light workers spread next to them.
#ifndef GTTS_ROLE_TABLE2
#if GTTS_ROLE_PRESET == 1
#define GTTS_ROLE_TABLE2 0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 6, 5, 4, 22, 21
#elif GTTS_ROLE_PRESET == 2
#define GTTS_ROLE_TABLE2 0, 1, 2, 3, 7, 11, 15, 18, 8, 12, 16, 19, 9, 13, 17, 20, 10, 14, 6, 5, 4, 22, 21
#elif GTTS_ROLE_PRESET == 3
#define GTTS_ROLE_TABLE2 0, 2, 7, 8, 1, 3, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 6, 5, 4, 22, 21
#else
#define GTTS_ROLE_TABLE2 0, 1, 2, 3, 5, 22, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 4, 21
#endif
#endif
	const int hw = tid >> 5;
	int role;
	{
		constexpr int roleOfHw[kWarps] = {GTTS_ROLE_TABLE2};
		// through a shuffle: the compiler then knows that the role is the same for the whole warp -- without it every
		// warp-collective instruction of the kernel is guarded against divergence (750 instructions of code)
		role = __shfl_sync(0xffffffffu, roleOfHw[hw], 0, 32);
	}
	for (int i = tid; i < kSrcFilterLen + kSrcPad; i += kThreads) C->tab[i] = i < kSrcFilterLen ? P.src_tab[i] : make_double2(0.0, 0.0);
	if (tid < kCounters) C->done[tid] = 0;
	if (tid < kSlots) {
		for (int b = 0; b < kCtlRing; ++b) { C->slot[tid].ctl[b].it = -1; C->slot[tid].ctl[b].voice = 0; C->slot[tid].ctl[b].nblocks = 0; C->slot[tid].ctl[b].vbuf = 0; }
	}
	__syncthreads();
	if (role == kRoleChainA) {
		schedule_slots<ST>(C, P, lane, -1, true);      // fills ctl[0] / sched[0]
		sched_signal(C, lane, 1);
	}
	__syncthreads();

	GTTS_SKIP_BITS2(P);
	if (role < kTubeWarps) {
		TubeCell tl = {0.0, 0.0, 0.0, 0.0};
		GTTS_ROLE_LOOP2(kTypeTube, if (!(skip & 32)) tube_iteration<ST>(C, P, role, lane, tl, p);)
	} else if (role == kRoleChainA) {
		// The scheduler publishes control block it + 1 BEFORE it waits for the helpers of its own chain: the control
		// blocks then run up to kFar iterations ahead of the slowest role, and no role ever waits for one.
		double pos = 0.0;
		GTTS_PROF_DECL;
		for (int it = 0;; ++it) {
			GTTS_PROF_T0();
			poll_counters(C, lane, it - kFar, lane < kWarps && lane != role);      // ring of control blocks: nobody more than kFar behind
			const int p = it & (kCtlRing - 1);
			if (!C->sched[p].live) break;
			schedule_slots<ST>(C, P, lane, it, false);
			sched_signal(C, lane, it + 2);
			if (it > 0) bar_wait(barrier_id(kTypeA, it), barrier_threads(kTypeA));   // the helpers' oscillator increments of iteration it - 1
			GTTS_PROF_T1();
			if (!(skip & 8)) chain_a_iteration<ST>(C, P, lane, pos, p);
			role_signal<kTypeA>(C, role, lane, it);
			GTTS_PROF_T2();
		}
		GTTS_PROF_STORE(P, role, lane);
	} else if (role == kRoleChainA2) {
		BandpassState bp = {0.0, 0.0, 0.0, 0.0};
		GTTS_ROLE_LOOP2(kTypeA2, if (!(skip & 8)) chain_a2_iteration<ST>(C, P, lane, bp, p);)
	} else if (role == kRoleChainB) {
		ChainBRegs cb = {0.0, 0.0};
		GTTS_ROLE_LOOP2(kTypeB, if (!(skip & 16)) chain_b_iteration<ST>(C, P, lane, cb, p);)
	} else if (role < kRoleCoef0) {
		HelperRegs hr = {};
		hr.mult = c_lcg[lane];
		hr.low = kNoLowMark;
		const int slot = role - kRoleHelper0;
		SlotSm* S = &C->slot[slot];
		GTTS_ROLE_LOOP2(kTypeHelper,
			if (!(skip & 4)) helper_iteration<ST>(S, P, lane, hr, p);
			if (!(skip & 1) && !C->sched[p].src_shared && ((C->sched[p].src_mask >> slot) & 1)) src_slot_task(C, P, lane, slot, p, 1);)
	} else if (role < kRoleSrcB) {
		const int slot = role - kRoleCoef0;
		WalkRegs wr = {0.f, 0.f, 0.f, 0.f, 0, 0};
		GTTS_ROLE_LOOP2(kTypeCoef,
			if (!(skip & 64)) walk_slot<ST>(&C->slot[slot], P, lane, p, wr);
			if (!(skip & 2)) coef_task(&C->slot[slot], lane, p);
			if (!(skip & 1) && !C->sched[p].src_shared && ((C->sched[p].src_mask >> slot) & 1)) src_slot_task(C, P, lane, slot, p, 0);)
	} else {
		// The SRC warps.  Slots in lockstep (one voice, equal lengths: the batch case): the shared SRC tasks, windows
		// of slots 0..3 | slots 3..6 with slot 3 left out (one instantiation).  Otherwise every slot converts its own
		// outputs, a third of the rows here, the others on the slot's coefficient worker and helper (src_slot_task).
		const int slot0 = role == kRoleSrcA ? 0 : 3;
		const int keep = role == kRoleSrcA ? 0xf : 0xe;
		const int own0 = role == kRoleSrcA ? 0 : 4, own1 = role == kRoleSrcA ? 4 : kSlots;     // per-slot SRC: the third share of slots 0..3 | 4..6
		GTTS_ROLE_LOOP2(kTypeSrc,
			if (!(skip & 1)) {
				if (C->sched[p].src_shared) {
					src_shared_group<4>(C, P, lane, slot0, keep, p);
				} else {
#pragma unroll 1
					for (int q = own0; q < own1; ++q) if ((C->sched[p].src_mask >> q) & 1) src_slot_task(C, P, lane, q, p, 2);
				}
			})
	}
}

#ifndef GTTS_EMU
__global__ void __launch_bounds__(kThreads, 1) tube_kernel_v2(const KernelParamsV2 P)
{
	extern __shared__ __align__(16) unsigned char gtts_smem_v2[];
	tube_v2_cta_body<false>(P, gtts_smem_v2, threadIdx.x);
}

// the same pipeline for the chunks of a stream (gtts_stream_*): roles start from / save to UttStateV2
__global__ void __launch_bounds__(kThreads, 1) tube_kernel_v2_stream(const KernelParamsV2 P)
{
	extern __shared__ __align__(16) unsigned char gtts_smem_v2s[];
	tube_v2_cta_body<true>(P, gtts_smem_v2s, threadIdx.x);
}
#endif

inline size_t smem_bytes() { return sizeof(CtaSm); }

} // namespace v2
} // namespace gtts
#endif
