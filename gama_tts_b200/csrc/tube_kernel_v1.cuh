// Pipelined, warp-specialised tube kernel (sm_100a): one persistent CTA per SM steps kSlots = 7
// utterances in lockstep, 32 internal samples ("block") per iteration, through a software pipeline
// whose stages run concurrently on different warps and different blocks:
//
//   it = b     slot helper   float32 walk of parameter 0 -> f0 -> oscillator increments        (lane = sample)
//   it = b+1   chain A       oscillator phase recurrence (serial, lane = slot) + slot bookkeeping
//   it = b+2   slot helper   float32 walk of parameters 1..6 -> amplitudes, frication taps, bandpass
//                            coefficients; noise (LCG jump); wavetable lookup; 49-tap FIR; mixing   (lane = sample)
//   it = b+3   chain A2      frication bandpass biquad + tap signals (serial, lane = slot)
//              task worker   float32 walk of parameters 7..15 -> junction coefficients            (lane = sample)
//   it = b+4   tube warps    the waveguide itself: one cell per lane, 16 lanes per utterance, waves in registers,
//                            three shuffles per sample
//   it = b+5   chain B       radiation filters + throat lowpass (serial, lane = slot x filter) + output sum
//   it = b+6   task worker   windowed-sinc sample-rate conversion, lane = output sample, whole 128-byte rows
//
// Warps: 0-3 tube (two slots each), 4 chain A, 5 chain B, 6 chain A2, 7-13 slot helpers, 14-23 task workers
// (static task list per iteration).  One CTA barrier per iteration: the slot control block and the task list
// are double-buffered by iteration parity.  Every role has its own loop, so a warp keeps only its own state
// in registers (80 per thread).
// The per-sample arithmetic follows tube_kernel.cuh (v0), which stays as the general kernel for streaming /
// resumed utterances and control periods shorter than one block.  DESIGN.md section 4 has the measurements
// behind the choices made here.
#ifndef GTTS_TUBE_KERNEL_V1_CUH_
#define GTTS_TUBE_KERNEL_V1_CUH_

#include "tube_kernel.cuh"

// Unrolling knobs: the kernel is sensitive to its instruction-cache footprint (tools/ab_build.sh NAME -D...)
#ifndef GTTS_WALK_UNROLL
#define GTTS_WALK_UNROLL 1
#endif
#ifndef GTTS_FIR_UNROLL
#define GTTS_FIR_UNROLL 4
#endif
#ifndef GTTS_COEF_UNROLL
#define GTTS_COEF_UNROLL 1
#endif
#ifndef GTTS_CHAINA_CHUNK
#define GTTS_CHAINA_CHUNK 4
#endif
#ifndef GTTS_CHAINB_CHUNK
#define GTTS_CHAINB_CHUNK 4
#endif

namespace gtts {
namespace v1 {

constexpr int kFirUnroll = GTTS_FIR_UNROLL, kCoefUnroll = GTTS_COEF_UNROLL, kWalkUnroll = GTTS_WALK_UNROLL;

// keeps the first use of a register inside the branch it is written in (the compiler would otherwise hoist
// cheap arithmetic on it above the branch)
#ifndef GTTS_EMU
#define GTTS_KEEP_IN_BRANCH(x) asm volatile("" : "+f"(x))
#else
#define GTTS_KEEP_IN_BRANCH(x) ((void) 0)
#endif

enum {
	kSlots = 7,
	kWarps = 24,
	kThreads = kWarps * 32,
	kTubeLanes = 16,              // one cell per lane
	kParamRow = 36,               // floats per walked-parameter row: 16-byte aligned, rows 4 banks apart (float4 stores by 7-9 lanes)
	kTubeWarps = (kSlots * kTubeLanes + 31) / 32,
	kChainAWarp = kTubeWarps,
	kChainBWarp = kTubeWarps + 1,
	kChainA2Warp = kTubeWarps + 2,
	kHelper0 = kTubeWarps + 3,    // one helper warp per slot
	kPool0 = kHelper0 + kSlots,   // the rest: task workers
	kPoolWarps = kWarps - kPool0,
	kStages = 6,                  // last stage (SRC) runs at it = b + 6
	kRow = 33,
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
	double2 kab[2][8][kRow];      // per tube lane g: {kA, kB} (lane 1: {k2, alpha left/right}; lanes 5-7 hold the fixed nasal k)
	double au[2][kRow];           // alpha upper of the 3-way junction
	double onepk7[3][kRow];
	double endm[2][kRow], endn[2][kRow];
	double rad[3][kRow];          // mouth radiation, nose radiation, throat outputs of one block
	double ve[kVRing], vo[kVRing];
	double xring[2 * kSrcRing];   // tube output ring, every sample stored twice (i and i + 128): any 26-sample window is contiguous
	alignas(16) float cur[7][kParamRow];     // slot helper scratch: parameters 0..6 of the block being converted, [parameter][sample]
	signed char ip[3][kBlock];    // integer part of the frication position (-50: none)
	VoiceDev V;                   // the slot's voice constants (copied from global memory when an utterance starts)
	int    fric[2];               // block has frication (some tap * bandpassed noise != 0), per pab buffer
	float  ckey[9];               // parameters 7..15 at the last sample of the previous block (NaN: none)
	// control block, double-buffered by iteration parity: the scheduler lane writes ctl[p ^ 1] while the
	// roles read ctl[p], so that one CTA barrier per iteration is enough
	struct Ctl {
		UttDesc U;
		int it;                   // iteration counter of the current utterance; -1: idle
		int nblocks;
		int voice;
		int last_len;             // samples in the last block
		// output rows of the block that is in the SRC stage this iteration (see src_rows), and the running
		// count they are derived from: outputs_before(32 b') = ceil(32 b' 65536 / inc) for the next SRC block b',
		// kept as quotient / remainder of (32 b' 65536 + inc - 1) / inc and advanced by (eQ, eR) per block
		long long src_k0, src_k1;
		long long eq;
		unsigned erem, eQ, eR, inc;
	} ctl[2];
	int     pad_[1];              // slot stride = 16 mod 128 bytes: the 7 slots start in different banks
};
static_assert(sizeof(SlotSm) % 128 == 16, "slot stride should be 16 mod 128 bytes (bank spreading); adjust pad_");

struct CtaSm {
	double2 tab[kSrcFilterLen];
	SlotSm slot[kSlots];
	float  pscratch[kSlots][9][kParamRow];    // per-slot scratch of the coefficient task: parameters 7..15 of one block, [parameter][sample]
	struct Sched {
		int live;                 // some slot has work
		int src_shared;           // 1: every slot in the SRC stage is at the same block of an equally long utterance of the same rate
		int src_mask;             // slots with a block in the SRC stage
		int src_tasks;            // SRC tasks this iteration
		long long src_k0, src_k1; // shared SRC tasks: first output of the first row, end of the range
		long long src_k0w;        // first output to write (>= src_k0: the first row of an unaligned utterance is partial)
	} sched[2];
};

struct KernelParamsV1 {
	const VoiceDev* voices;
	const double* tables;         // per-voice glottal wavetables, 512 doubles each (host-built)
	const UttDesc* utts;
	const int32_t* order;
	const float* frames;
	float* out;
	const double2* src_tab;
	int32_t* queue;
	int32_t n_utt;
	int32_t debug_skip;           // development builds only (-DGTTS_EXPERIMENTS + GTTS_DEBUG_SKIP): bit0 no SRC, bit1 no coef task, bit2 no helper, bit3 no chain A, bit4 no chain B, bit5 no tube
	long long* prof;              // optional [grid][kWarps + 1]: busy cycles per warp + iteration count (GTTS_PROFILE=1)
};

#ifndef GTTS_EMU
#define GTTS_CLOCK() clock64()
// clock read that cannot be scheduled before `dep` (a value loaded after a barrier) is available
#define GTTS_CLOCK_AFTER(t, dep) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory")
#else
#define GTTS_CLOCK() 0ll
#define GTTS_CLOCK_AFTER(t, dep) ((t) = 0)
#endif

// Role skipping exists in development builds only (-DGTTS_EXPERIMENTS, tools/ab_build.sh): the shipped kernel has no
// switch that changes its output.
#ifdef GTTS_EXPERIMENTS
#define GTTS_SKIP_BITS(P) const int skip = (P).debug_skip
#else
#define GTTS_SKIP_BITS(P) constexpr int skip = 0
#endif

GTTS_DEV int block_len(const SlotSm::Ctl& k, int b)
{
	return b == k.nblocks - 1 ? k.last_len : kBlock;
}

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

// ---- float32 walk of a group of parameters over one block (Controller.cpp:297-311) -----------------
// lane k walks parameter base + k: cur/delta/off/frame are the lane's cursor (registers or shared),
// out[j] (the lane's own row of kParamRow floats, 16-byte aligned) receives the value used for sample j.  Control periods are >= one block, so at most one
// frame boundary falls inside the block.
// The cursor carries the next two frame values (fn1 = frame[f+1], fn2 = frame[f+2], clamped to the last
// frame): they are fetched from global memory one control period ahead, so no load sits on the walk.
GTTS_DEV void walk_block(const float* frames, long long nFrames, int steps, float invSteps, int param,
			int nb, float& cur, float& delta, float& fn1, float& fn2, int& off, int& frame,
			float* out, bool active)
{
	// `first` samples of the block still belong to the current control period; the period ends inside
	// (or exactly at the end of) this block iff off + first == steps.  At step `first` the walk restarts
	// from the next frame value (the reference restarts from the frame value, not from the accumulated
	// one: Controller.cpp:297-300).
	if (steps == 1) {
		// one frame per internal sample (the plugin shim records the reference's per-sample parameters:
		// Controller.cpp:303-311 runs one step on cur = frame[i]): the values ARE the frames, no walk.
		// `frame` counts the samples consumed so far.  Out of line: it must not sit in the hot instruction stream.
		if (active) {
			copy_frames(frames + (long long) frame * kNumParams + param, nb, out);
			frame += nb;
		}
		return;
	}
	const int first = (steps - off) < nb ? (steps - off) : nb;
	const bool reaches = active && (off + first == steps);
	const int restart = reaches ? first : -1;
	float c = cur, d = delta;
	// The lanes of one call belong to at most two groups with a common cursor offset (lane 0 | lanes 1..), so
	// at most two chunks of four samples contain a restart point: only those test every sample (warp-uniform
	// branch), the others are four stores and four adds.  A lane restarts from the next frame value; fn2 was
	// fetched one control period ago and is not touched on any other path, so that a fetch still in flight
	// (pinned host memory: microseconds over PCIe) never stalls the walk.
	const int mine = reaches ? first : 2 * kBlock;
	const int ra = __shfl_sync(0xffffffffu, mine, 0), rb = __shfl_sync(0xffffffffu, mine, 1);
	float4* o = reinterpret_cast<float4*>(out);
#pragma unroll kWalkUnroll
	for (int j0 = 0; j0 < kBlock; j0 += 4) {
		float v[4];
		if ((unsigned) (ra - j0) >= 4u && (unsigned) (rb - j0) >= 4u) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				v[q] = c;
				c = __fadd_rn(c, d);
			}
		} else {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (j0 + q == restart) {
					GTTS_KEEP_IN_BRANCH(fn2);
					c = fn1;
					d = __fmul_rn(__fsub_rn(fn2, fn1), invSteps);
				}
				v[q] = c;
				c = __fadd_rn(c, d);
			}
		}
		if (active) *o = make_float4(v[0], v[1], v[2], v[3]);      // one 16-byte store per four samples
		o += 1;
	}
	if (restart == kBlock) {
		GTTS_KEEP_IN_BRANCH(fn2);
		c = fn1;
		d = __fmul_rn(__fsub_rn(fn2, fn1), invSteps);
	}
	if (active) {
		cur = c;
		delta = d;
		if (reaches) {
			frame += 1; off = nb - first; fn1 = fn2;
			// refill for the boundary after the next one: frame[f+3], needed a whole control period from now;
			// loaded straight into the cursor register, which nothing reads before that boundary
			if ((long long) frame + 2 < nFrames) fn2 = frames[((long long) frame + 2) * kNumParams + param];
		} else {
			off += nb;
		}
	}
}

GTTS_DEV void cursor_init(const float* frames, long long nFrames, float invSteps, int param,
			float& cur, float& delta, float& fn1, float& fn2)
{
	const float a = frames[param];
	const float b = (nFrames > 1) ? frames[kNumParams + param] : a;
	const float c = (nFrames > 2) ? frames[2 * kNumParams + param] : b;
	cur = a;
	delta = __fmul_rn(__fsub_rn(b, a), invSteps);
	fn1 = b;
	fn2 = c;
}

// ---- slot helper (warp 4 + s): stages at it = b and it = b + 2 ------------------------------------------
struct HelperRegs {
	float cur, delta;             // lane 0: parameter 0 (block it); lanes 1..6: parameters 1..6 (block it - 2)
	float fn1, fn2;               // next two frame values of the lane's parameter (prefetched)
	int off, frame;
	unsigned long long mult;      // 377^(lane+1) mod 2^44, this lane's jump-ahead multiplier (loaded once)
	unsigned long long lcg;       // noise generator state on the 2^-44 grid: next samples are lcg * 377^(j+1) mod 2^44
	double noise_x1;
	// conversions of the previous block, re-used while the parameter does not change (the reference
	// caches the same way: BandpassFilter.h:93, WavetableGlottalSource.h:164)
	float c_p1, c_p2, c_p3, c_p5, c_p6;
	double c_ax, c_ah1, c_fa, c_a2, c_a1, c_b0;
};

GTTS_DEV void helper_iteration(CtaSm* C, SlotSm* S, const KernelParamsV1& P, int lane, HelperRegs& h, int p)
{
	const SlotSm::Ctl& K = S->ctl[p];
	const int it = K.it;
	if (it < 0) return;
	const VoiceDev& V = S->V;
	const float* frames = P.frames + K.U.frame_begin * kNumParams;
	const long long nFrames = K.U.n_frames;
	if (it == 0) {
		// voice constants into shared memory: every later stage of this utterance reads them from there
		{
			const double* src = reinterpret_cast<const double*>(&P.voices[K.voice]);
			double* dst = reinterpret_cast<double*>(&S->V);
			for (int i = lane; i < (int) (sizeof(VoiceDev) / sizeof(double)); i += 32) dst[i] = src[i];
		}
		// new utterance: clear the rings, reset cursors and the noise generator
		for (int i = lane; i < kVRing; i += 32) { S->ve[i] = 0.0; S->vo[i] = 0.0; }
		for (int i = lane; i < 2 * kSrcRing; i += 32) S->xring[i] = 0.0;
		if (lane < 7) cursor_init(frames, nFrames, K.U.inv_steps, lane, h.cur, h.delta, h.fn1, h.fn2);
		h.off = 0; h.frame = 0;
		h.lcg = c_lcg_init; h.noise_x1 = 0.0;
		h.c_p1 = h.c_p2 = h.c_p3 = h.c_p5 = h.c_p6 = __int_as_float(0x7fc00000);     // NaN: nothing cached
		__syncwarp();
	}
	const int b0 = it, b2 = it - 2;
	const bool do0 = b0 < K.nblocks, do2 = b2 >= 0 && b2 < K.nblocks;
	// one walk loop serves both cursors: lane 0 (block b0), lanes 1..6 (block b2)
	{
		const bool mine = (lane == 0) ? do0 : (lane < 7 && do2);
		const int nb = (lane == 0) ? block_len(K, b0) : block_len(K, b2);
		walk_block(frames, nFrames, K.U.steps, K.U.inv_steps, lane, nb, h.cur, h.delta, h.fn1, h.fn2, h.off, h.frame,
				S->cur[lane < 7 ? lane : 0], mine);
	}
	__syncwarp();
	if (do0) {
		const int nb = block_len(K, b0);
		if (lane < nb) {
			const double f0 = 220.0 * gtts_exp2(((double) S->cur[0][lane] + 3.0) * (1.0 / 12.0));
			S->osc[b0 & 1][lane] = (f0 / 2.0) * V.basic_inc;
		}
	}
	if (do2) {
		const int nb = block_len(K, b2);
		const int buf = b2 & 1;
		double ax = 0.0, ah1 = 0.0;
		{
			const bool live = lane < nb;
			const int col = lane < nb ? lane : 0;
			const float p1 = S->cur[1][col], p2 = S->cur[2][col], p3 = S->cur[3][col], p5 = S->cur[5][col], p6 = S->cur[6][col];
			// glottal, aspiration and frication amplitudes (VTMUtil.h:50-67): one rolled loop over the three
			// parameters.  A value is re-used while its parameter is uniform over the block and unchanged
			// since the previous block (the reference caches the same way, e.g. BandpassFilter.h:93).
			// All three are converted in ONE straight-line region when any of them has to be (three independent
			// exp10 chains overlap; one after the other they were a quarter of the helper's time).
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
			const double fpos = (double) S->cur[4][col];
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
				S->ip[b2 % 3][lane] = (signed char) ip;
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
		// wavetable lookup of both half samples (WavetableGlottalSource.h:212-228)
		if (lane < nb) {
			const double* table = P.tables + (size_t) K.voice * kTableLen;
			const bool dynamic = (V.waveform == 0) && (V.tn_delta != 0.0);
			double nd2 = 0.0, inv = 0.0;
			if (dynamic) {
				nd2 = (double) V.div2 - rint(ax * V.tn_delta);
				nd2 = nd2 > 0.0 ? nd2 : 0.0;
				inv = 1.0 / (nd2 - (double) V.div1);
			}
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
					tl = table[lo & (kTableLen - 1)];
				}
				if (dynamic && up >= (unsigned) V.div1 && up < (unsigned) V.div2) {
					const double x = (double) (int) (up - V.div1) * inv;
					tu = (up >= (unsigned) nd2) ? 0.0 : 1.0 - (x * x);
				} else {
					tu = table[up & (kTableLen - 1)];
				}
				v[s] = tl + ((pos - (double) lo) * (tu - tl));
			}
			// linear window: [0, 24) = the last 24 samples of the previous blocks, [24, 56) = this block
			S->ve[24 + lane] = v[0];
			S->vo[24 + lane] = v[1];
		}
		__syncwarp();
		double firOut = 0.0;
		{
			// y = sum_i c[i] x2[2n+1-i], i ascending: taps 2m and 2m+1 read the odd and the even phase of
			// sample n - m (m = 0..23), tap 48 the odd phase of n - 24.  Groups of 4 pairs, the next group
			// is loaded while the current one is accumulated.
			const double* pe = S->ve + 24 + lane;
			const double* po = S->vo + 24 + lane;
#ifndef GTTS_EMU
			// two accumulator chains (odd-phase taps | even-phase taps): 25 dependent multiply-adds instead of 49
			double accO = 0.0, accE = 0.0;
#pragma unroll kFirUnroll
			for (int m = 0; m < 24; ++m) {
				accO += po[-m] * c_fir[2 * m];
				accE += pe[-m] * c_fir[2 * m + 1];
			}
			accO += po[-24] * c_fir[48];
			firOut = accO + accE;
#else
			// CPU emulation (tests/simt_emu): the reference's single ascending sum, so that the emulated kernel
			// stays bit-identical to the oracle
			double acc = 0.0;
			for (int m = 0; m < 24; ++m) {
				acc += po[-m] * c_fir[2 * m];
				acc += pe[-m] * c_fir[2 * m + 1];
			}
			acc += po[-24] * c_fir[48];
			firOut = acc;
#endif
		}
		__syncwarp();
		if (lane < 24) {
			// slide the window: the last 24 samples become the history of the next block
			S->ve[lane] = S->ve[32 + lane];      // source [32, 56) and destination [0, 24) do not overlap
			S->vo[lane] = S->vo[32 + lane];
		}
		if (lane < nb) {
			const double acc = firOut;
			double pulse = acc;
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
			S->in[b2 % 3][lane] = (pulse + (ah1 * sig)) * 0.125;
			S->thr[b2 & 3][lane] = pulse * 0.125;
		}
	}
	(void) C;
}

// ---- pool task: junction coefficients of block b = it - 3 (VocalTractModel0.h:484-512, 698-716) ---------
// The walk cursor of parameters 7..15 (lanes 0..8) lives in the registers of pool worker `slotIndex`, which runs
// this slot's task in every iteration (tasks 0..6 of the list are the coefficient tasks).
struct CoefRegs { float cur, delta, fn1, fn2; int off, frame; };

GTTS_DEV void coef_task(CtaSm* C, SlotSm* S, const KernelParamsV1& P, int lane, int slotIndex, int p, CoefRegs& r)
{
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - 3;
	if (K.it < 0 || b < 0 || b >= K.nblocks) return;
	const VoiceDev& V = S->V;
	const float* frames = P.frames + K.U.frame_begin * kNumParams;
	const int nb = block_len(K, b);
	float (*scr)[kParamRow] = C->pscratch[slotIndex];
	if (b == 0) {
		r.cur = r.delta = r.fn1 = r.fn2 = 0.f;
		if (lane < 9) cursor_init(frames, K.U.n_frames, K.U.inv_steps, 7 + lane, r.cur, r.delta, r.fn1, r.fn2);
		r.off = 0; r.frame = 0;
	}
	walk_block(frames, K.U.n_frames, K.U.steps, K.U.inv_steps, 7 + lane, nb, r.cur, r.delta, r.fn1, r.fn2, r.off, r.frame,
			scr[lane < 9 ? lane : 0], lane < 9);
	__syncwarp();
	// Radii and velum unchanged since the last sample of the previous block (a held posture): the
	// coefficients are the previous block's last row, copied instead of recomputed (9 divisions saved).
	const int buf0 = b & 1;
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
			// rows 6, 7 hold per-voice constants: written with blocks 0 and 1 (one per buffer), never again
#pragma unroll
			for (int r = 0; r < 6; ++r) S->kab[buf0][r][lane] = S->kab[buf0 ^ 1][r][kBlock - 1];
			if (b < 2) {
				S->kab[buf0][6][lane] = S->kab[buf0 ^ 1][6][kBlock - 1];
				S->kab[buf0][7][lane] = S->kab[buf0 ^ 1][7][kBlock - 1];
			}
			S->au[buf0][lane] = S->au[buf0 ^ 1][kBlock - 1];
			S->onepk7[b % 3][lane] = S->onepk7[(b + 2) % 3][kBlock - 1];
		}
	} else if (lane < nb) {
		// Nine scattering coefficients k = (a - b) / (a + b) in ONE rolled loop (the kernel is bound by its
		// instruction footprint): i = 0..6 oral junctions r_i | r_{i+1}, i = 7 mouth r_8 | aperture,
		// i = 8 velum | first nasal section.  Radii as in setAllParameters: max(r * coef, 0.01).
		const int buf = b & 1;
		double* kabFlat = &S->kab[buf][0][0].x;         // [row][lane]{x, y} -> (row * kRow + lane) * 2 + c
		const double vel = (double) scr[8][lane];
		const double v2 = vel * vel;
		const double dmp = V.damping;
		double a2, r2_3 = 0.0, k7 = 0.0;
		{
			double r = (double) scr[0][lane] * V.radius_coef[0];
			r = r > 0.01 ? r : 0.01;
			a2 = r * r;
		}
#pragma unroll kCoefUnroll
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
			const int dst = i <= 2 ? i : (i == 3 ? 4 : i + 2);     // (row, component) of junction i in the kab rows
			kabFlat[((dst >> 1) * kRow + lane) * 2 + (dst & 1)] = ((i == 7) ? k * V.refl_b0_m : k) * dmp;   // damping (and the mouth end's b0) folded in: tube_iteration
			a2 = b2;
		}
		const double sum = div_fast(2.0, r2_3 + r2_3 + v2);
		S->kab[buf][1][lane].y = ((sum * r2_3) - 1.0) * dmp;   // alpha left == alpha right, stored as (alpha - 1) d (tube_iteration)
		if (b < 2) {
			// the fixed junctions: the same in every block, so written once per buffer and utterance
			S->kab[buf][2][lane].y = 0.0;                      // S6-S7 is a pure damped delay: k = 0
			S->kab[buf][5][lane].y = V.nasal_k[1] * dmp;
			S->kab[buf][6][lane] = make_double2(V.nasal_k[2] * dmp, V.nasal_k[3] * dmp);
			S->kab[buf][7][lane] = make_double2(V.nasal_k[4] * dmp, (V.nasal_k[5] * V.refl_b0_n) * dmp);   // nose end: b0 folded in
		}
		S->au[buf][lane] = (sum * v2) * dmp;
		S->onepk7[b % 3][lane] = 1.0 + k7;
	}
	__syncwarp();
}

// ---- output rows of the SRC stage --------------------------------------------------------------------
// Block b completes the outputs [e(b), e(b + 1)), e(b) = ceil(32 b 65536 / inc) (closed form of
// SampleRateConverter.h:295-361).  The SRC tasks write whole 128-byte rows instead: every block but the last
// stops at the last row boundary (an absolute multiple of 32 samples in the output buffer) and leaves the
// remainder to the next block, whose window still holds the inputs (the ring keeps 128: 57 back + 64 ahead
// being written).  Row-aligned 128-byte stores are what lets the kernel write straight into pinned host memory
// at PCIe rate (52.7 GB/s measured against 30 GB/s for unaligned rows, tools/microbench/zerocopy_store.cu).
// The boundary between the last two blocks stays exact: the 26 flush zeros that chain B appends after the last
// block would otherwise reach, around the ring, the oldest inputs of the deferred outputs.
// Called by the scheduler lane of the slot for the block that enters the SRC stage; e0 = e(b), e1 = e(b + 1).
GTTS_DEV void src_rows(SlotSm::Ctl& N, int b, long long e0, long long e1)
{
	const long long a = N.U.out_begin & 31;        // 0 with the planner's padded layout
	long long k0 = (b == 0) ? 0 : (b == N.nblocks - 2 ? e0 : ((e0 + a) & ~31ll) - a);
	long long k1;
	if (b == N.nblocks - 1) {
		k1 = N.U.n_out;                             // flush: chain B appended the 26 zeros
	} else {
		k1 = (b == N.nblocks - 3) ? e1 : ((e1 + a) & ~31ll) - a;
		if (k1 > N.U.n_out) k1 = N.U.n_out;
	}
	if (k0 < 0) k0 = 0;
	if (k1 < k0) k1 = k0;
	N.src_k0 = k0;
	N.src_k1 = k1;
}

// ---- pool task: sample-rate conversion of the outputs that block b = it - 6 completes ---------------
GTTS_DEV void src_task(CtaSm* C, SlotSm* S, const KernelParamsV1& P, int lane, int p)
{
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = K.it - kStages;
	if (K.it < 0 || b < 0 || b >= K.nblocks) return;
	const VoiceDev& V = S->V;
	const unsigned inc = V.src_inc;
	const long long k0 = K.src_k0, k1 = K.src_k1;
	// rows start on absolute multiples of 32 samples: lane = (out_begin + k) mod 32
	const long long kRow0 = ((k0 + (K.U.out_begin & 31)) & ~31ll) - (K.U.out_begin & 31);
	float* out = P.out + K.U.out_begin;
	// Up to 96 outputs per block (ratio < 3): each lane carries three independent accumulator chains
	// (outputs k, k + 32, k + 64) so that the 52 dependent multiply-adds of one output overlap with the
	// other two; every chain still sums its own taps in the reference's order (left wing, then right).
	for (long long kb = kRow0; kb < k1; kb += 96) {
		double acc[3] = {0.0, 0.0, 0.0};
		double interpL[3], interpR[3];
		const double2* tabL[3];
		const double2* tabR[3];
		const double* xw[3];
		long long kk[3];
#pragma unroll
		for (int o = 0; o < 3; ++o) {
			kk[o] = kb + lane + 32 * o;
			const unsigned long long t = (unsigned long long) (kk[o] < 0 ? 0 : kk[o]) * inc;
			const int e = (int) (t >> 16);
			const unsigned f = (unsigned) (t & 0xFFFFu);
			const unsigned gph = (~f) & 0xFFFFu;
			interpL[o] = (double) (f & 0xFFu) / 256;
			interpR[o] = (double) (gph & 0xFFu) / 256;
			tabL[o] = C->tab + (f >> 8);
			tabR[o] = C->tab + (gph >> 8);
			xw[o] = S->xring + ((e - 25) & (kSrcRing - 1));      // window [e-25, e], contiguous in the doubled ring
		}
#pragma unroll 1
		for (int j = 0; j < kSrcZeroCrossings; ++j) {
#pragma unroll
			for (int o = 0; o < 3; ++o) {
				const double2 c = tabL[o][256 * j];
				const double x = xw[o][12 - j];                     // x[e - 13 - j]
				acc[o] += (x * (c.x + (c.y * interpL[o])));
			}
		}
#pragma unroll 1
		for (int j = 0; j < kSrcZeroCrossings; ++j) {
#pragma unroll
			for (int o = 0; o < 3; ++o) {
				const double2 c = tabR[o][256 * j];
				const double x = xw[o][13 + j];                     // x[e - 12 + j]
				acc[o] += (x * (c.x + (c.y * interpR[o])));
			}
		}
#pragma unroll
		for (int o = 0; o < 3; ++o) if (kk[o] >= k0 && kk[o] < k1) out[kk[o]] = (float) acc[o];
	}
}

// ---- pool task: SRC of 32 outputs for ALL slots at once (slots aligned: same rate, same block) ---------
// When the slots step equally long utterances of one voice in lockstep (the batch case), output k of
// every slot has the same phase, hence the same 26 interpolated coefficients h + deltaH * frac: they are
// fetched and formed once and applied to each slot's own window.  This divides the coefficient traffic
// (the dominant shared-memory load of the whole kernel) by the number of slots; each slot's sum still
// runs over its own taps in the reference's order.
GTTS_DEV void src_shared_task(CtaSm* C, const KernelParamsV1& P, int lane, int pass, int p)
{
	const int mask = C->sched[p].src_mask;
	int ref = 0;
	while (!((mask >> ref) & 1)) ++ref;
	const unsigned inc = C->slot[ref].V.src_inc;
	// src_k0 is the first output of the first ROW (it may lie before the first output to write, k0w)
	const long long k0 = C->sched[p].src_k0, k1 = C->sched[p].src_k1, k0w = C->sched[p].src_k0w;
	const long long k = k0 + 32ll * pass + lane;
	const unsigned long long t = (unsigned long long) (k < 0 ? 0 : k) * inc;
	const int e = (int) (t >> 16);
	const unsigned f = (unsigned) (t & 0xFFFFu);
	const unsigned gph = (~f) & 0xFFFFu;
	const double interpL = (double) (f & 0xFFu) / 256, interpR = (double) (gph & 0xFFu) / 256;
	const double2* tabL = C->tab + (f >> 8);
	const double2* tabR = C->tab + (gph >> 8);
	const int woff = (e - 25) & (kSrcRing - 1);
	double acc[kSlots];
#pragma unroll
	for (int q = 0; q < kSlots; ++q) acc[q] = 0.0;
#pragma unroll 1
	for (int j = 0; j < kSrcZeroCrossings; ++j) {
		const double2 c = tabL[256 * j];
		const double cc = c.x + (c.y * interpL);
#pragma unroll
		for (int q = 0; q < kSlots; ++q) acc[q] += (C->slot[q].xring[woff + 12 - j] * cc);
	}
#pragma unroll 1
	for (int j = 0; j < kSrcZeroCrossings; ++j) {
		const double2 c = tabR[256 * j];
		const double cc = c.x + (c.y * interpR);
#pragma unroll
		for (int q = 0; q < kSlots; ++q) acc[q] += (C->slot[q].xring[woff + 13 + j] * cc);
	}
	if (k >= k0w && k < k1) {
#pragma unroll
		for (int q = 0; q < kSlots; ++q) {
			if ((mask >> q) & 1) P.out[C->slot[q].ctl[p].U.out_begin + k] = (float) acc[q];
		}
	}
}

// ---- chain A / chain A2 (one warp each, lane = slot): the two serial recurrences ahead of the tube -------------
// Lanes whose slot has no such block run on dummy data (their results are never read), which keeps the loops
// free of divergent branches.
struct ChainARegs { double pos; };
struct ChainA2Regs { BandpassState bp; };

// chain A, lane = slot: oscillator phase of block it - 1 (WavetableGlottalSource.h:196-199, 265-272: two
// half-sample increments per sample, wrap above 511).  42 dependent cycles per sample; the operands of a chunk
// of samples are loaded before it is stepped.
GTTS_DEV void chain_a_iteration(CtaSm* C, const KernelParamsV1& P, int lane, ChainARegs& r, int p)
{
	(void) P;
	if (lane >= kSlots) return;
	SlotSm* S = &C->slot[lane];
	const int b1 = S->ctl[p].it - 1;
	if (b1 == 0) r.pos = 0.0;
	const int buf1 = b1 & 1;
	const double* osc = S->osc[buf1];
	double* p0 = S->pos[buf1][0];
	double* p1 = S->pos[buf1][1];
	double pos = r.pos;
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock; j0 += GTTS_CHAINA_CHUNK) {
		double inc[GTTS_CHAINA_CHUNK], o0[GTTS_CHAINA_CHUNK], o1[GTTS_CHAINA_CHUNK];
#pragma unroll
		for (int q = 0; q < GTTS_CHAINA_CHUNK; ++q) inc[q] = osc[j0 + q];
#pragma unroll
		for (int q = 0; q < GTTS_CHAINA_CHUNK; ++q) {
			double s = pos + inc[q];
			pos = (s > 511.0) ? s - 512.0 : s;
			o0[q] = pos;
			s = pos + inc[q];
			pos = (s > 511.0) ? s - 512.0 : s;
			o1[q] = pos;
		}
#pragma unroll
		for (int q = 0; q < GTTS_CHAINA_CHUNK; ++q) { p0[j0 + q] = o0[q]; p1[j0 + q] = o1[q]; }
	}
	r.pos = pos;
}

// chain A2, lane = slot: frication bandpass of block it - 3 (BandpassFilter.h:114-122) and the two tap signals.
// Lanes whose slot has no such block run on dummy data (their results are never read).
GTTS_DEV void chain_a2_iteration(CtaSm* C, const KernelParamsV1& P, int lane, ChainA2Regs& r, int p)
{
	(void) P;
	if (lane >= kSlots) return;
	SlotSm* S = &C->slot[lane];
	const int b3 = S->ctl[p].it - 3;
	if (b3 == 0) { r.bp.x1 = r.bp.x2 = r.bp.y1 = r.bp.y2 = 0.0; }
	const int buf3 = b3 & 1;
	const double* sig = S->sig[buf3];
	const double* c0 = S->bp[buf3][0];
	const double* c1 = S->bp[buf3][1];
	const double* c2 = S->bp[buf3][2];
	const double* ta = S->tapa[buf3];
	const double* tb = S->tapb[buf3];
	double2* pab = S->pab[buf3];
	double x1 = r.bp.x1, x2 = r.bp.x2, y1 = r.bp.y1, y2 = r.bp.y2;
	bool anyFric = false;
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock; j0 += 4) {
		double x[4], b0[4], a1[4], a2[4], tA[4], tB[4], oa[4], ob[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			x[q] = sig[j0 + q]; b0[q] = c0[j0 + q]; a1[q] = c1[j0 + q]; a2[q] = c2[j0 + q];
			tA[q] = ta[j0 + q]; tB[q] = tb[j0 + q];
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const double y = b0[q] * (x[q] - x2) - a1[q] * y1 - a2[q] * y2;
			x2 = x1; x1 = x[q]; y2 = y1; y1 = y;
			oa[q] = tA[q] * y;
			ob[q] = tB[q] * y;
			anyFric = anyFric || oa[q] != 0.0 || ob[q] != 0.0;
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) pab[j0 + q] = make_double2(oa[q], ob[q]);
	}
	S->fric[buf3] = anyFric ? 1 : 0;
	r.bp.x1 = x1; r.bp.x2 = x2; r.bp.y1 = y1; r.bp.y2 = y2;
}

// ---- chain B (one warp, lane = slot * 4 + filter): radiation filters + throat lowpass of block it - 5 ----
// One code path for the three one-pole filters: y = b0 x + b1 x1 - a1 y1, out = y * gain
//   f = 0 mouth radiation (b0 = A, b1 = a1 = -A, gain 1) on (1 + k7) T[S10]     (RadiationFilter.h:73-79)
//   f = 1 nose radiation on (1 + nk5) NT[N6]
//   f = 2 throat lowpass (b0, b1 = 0, a1, gain = throat gain) on pulse * 0.125   (Throat.h:80-85)
struct ChainBRegs { double x1, y1; };

GTTS_DEV void chain_b_iteration(CtaSm* C, const KernelParamsV1& P, int lane, ChainBRegs& r, int p)
{
	const int s = lane >> 2, f = lane & 3;
	if (s < kSlots && f < 3) {
		SlotSm* S = &C->slot[s];
		const SlotSm::Ctl& K = S->ctl[p];
		const int b = K.it - 5;
		const VoiceDev& V = S->V;
		if (b == 0) { r.x1 = 0.0; r.y1 = 0.0; }
		const double b0 = f == 0 ? V.rad_m : (f == 1 ? V.rad_n : V.throat_b0);
		const double b1 = f == 0 ? -V.rad_m : (f == 1 ? -V.rad_n : 0.0);
		const double a1 = f == 0 ? -V.rad_m : (f == 1 ? -V.rad_n : V.throat_a1);
		const double gain = f == 2 ? V.throat_gain : 1.0;
		const double onePlusN = 1.0 + V.nasal_k[5];
		const double* in = f == 0 ? S->endm[b & 1] : (f == 1 ? S->endn[b & 1] : S->thr[b & 3]);
		const double* scale = S->onepk7[(b % 3 + 3) % 3];
		double* out = S->rad[f];
		double x1 = r.x1, y1 = r.y1;
#pragma unroll 1
		for (int j0 = 0; j0 < kBlock; j0 += GTTS_CHAINB_CHUNK) {
			double raw[GTTS_CHAINB_CHUNK], sc[GTTS_CHAINB_CHUNK], o[GTTS_CHAINB_CHUNK];
#pragma unroll
			for (int q = 0; q < GTTS_CHAINB_CHUNK; ++q) { raw[q] = in[j0 + q]; sc[q] = scale[j0 + q]; }
#pragma unroll
			for (int q = 0; q < GTTS_CHAINB_CHUNK; ++q) {
				const double m = f == 0 ? sc[q] : onePlusN;
				const double x = f == 2 ? raw[q] : m * raw[q];
				const double y = b0 * x + b1 * x1 - a1 * y1;
				x1 = x;
				y1 = y;
				o[q] = y * gain;
			}
#pragma unroll
			for (int q = 0; q < GTTS_CHAINB_CHUNK; ++q) out[j0 + q] = o[q];
		}
		r.x1 = x1; r.y1 = y1;
	}
	__syncwarp();
	// output sum, lane = sample: (mouth + nose) + throat (VocalTractModel0.h:657-660, 441)
#pragma unroll 1
	for (int q = 0; q < kSlots; ++q) {
		SlotSm* Q = &C->slot[q];
		const SlotSm::Ctl& QK = Q->ctl[p];
		const int qb = QK.it - 5;
		if (QK.it < 0 || qb < 0 || qb >= QK.nblocks) continue;
		const int qn = block_len(QK, qb);
		const long long n0 = (long long) qb * kBlock;
		if (lane < qn) {
			const int idx = (int) ((n0 + lane) & (kSrcRing - 1));
			const double v = (Q->rad[0][lane] + Q->rad[1][lane]) + Q->rad[2][lane];
			Q->xring[idx] = v;
			Q->xring[idx + kSrcRing] = v;
		}
		if (qb == QK.nblocks - 1 && lane < 2 * kSrcZeroCrossings) {
			// flushBuffer(): 26 zeros after the last input (SampleRateConverter.h:462-471)
			const int idx = (int) ((QK.U.n_internal + lane) & (kSrcRing - 1));
			Q->xring[idx] = 0.0;
			Q->xring[idx + kSrcRing] = 0.0;
		}
	}
}

// ---- tube warps: block it - 4, one cell per lane, 16 lanes per utterance, two utterances per warp ---------
// Cells (same wiring as stage_tube in tube_kernel.cuh): u = 0..9 oral S1..S10 (u = 3: the 3-way junction after
// S4, u = 9: mouth end), u = 10..15 nasal N1..N6 (u = 15: nose end).  The forward wave goes u -> u + 1, the
// backward wave u + 1 -> u, the velum branch links u = 3 <-> u = 10: three 64-bit shuffles per sample.
//
// The loop is one dependent chain per sample (shuffle -> T -> outputs -> shuffle); a single warp issues in
// order, so both the length of that chain and the number of instructions around it set the iteration time of
// the whole CTA.  Hence (measured steps in DESIGN.md section 4.4):
//  * the three kinds of cell are ONE formula with per-lane constants -- a 64-bit select costs two issue slots on
//    top of the arithmetic it chooses between:
//        dl = e3 nb + k (T + sigma Bn)      Tout = (T + dl) d + tap      Bout = (cA last + cW Bn + dl) d
//      2-port junction  sigma = -1, cW = 1, cA = 0, e3 = 0: dl = k (T - Bn)                (VocalTractModel0.h:575-600)
//      3-way junction   sigma = +1, cW = 1, e3 = alpha_u, k = alpha - 1: dl = jp - T - Bn with
//                       jp = alpha (T + Bn) + alpha_u nb; Tout = (jp - Bn) d, Bout = (jp - T) d, and the wave into
//                       the nose (jp - nb) d = (T + dl + Bn - nb) d                           (:602-617)
//      open end         sigma = 0, cW = 0, cA = -a1 / d, k = b0 k_end: Bout / d = b0 (k T) - a1 y1 is the reflection
//                       lowpass, whose state y1 is last / d (last = the lane's own Bout of the previous sample)  (:619-630)
//  * damping is folded into the coefficients (the coefficient task stores k d and alpha_u d), so that every output is
//    two dependent FMAs behind T:  out = kd ts + (d T + early), ts = T + sigma Bn, `early` known before T is;
//  * the input of the cell is mP fromPrev + mL link + (mG last + input) with 0 / 1 / d masks (input = 0 off lane 0);
//  * per-sample operands are prefetched four samples at a time, frication taps sit behind a warp-uniform branch.
// Lanes of slots without a block at this stage run on dummy data: their state is reset when their block 0 arrives.
struct TubeCell { double T, Bn, nb, last; };      // state of a tube lane between blocks (see tube_iteration)

GTTS_DEV void tube_iteration(CtaSm* C, const KernelParamsV1& P, int warp, int lane, TubeCell& t, int p)
{
	(void) P;
	// lane = 2 u + s: cell u of the warp's slot s.  The two slots are interleaved so that the few lanes that load
	// or store something of their own (u = 0, 3, 9, 15) sit in the same half-warp for both slots: a 64-bit
	// shared-memory access costs one wavefront per half-warp with an active lane, and the shared-memory pipe
	// (71 % busy) is what the loop's shuffles queue behind.
	const int u = lane >> 1;
	const int sbit = lane & 1;
	const int slot = warp + kTubeWarps * sbit;         // slots w and w + 4: their rows are 64 bytes (16 banks) apart modulo 128
	SlotSm* S = &C->slot[slot < kSlots ? slot : 0];
	const SlotSm::Ctl& K = S->ctl[p];
	const int b = (slot < kSlots) ? K.it - 4 : -1;
	const bool hasBlock = slot < kSlots && K.it >= 0 && b >= 0 && b < K.nblocks;
	const int buf = b & 1, b3 = (b % 3 + 3) % 3;
	const bool fricBlock = __any_sync(0xffffffffu, hasBlock && S->fric[buf] != 0);
	const VoiceDev& V = S->V;
	if (b == 0) { t.T = t.Bn = t.nb = t.last = 0.0; }
	const double d = V.damping;
	const bool is3 = u == 3, isEnd = (u == 9) || (u == 15);
	const bool storesEnd = isEnd && slot < kSlots;     // the second group of the last warp is a dummy: it must not store
	const double sigma = is3 ? 1.0 : (isEnd ? 0.0 : -1.0);
	const double cWd = isEnd ? 0.0 : d;
	const double cAd = isEnd ? -(u == 9 ? V.refl_a1_m : V.refl_a1_n) : 0.0;
	const double mP = (u == 0 || u == 10) ? 0.0 : 1.0, mL = (u == 10) ? 1.0 : 0.0, mG = (u == 0) ? d : 0.0;
	const int tap = (u >= 1 && u <= 8) ? u - 1 : -100;
	const double* kRowp = &S->kab[buf][u >> 1][0].x + (u & 1);     // component u & 1 of the pair row, stride 2 doubles
	const double2* pabRow = S->pab[buf];
	const bool isGlot = u == 0;
	const double* e3Row = S->au[buf];                  // read by lane u = 3 only
	const double* inRow = S->in[b3];                   // read by lane u = 0 only
	const signed char* ipRow = S->ip[b3];
	double* endRow = (u == 15) ? S->endn[buf] : S->endm[buf];
	const int srcPrev = 2 * ((u + 15) & 15) + sbit, srcNext = 2 * ((u + 1) & 15) + sbit, srcLink = 2 * (is3 ? 10 : 3) + sbit;
	// T = forward wave into the cell, Bn = backward wave from the next cell, nb = wave on the velum link,
	// last = the cell's own backward output of the previous sample (glottis reflection, end-filter state)
	double T = t.T, Bn = t.Bn, nb = t.nb, last = t.last;
	// alpha_u d (lane u = 3) and the glottal input (lane u = 0): zero on every other lane, loaded under predicate
	double e3v[4] = {0.0, 0.0, 0.0, 0.0}, inv[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
	for (int j0 = 0; j0 < kBlock; j0 += 4) {
		double kv[4], tf[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			kv[q] = kRowp[2 * (j0 + q)];
			if (is3) e3v[q] = e3Row[j0 + q];
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
			// known before T: everything that depends only on the neighbours' waves and the lane's own state
			const double e = e3v[q] * nb;                            // alpha_u d NB[N1] (3-way junction only)
			const double pre = (mG * last) + inv[q];                 // glottis: T[S1] = B[S1] d + input, B[S1] of the previous sample
			const double eB = (cAd * last) + ((cWd * Bn) + e);
			const double eT = e + tf[q];
			const double eX = e + (d * (Bn - nb));
			// behind T: two dependent FMAs per output
			const double ts = (sigma * Bn) + T;
			const double dT = d * T;
			if (storesEnd) endRow[j0 + q] = T;
			const double Tout = (kv[q] * ts) + (dT + eT);
			const double Bout = (kv[q] * ts) + eB;
			const double Xd = (kv[q] * ts) + (dT + eX);
			const double linkOut = is3 ? Xd : Bout;
			const double fromPrev = shfl_d(Tout, srcPrev, 32);
			const double fromNext = shfl_d(Bout, srcNext, 32);
			const double link = shfl_d(linkOut, srcLink, 32);
			last = Bout;
			T = (mP * fromPrev) + ((mL * link) + pre);
			Bn = fromNext;
			nb = link;
		}
	}
	t.T = T; t.Bn = Bn; t.nb = nb; t.last = last;
}

// ---- slot bookkeeping for the NEXT iteration (chain A warp, lane = slot, after its chain work) ----------
// Reads ctl[p] / writes ctl[p ^ 1] and sched[p ^ 1]; everybody switches to p ^ 1 after the barrier.
GTTS_DEV void schedule_slots(CtaSm* C, const KernelParamsV1& P, int lane, int p, bool first)
{
	const int q = p ^ 1;
	int alive = 0, valid = 0, it = -1;
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
			N.eq = K.eq; N.erem = K.erem; N.eQ = K.eQ; N.eR = K.eR; N.inc = K.inc;
		} else {
			N.it = -1; N.nblocks = 0; N.voice = first ? 0 : K.voice; N.U = K.U;
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
				N.nblocks = (int) ((U.n_internal + kBlock - 1) / kBlock);
				N.last_len = (int) (U.n_internal - (long long) (N.nblocks - 1) * kBlock);
				N.it = 0;
				// output counter: e(0) = 0 = (0 + inc - 1) / inc, remainder inc - 1; one block adds 32 * 65536
				const unsigned uinc = P.voices[U.voice].src_inc;
				N.inc = uinc;
				N.eQ = (unsigned) (kBlock << 16) / uinc;
				N.eR = (unsigned) (kBlock << 16) % uinc;
				N.eq = 0;
				N.erem = uinc - 1;
				it = 0;
				break;
			}
		}
		alive = it >= 0;
		valid = it >= 0 && it - kStages >= 0 && it - kStages < N.nblocks;
		if (valid) {
			nInternal = N.U.n_internal; inc = N.inc; phase = (int) (N.U.out_begin & 31);
			// the block entering the SRC stage: advance the output counter by one block, derive its rows
			const long long e0 = N.eq;
			unsigned rem = N.erem + N.eR;
			long long e1 = e0 + N.eQ;
			if (rem >= inc) { rem -= inc; e1 += 1; }
			N.eq = e1; N.erem = rem;
			src_rows(N, it - kStages, e0, e1);
		}
	}
	const unsigned any = __ballot_sync(0xffffffffu, alive);
	// SRC stage alignment: all slots that have a block at it - 6 are at the same block of equally long
	// utterances with the same SRC increment -> one shared SRC task per 32 outputs
	const unsigned vmask = __ballot_sync(0xffffffffu, valid);
	const int refLane = vmask ? __ffs((int) vmask) - 1 : 0;
	const int itRef = __shfl_sync(0xffffffffu, it, refLane, 32);
	const long long nRef = __shfl_sync(0xffffffffu, nInternal, refLane, 32);
	const unsigned incRef = __shfl_sync(0xffffffffu, inc, refLane, 32);
	const int phaseRef = __shfl_sync(0xffffffffu, phase, refLane, 32);
	const int same = !valid || (it == itRef && nInternal == nRef && inc == incRef && phase == phaseRef);
	const bool allSame = __ballot_sync(0xffffffffu, same) == 0xffffffffu;
	__syncwarp();                                  // lane 0 reads what the reference slot's lane wrote into ctl[q]
	if (lane == 0) {
		CtaSm::Sched& D = C->sched[q];
		D.live = any != 0;
		D.src_mask = (int) vmask;
		D.src_shared = (allSame && __popc(vmask) >= 2) ? 1 : 0;
		D.src_tasks = kSlots;
		if (D.src_shared) {
			const SlotSm::Ctl& RK = C->slot[refLane].ctl[q];
			long long k0 = RK.src_k0;
			const long long k1 = RK.src_k1;
			const long long a = RK.U.out_begin & 31;
			D.src_k0w = k0;
			k0 = ((k0 + a) & ~31ll) - a;
			D.src_k0 = k0;
			D.src_k1 = k1;
			D.src_tasks = (int) ((k1 - k0 + 31) / 32);
		}
	}
}

// One task of the per-iteration task list: the coefficient task of every slot first (task t runs on worker t in
// every iteration, which is what lets its cursor stay in registers), then the SRC tasks.
GTTS_DEV void run_task(CtaSm* C, const KernelParamsV1& P, int lane, int task, int p, CoefRegs& cr)
{
	GTTS_SKIP_BITS(P);
	if (task < kSlots) {
		if (!(skip & 2)) coef_task(C, &C->slot[task], P, lane, task, p, cr);
	} else if (task < kSlots + C->sched[p].src_tasks) {
		if (!(skip & 1)) {
			if (C->sched[p].src_shared) src_shared_task(C, P, lane, task - kSlots, p);
			else src_task(C, &C->slot[task - kSlots], P, lane, p);
		}
	}
}

// Every role runs its OWN iteration loop (one CTA barrier per iteration each), so that a warp keeps only
// its own role's state in registers: ~80 registers per thread instead of 128, which is what lets 24 warps
// share the register file.
#ifdef GTTS_ROLE_PROFILE
#define GTTS_ROLE_LOOP(BODY)                                                     \
	{                                                                            \
		int p = 0;                                                               \
		long long busy = 0, iters = 0, lastIn = 0;                               \
		int live = C->sched[0].live;                                             \
		while (live) {                                                           \
			const long long tStart = GTTS_CLOCK();                               \
			BODY                                                                 \
			const long long tEnd = GTTS_CLOCK();                                 \
			busy += tEnd - tStart;                                               \
			iters += 1;                                                          \
			__syncthreads();                                                     \
			p ^= 1;                                                              \
			live = C->sched[p].live;                                             \
			long long tRel;                                                      \
			GTTS_CLOCK_AFTER(tRel, live);                                        \
			if (tRel - tEnd < 300) lastIn += 1;                                  \
		}                                                                        \
		role_profile(P, warp, lane, busy, iters, lastIn);                        \
	}
#else
// Default build: no per-role cycle counters (they cost 2.7 % through the instruction cache alone).  Build with
// -DGTTS_ROLE_PROFILE (tools/ab_build.sh NAME -DGTTS_ROLE_PROFILE) for GTTS_PROFILE=1 to report them.
#define GTTS_ROLE_LOOP(BODY)                                                     \
	{                                                                            \
		int p = 0;                                                               \
		while (C->sched[p].live) {                                               \
			BODY                                                                 \
			__syncthreads();                                                     \
			p ^= 1;                                                              \
		}                                                                        \
	}
#endif

// GTTS_PROFILE=1: per CTA [0, kWarps) busy cycles of every role, [kWarps] iterations, then per role the number of
// iterations in which it was among the last to reach the barrier (it waited less than 200 cycles there)
GTTS_DEV void role_profile(const KernelParamsV1& P, int warp, int lane, long long busy, long long iters, long long lastIn)
{
#ifndef GTTS_EMU
	if (P.prof != nullptr && lane == 0) {
		long long* row = P.prof + (size_t) blockIdx.x * (2 * kWarps + 1);
		row[warp] = busy;
		row[kWarps + 1 + warp] = lastIn;
		if (warp == 0) row[kWarps] = iters;
	}
#else
	(void) P; (void) warp; (void) lane; (void) busy; (void) iters; (void) lastIn;
#endif
}

GTTS_DEV void tube_v1_cta_body(const KernelParamsV1& P, unsigned char* smem, int tid)
{
	CtaSm* C = reinterpret_cast<CtaSm*>(smem);
	const int lane = tid & 31;
	// Hardware warp w runs on SM sub-partition w % 4 and the roles slow each other down through its issue slots
	// and FP64 pipe, so which role runs where matters by a few percent (measured with tools/ab_build.sh
	// -DGTTS_ROLE_PRESET=n): chain A (role 4) trades places with the lightest worker (role 23, the third SRC row).
#ifndef GTTS_ROLE_PRESET
#define GTTS_ROLE_PRESET 1      // measured: 0: 2.132 ms, 1: 2.111 ms, 2: 2.128 ms, 3: 2.119 ms (profile_run 1036 x 200)
#endif
#if GTTS_ROLE_PRESET == 0
#define GTTS_ROLE_TABLE 0, 1, 2, 3, 23, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 4
#elif GTTS_ROLE_PRESET == 1      // chain B next to tube 0, the light worker next to tube 1
#define GTTS_ROLE_TABLE 0, 1, 2, 3, 5, 23, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 4
#elif GTTS_ROLE_PRESET == 2      // a helper of sub-partition 1 trades places with a coefficient worker of sub-partition 2
#define GTTS_ROLE_TABLE 0, 1, 2, 3, 23, 5, 6, 7, 8, 9, 10, 11, 12, 18, 14, 15, 16, 17, 13, 19, 20, 21, 22, 4
#else                            // both
#define GTTS_ROLE_TABLE 0, 1, 2, 3, 5, 23, 6, 7, 8, 9, 10, 11, 12, 18, 14, 15, 16, 17, 13, 19, 20, 21, 22, 4
#endif
	const int hw = tid >> 5;
	int warp;
	{
		constexpr int roleOfHw[kWarps] = {GTTS_ROLE_TABLE};
		warp = roleOfHw[hw];
	}
	for (int i = tid; i < kSrcFilterLen; i += kThreads) C->tab[i] = P.src_tab[i];
	if (tid < kSlots) {
		for (int b = 0; b < 2; ++b) { C->slot[tid].ctl[b].it = -1; C->slot[tid].ctl[b].voice = 0; C->slot[tid].ctl[b].nblocks = 0; }
	}
	__syncthreads();
	if (warp == kChainAWarp) schedule_slots(C, P, lane, 1, true);      // fills ctl[0] / sched[0]
	__syncthreads();

	// Warp roles (warp id % 4 selects the SM sub-partition; the heavy issuers are spread evenly):
	//   0-3 tube, 4 chain A (+ slot bookkeeping), 5 chain B, 6-12 slot helpers, 13-23 task workers.
	// Task t of an iteration goes to worker t mod 11: with 9-14 tasks nearly every worker has one.
	GTTS_SKIP_BITS(P);
	if (warp < kTubeWarps) {
		TubeCell tl = {0.0, 0.0, 0.0, 0.0};
		GTTS_ROLE_LOOP(if (!(skip & 32)) tube_iteration(C, P, warp, lane, tl, p);)
	} else if (warp == kChainAWarp) {
		ChainARegs ca = {0.0};
		GTTS_ROLE_LOOP(if (!(skip & 8)) chain_a_iteration(C, P, lane, ca, p); schedule_slots(C, P, lane, p, false);)
	} else if (warp == kChainA2Warp) {
		ChainA2Regs ca2 = {{0.0, 0.0, 0.0, 0.0}};
		GTTS_ROLE_LOOP(if (!(skip & 8)) chain_a2_iteration(C, P, lane, ca2, p);)
	} else if (warp == kChainBWarp) {
		ChainBRegs cb = {0.0, 0.0};
		GTTS_ROLE_LOOP(if (!(skip & 16)) chain_b_iteration(C, P, lane, cb, p);)
	} else if (warp < kPool0) {
		HelperRegs hr = {};
		hr.mult = c_lcg[lane];
		SlotSm* S = &C->slot[warp - kHelper0];
		GTTS_ROLE_LOOP(if (!(skip & 4)) helper_iteration(C, S, P, lane, hr, p);)
	} else {
		const int worker = warp - kPool0;
		CoefRegs cr = {0.f, 0.f, 0.f, 0.f, 0, 0};
		GTTS_ROLE_LOOP(for (int task = worker; task < C->sched[p].src_tasks + kSlots; task += kPoolWarps) run_task(C, P, lane, task, p, cr);)
	}
}

#ifndef GTTS_EMU
__global__ void __launch_bounds__(kThreads, 1) tube_kernel_v1(const KernelParamsV1 P)
{
	extern __shared__ __align__(16) unsigned char gtts_smem_v1[];
	tube_v1_cta_body(P, gtts_smem_v1, threadIdx.x);
}
#endif

inline size_t smem_bytes() { return sizeof(CtaSm); }

} // namespace v1
} // namespace gtts
#endif
