// Device code of the tube path (sm_100a).  One warp owns one utterance at a time and walks it in
// frame-aligned blocks of <= 32 internal samples; inside a block the work is split into
//   time-parallel stages (lane = sample):  parameter conversion, noise, wavetable lookup, FIR, mixing, SRC
//   serial stages (the true recurrences):   float32 parameter interpolation (lane = parameter),
//                                            oscillator phase, bandpass biquad, the tube itself
//                                            (lane = pair of scattering junctions, waves in registers,
//                                            neighbours by warp shuffle), radiation / throat IIRs
// All stages communicate through the per-warp shared-memory block `WarpSm`.
//
// Arithmetic contract: every float32 operation of the reference is reproduced
// bit-exactly (__fadd_rn/__fsub_rn/__fmul_rn), the noise generator is reproduced bit-exactly (as the
// integer LCG it is), every float64 operation is IEEE double in the reference's order of
// evaluation, with FMA contraction allowed (the reference itself is built FMA-contracted).
//
// The file is also compiled for the host by tests/simt_emu (GTTS_EMU) so that the kernel logic can be
// checked against the oracle without a GPU; that build is test infrastructure, not a product path.
#ifndef GTTS_TUBE_KERNEL_CUH_
#define GTTS_TUBE_KERNEL_CUH_

#include "tube_types.h"

#ifndef GTTS_EMU
#include <cuda_runtime.h>
#define GTTS_DEV __device__ __forceinline__
#define GTTS_DEV_NOINLINE __device__ __noinline__
#define GTTS_CONST __constant__
// pow(2,x) / pow(10,x) of the reference (VTMUtil.h:50-84) for the argument ranges of this path (|x| < 16: note
// numbers / 12 and dB / 20) without branches or special cases, so that several conversions of one sample
// overlap in one instruction stream: 2^n e^y with n = rint(x log2 b), |y| <= ln(2) / 2, e^y by its degree-13
// Taylor polynomial (relative error 2.2e-16 before rounding; measured against libm in tests/test_gpu_parity.py).
// ~22 instructions each instead of ~45 (exp2 / exp10 of libdevice) or ~100 (pow).
__device__ __forceinline__ double gtts_exp_core(double y)
{
	double p = 1.0 / 6227020800.0;
	p = fma(p, y, 1.0 / 479001600.0);
	p = fma(p, y, 1.0 / 39916800.0);
	p = fma(p, y, 1.0 / 3628800.0);
	p = fma(p, y, 1.0 / 362880.0);
	p = fma(p, y, 1.0 / 40320.0);
	p = fma(p, y, 1.0 / 5040.0);
	p = fma(p, y, 1.0 / 720.0);
	p = fma(p, y, 1.0 / 120.0);
	p = fma(p, y, 1.0 / 24.0);
	p = fma(p, y, 1.0 / 6.0);
	p = fma(p, y, 0.5);
	p = fma(p, y, 1.0);
	p = fma(p, y, 1.0);
	return p;
}
// 10^x for integer x, |x| <= 22: exact 10^|x|, then one IEEE division (out of line: rare)
__device__ __noinline__ double gtts_exp10_int(double x)
{
	double p = 1.0;
	for (int i = (int) fabs(x); i > 0; --i) p *= 10.0;
	return x < 0.0 ? 1.0 / p : p;
}
__device__ __forceinline__ double gtts_exp10(double x)
{
	const double magic = 6755399441055744.0;                     // 2^52 + 2^51: the sum's low word is rint(t)
	// Integer x (40 dB, 20 dB: 10^-1, 10^-2) must give the correctly rounded power like the reference's pow():
	// the amplitude feeds rint(amplitude * tnDelta) in the wavetable (WavetableGlottalSource.h:162-184), and
	// 0.1 * 15 is a tie that one ulp decides.
	const double xi = (x + magic) - magic;
	if (x == xi && fabs(x) <= 22.0) return gtts_exp10_int(x);
	const double tt = fma(x, 3.321928094887362, magic);
	const int n = __double2loint(tt);
	const double nd = tt - magic;
	double r = fma(nd, -0.3010299956639812, x);                  // x - n log10(2), log10(2) in two parts
	r = fma(nd, 2.8037281277851704e-18, r);
	const double y = fma(r, 2.302585092994046, r * -2.1707562233822494e-16);
	return gtts_exp_core(y) * __hiloint2double((n + 1023) << 20, 0);
}
__device__ __forceinline__ double gtts_exp2(double x)
{
	const double magic = 6755399441055744.0;
	const double tt = x + magic;
	const int n = __double2loint(tt);
	const double r = x - (tt - magic);
	const double y = fma(r, 0.6931471805599453, r * 2.3190468138462996e-17);
	return gtts_exp_core(y) * __hiloint2double((n + 1023) << 20, 0);
}
#endif

namespace gtts {

enum {
	kBlock = 32,                // internal samples per block (one per lane)
	kRowStride = 33,            // padded row length of the per-sample arrays (bank-conflict free)
	kSrcRing = 128,             // tube-output ring (doubles), power of two
	kVRing = 64,                // 2x-oversampled oscillator stream rings (even / odd), power of two
	kCurStride = 17,            // padded float row (16 parameters)
	kNoLowMark = 1 << 30,       // wavetable rise segment intact (see stage_lookup)
};

// Rows of WarpSm::row -- one double per sample of the block.
enum {
	R_K0 = 0,                   // R_K0..R_K0+6: oral junction coefficients k[0..6] (VocalTractModel0.h:487-491)
	R_K7 = 7,                   // mouth aperture coefficient (:494-496)
	R_NK0 = 8,                  // velum / first nasal junction (:509-511)
	R_AL = 9, R_AU = 10,        // 3-way junction alphas (left == right) (:500-506)
	R_PA = 11, R_PB = 12,       // frication injected at tap ip and ip+1: tap * bandpassed noise (:524-552)
	R_IN = 13,                  // glottal + aspiration input to the tube, already * 0.125 (:437)
	R_THR = 14,                 // throat input, pulse * 0.125 (:441)
	R_OSC = 15,                 // wavetable increment per half sample (WavetableGlottalSource.h:196-199)
	R_AX = 16, R_AH1 = 17,
	R_TAPA = 18, R_TAPB = 19,   // (1-c)*fricAmp, c*fricAmp
	R_BPB0 = 20, R_BPA1 = 21, R_BPA2 = 22,
	R_SIG = 23,                 // noise signal fed to the bandpass
	R_ENDM = 24, R_ENDN = 25,   // T[S10], NT[N6] (previous-sample state) for the radiation filters
	R_POST0 = 26, R_POST1 = 27, R_POST2 = 28,
	R_P0 = 29, R_P1 = 30,       // oscillator positions of the two half samples
	R_COUNT = 31
};

struct WarpSm {
	double table[kTableLen];            // glottal wavetable of the current utterance
	double row[R_COUNT][kRowStride];
	double ve[kVRing], vo[kVRing];      // 2x stream, even / odd phase, indexed by internal sample & 63
	double xring[kSrcRing];             // tube output (SRC input), indexed by internal sample & 127
	float  cur[kBlock][kCurStride];     // interpolated float32 parameters of the block
	int    ip[kBlock];                  // frication tap index
};

struct KernelParams {
	const VoiceDev* voices;
	const UttDesc* utts;
	const int32_t* order;               // processing order (longest first)
	const float* frames;
	float* out;
	UttState* states;
	const double2* src_tab;             // {h, deltaH}[3328]
	int32_t* queue;                     // atomic work counter
	int32_t n_utt;
};

#ifndef GTTS_EMU
GTTS_CONST double c_fir[kFirMaxTaps];
GTTS_CONST unsigned long long c_lcg[kBlock];       // 377^(j+1) mod 2^44
GTTS_CONST unsigned long long c_lcg_init;          // state s with 377 s = (first seed after 0.7892347) on the 2^-44 grid
#else
extern double c_fir[kFirMaxTaps];
extern unsigned long long c_lcg[kBlock];
extern unsigned long long c_lcg_init;
#endif

// ---- small helpers ---------------------------------------------------------------------------------

GTTS_DEV double shfl_d(double v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }

// Util::amplitude60dB (VTMUtil.h:50-67)
#ifdef GTTS_AMP60_NOINLINE
GTTS_DEV_NOINLINE double amp60(double db)
#else
GTTS_DEV double amp60(double db)
#endif
{
	if (db <= 0.0) return 0.0;
	if (db == 60.0) return 1.0;
	return gtts_exp10((db - 60.0) * (1.0 / 20.0));
}

// sin(x) and cos(x) for the bandpass coefficients (x = pi bw / fs, 2 pi cf / fs: a few radians at most; valid for
// |x| < 1e5): reduction by pi / 2 in two parts, Taylor cores on |r| <= pi / 4 (degree 17 / 18), branch-free.
// ~33 instructions for both values instead of the ~170 of libdevice's tan() + cos() with their special-case paths.
#ifndef GTTS_EMU
#ifdef GTTS_SINCOS_NOINLINE
__device__ __noinline__ void gtts_sincos(double x, double& s, double& c)
#else
__device__ __forceinline__ void gtts_sincos(double x, double& s, double& c)
#endif
{
	const double magic = 6755399441055744.0;
	const double tt = fma(x, 0.6366197723675814, magic);
	const int n = __double2loint(tt);
	const double nd = tt - magic;
	double r = fma(nd, -1.5707963267948966, x);
	r = fma(nd, -6.123233995736766e-17, r);
	const double r2 = r * r;
	double ps = 1.0 / 355687428096000.0;                        // 1 / 17!
	ps = fma(ps, r2, -1.0 / 1307674368000.0);
	ps = fma(ps, r2, 1.0 / 6227020800.0);
	ps = fma(ps, r2, -1.0 / 39916800.0);
	ps = fma(ps, r2, 1.0 / 362880.0);
	ps = fma(ps, r2, -1.0 / 5040.0);
	ps = fma(ps, r2, 1.0 / 120.0);
	ps = fma(ps, r2, -1.0 / 6.0);
	const double sp = fma(ps * r2, r, r);
	double pc = -1.0 / 6402373705728000.0;                       // -1 / 18!
	pc = fma(pc, r2, 1.0 / 20922789888000.0);
	pc = fma(pc, r2, -1.0 / 87178291200.0);
	pc = fma(pc, r2, 1.0 / 479001600.0);
	pc = fma(pc, r2, -1.0 / 3628800.0);
	pc = fma(pc, r2, 1.0 / 40320.0);
	pc = fma(pc, r2, -1.0 / 720.0);
	pc = fma(pc, r2, 1.0 / 24.0);
	pc = fma(pc, r2, -0.5);
	const double cp = fma(pc, r2, 1.0);
	// quadrant n mod 4: (s, c) = (sp, cp), (cp, -sp), (-sp, -cp), (-cp, sp)
	const bool swap = n & 1;
	const double a = swap ? cp : sp, b = swap ? sp : cp;
	s = (n & 2) ? -a : a;
	c = ((n + 1) & 2) ? -b : b;
}
#endif

// (a - b) / (a + b): scattering coefficient from two squared radii.
GTTS_DEV double kcoef(double a2, double b2) { return (a2 - b2) / (a2 + b2); }

// a / b for well-scaled operands (squared radii: 1e-4 .. 1e2), branch-free: hardware reciprocal estimate, two
// Newton steps, one residual correction -- 8 instructions without the special-case path of the IEEE division, so
// that the ten divisions of a sample overlap instead of running one after the other.  Within 1 ulp of a / b.
GTTS_DEV double div_fast(double a, double b)
{
#ifndef GTTS_EMU
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
	double e = fma(-b, r, 1.0);
	r = fma(r, e, r);
	e = fma(-b, r, 1.0);
	r = fma(r, e, r);
	const double q = a * r;
	return fma(fma(-b, q, a), r, q);
#else
	return a / b;
#endif
}

// ---- stage: float32 interpolation, lane = parameter (Controller.cpp:297-311) ----------------------------
// Writes cur[j][k] for j < nb and advances the running value by nb sequential float additions.
GTTS_DEV void stage_interp(WarpSm* S, int lane, int nb, float& cur, float delta)
{
	if (lane < kNumParams) {
		for (int j = 0; j < nb; ++j) {
			S->cur[j][lane] = cur;
			cur = __fadd_rn(cur, delta);
		}
	}
	__syncwarp();
}

// ---- stage: parameter conversion, lane = sample (VocalTractModel0.h:396-404, 484-552, 698-716) ---------
GTTS_DEV void stage_convert(WarpSm* S, const VoiceDev& V, int lane, int nb)
{
	if (lane < nb) {
		const float* p = S->cur[lane];
		// Util::frequency (VTMUtil.h:76-84) and the oscillator increment of one half sample
		const double f0 = 220.0 * gtts_exp2(((double) p[0] + 3.0) * (1.0 / 12.0));
		S->row[R_OSC][lane] = (f0 / 2.0) * V.basic_inc;
		S->row[R_AX][lane] = amp60((double) p[1]);
		S->row[R_AH1][lane] = amp60((double) p[2]);
		// radii (setAllParameters: max(r * coef, GS_VTM0_MIN_RADIUS)) and tube coefficients
		double r2[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			double r = (double) p[7 + i] * V.radius_coef[i];
			r = r > 0.01 ? r : 0.01;
			r2[i] = r * r;
		}
#pragma unroll
		for (int i = 0; i < 7; ++i) S->row[R_K0 + i][lane] = kcoef(r2[i], r2[i + 1]);
		S->row[R_K7][lane] = kcoef(r2[7], V.ap2);
		const double vel = (double) p[15];
		const double v2 = vel * vel;
		const double sum = 2.0 / (r2[3] + r2[3] + v2);
		S->row[R_AL][lane] = sum * r2[3];
		S->row[R_AU][lane] = sum * v2;
		S->row[R_NK0][lane] = kcoef(v2, V.nr1_2);
		// frication taps
		const double fa = amp60((double) p[3]);
		const double fpos = (double) p[4];
		int ip = (int) fpos;
		const double comp = fpos - ip;
		double ta = (1.0 - comp) * fa, tb = comp * fa;
		if (ip < 0 || ip > 7) { ta = 0.0; tb = 0.0; ip = -50; }
		S->row[R_TAPA][lane] = ta;
		S->row[R_TAPB][lane] = tb;
		S->ip[lane] = ip;
		// bandpass coefficients (BandpassFilter.h:91-110); pure functions of (fs, bw, cf)
		const double pi = 3.14159265358979323846;
		const double tv = tan(pi * (double) p[6] * V.Ts);
		const double cv = cos(2.0 * pi * (double) p[5] * V.Ts);
		const double a2 = (1.0 - tv) / (1.0 + tv);
		S->row[R_BPA2][lane] = a2;
		S->row[R_BPA1][lane] = -(1.0 + a2) * cv;
		S->row[R_BPB0][lane] = 0.5 - 0.5 * a2;
	}
	__syncwarp();
}

// ---- stage: noise, lane = sample (NoiseSource.h:40-44, NoiseFilter.h:63-68) -------------------------
// seed' = frac(seed * 377) in double is exactly s' = 377 s mod 2^44 on the 2^-44 grid (every product
// < 512 is representable), so lane j jumps ahead with 377^(j+1) mod 2^44.  Off the grid (only the
// initial seed 0.7892347) the recurrence is stepped in double.  Returns the low-passed noise of
// sample `lane`; seed / x1 are warp-uniform state.
GTTS_DEV double stage_noise(int lane, int nb, double& seed, double& x1)
{
	const double two44 = 17592186044416.0;
	const double scaled = seed * two44;
	double mine;
	if (scaled == rint(scaled)) {
		const unsigned long long s = (unsigned long long) scaled;
		const unsigned long long sj = (s * c_lcg[lane]) & ((1ull << 44) - 1);
		mine = (double) sj * (1.0 / two44);
	} else {
		double sd = seed;
		mine = 0.0;
		for (int j = 0; j < nb; ++j) {
			const double prod = __dmul_rn(sd, 377.0);
			sd = __dsub_rn(prod, (double) (int) prod);
			if (j == lane) mine = sd;
		}
	}
	const double n = mine - 0.5;
	double prev = shfl_d(n, (lane + 31) & 31, 32);
	if (lane == 0) prev = x1;
	x1 = shfl_d(n, nb - 1, 32);
	seed = shfl_d(mine, nb - 1, 32);
	return n + prev;
}

// ---- stage: oscillator phase, serial (WavetableGlottalSource.h:196-199, 265-272) --------------------
GTTS_DEV void stage_phase(WarpSm* S, int lane, int nb, double& pos)
{
	double p0 = 0.0, p1 = 0.0;
	for (int j = 0; j < nb; ++j) {
		const double inc = S->row[R_OSC][j];
		double s = pos + inc;
		pos = (s > 511.0) ? s - 512.0 : s;
		const double a = pos;
		s = pos + inc;
		pos = (s > 511.0) ? s - 512.0 : s;
		if (j == lane) { p0 = a; p1 = pos; }
	}
	if (lane < nb) { S->row[R_P0][lane] = p0; S->row[R_P1][lane] = p1; }
	__syncwarp();
}

// Wavetable entry i.  The stored table is exact for the rise segment, the closed part and (when the
// fall time does not depend on the amplitude) the fall segment; with tn_min != tn_max the fall
// segment is a function of the current amplitude (WavetableGlottalSource.h:162-184) and is evaluated
// analytically: 1 - ((i - div1) / (newDiv2 - div1))^2 below newDiv2, 0 up to div2.
GTTS_DEV double table_at(const WarpSm* S, const VoiceDev& V, unsigned i, bool dynamic, double nd2, double inv, int low)
{
	if (dynamic && i >= (unsigned) V.div1 && i < (unsigned) V.div2) {
		if (i >= (unsigned) nd2) return 0.0;
		const double x = (double) (int) (i - V.div1) * inv;
		return 1.0 - (x * x);
	}
	if ((int) i >= low && i < (unsigned) V.div1) return 0.0;      // rise segment zeroed by an earlier closure point below div1
	return S->table[i & (kTableLen - 1)];     // in range already unless the pitch is absurd: no read outside the table then either
}

// ---- stage: wavetable lookup of both half samples, lane = sample (:212-228) --------------------------
// `low`: the lowest closure point seen so far if one fell below div1 (glottal volume above 60 dB with
// tn_min != tn_max: setup() then zeroes [newDiv2, div2), part of the rise segment, for good; :176-183).
GTTS_DEV void stage_lookup(WarpSm* S, const VoiceDev& V, int lane, int nb, long long n0, int& low)
{
	const bool dynamic = (V.waveform == 0) && (V.tn_delta != 0.0);
	double nd2 = 0.0, inv = 0.0;
	int lowHere = low;
	if (dynamic) {
		const double ax = S->row[R_AX][lane < nb ? lane : 0];
		nd2 = (double) V.div2 - rint(ax * V.tn_delta);
		nd2 = nd2 > 0.0 ? nd2 : 0.0;
		inv = 1.0 / (nd2 - (double) V.div1);
		int mine = (lane < nb && nd2 < (double) V.div1) ? (int) nd2 : kNoLowMark;
		for (int dlt = 1; dlt < 32; dlt <<= 1) {
			const int o = __shfl_up_sync(0xffffffffu, mine, dlt, 32);
			if (lane >= dlt) mine = mine < o ? mine : o;
		}
		lowHere = mine < low ? mine : low;
		low = __shfl_sync(0xffffffffu, lowHere, 31, 32);
	}
	if (lane < nb) {
		double v[2];
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			const double pos = S->row[s ? R_P1 : R_P0][lane];
			const unsigned lo = __double2uint_rz(pos);
			const unsigned up = (lo + 1 > 511u) ? lo + 1 - 512u : lo + 1;
			const double tl = table_at(S, V, lo, dynamic, nd2, inv, lowHere);
			const double tu = table_at(S, V, up, dynamic, nd2, inv, lowHere);
			v[s] = tl + ((pos - (double) lo) * (tu - tl));
		}
		const int slot = (int) ((n0 + lane) & (kVRing - 1));
		S->ve[slot] = v[0];
		S->vo[slot] = v[1];
	}
	__syncwarp();
}

// ---- stage: decimating FIR + source mixing, lane = sample -------------------------------------------
// (WavetableGlottalSourceFIRFilter.h:276-304; VocalTractModel0.h:408-438)
GTTS_DEV void stage_fir_mix(WarpSm* S, const VoiceDev& V, int lane, int nb, long long n0, double lp)
{
	if (lane < nb) {
		// y = sum_i c[i] * x2[2n+1-i], i ascending from an accumulator of 0
		const long long n = n0 + lane;
		double acc = 0.0;
#pragma unroll
		for (int i = 0; i < kFirTaps; ++i) {
			// x2 index 2n+1-i: i even -> odd phase of sample n - i/2; i odd -> even phase of n - (i-1)/2
			const int idx = (int) ((n - ((i & 1) ? (i - 1) / 2 : i / 2)) & (kVRing - 1));
			const double x = (i & 1) ? S->ve[idx] : S->vo[idx];
			acc += x * c_fir[i];
		}
		double pulse = acc;
		const double ax = S->row[R_AX][lane];
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
		S->row[R_SIG][lane] = sig;
		S->row[R_IN][lane] = (pulse + (S->row[R_AH1][lane] * sig)) * 0.125;
		S->row[R_THR][lane] = pulse * 0.125;
	}
	__syncwarp();
}

// ---- stage: frication bandpass, serial (BandpassFilter.h:114-122) -----------------------------------
struct BandpassState { double x1, x2, y1, y2; };

GTTS_DEV void stage_bandpass(WarpSm* S, int lane, int nb, BandpassState& b)
{
	double mine = 0.0;
	for (int j = 0; j < nb; ++j) {
		const double x = S->row[R_SIG][j];
		const double y = S->row[R_BPB0][j] * (x - b.x2) - S->row[R_BPA1][j] * b.y1 - S->row[R_BPA2][j] * b.y2;
		b.x2 = b.x1; b.x1 = x; b.y2 = b.y1; b.y1 = y;
		if (j == lane) mine = y;
	}
	if (lane < nb) {
		S->row[R_PA][lane] = S->row[R_TAPA][lane] * mine;
		S->row[R_PB][lane] = S->row[R_TAPB][lane] * mine;
	}
	__syncwarp();
}

// ---- stage: the tube, serial over samples, 8 lanes per utterance (VocalTractModel0.h:565-661) ---------
// Lane g = lane & 7 owns two "cells"; a cell reads two waves of the previous sample and writes two
// waves of the next one.  A cell = a 2-port scattering junction
//        dl = k (Tin - Bin);  Tout = (Tin + dl) d + tap * fric;  Bout = (Bin + dl) d
//   g   cell A (k, tap)                    cell B
//   0   S1-S2  (k0, -)  + glottis input    S2-S3 (k1, FC1)
//   1   S3-S4  (k2, FC2)                   3-way junction S4/S5/velum (alphas, FC3)
//   2   S5-S6  (k3, FC4)                   S6-S7 pure damped delay (k = 0, FC5)
//   3   S7-S8  (k4, FC6)                   S8-S9 (k5, FC7)
//   4   S9-S10 (k6, FC8)                   mouth: reflection filter on k7 * T[S10]
//   5   N1-N2  (velum coefficient)         N2-N3
//   6   N3-N4                              N4-N5
//   7   N5-N6                              nose: reflection filter on nk5 * NT[N6]
// Wiring per sample: A.Tout -> B.Tin and B.Bout -> A.Bin stay in the lane; B.Tout -> next lane's A.Tin
// and A.Bout -> previous lane's B.Bin go through one shuffle each; the nasal branch (lane 1 <-> lane 5)
// needs a third.  The radiation filters are feed-forward from T[S10] / NT[N6] and run in stage_post.
struct TubeLane {
	double aT, aB, bT, bB;      // inputs of cell A and cell B (state of the previous sample)
	double extra;               // g0: B[S1] (glottis end), g1: NB[N1] (velum branch), g4/g7: reflection y1
};

struct TubeRole {
	int rowA, rowB;             // rows of the per-sample coefficient (or -1: constant)
	double constA, constB;
	int tapA, tapB;             // frication tap index fed into Tout of the cell (or -100)
	double refl_b0, refl_a1;    // end lanes
	int g;
};

GTTS_DEV TubeRole tube_role(const VoiceDev& V, int lane)
{
	TubeRole r;
	const int g = lane & 7;
	r.g = g;
	r.rowA = -1; r.rowB = -1; r.constA = 0.0; r.constB = 0.0; r.tapA = -100; r.tapB = -100;
	r.refl_b0 = 0.0; r.refl_a1 = 0.0;
	switch (g) {
	case 0: r.rowA = R_K0 + 0; r.rowB = R_K0 + 1; r.tapB = 0; break;
	case 1: r.rowA = R_K0 + 2; r.tapA = 1; r.tapB = 2; break;
	case 2: r.rowA = R_K0 + 3; r.tapA = 3; r.constB = 0.0; r.tapB = 4; break;
	case 3: r.rowA = R_K0 + 4; r.rowB = R_K0 + 5; r.tapA = 5; r.tapB = 6; break;
	case 4: r.rowA = R_K0 + 6; r.rowB = R_K7; r.tapA = 7; r.refl_b0 = V.refl_b0_m; r.refl_a1 = V.refl_a1_m; break;
	case 5: r.rowA = R_NK0; r.constB = V.nasal_k[1]; break;
	case 6: r.constA = V.nasal_k[2]; r.constB = V.nasal_k[3]; break;
	default: r.constA = V.nasal_k[4]; r.constB = V.nasal_k[5]; r.refl_b0 = V.refl_b0_n; r.refl_a1 = V.refl_a1_n; break;
	}
	return r;
}

// one internal sample of the tube (the waves of sample j from those in `t`)
GTTS_DEV void tube_step(WarpSm* S, const TubeRole& R, double d, int g, int base, int rowA, int rowB, int j, TubeLane& t)
{
	const double kA = R.rowA < 0 ? R.constA : S->row[rowA][j];
	const double kB = R.rowB < 0 ? R.constB : S->row[rowB][j];
	const int ip = S->ip[j];
	const double pa = S->row[R_PA][j], pb = S->row[R_PB][j];
	const double tfA = (R.tapA == ip) ? pa : ((R.tapA == ip + 1) ? pb : 0.0);
	const double tfB = (R.tapB == ip) ? pa : ((R.tapB == ip + 1) ? pb : 0.0);

	// cell A: always a 2-port junction
	const double dlA = kA * (t.aT - t.aB);
	const double aTo = ((t.aT + dlA) * d) + tfA;
	const double aBo = (t.aB + dlA) * d;

	double bTo, bBo, linkOut = aBo;
	if (g == 1) {
		// 3-way junction (:595-604): bT = T[S4], bB = B[S5], extra = NB[N1]
		const double aL = S->row[R_AL][j], aU = S->row[R_AU][j];
		const double jp = (aL * t.bT) + (aL * t.bB) + (aU * t.extra);
		bBo = (jp - t.bT) * d;
		bTo = ((jp - t.bB) * d) + tfB;
		linkOut = (jp - t.extra) * d;              // NT[N1] of the next sample
	} else if (g == 4 || g == 7) {
		// open end (:634-636, 653-654): reflection lowpass, y = b0 x - a1 y1
		if (g == 4) S->row[R_ENDM][j] = t.bT; else S->row[R_ENDN][j] = t.bT;
		const double y = R.refl_b0 * (kB * t.bT) - R.refl_a1 * t.extra;
		t.extra = y;
		bBo = d * y;
		bTo = 0.0;
	} else {
		const double dlB = kB * (t.bT - t.bB);
		bTo = ((t.bT + dlB) * d) + tfB;
		bBo = (t.bB + dlB) * d;
	}

	// exchange with the neighbouring lanes
	const double fromPrev = shfl_d(bTo, base + ((g + 7) & 7), 32);       // B.Tout of lane g-1
	const double fromNext = shfl_d(aBo, base + ((g + 1) & 7), 32);       // A.Bout of lane g+1
	const double link = shfl_d(linkOut, base + ((g == 1) ? 5 : 1), 32);  // lane 1 <-> lane 5
	double nextAT = fromPrev;
	if (g == 0) {
		nextAT = (t.extra * d) + S->row[R_IN][j];   // T[S1] = B[S1] d + input (:572-573)
		t.extra = aBo;                             // B[S1] of the next sample
	} else if (g == 5) {
		nextAT = link;                             // NT[N1] from the 3-way junction
	} else if (g == 1) {
		t.extra = link;                            // NB[N1] from lane 5's cell A
	}
	t.aT = nextAT;
	t.aB = bBo;
	t.bT = aTo;
	t.bB = fromNext;
}

GTTS_DEV void stage_tube(WarpSm* S, const VoiceDev& V, const TubeRole& R, int lane, int nb, TubeLane& t)
{
	const double d = V.damping;
	const int g = R.g;
	const int base = lane & ~7;
	const int rowA = R.rowA < 0 ? 0 : R.rowA, rowB = R.rowB < 0 ? 0 : R.rowB;
	for (int j = 0; j < nb; ++j) tube_step(S, R, d, g, base, rowA, rowB, j, t);
	__syncwarp();
}

// Model 3 (VocalTractModel2<double, 3>, VocalTractModel2.h:234-268, 626-670): every section is a delay line of three
// samples -- what a step writes is read three steps later -- so the wave state is three interleaved copies of the
// model-0 state, each advanced every third sample; t[0] is the copy of the current sample.  The reflection filters
// of the two open ends run on every sample: their state (`extra` of lanes 4 and 7) is shared by the copies.
GTTS_DEV void stage_tube_delay3(WarpSm* S, const VoiceDev& V, const TubeRole& R, int lane, int nb, TubeLane* t)
{
	const double d = V.damping;
	const int g = R.g;
	const int base = lane & ~7;
	const int rowA = R.rowA < 0 ? 0 : R.rowA, rowB = R.rowB < 0 ? 0 : R.rowB;
	const bool isEnd = g == 4 || g == 7;
	for (int j = 0; j < nb; ++j) {
		TubeLane cur = t[0];
		tube_step(S, R, d, g, base, rowA, rowB, j, cur);
		t[0] = t[1];
		t[1] = t[2];
		t[2] = cur;
		if (isEnd) { t[0].extra = cur.extra; t[1].extra = cur.extra; }
	}
	__syncwarp();
}

// Model 4 (VocalTractModel4<double, 1>, VocalTractModel4.h:671-744): 30 oropharynx sections on lanes 0..29, the 18
// nasal sections on lanes 12..29 (N1 shares a lane with S13 -- the three waves of the velum junction are where the four
// neighbour shuffles of a sample put them -- and N18 shares lane 29 with S30: both open ends on one lane).  The
// junctions of model 0 (pressure waves) at the region boundaries S3|S4, S5|S6, S9|S10, S15|S16, S21|S22, S25|S26,
// S27|S28 and every third nasal section; sections inside a region are plain copies without damping (:303-307), except
// S18 -> S19, which carries frication tap FC5 and is damped (:308-312).  Coefficients, taps and end filters are the
// ones the model-0 stages of this kernel produce.
struct Tube4Lane { double oT, oB, nT, nB, yM, yN; };

GTTS_DEV void stage_tube4(WarpSm* S, const VoiceDev& V, int lane, int nb, Tube4Lane& t)
{
	const double d = V.damping;
	// boundary (lane - 1 | lane): junction row / damping / frication tap of this section's top wave; boundary
	// (lane | lane + 1): junction row / damping of its bottom wave
	int rowL = -1, rowR = -1, tapL = -100;
	double dL = 1.0, dR = 1.0;
	{
		const int jl[7] = {2, 4, 8, 14, 20, 24, 26};          // left section of J1..J7
		const int jt[7] = {-100, 0, 1, 3, 5, 6, 7};           // frication tap injected behind the junction (FC1 at J2 ... FC8 at J7)
#pragma unroll
		for (int q = 0; q < 7; ++q) {
			if (jl[q] + 1 == lane) { rowL = R_K0 + q; dL = d; tapL = jt[q]; }
			if (jl[q] == lane) { rowR = R_K0 + q; dR = d; }
		}
		if (lane == 18) { dL = d; tapL = 4; }                 // S18 -> S19: damped, FC5
		if (lane == 17) dR = d;
		if (lane == 12) tapL = 2;                             // FC3 behind the velum junction
	}
	const int ni = lane - 12;                                 // nasal section (valid 0..17)
	double nkLc = 0.0, nkRc = 0.0, ndL = 1.0, ndR = 1.0;
	bool nkLrow = false, nkRrow = false;
	if (ni >= 0 && ni < 18) {
		if (ni >= 3 && ni % 3 == 0) { const int q = ni / 3 - 1; ndL = d; if (q == 0) nkLrow = true; else nkLc = V.nasal_k[q]; }
		if (ni % 3 == 2 && ni < 17) { const int q = ni / 3; ndR = d; if (q == 0) nkRrow = true; else nkRc = V.nasal_k[q]; }
	}
	const int prev = (lane + 31) & 31, next = (lane + 1) & 31;
	for (int j = 0; j < nb; ++j) {
		const double kL = rowL >= 0 ? S->row[rowL][j] : 0.0;
		const double kR = rowR >= 0 ? S->row[rowR][j] : 0.0;
		const double nkL = nkLrow ? S->row[R_NK0][j] : nkLc;
		const double nkR = nkRrow ? S->row[R_NK0][j] : nkRc;
		const int ip = S->ip[j];
		const double tf = (tapL == ip) ? S->row[R_PA][j] : ((tapL == ip + 1) ? S->row[R_PB][j] : 0.0);
		const double oTL = shfl_d(t.oT, prev, 32), oBR = shfl_d(t.oB, next, 32);
		const double nTL = shfl_d(t.nT, prev, 32), nBR = shfl_d(t.nB, next, 32);
		// propagate / propagateJunction (VocalTractModel4.h:303-318)
		double noT = (oTL + kL * (oTL - t.oB)) * dL + tf;
		double noB = (oBR + kR * (t.oT - oBR)) * dR;
		double nnT = (nTL + nkL * (nTL - t.nB)) * ndL;
		double nnB = (nBR + nkR * (t.nT - nBR)) * ndR;
		if (lane == 0) {
			noT = t.oB * d + S->row[R_IN][j];                                   // :676
		} else if (lane == 11) {
			// 3-way junction (:319-328), S12 side
			const double aL = S->row[R_AL][j], aU = S->row[R_AU][j];
			const double jp = aL * t.oT + aL * oBR + aU * nBR;
			noB = (jp - t.oT) * d;
		} else if (lane == 12) {
			const double aL = S->row[R_AL][j], aU = S->row[R_AU][j];
			const double jp = aL * oTL + aL * t.oB + aU * t.nB;
			noT = (jp - t.oB) * d + tf;
			nnT = (jp - t.nB) * d;
		} else if (lane == 29) {
			// both open ends (:712-716, 737-741): reflection low-pass; the radiation filters run in stage_post
			S->row[R_ENDM][j] = t.oT;
			S->row[R_ENDN][j] = t.nT;
			const double ym = V.refl_b0_m * (S->row[R_K7][j] * t.oT) - V.refl_a1_m * t.yM;
			t.yM = ym;
			noB = d * ym;
			const double yn = V.refl_b0_n * (V.nasal_k[5] * t.nT) - V.refl_a1_n * t.yN;
			t.yN = yn;
			nnB = d * yn;
		}
		t.oT = noT; t.oB = noB; t.nT = nnT; t.nB = nnB;
	}
	__syncwarp();
}

// ---- stage: radiation filters + throat, serial, then the output sum (lane = sample) -----------------
// (RadiationFilter.h:73-79 on (1 + k) T; Throat.h:80-85; VocalTractModel0.h:639-641, 657-660, 441)
struct PostState { double x1, y1; };    // lane 0 mouth radiation, lane 1 nose radiation, lane 2 throat

GTTS_DEV void stage_post(WarpSm* S, const VoiceDev& V, int lane, int nb, long long n0, PostState& p)
{
	if (lane < 3) {
		const double b0 = lane == 0 ? V.rad_m : (lane == 1 ? V.rad_n : V.throat_b0);
		const double b1 = lane == 0 ? -V.rad_m : (lane == 1 ? -V.rad_n : 0.0);
		const double a1 = lane == 0 ? -V.rad_m : (lane == 1 ? -V.rad_n : V.throat_a1);
		const double gain = lane == 2 ? V.throat_gain : 1.0;
		const int inRow = lane == 0 ? R_ENDM : (lane == 1 ? R_ENDN : R_THR);
		const double onePlusN = 1.0 + V.nasal_k[5];
		for (int j = 0; j < nb; ++j) {
			double x = S->row[inRow][j];
			if (lane == 0) x = (1.0 + S->row[R_K7][j]) * x;
			else if (lane == 1) x = onePlusN * x;
			const double y = b0 * x + b1 * p.x1 - a1 * p.y1;
			p.x1 = x;
			p.y1 = y;
			S->row[R_POST0 + lane][j] = y * gain;
		}
	}
	__syncwarp();
	if (lane < nb) {
		const double s = (S->row[R_POST0][lane] + S->row[R_POST1][lane]) + S->row[R_POST2][lane];
		S->xring[(n0 + lane) & (kSrcRing - 1)] = s;
	}
	__syncwarp();
}

// ---- stage: sample-rate conversion, lane = output sample (SampleRateConverter.h:295-361) -------------
// Output k is centred on ring position e = (k inc) >> 16 (input index e - 13) with phase
// f = (k inc) & 0xFFFF; left wing taps h[L + 256 j] on x[e-13-j], right wing taps (from ~f) on
// x[e-12+j]; one accumulator, left wing first.  Produces outputs [kDone, kEnd).
GTTS_DEV void stage_src(const WarpSm* S, const double2* tab, const VoiceDev& V, int lane,
			long long kDone, long long kEnd, float* out, long long nEnd)
{
	for (long long k = kDone + lane; k < kEnd; k += 32) {
		const unsigned long long t = (unsigned long long) k * V.src_inc;
		const int e = (int) (t >> 16);
		const unsigned f = (unsigned) (t & 0xFFFFu);
		double acc = 0.0;
		if (!V.src_upsample) {
			// down-sampling (SampleRateConverter.h:362-415): input i sits at ring position pad + i; phases from
			// rint(f ratio), advancing by phaseIncrement per tap while the filter index stays inside the table
			unsigned ph = (unsigned) rint((double) f * V.src_ratio);
			long long pos = (long long) e - V.src_pad;
			unsigned ii;
			while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
				const double2 c = tab[ii];
				acc += (pos < nEnd ? S->xring[(int) (pos & (kSrcRing - 1))] : 0.0) * (c.x + (c.y * ((double) (ph & 0xFFu) / 256)));
				pos -= 1;
				ph += V.src_phase_inc;
			}
			ph = (unsigned) rint((double) ((~f) & 0xFFFFu) * V.src_ratio);
			pos = (long long) e - V.src_pad + 1;
			while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
				const double2 c = tab[ii];
				acc += (pos < nEnd ? S->xring[(int) (pos & (kSrcRing - 1))] : 0.0) * (c.x + (c.y * ((double) (ph & 0xFFu) / 256)));
				pos += 1;
				ph += V.src_phase_inc;
			}
			out[k] = (float) acc;
			continue;
		}
		{
			const double interp = (double) (f & 0xFFu) / 256;
			const unsigned L = f >> 8;
#pragma unroll
			for (int j = 0; j < kSrcZeroCrossings; ++j) {
				const double2 c = tab[L + 256 * j];
				const double x = S->xring[(e - 13 - j) & (kSrcRing - 1)];
				acc += (x * (c.x + (c.y * interp)));
			}
		}
		{
			const unsigned gph = (~f) & 0xFFFFu;
			const double interp = (double) (gph & 0xFFu) / 256;
			const unsigned L = gph >> 8;
#pragma unroll
			for (int j = 0; j < kSrcZeroCrossings; ++j) {
				const double2 c = tab[L + 256 * j];
				const double x = S->xring[(e - 12 + j) & (kSrcRing - 1)];
				acc += (x * (c.x + (c.y * interp)));
			}
		}
		out[k] = (float) acc;
	}
}

// ---- wavetable construction (WavetableGlottalSource.h:111-136) --------------------------------------
GTTS_DEV void build_table(WarpSm* S, const VoiceDev& V, int lane)
{
	for (int i = lane; i < kTableLen; i += 32) {
		double v;
		if (V.waveform == 0) {
			if (i < V.div1) {
				const double x = (double) i / (double) V.div1;
				const double x2 = x * x;
				const double x3 = x2 * x;
				v = (3.0 * x2) - (2.0 * x3);
			} else if (i < V.div2) {
				const double x = (double) (i - V.div1) / V.tn_length;
				v = 1.0 - (x * x);
			} else {
				v = 0.0;
			}
		} else {
			v = sin(((double) i / kTableLen) * 2.0 * 3.14159265358979323846);
		}
		S->table[i] = v;
	}
}

// ---- per-utterance driver: one warp, all stages in sequence -----------------------------------------
// TM: the voice's tube (VoiceDev::tube_model): 0 models 0 / 2, 3 model 3, 4 model 4.  Streaming state (UttState) exists
// for TM == 0 only; the host refuses streams of the other models.
template<int TM>
GTTS_DEV void run_utterance(WarpSm* S, const double2* tab, const KernelParams& P, const UttDesc& U, int lane)
{
	const VoiceDev V = P.voices[U.voice];
	const bool resume = (U.flags & 1) != 0;
	const bool noFlush = (U.flags & 2) != 0;
	const bool lookahead = (U.flags & 4) != 0;   // frames[n_frames] exists and is the end of the last period
	UttState* st = (U.state_index >= 0) ? &P.states[U.state_index] : nullptr;

	// state
	double seed = 0.7892347, noiseX1 = 0.0, pos = 0.0;
	BandpassState bp = {0.0, 0.0, 0.0, 0.0};
	TubeLane tl = {0.0, 0.0, 0.0, 0.0, 0.0};
	TubeLane tl3[TM == 3 ? 3 : 1];
	for (int q = 0; q < (TM == 3 ? 3 : 1); ++q) tl3[q] = tl;
	Tube4Lane tl4 = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
	PostState ps = {0.0, 0.0};
	long long nDone = 0, kDone = 0;
	int tableLow = kNoLowMark;
	const TubeRole role = tube_role(V, lane);
	const int g = lane & 7;

	build_table(S, V, lane);
	for (int i = lane; i < kVRing; i += 32) { S->ve[i] = 0.0; S->vo[i] = 0.0; }
	for (int i = lane; i < kSrcRing; i += 32) S->xring[i] = 0.0;
	__syncwarp();

	if (resume && st != nullptr && st->started) {
		seed = st->seed; noiseX1 = st->noise_x1; pos = st->pos;
		bp.x1 = st->bp_x1; bp.x2 = st->bp_x2; bp.y1 = st->bp_y1; bp.y2 = st->bp_y2;
		nDone = st->n_in_done; kDone = st->n_out_done;
		tableLow = st->table_low;
		// waves: cell inputs from the section arrays
		const double* ot = st->oral_t; const double* ob = st->oral_b;
		const double* nt = st->nasal_t; const double* nbw = st->nasal_b;
		switch (g) {
		case 0: tl.aT = ot[0]; tl.aB = ob[1]; tl.bT = ot[1]; tl.bB = ob[2]; tl.extra = ob[0]; break;
		case 1: tl.aT = ot[2]; tl.aB = ob[3]; tl.bT = ot[3]; tl.bB = ob[4]; tl.extra = nbw[0]; break;
		case 2: tl.aT = ot[4]; tl.aB = ob[5]; tl.bT = ot[5]; tl.bB = ob[6]; break;
		case 3: tl.aT = ot[6]; tl.aB = ob[7]; tl.bT = ot[7]; tl.bB = ob[8]; break;
		case 4: tl.aT = ot[8]; tl.aB = ob[9]; tl.bT = ot[9]; tl.extra = st->refl_y1_m; break;
		case 5: tl.aT = nt[0]; tl.aB = nbw[1]; tl.bT = nt[1]; tl.bB = nbw[2]; break;
		case 6: tl.aT = nt[2]; tl.aB = nbw[3]; tl.bT = nt[3]; tl.bB = nbw[4]; break;
		default: tl.aT = nt[4]; tl.aB = nbw[5]; tl.bT = nt[5]; tl.extra = st->refl_y1_n; break;
		}
		if (lane == 0) { ps.x1 = st->rad_x1_m; ps.y1 = st->rad_y1_m; }
		else if (lane == 1) { ps.x1 = st->rad_x1_n; ps.y1 = st->rad_y1_n; }
		else if (lane == 2) { ps.x1 = 0.0; ps.y1 = st->throat_y1; }
		// histories: the last 24 internal samples of the 2x stream, the last 26 tube outputs
		if (lane < 24) {
			const long long n = nDone - 24 + lane;
			S->ve[n & (kVRing - 1)] = st->fir_hist[2 * lane];
			S->vo[n & (kVRing - 1)] = st->fir_hist[2 * lane + 1];
		}
		for (int i = lane; i < 64; i += 32) S->xring[(nDone - 64 + i) & (kSrcRing - 1)] = st->src_hist[i];      // 2 pad <= 64 inputs back
		__syncwarp();
	}

	float* out = P.out + U.out_begin;
	const long long nOutTotal = U.n_out;
	const long long nEnd = nDone + U.n_internal;      // inputs of this utterance (stream: up to the end of this chunk)
	const float* frames = P.frames + U.frame_begin * kNumParams;

	for (long long p = 0; p < U.n_frames; ++p) {
		// Controller.cpp:297-300
		float cur = 0.0f, delta = 0.0f;
		if (lane < kNumParams) {
			cur = frames[p * kNumParams + lane];
			const bool haveNext = (p + 1 < U.n_frames) || lookahead;
			const float nxt = haveNext ? frames[(p + 1) * kNumParams + lane] : cur;
			delta = __fmul_rn(__fsub_rn(nxt, cur), U.inv_steps);
		}
		for (int off = 0; off < U.steps; off += kBlock) {
			const int nb = (U.steps - off) < kBlock ? (U.steps - off) : kBlock;
			stage_interp(S, lane, nb, cur, delta);
			stage_convert(S, V, lane, nb);
			const double lp = stage_noise(lane, nb, seed, noiseX1);
			stage_phase(S, lane, nb, pos);
			stage_lookup(S, V, lane, nb, nDone, tableLow);
			stage_fir_mix(S, V, lane, nb, nDone, lp);
			stage_bandpass(S, lane, nb, bp);
			if (TM == 3) stage_tube_delay3(S, V, role, lane, nb, tl3);
			else if (TM == 4) stage_tube4(S, V, lane, nb, tl4);
			else stage_tube(S, V, role, lane, nb, tl);
			stage_post(S, V, lane, nb, nDone, ps);
			nDone += nb;
			// outputs whose right wing is complete: (k inc) >> 16 <= nDone - 1
			long long kEnd = (long long) ((((unsigned long long) nDone << 16) + V.src_inc - 1) / V.src_inc);
			if (kEnd > nOutTotal) kEnd = nOutTotal;
			stage_src(S, tab, V, lane, kDone, kEnd, out, nEnd);
			if (kEnd > kDone) kDone = kEnd;
			__syncwarp();
		}
	}

	if (!noFlush) {
		// flushBuffer(): 2*pad zeros, then everything that is left (SampleRateConverter.h:462-471).  The down-sampling
		// converter takes the inputs past the end as zero instead (up to 96 zeros do not fit the ring next to its window).
		if (V.src_upsample) for (int i = lane; i < 2 * V.src_pad; i += 32) S->xring[(nDone + i) & (kSrcRing - 1)] = 0.0;
		__syncwarp();
		stage_src(S, tab, V, lane, kDone, nOutTotal, out, nEnd);
		kDone = nOutTotal;
	}

	if (st != nullptr) {
		// save (streaming): section-indexed waves so that the layout is independent of the lane mapping
		double* ot = st->oral_t; double* ob = st->oral_b; double* nt = st->nasal_t; double* nbw = st->nasal_b;
		if (lane < 8) {
			switch (g) {
			case 0: ot[0] = tl.aT; ob[1] = tl.aB; ot[1] = tl.bT; ob[2] = tl.bB; ob[0] = tl.extra; break;
			case 1: ot[2] = tl.aT; ob[3] = tl.aB; ot[3] = tl.bT; ob[4] = tl.bB; nbw[0] = tl.extra; break;
			case 2: ot[4] = tl.aT; ob[5] = tl.aB; ot[5] = tl.bT; ob[6] = tl.bB; break;
			case 3: ot[6] = tl.aT; ob[7] = tl.aB; ot[7] = tl.bT; ob[8] = tl.bB; break;
			case 4: ot[8] = tl.aT; ob[9] = tl.aB; ot[9] = tl.bT; st->refl_y1_m = tl.extra; break;
			case 5: nt[0] = tl.aT; nbw[1] = tl.aB; nt[1] = tl.bT; nbw[2] = tl.bB; break;
			case 6: nt[2] = tl.aT; nbw[3] = tl.aB; nt[3] = tl.bT; nbw[4] = tl.bB; break;
			default: nt[4] = tl.aT; nbw[5] = tl.aB; nt[5] = tl.bT; st->refl_y1_n = tl.extra; break;
			}
		}
		if (lane == 0) {
			st->rad_x1_m = ps.x1; st->rad_y1_m = ps.y1;
			st->seed = seed; st->noise_x1 = noiseX1; st->pos = pos;
			st->bp_x1 = bp.x1; st->bp_x2 = bp.x2; st->bp_y1 = bp.y1; st->bp_y2 = bp.y2;
			st->n_in_done = nDone; st->n_out_done = kDone; st->started = 1;
			st->table_low = tableLow;
		} else if (lane == 1) {
			st->rad_x1_n = ps.x1; st->rad_y1_n = ps.y1;
		} else if (lane == 2) {
			st->throat_y1 = ps.y1;
		}
		if (lane < 24) {
			const long long n = nDone - 24 + lane;
			st->fir_hist[2 * lane] = S->ve[n & (kVRing - 1)];
			st->fir_hist[2 * lane + 1] = S->vo[n & (kVRing - 1)];
		}
		for (int i = lane; i < 64; i += 32) st->src_hist[i] = S->xring[(nDone - 64 + i) & (kSrcRing - 1)];
	}
	__syncwarp();
}

// ---- CTA body: warps are independent workers pulling utterances (longest first) from a queue ---------
GTTS_DEV void tube_cta_body(const KernelParams& P, unsigned char* smem, int tid, int nthreads)
{
	double2* tab = reinterpret_cast<double2*>(smem);
	WarpSm* S = reinterpret_cast<WarpSm*>(smem + sizeof(double2) * kSrcFilterLen) + (tid >> 5);
	const int lane = tid & 31;
	for (int i = tid; i < kSrcFilterLen; i += nthreads) tab[i] = P.src_tab[i];
	__syncthreads();
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(P.queue, 1);
		q = __shfl_sync(0xffffffffu, q, 0, 32);
		if (q >= P.n_utt) break;
		const UttDesc U = P.utts[P.order[q]];
		const int tm = P.voices[U.voice].tube_model;      // warp-uniform
		if (tm == 3) run_utterance<3>(S, tab, P, U, lane);
		else if (tm == 4) run_utterance<4>(S, tab, P, U, lane);
		else run_utterance<0>(S, tab, P, U, lane);
	}
}

#ifndef GTTS_EMU
template<int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) tube_kernel_v0(const KernelParams P)
{
	extern __shared__ __align__(16) unsigned char gtts_smem[];
	tube_cta_body(P, gtts_smem, threadIdx.x, kWarps * 32);
}
#endif

inline size_t tube_smem_bytes(int warps) { return sizeof(double2) * kSrcFilterLen + sizeof(WarpSm) * warps; }

} // namespace gtts
#endif
