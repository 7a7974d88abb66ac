// tube5_kernel -- the reference's model 5 (VocalTractModel5<double, 1>, the voice its documentation uses by
// default) on sm_100a: one warp per utterance, frame-aligned blocks of <= 32 internal samples.
//
//   lane = sample    parameter conversion (VocalTractModel5.h:527-533, 610-630, 776-792: 2^x, 10^x, 9 junction
//                    divisions, the mouth radiation impedance from R8 -- cos, sqrt, 2 divisions per sample --, bandpass
//                    coefficients), noise by LCG jump-ahead, the Rosenberg-B pulse value from its phase, mixing
//   lane = filter    the serial recurrences: the pulse phase (one lane); the three Butterworth low-passes -- glottal
//                    noise, frication noise, glottal wave -- on three lanes in one instruction stream; the frication
//                    bandpass (one lane)
//   lane = section   the tube (:646-730): 30 oropharynx sections on lanes 0..29 and the 21 nasal sections on lanes
//                    12..31, 0 -- N1 shares a lane with S13, so the three waves of the velum junction are already
//                    where the four neighbour shuffles of a sample put them.  Flow-equation junctions and plain
//                    damped delays are one formula (coefficient 0 for a delay: (x - 0 * s) * d is x * d exactly).
//   lane = output    down-sampling windowed-sinc converter (SampleRateConverter.h:362-415), then the float32
//                    difference filter * output rate of the output callback (:497-513, DifferenceFilter.h:62-69)
//
// Arithmetic: IEEE double in the reference's order of evaluation, FMA contraction allowed; 2^x / 10^x / sin / cos and
// the coefficient divisions by the branch-free forms of tube_kernel.cuh (<= 2 ulp; tan as sin / cos folded into the one
// division that follows it); sqrt and the pulse-shape divisions exact.  Noise generator and float32 interpolation
// bit-exact.
//
// Also compiled for the host by tests/simt_emu (GTTS_EMU): test infrastructure, never linked into the product.
#ifndef GTTS_TUBE5_KERNEL_CUH_
#define GTTS_TUBE5_KERNEL_CUH_

#include "tube_kernel.cuh"
#include "tube5_types.h"

#ifndef GTTS_M5_TUBE_CHUNK
#define GTTS_M5_TUBE_CHUNK 1     // measured on B200: 1 -> 46.9 k audio-s/s, 2 -> 44.6 k, 4 -> 42.8 k, 8 -> 33.2 k (registers, not latency, are what is short)
#endif

namespace gtts {
namespace m5 {

enum {
	kRow = 33,
	kRing = 128,                // tube-output ring (doubles)
	kYRing = 64,                // raw converter outputs (float) for the difference filter
	kTubeChunk = GTTS_M5_TUBE_CHUNK,   // samples whose operands the tube loop fetches at once
};

enum {
	R_DT = 0,                   // f0 / fs
	R_GA, R_AA, R_FA,           // glottal, aspiration, frication amplitude
	R_J0,                       // R_J0 .. R_J0 + 6: oropharynx junctions J1..J7
	R_VL = R_J0 + 7, R_VR, R_VU,   // velum junction
	R_NJ1,
	R_CT1, R_CT2, R_CT3, R_CR2, R_CR3,   // mouth radiation impedance
	R_BPA2, R_BPA1, R_BPB0,
	R_FL, R_FR,                 // frication weights of sections S6 + fo, S6 + fo + 1
	R_FVL, R_FVR,               // frication values injected into those sections
	R_MINGL, R_DGL,
	R_NT2,                      // t2 the source takes at its next wrap
	R_NOISE, R_GN, R_FN,
	R_T, R_T2,                  // pulse phase and fall end per sample
	R_PULSE,                    // source value, then the low-passed pulse
	R_IN, R_FRIC, R_GL,
	R_FLOWM, R_FLOWN,
	R_COUNT
};

struct WarpSm5 {
	Voice5Dev V;                        // the current utterance's voice constants
	double row[R_COUNT][kRow];
	double xring[kRing];
	float  yraw[kYRing];
	float  cur[kBlock][kCurStride];
	int    fo[kBlock];
};

struct KernelParams5 {
	const Voice5Dev* voices;
	const UttDesc* utts;
	const int32_t* order;
	const float* frames;
	float* out;
	const double2* src_tab;
	int32_t* queue;
	int32_t n_utt;
};

GTTS_DEV void stage_convert5(WarpSm5* S, const Voice5Dev& V, int lane, int nb)
{
	if (lane < nb) {
		const float* p = S->cur[lane];
		const double f0 = 220.0 * gtts_exp2(((double) p[0] + 3.0) * (1.0 / 12.0));
		S->row[R_DT][lane] = f0 / V.fs;
		const double ga = amp60((double) p[1]);
		S->row[R_GA][lane] = ga;
		S->row[R_AA][lane] = amp60((double) p[2]);
		S->row[R_FA][lane] = amp60((double) p[3]);
		double r[8], r2[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const double v = (double) p[7 + i] * V.radius_coef[i];
			r[i] = v > 0.01 ? v : 0.01;
			r2[i] = r[i] * r[i];
		}
#pragma unroll
		for (int i = 0; i < 7; ++i) S->row[R_J0 + i][lane] = div_fast(r2[i] - r2[i + 1], r2[i] + r2[i + 1]);
		const double vel = (double) p[15];
		const double v2 = vel * vel;
		{
			// Junction3::configure (:281-292) with left == right == R4
			const double c = div_fast(1.0, r2[3] + r2[3] + v2);
			S->row[R_VL][lane] = c * (r2[3] - r2[3] - v2);
			S->row[R_VR][lane] = c * (r2[3] - r2[3] - v2);
			S->row[R_VU][lane] = c * (v2 - r2[3] - r2[3]);
		}
		S->row[R_NJ1][lane] = div_fast(v2 - V.nr2_2, v2 + V.nr2_2);
		if (!V.const_mouth) {
			// PoleZeroRadiationImpedance::update (PoleZeroRadiationImpedance.h:143-176), radius in metres
			const double radius = r[7] * (double) 1.0e-2f;
			const double rr = radius < 0.5e-2 ? 0.5e-2 : radius;
			const double transFreq = div_fast(62.3371, rr) + 320.204;
#ifndef GTTS_EMU
			double sinWT, cosWT;
			gtts_sincos((2.0 * 3.14159265358979323846) * transFreq * V.Ts, sinWT, cosWT);
#else
			const double cosWT = cos((2.0 * 3.14159265358979323846) * transFreq * V.Ts);
#endif
			const double qa = 2.0 * cosWT;
			const double qb = -2.0 * (cosWT + 1.0);
			const double qc = cosWT + 1.0;
			const double delta = qb * qb - 4.0 * qa * qc;
			double a = div_fast(-qb - sqrt(delta), 2.0 * qa);
			const double b = 2.0 * a - 1.0;
			if (radius < 0.5e-2) a *= 40391.2 * (radius * radius);
			const double coef = div_fast(1.0, a + 1.0);
			S->row[R_CT1][lane] = (a + b) * coef;
			S->row[R_CT2][lane] = 2.0 * coef;
			S->row[R_CT3][lane] = -2.0 * b * coef;
			S->row[R_CR2][lane] = (a - 1.0) * coef;
			S->row[R_CR3][lane] = (b - a) * coef;
		} else {
			S->row[R_CT1][lane] = V.rad_m[0]; S->row[R_CT2][lane] = V.rad_m[1]; S->row[R_CT3][lane] = V.rad_m[2];
			S->row[R_CR2][lane] = V.rad_m[3]; S->row[R_CR3][lane] = V.rad_m[4];
		}
		{
			// BandpassFilter::update (BandpassFilter.h:88-110)
			const double pi = 3.14159265358979323846;
#ifndef GTTS_EMU
			// a2 = (1 - tan x) / (1 + tan x) = (cos x - sin x) / (cos x + sin x): one division
			double sx, cx, sy, cv;
			gtts_sincos(pi * (double) p[6] * V.Ts, sx, cx);
			gtts_sincos(2.0 * pi * (double) p[5] * V.Ts, sy, cv);
			const double a2 = div_fast(cx - sx, cx + sx);
#else
			const double tv = tan(pi * (double) p[6] * V.Ts);
			const double cv = cos(2.0 * pi * (double) p[5] * V.Ts);
			const double a2 = (1.0 - tv) / (1.0 + tv);
#endif
			S->row[R_BPA2][lane] = a2;
			S->row[R_BPA1][lane] = -(1.0 + a2) * cv;
			S->row[R_BPB0][lane] = 0.5 - 0.5 * a2;
		}
		{
			// frication position (:711-715): offset into S6..S28
			const double fricOffset = 22.0 * ((double) p[4] / 7.0);
			int fo = (int) fricOffset;
			const double fr = fricOffset - fo;
			double fl = 1.0 - fr, frr = fr;
			if (fo < 0 || fo > 22) { fo = -100; fl = 0.0; frr = 0.0; }    // outside the tube: undefined in the reference
			S->fo[lane] = fo;
			S->row[R_FL][lane] = fl;
			S->row[R_FR][lane] = frr;
		}
		{
			const double minGl = 1.0 - ga * V.min_loss, maxGl = 1.0 - ga * V.max_loss;
			S->row[R_MINGL][lane] = minGl;
			S->row[R_DGL][lane] = maxGl - minGl;
		}
		// RosenbergBGlottalSource::setup (:112-120): a pure function of the amplitude
		S->row[R_NT2][lane] = V.t1 + V.tn_max - ga * (V.tn_max - V.tn_min);
	}
	__syncwarp();
}

// noise (NoiseSource.h:40-44): exact integer LCG mod 2^44, lane j jumps ahead by 377^(j + 1)
GTTS_DEV void stage_noise5(WarpSm5* S, int lane, int nb, unsigned long long& lcg)
{
	const unsigned long long sj = (lcg * c_lcg[lane]) & ((1ull << 44) - 1);
	if (lane < nb) S->row[R_NOISE][lane] = (double) sj * (1.0 / 17592186044416.0) - 0.5;
	lcg = __shfl_sync(0xffffffffu, sj, nb - 1, 32);
	__syncwarp();
}

struct Serial5 {
	double x1, x2, y1, y2;      // lane 0: glottal-noise low-pass, lane 1: frication-noise low-pass, lane 2: glottal low-pass
	double t, t2;               // lane 0: pulse phase and the end of its fall
	double bpX1, bpX2, bpY1, bpY2;   // lane 0: frication bandpass
};

// RosenbergBGlottalSource::getSample (:122-150): the phase and the fall end each sample sees, one lane
GTTS_DEV void stage_phase5(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane == 0) {
		const bool dynamic = V.waveform == 0 && V.tn_min != V.tn_max;
		double t = s.t, t2 = s.t2;
		for (int j0 = 0; j0 < nb; j0 += 4) {
			double dt[4], nt2[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) { dt[q] = S->row[R_DT][j0 + q]; nt2[q] = S->row[R_NT2][j0 + q]; }
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (j0 + q < nb) {
					S->row[R_T][j0 + q] = t;
					S->row[R_T2][j0 + q] = t2;
					t += dt[q];
					if (t > 1.0) {
						t -= 1.0;
						if (dynamic) t2 = nt2[q];
					}
				}
			}
		}
		s.t = t; s.t2 = t2;
	}
	__syncwarp();
}

GTTS_DEV void stage_pulse_value(WarpSm5* S, const Voice5Dev& V, int lane, int nb)
{
	if (lane < nb) {
		const double t = S->row[R_T][lane], t2 = S->row[R_T2][lane];
		double value;
		if (V.waveform == 0) {
			if (t < V.t1) {
				const double x = t / V.t1;
				value = (x * x) * (3.0 - 2.0 * x);
			} else if (t < t2) {
				const double x = (t - V.t1) / (t2 - V.t1);
				value = 1.0 - x * x;
			} else {
				value = 0.0;
			}
		} else {
			value = sin(t * (2.0 * 3.14159265358979323846));
		}
		S->row[R_PULSE][lane] = value;
	}
	__syncwarp();
}

// The three Butterworth low-passes, one per lane, in ONE instruction stream: the first-order ones
// (Butterworth1LowpassFilter.h:89-96, y = b0 (x + x1) - a1 y1) are the second-order form
// (Butterworth2LowpassFilter.h:104-113, y = b0 (x + x2) + b1 x1 - a1 y1 - a2 y2) with b1 = a2 = 0 and x1 in the place of
// x2 -- the extra terms are exact zeros.  Lane 0: noise -> glottal noise, lane 1: noise -> frication noise,
// lane 2: pulse value -> low-passed pulse (in place).
GTTS_DEV void stage_lowpass5(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane < 3) {
		const bool second = lane == 1;
		const double b0 = lane == 0 ? V.gn_b0 : (lane == 1 ? V.fn_b0 : V.gl_b0);
		const double a1 = lane == 0 ? V.gn_a1 : (lane == 1 ? V.fn_a1 : V.gl_a1);
		const double b1 = second ? V.fn_b1 : 0.0, a2 = second ? V.fn_a2 : 0.0;
		const int inRow = lane == 2 ? R_PULSE : R_NOISE;
		const int outRow = lane == 0 ? R_GN : (lane == 1 ? R_FN : R_PULSE);
		double x1 = s.x1, x2 = s.x2, y1 = s.y1, y2 = s.y2;
		for (int j0 = 0; j0 < nb; j0 += 4) {
			double x[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) x[q] = S->row[inRow][j0 + q];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (j0 + q < nb) {
					const double xa = second ? x2 : x1;
					const double y = b0 * (x[q] + xa) + b1 * x1 - a1 * y1 - a2 * y2;
					x2 = x1; x1 = x[q]; y2 = y1; y1 = y;
					S->row[outRow][j0 + q] = y;
				}
			}
		}
		s.x1 = x1; s.x2 = x2; s.y1 = y1; s.y2 = y2;
	}
	__syncwarp();
}

// mixing (VocalTractModel5.h:546-575), lane = sample; returns the bypass signal
GTTS_DEV double stage_mix5(WarpSm5* S, const Voice5Dev& V, int lane, int nb)
{
	double bypassSignal = 0.0;
	if (lane < nb) {
		const double pulse = S->row[R_PULSE][lane];
		const double ga = S->row[R_GA][lane];
		const double pulsedNoise = S->row[R_GN][lane] * pulse;
		const double noisyPulse = ga * (pulse * V.one_minus_breath + pulsedNoise * V.breath);
		double fric = S->row[R_FN][lane];
		if (V.modulation) {
			double crossmix = ga * V.crossmix;
			crossmix = (crossmix < 1.0) ? crossmix : 1.0;
			fric = fric * (noisyPulse * crossmix + (1.0 - crossmix));
		}
		const double in = noisyPulse + S->row[R_AA][lane] * fric;
		S->row[R_IN][lane] = in;
		S->row[R_FRIC][lane] = fric;
		S->row[R_GL][lane] = S->row[R_MINGL][lane] + S->row[R_DGL][lane] * pulse;
		bypassSignal = in;
	}
	__syncwarp();
	return bypassSignal;
}

// frication bandpass (BandpassFilter.h:112-122) and the values injected into the tube
GTTS_DEV void stage_bandpass5(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane == 0) {
		double x1 = s.bpX1, x2 = s.bpX2, y1 = s.bpY1, y2 = s.bpY2;
		for (int j0 = 0; j0 < nb; j0 += 4) {
			double x[4], b0[4], a1[4], a2[4], fa[4], wl[4], wr[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int j = j0 + q;
				x[q] = S->row[R_FRIC][j]; b0[q] = S->row[R_BPB0][j]; a1[q] = S->row[R_BPA1][j]; a2[q] = S->row[R_BPA2][j];
				fa[q] = S->row[R_FA][j]; wl[q] = S->row[R_FL][j]; wr[q] = S->row[R_FR][j];
			}
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (j0 + q < nb) {
					const double y = b0[q] * (x[q] - x2) - a1[q] * y1 - a2[q] * y2;
					x2 = x1; x1 = x[q]; y2 = y1; y1 = y;
					const double fv = fa[q] * (V.fric_factor * y);
					S->row[R_FVL][j0 + q] = fv * wl[q];
					S->row[R_FVR][j0 + q] = fv * wr[q];
				}
			}
		}
		s.bpX1 = x1; s.bpX2 = x2; s.bpY1 = y1; s.bpY2 = y2;
	}
	__syncwarp();
}

struct Tube5 { double oT, oB, nT, nB, in1, outT1, outR1; };

// What a lane does in the tube loop: five per-sample operand rows (a..e) and its role flags.
enum {
	F_KL = 1,                   // a = junction coefficient on the left boundary of the lane's oropharynx section
	F_KR = 2,                   // b = junction coefficient on its right boundary
	F_NKL = 4, F_NKR = 8,       // c = first nasal junction (per sample) on the left / right boundary of the nasal section
	F_GLOT = 16,                // lane 0 (S1): a = glottal loss factor, d = input
	F_V11 = 32,                 // lane 11 (S12): d = velum junction, left coefficient
	F_V12 = 64,                 // lane 12 (S13 and N1): d = right coefficient, e = upper coefficient
	F_MOUTH = 128,              // lane 29 (S30): a..e = radiation impedance cT1, cT2, cT3, cR2, cR3
	F_NOSE = 256,               // lane 0 (N21): constant radiation impedance
};

struct Tube5Role {
	int ra, rb, rc, rd, re;     // rows
	int flags;
	int flowRow;                // where an end lane stores its output flow
	double nkLc, nkRc;          // fixed nasal junction coefficients (0: plain delay)
};

GTTS_DEV Tube5Role tube5_role(const Voice5Dev& V, int lane)
{
	Tube5Role r;
	r.ra = r.rb = r.rc = r.rd = r.re = R_DT;        // any row: the value is not used
	r.flags = 0; r.flowRow = R_FLOWM; r.nkLc = 0.0; r.nkRc = 0.0;
	// oropharynx junctions J1..J7 between S3|S4, S5|S6, S9|S10, S15|S16, S21|S22, S25|S26, S27|S28
	const int jl[7] = {2, 4, 8, 14, 20, 24, 26};
#pragma unroll
	for (int q = 0; q < 7; ++q) {
		if (jl[q] + 1 == lane) { r.ra = R_J0 + q; r.flags |= F_KL; }
		if (jl[q] == lane) { r.rb = R_J0 + q; r.flags |= F_KR; }
	}
	// nasal junctions NJ1..NJ6 between N3|N4, N6|N7, ... N18|N19: NJ1 per sample, the others fixed
	const int ni = (lane - 12) & 31;
	if (ni < kNasal) {
		if (ni >= 3 && ni % 3 == 0) { const int q = ni / 3 - 1; if (q == 0) { r.rc = R_NJ1; r.flags |= F_NKL; } else if (q < 6) r.nkLc = V.nasal_k[q]; }
		if (ni % 3 == 2 && ni < 18) { const int q = ni / 3; if (q == 0) { r.rc = R_NJ1; r.flags |= F_NKR; } else r.nkRc = V.nasal_k[q]; }
	}
	if (lane == 0) { r.ra = R_GL; r.rd = R_IN; r.flags |= F_GLOT | F_NOSE; r.flowRow = R_FLOWN; }
	if (lane == 11) { r.rd = R_VL; r.flags |= F_V11; }
	if (lane == 12) { r.rd = R_VR; r.re = R_VU; r.flags |= F_V12; }
	if (lane == kOral - 1) { r.ra = R_CT1; r.rb = R_CT2; r.rc = R_CT3; r.rd = R_CR2; r.re = R_CR3; r.flags |= F_MOUTH; }
	return r;
}

// The tube, lane = section (see the header of this file), without branches: every lane evaluates the same expressions
// on its own operands, the few special sections (glottis, velum, the two open ends) by selecting operands and
// results.  Operands are fetched four samples at a time, so a shared-memory latency is paid once per four steps.
//   top    (X + cT Y) dT + Z        X = left neighbour's top, Y = X + own bottom, cT = -k, dT = d         propagate / junction
//                                   X = own bottom, cT = 0, dT = glottal loss, Z = input                  S1
//                                   X, Y += N1's bottom, cT = right coefficient                           S13
//   bottom (W + cB V) d             W = right neighbour's bottom, V = own top + W, cB = k; W, V += N1's bottom on S12
GTTS_DEV void stage_tube5(WarpSm5* S, const Voice5Dev& V, const Tube5Role& R, int lane, int nb, Tube5& t)
{
	const double d = V.damping;
	const int fl = R.flags;
	const bool fGlot = fl & F_GLOT, fV11 = fl & F_V11, fV12 = fl & F_V12, fMouth = fl & F_MOUTH, fNose = fl & F_NOSE;
	const bool fEnd = fMouth || fNose;
	const int prev = (lane + 31) & 31, next = (lane + 1) & 31;
	const double n1 = V.rad_n[0], n2 = V.rad_n[1], n3 = V.rad_n[2], n4 = V.rad_n[3], n5 = V.rad_n[4];
	double oT = t.oT, oB = t.oB, nT = t.nT, nB = t.nB, in1 = t.in1, outT1 = t.outT1, outR1 = t.outR1;
	for (int j0 = 0; j0 < nb; j0 += kTubeChunk) {
		double ra[kTubeChunk], rb[kTubeChunk], rc[kTubeChunk], rd[kTubeChunk], re[kTubeChunk], fvl[kTubeChunk], fvr[kTubeChunk];
		int fo[kTubeChunk];
#pragma unroll
		for (int q = 0; q < kTubeChunk; ++q) {
			const int j = j0 + q;
			ra[q] = S->row[R.ra][j]; rb[q] = S->row[R.rb][j]; rc[q] = S->row[R.rc][j]; rd[q] = S->row[R.rd][j]; re[q] = S->row[R.re][j];
			fo[q] = S->fo[j];
			fvl[q] = S->row[R_FVL][j];
			fvr[q] = S->row[R_FVR][j];
		}
#pragma unroll
		for (int q = 0; q < kTubeChunk; ++q) {
			if (j0 + q < nb) {
				const double a = ra[q], b = rb[q], c = rc[q], dd = rd[q], e = re[q];
				const double cT = fV12 ? dd : ((fl & F_KL) ? -a : 0.0);
				const double dT = fGlot ? a : d;
				const double Z = fGlot ? dd : 0.0;
				const double cB = fV11 ? dd : ((fl & F_KR) ? b : 0.0);
				const double nkL = (fl & F_NKL) ? c : R.nkLc;
				const double nkR = (fl & F_NKR) ? c : R.nkRc;
				const double c1 = fMouth ? a : n1, c2 = fMouth ? b : n2, c3 = fMouth ? c : n3, c4 = fMouth ? dd : n4, c5 = fMouth ? e : n5;
				const double oTL = shfl_d(oT, prev, 32), oBR = shfl_d(oB, next, 32);
				const double nTL = shfl_d(nT, prev, 32), nBR = shfl_d(nB, next, 32);
				// oropharynx, top wave
				const double up12 = fV12 ? nB : 0.0;                 // N1's bottom wave at the velum junction, seen from S13
				const double X = (fGlot ? oB : oTL) + up12;
				const double Y = (oTL + oB) + up12;
				double noT = (X + cT * Y) * dT + Z;
				// oropharynx, bottom wave
				const double up11 = fV11 ? nBR : 0.0;                // the same wave seen from S12
				const double W = oBR + up11;
				const double Vv = (oT + oBR) + up11;
				double noB = (W + cB * Vv) * d;
				// nasal tract
				const double nnTg = (nTL - nkL * (nTL + nB)) * d;
				const double nnT12 = ((oTL + oB) + e * Y) * d;      // N1's top wave out of the velum junction
				const double nnT = fV12 ? nnT12 : nnTg;
				double nnB = (nBR + nkR * (nT + nBR)) * d;
				// open ends: PoleZeroRadiationImpedance::process (:178-189), S30 on lane 29, N21 on lane 0
				const double in = fMouth ? oT : nT;
				const double outT = c1 * outT1 + c2 * in + c3 * in1;
				const double outR = c1 * outR1 + c4 * in + c5 * in1;
				in1 = in; outT1 = outT; outR1 = outR;
				const double refl = outR * d;
				noB = fMouth ? refl : noB;
				nnB = fNose ? refl : nnB;
				if (fEnd) S->row[R.flowRow][j0 + q] = outT;
				// frication (:716-723)
				const int f = fo[q];
				const double add = (lane == 5 + f) ? fvl[q] : fvr[q];
				if (lane == 5 + f || (lane == 6 + f && f < 22)) noT += add;
				oT = noT; oB = noB; nT = nnT; nB = nnB;
			}
		}
	}
	t.oT = oT; t.oB = oB; t.nT = nT; t.nB = nB; t.in1 = in1; t.outT1 = outT1; t.outR1 = outR1;
	__syncwarp();
}

// Down-sampling converter (SampleRateConverter.h:362-415), lane = output, outputs [kDone, kEnd); then the output
// callback: float difference filter y = x - x[k - 2], times the output rate (VocalTractModel5.h:506-512).
GTTS_DEV void stage_src5(WarpSm5* S, const double2* tab, const Voice5Dev& V, int lane, long long kDone, long long kEnd,
			long long nEnd, float* out)
{
	for (long long k0 = kDone; k0 < kEnd; k0 += 32) {
		const long long k = k0 + lane;
		const bool live = k < kEnd;
		if (live) {
			const unsigned long long tt = (unsigned long long) k * V.src_inc;
			const long long e = (long long) (tt >> 16);
			const unsigned f = (unsigned) (tt & 0xFFFFu);
			double acc = 0.0;
			unsigned ph = (unsigned) rint((double) f * V.src_ratio);
			long long pos = e - V.src_pad;
			unsigned ii;
			while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
				const double2 c = tab[ii];
				const double x = (pos >= 0 && pos < nEnd) ? S->xring[(int) (pos & (kRing - 1))] : 0.0;
				acc += x * (c.x + (c.y * ((double) (ph & 0xFFu) / 256)));
				pos -= 1;
				ph += V.src_phase_inc;
			}
			ph = (unsigned) rint((double) ((~f) & 0xFFFFu) * V.src_ratio);
			pos = e - V.src_pad + 1;
			while ((ii = (ph >> 8)) < (unsigned) kSrcFilterLen) {
				const double2 c = tab[ii];
				const double x = (pos >= 0 && pos < nEnd) ? S->xring[(int) (pos & (kRing - 1))] : 0.0;
				acc += x * (c.x + (c.y * ((double) (ph & 0xFFu) / 256)));
				pos += 1;
				ph += V.src_phase_inc;
			}
			S->yraw[k & (kYRing - 1)] = (float) acc;
		}
		__syncwarp();
		if (live) {
			const float x = S->yraw[k & (kYRing - 1)];
			if (V.bypass == 1) {
				out[k] = x;
			} else {
				const float x2 = k >= 2 ? S->yraw[(k - 2) & (kYRing - 1)] : 0.0f;
				const float y = __fsub_rn(x, x2);
				out[k] = (float) ((double) y * V.output_rate);
			}
		}
		__syncwarp();
	}
}

GTTS_DEV void run_utterance5(WarpSm5* S, const double2* tab, const KernelParams5& P, const UttDesc& U, int lane)
{
	{
		const double* src = reinterpret_cast<const double*>(&P.voices[U.voice]);
		double* dst = reinterpret_cast<double*>(&S->V);
		for (int i = lane; i < (int) (sizeof(Voice5Dev) / sizeof(double)); i += 32) dst[i] = src[i];
		__syncwarp();
	}
	const Voice5Dev& V = S->V;
	unsigned long long lcg = c_lcg_init;
	Serial5 s;
	s.x1 = s.x2 = s.y1 = s.y2 = 0.0;
	s.t = 0.0;
	s.t2 = V.t1 + V.tn_max;                         // RosenbergBGlottalSource.h:76-78
	s.bpX1 = s.bpX2 = s.bpY1 = s.bpY2 = 0.0;
	Tube5 t = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
	const Tube5Role role = tube5_role(V, lane);
	long long nDone = 0, kDone = 0;
	for (int i = lane; i < kRing; i += 32) S->xring[i] = 0.0;
	for (int i = lane; i < kYRing; i += 32) S->yraw[i] = 0.0f;
	__syncwarp();
	float* out = P.out + U.out_begin;
	const float* frames = P.frames + U.frame_begin * kNumParams;
	const long long nEnd = U.n_internal;
	for (long long p = 0; p < U.n_frames; ++p) {
		// Controller.cpp:297-300
		float cur = 0.0f, delta = 0.0f;
		if (lane < kNumParams) {
			cur = frames[p * kNumParams + lane];
			const float nxt = (p + 1 < U.n_frames) ? frames[(p + 1) * kNumParams + lane] : cur;
			delta = __fmul_rn(__fsub_rn(nxt, cur), U.inv_steps);
		}
		for (int off = 0; off < U.steps; off += kBlock) {
			const int nb = (U.steps - off) < kBlock ? (U.steps - off) : kBlock;
			if (lane < kNumParams) {
				for (int j = 0; j < nb; ++j) {
					S->cur[j][lane] = cur;
					cur = __fadd_rn(cur, delta);
				}
			}
			__syncwarp();
			stage_convert5(S, V, lane, nb);
			stage_noise5(S, lane, nb, lcg);
			stage_phase5(S, V, lane, nb, s);
			stage_pulse_value(S, V, lane, nb);
			stage_lowpass5(S, V, lane, nb, s);
			const double bypassSignal = stage_mix5(S, V, lane, nb);
			double x;
			if (V.bypass == 1) {
				x = bypassSignal;
			} else {
				stage_bandpass5(S, V, lane, nb, s);
				stage_tube5(S, V, role, lane, nb, t);
				x = lane < nb ? S->row[R_FLOWM][lane] + S->row[R_FLOWN][lane] : 0.0;
			}
			if (lane < nb) S->xring[(nDone + lane) & (kRing - 1)] = x;
			__syncwarp();
			nDone += nb;
			// outputs whose right wing is complete: (k inc) >> 16 <= nDone - 1
			long long kEnd = (long long) ((((unsigned long long) nDone << 16) + V.src_inc - 1) / V.src_inc);
			if (kEnd > U.n_out) kEnd = U.n_out;
			stage_src5(S, tab, V, lane, kDone, kEnd, nEnd, out);
			if (kEnd > kDone) kDone = kEnd;
		}
	}
	// flushBuffer() (SampleRateConverter.h:462-471): the inputs past the end are zeros (stage_src5 takes them as such)
	stage_src5(S, tab, V, lane, kDone, U.n_out, nEnd, out);
}

GTTS_DEV void tube5_cta_body(const KernelParams5& P, unsigned char* smem, int tid, int nthreads)
{
	double2* tab = reinterpret_cast<double2*>(smem);
	WarpSm5* S = reinterpret_cast<WarpSm5*>(smem + sizeof(double2) * kSrcFilterLen) + (tid >> 5);
	const int lane = tid & 31;
	for (int i = tid; i < kSrcFilterLen; i += nthreads) tab[i] = P.src_tab[i];
	__syncthreads();
	for (;;) {
		int u = 0;
		if (lane == 0) u = atomicAdd(P.queue, 1);
		u = __shfl_sync(0xffffffffu, u, 0, 32);
		if (u >= P.n_utt) break;
		const UttDesc U = P.utts[P.order[u]];
		run_utterance5(S, tab, P, U, lane);
		__syncwarp();
	}
}

inline size_t smem_bytes(int warps) { return sizeof(double2) * kSrcFilterLen + sizeof(WarpSm5) * warps; }

#ifndef GTTS_EMU
template<int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) tube5_kernel(const KernelParams5 P)
{
	extern __shared__ __align__(16) unsigned char smem_m5[];
	tube5_cta_body(P, smem_m5, (int) threadIdx.x, WARPS * 32);
}
#endif

} // namespace m5
} // namespace gtts
#endif
