// tube5_kernel -- the reference's model 5 (VocalTractModel5<double, 1>, the voice its documentation uses by
// default) on sm_100a: one warp per utterance, frame-aligned blocks of <= 32 internal samples.
//
//   lane = sample    parameter conversion (VocalTractModel5.h:527-533, 610-630, 776-792: 2^x, 10^x, 9 junction
//                    divisions, the mouth radiation impedance from R8 -- cos, sqrt, 2 divisions per sample --, bandpass
//                    coefficients), noise by LCG jump-ahead, the Rosenberg-B pulse value from its phase, mixing
//   lane = filter    the serial one- and two-pole filters, three at a time on three lanes: glottal-noise and
//                    frication-noise Butterworth low-passes and the pulse phase; then the glottal low-pass; then the
//                    frication bandpass
//   lane = section   the tube (:646-730): 30 oropharynx sections on lanes 0..29 and the 21 nasal sections on lanes
//                    12..31, 0 -- N1 shares a lane with S13, so the three waves of the velum junction are already
//                    where the four neighbour shuffles of a sample put them.  Flow-equation junctions and plain
//                    damped delays are one formula (coefficient 0 for a delay: (x - 0 * s) * d is x * d exactly).
//   lane = output    down-sampling windowed-sinc converter (SampleRateConverter.h:362-415), then the float32
//                    difference filter * output rate of the output callback (:497-513, DifferenceFilter.h:62-69)
//
// Arithmetic: IEEE double in the reference's order of evaluation, FMA contraction allowed; 2^x / 10^x by the
// branch-free forms of tube_kernel.cuh (<= 2 ulp); cos / tan / sqrt / division as CUDA provides them (<= 2 ulp, sqrt
// and division exact).  Noise generator and float32 interpolation bit-exact.
//
// Also compiled for the host by tests/simt_emu (GTTS_EMU): test infrastructure, never linked into the product.
#ifndef GTTS_TUBE5_KERNEL_CUH_
#define GTTS_TUBE5_KERNEL_CUH_

#include "tube_kernel.cuh"
#include "tube5_types.h"

namespace gtts {
namespace m5 {

enum {
	kRow = 33,
	kRing = 128,                // tube-output ring (doubles)
	kYRing = 64,                // raw converter outputs (float) for the difference filter
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
	R_MINGL, R_DGL,
	R_NT2,                      // t2 the source takes at its next wrap
	R_NOISE, R_GN, R_FN,
	R_T, R_T2,                  // pulse phase and fall end per sample
	R_PULSE,                    // source value, then the low-passed pulse
	R_IN, R_FRIC, R_GL, R_FV,
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
		for (int i = 0; i < 7; ++i) S->row[R_J0 + i][lane] = (r2[i] - r2[i + 1]) / (r2[i] + r2[i + 1]);
		const double vel = (double) p[15];
		const double v2 = vel * vel;
		{
			// Junction3::configure (:281-292) with left == right == R4
			const double c = 1.0 / (r2[3] + r2[3] + v2);
			S->row[R_VL][lane] = c * (r2[3] - r2[3] - v2);
			S->row[R_VR][lane] = c * (r2[3] - r2[3] - v2);
			S->row[R_VU][lane] = c * (v2 - r2[3] - r2[3]);
		}
		S->row[R_NJ1][lane] = (v2 - V.nr2_2) / (v2 + V.nr2_2);
		if (!V.const_mouth) {
			// PoleZeroRadiationImpedance::update (PoleZeroRadiationImpedance.h:143-176), radius in metres
			const double radius = r[7] * (double) 1.0e-2f;
			const double rr = radius < 0.5e-2 ? 0.5e-2 : radius;
			const double transFreq = 62.3371 / rr + 320.204;
			const double cosWT = cos((2.0 * 3.14159265358979323846) * transFreq * V.Ts);
			const double qa = 2.0 * cosWT;
			const double qb = -2.0 * (cosWT + 1.0);
			const double qc = cosWT + 1.0;
			const double delta = qb * qb - 4.0 * qa * qc;
			double a = (-qb - sqrt(delta)) / (2.0 * qa);
			const double b = 2.0 * a - 1.0;
			if (radius < 0.5e-2) a *= 40391.2 * (radius * radius);
			const double coef = 1.0 / (a + 1.0);
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
			const double tv = tan(pi * (double) p[6] * V.Ts);
			const double cv = cos(2.0 * pi * (double) p[5] * V.Ts);
			const double a2 = (1.0 - tv) / (1.0 + tv);
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
	double a, b, c, d;          // lane 0: glottal-noise filter x1, y1; lane 1: frication-noise x1, x2, y1, y2; lane 2: t, t2
	double glX1, glY1;          // lane 0, second pass: glottal low-pass
	double bpX1, bpX2, bpY1, bpY2;   // lane 0, third pass: frication bandpass
};

// the three feed-forward recurrences, one per lane
GTTS_DEV void stage_serial_a(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane == 0) {
		// Butterworth1LowPassFilter::filter (:89-96)
		for (int j = 0; j < nb; ++j) {
			const double x = S->row[R_NOISE][j];
			const double y = V.gn_b0 * (x + s.a) - V.gn_a1 * s.b;
			s.a = x; s.b = y;
			S->row[R_GN][j] = y;
		}
	} else if (lane == 1) {
		// Butterworth2LowPassFilter::filter (:104-113)
		for (int j = 0; j < nb; ++j) {
			const double x = S->row[R_NOISE][j];
			const double y = V.fn_b0 * (x + s.b) + V.fn_b1 * s.a - V.fn_a1 * s.c - V.fn_a2 * s.d;
			s.b = s.a; s.a = x; s.d = s.c; s.c = y;
			S->row[R_FN][j] = y;
		}
	} else if (lane == 2) {
		// RosenbergBGlottalSource::getSample (:122-150): the phase and the fall end each sample sees
		const bool dynamic = V.waveform == 0 && V.tn_min != V.tn_max;
		for (int j = 0; j < nb; ++j) {
			S->row[R_T][j] = s.a;
			S->row[R_T2][j] = s.b;
			s.a += S->row[R_DT][j];
			if (s.a > 1.0) {
				s.a -= 1.0;
				if (dynamic) s.b = S->row[R_NT2][j];
			}
		}
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

GTTS_DEV void stage_serial_b(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane == 0) {
		for (int j = 0; j < nb; ++j) {
			const double x = S->row[R_PULSE][j];
			const double y = V.gl_b0 * (x + s.glX1) - V.gl_a1 * s.glY1;
			s.glX1 = x; s.glY1 = y;
			S->row[R_PULSE][j] = y;
		}
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

// frication bandpass (BandpassFilter.h:112-122) and the value injected into the tube
GTTS_DEV void stage_serial_c(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Serial5& s)
{
	if (lane == 0) {
		for (int j = 0; j < nb; ++j) {
			const double x = S->row[R_FRIC][j];
			const double y = S->row[R_BPB0][j] * (x - s.bpX2) - S->row[R_BPA1][j] * s.bpY1 - S->row[R_BPA2][j] * s.bpY2;
			s.bpX2 = s.bpX1; s.bpX1 = x; s.bpY2 = s.bpY1; s.bpY1 = y;
			S->row[R_FV][j] = S->row[R_FA][j] * (V.fric_factor * y);
		}
	}
	__syncwarp();
}

struct Tube5 { double oT, oB, nT, nB, in1, outT1, outR1; };

// the tube, lane = section (see the header of this file)
GTTS_DEV void stage_tube5(WarpSm5* S, const Voice5Dev& V, int lane, int nb, Tube5& t)
{
	const double d = V.damping;
	const int ni = (lane - 12) & 31;                  // nasal section of this lane (N1 = 0 on lane 12 ... N21 = 20 on lane 0)
	// oropharynx junction on the left / right boundary of section `lane` (J1..J7 between S3|S4, S5|S6, S9|S10, S15|S16,
	// S21|S22, S25|S26, S27|S28), as a row of per-sample coefficients or -1
	int rowL = -1, rowR = -1;
	{
		const int jl[7] = {2, 4, 8, 14, 20, 24, 26};
#pragma unroll
		for (int q = 0; q < 7; ++q) {
			if (jl[q] + 1 == lane) rowL = R_J0 + q;
			if (jl[q] == lane) rowR = R_J0 + q;
		}
	}
	// nasal junctions NJ1..NJ6 between N3|N4, N6|N7, ... N18|N19: NJ1 per sample, the others fixed
	double nkLc = 0.0, nkRc = 0.0;
	bool nkLrow = false, nkRrow = false;
	if (ni < kNasal) {
		if (ni >= 3 && ni % 3 == 0) { const int q = ni / 3 - 1; if (q == 0) nkLrow = true; else if (q < 6) nkLc = V.nasal_k[q]; }
		if (ni % 3 == 2 && ni < 18) { const int q = ni / 3; if (q == 0) nkRrow = true; else nkRc = V.nasal_k[q]; }
	}
	const bool isMouth = lane == kOral - 1, isNose = lane == 0;
	const int prev = (lane + 31) & 31, next = (lane + 1) & 31;
	for (int j = 0; j < nb; ++j) {
		const double kL = rowL >= 0 ? S->row[rowL][j] : 0.0;
		const double kR = rowR >= 0 ? S->row[rowR][j] : 0.0;
		const double nkL = nkLrow ? S->row[R_NJ1][j] : nkLc;
		const double nkR = nkRrow ? S->row[R_NJ1][j] : nkRc;
		const double oTL = shfl_d(t.oT, prev, 32), oBR = shfl_d(t.oB, next, 32);
		const double nTL = shfl_d(t.nT, prev, 32), nBR = shfl_d(t.nB, next, 32);
		// propagate / propagateJunction (:316-325)
		double noT = (oTL - kL * (oTL + t.oB)) * d;
		double noB = (oBR + kR * (t.oT + oBR)) * d;
		double nnT = (nTL - nkL * (nTL + t.nB)) * d;
		double nnB = (nBR + nkR * (t.nT + nBR)) * d;
		if (lane == 0) {
			noT = t.oB * S->row[R_GL][j] + S->row[R_IN][j];       // :651
		} else if (lane == 11) {
			// 3-way junction (:326-332), S12 side: left.bottom
			const double partial = t.oT + oBR + nBR;
			noB = (oBR + nBR + S->row[R_VL][j] * partial) * d;
		} else if (lane == 12) {
			// S13 and N1: right.top and upper.top
			const double partial = oTL + t.oB + t.nB;
			noT = (oTL + t.nB + S->row[R_VR][j] * partial) * d;
			nnT = (oTL + t.oB + S->row[R_VU][j] * partial) * d;
		}
		if (isMouth || isNose) {
			// PoleZeroRadiationImpedance::process (:178-189): S30 on lane 29, N21 on lane 0
			const double in = isMouth ? t.oT : t.nT;
			const double cT1 = isMouth ? S->row[R_CT1][j] : V.rad_n[0], cT2 = isMouth ? S->row[R_CT2][j] : V.rad_n[1];
			const double cT3 = isMouth ? S->row[R_CT3][j] : V.rad_n[2], cR2 = isMouth ? S->row[R_CR2][j] : V.rad_n[3];
			const double cR3 = isMouth ? S->row[R_CR3][j] : V.rad_n[4];
			const double outT = cT1 * t.outT1 + cT2 * in + cT3 * t.in1;
			const double outR = cT1 * t.outR1 + cR2 * in + cR3 * t.in1;
			t.in1 = in; t.outT1 = outT; t.outR1 = outR;
			if (isMouth) { noB = outR * d; S->row[R_FLOWM][j] = outT; }
			else { nnB = outR * d; S->row[R_FLOWN][j] = outT; }
		}
		// frication (:716-723)
		{
			const int fo = S->fo[j];
			if (lane == 5 + fo) noT += S->row[R_FV][j] * S->row[R_FL][j];
			else if (lane == 6 + fo && fo < 22) noT += S->row[R_FV][j] * S->row[R_FR][j];
		}
		t.oT = noT; t.oB = noB; t.nT = nnT; t.nB = nnB;
	}
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
	s.a = s.b = s.c = s.d = 0.0;
	if (lane == 2) s.b = V.t1 + V.tn_max;           // t2 (RosenbergBGlottalSource.h:76-78)
	s.glX1 = s.glY1 = 0.0;
	s.bpX1 = s.bpX2 = s.bpY1 = s.bpY2 = 0.0;
	Tube5 t = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
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
			stage_serial_a(S, V, lane, nb, s);
			stage_pulse_value(S, V, lane, nb);
			stage_serial_b(S, V, lane, nb, s);
			const double bypassSignal = stage_mix5(S, V, lane, nb);
			double x;
			if (V.bypass == 1) {
				x = bypassSignal;
			} else {
				stage_serial_c(S, V, lane, nb, s);
				stage_tube5(S, V, lane, nb, t);
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
