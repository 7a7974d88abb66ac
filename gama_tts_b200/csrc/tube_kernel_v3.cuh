// tube_kernel_v3 -- the wide-batch kernel: ONE THREAD PER UTTERANCE (sm_100a).
//
// The pipelined kernel (tube_kernel_v2.cuh) finds its parallelism inside an utterance -- seven utterances per SM,
// stages on different warps, tube sections on lanes -- because BASELINE config 2 has only seven utterances per SM.
// A batch of tens of thousands of utterances (configs 3 and 4) has more parallelism between utterances than the
// GPU has FP64 lanes, and there the per-utterance machinery is overhead: a lane-per-section tube spends an FP64
// warp instruction and three shuffles on sixteen junctions, the serial filters run on 7 of 32 lanes, every
// stage passes its results through shared memory.  Here a thread owns an utterance from its first sample to its
// last: all recurrences (tube waves, filters, oscillator phase, noise generator, interpolation accumulators) stay
// in its registers, every FP64 instruction works for 32 utterances, and nothing is exchanged between lanes.
//
// Lockstep: the 32 utterances of a warp (a "group", formed by the host: same control period where possible, similar
// length) advance one internal sample per iteration; a lane whose utterance has ended idles until the group's
// longest one has.  What is indexed by time -- the tube-output ring of the SRC, the 2x oscillator stream of the FIR
// -- lives in shared memory as [position][thread] (conflict-free: all lanes of a warp are at the same position);
// per-utterance voice constants likewise as [constant][thread].  The SRC coefficient table is shared by the CTA.
//
// Arithmetic: the reference's expressions in the reference's order (VocalTractModel0.h:396-445, 484-661), float32
// interpolation and the noise generator bit-exact, FMA contraction allowed; departures, each a rounding error or
// two: 2^x / 10^x / sin / cos / the junction divisions by the branch-free forms of tube_kernel.cuh, the wavetable
// evaluated analytically (3x^2 - 2x^3, 1 - x^2 with x = i * (1 / n) instead of i / n), the 49-tap FIR and the
// 26-tap SRC sum on four / two accumulators.  Handles up-sampling voices (internal rate <= output rate) and
// control periods of any length; the host routes everything else to the other kernels.
#ifndef GTTS_TUBE_KERNEL_V3_CUH_
#define GTTS_TUBE_KERNEL_V3_CUH_

#include "tube_kernel.cuh"

namespace gtts {
namespace v3 {

enum {
	kThreads = 128,             // utterances per CTA (one CTA per SM: the per-thread arrays below fill its shared memory)
	kWarps = kThreads / 32,
	kRing = 32,                 // tube-output ring: the SRC window is 26 samples
	kFirRing = 32,              // 2x oscillator stream, even / odd phase: the FIR reaches 24 samples back
};

// per-utterance constants in shared memory, [constant][thread]
enum {
	C_RC0 = 0,                  // C_RC0 .. C_RC0 + 7: radius coefficients
	C_AP2 = 8, C_NR1_2, C_DAMP,
	C_REFL_B0_M, C_REFL_A1_M, C_RAD_M,
	C_REFL_B0_N, C_REFL_A1_N, C_RAD_N,
	C_NK1, C_NK2, C_NK3, C_NK4, C_NK5,
	C_THR_B0, C_THR_A1, C_THR_GAIN,
	C_TS, C_BASIC_INC, C_BREATH, C_ONE_MINUS_BREATH, C_CROSSMIX,
	C_TN_DELTA, C_DIV1, C_DIV2, C_INV_DIV1, C_INV_TNLEN,
	C_COUNT
};

struct CtaSm {
	double2 tab[kSrcFilterLen];               // {h, deltaH}
	double xring[kRing][kThreads];
	double fe[kFirRing][kThreads];            // first half sample of internal sample n at [n & 31]
	double fo[kFirRing][kThreads];            // second half sample
	double vc[C_COUNT][kThreads];
	float  delta[kNumParams][kThreads];       // interpolation increments of the current control period
};

struct KernelParamsV3 {
	const VoiceDev* voices;
	const double* tables;                     // per-voice wavetables (read for sine voices only)
	const UttDesc* utts;
	const int32_t* order;                     // n_groups x 32 utterance indices, -1: empty lane
	const float* frames;
	float* out;
	const double2* src_tab;
	int32_t* queue;
	int32_t n_groups;
};

inline size_t smem_bytes() { return sizeof(CtaSm); }

// One utterance, first sample to last.  `S` + tid addresses this thread's columns.
GTTS_DEV void run_thread(CtaSm* S, const KernelParamsV3& P, const UttDesc& U, int tid)
{
	const VoiceDev& V = P.voices[U.voice];
	// ---- per-utterance constants ----
#define VC(i) S->vc[i][tid]
#pragma unroll
	for (int i = 0; i < 8; ++i) VC(C_RC0 + i) = V.radius_coef[i];
	VC(C_AP2) = V.ap2; VC(C_NR1_2) = V.nr1_2; VC(C_DAMP) = V.damping;
	VC(C_REFL_B0_M) = V.refl_b0_m; VC(C_REFL_A1_M) = V.refl_a1_m; VC(C_RAD_M) = V.rad_m;
	VC(C_REFL_B0_N) = V.refl_b0_n; VC(C_REFL_A1_N) = V.refl_a1_n; VC(C_RAD_N) = V.rad_n;
	VC(C_NK1) = V.nasal_k[1]; VC(C_NK2) = V.nasal_k[2]; VC(C_NK3) = V.nasal_k[3]; VC(C_NK4) = V.nasal_k[4]; VC(C_NK5) = V.nasal_k[5];
	VC(C_THR_B0) = V.throat_b0; VC(C_THR_A1) = V.throat_a1; VC(C_THR_GAIN) = V.throat_gain;
	VC(C_TS) = V.Ts; VC(C_BASIC_INC) = V.basic_inc; VC(C_BREATH) = V.breath; VC(C_ONE_MINUS_BREATH) = V.one_minus_breath;
	VC(C_CROSSMIX) = V.crossmix;
	VC(C_TN_DELTA) = V.tn_delta; VC(C_DIV1) = (double) V.div1; VC(C_DIV2) = (double) V.div2;
	VC(C_INV_DIV1) = 1.0 / (double) V.div1; VC(C_INV_TNLEN) = 1.0 / V.tn_length;
	const int sine = V.waveform != 0;
	const int modulation = V.modulation;
	const double* sineTab = P.tables + (size_t) U.voice * kTableLen;
	const unsigned inc = V.src_inc;
	const int steps = U.steps;
	const float invSteps = U.inv_steps;
	const long long nInternal = U.n_internal, nFrames = U.n_frames;
	const float* frames = P.frames + U.frame_begin * kNumParams;
	float* out = P.out + U.out_begin;

#pragma unroll 1
	for (int i = 0; i < kRing; ++i) S->xring[i][tid] = 0.0;
#pragma unroll 1
	for (int i = 0; i < kFirRing; ++i) { S->fe[i][tid] = 0.0; S->fo[i][tid] = 0.0; }

	// ---- state (VocalTractModel0.h:309-326 reset) ----
	double oT[10], oB[10], nT[6], nB[6];
#pragma unroll
	for (int i = 0; i < 10; ++i) { oT[i] = 0.0; oB[i] = 0.0; }
#pragma unroll
	for (int i = 0; i < 6; ++i) { nT[i] = 0.0; nB[i] = 0.0; }
	double reflY1m = 0.0, radX1m = 0.0, radY1m = 0.0, reflY1n = 0.0, radX1n = 0.0, radY1n = 0.0, thrY1 = 0.0;
	double bpX1 = 0.0, bpX2 = 0.0, bpY1 = 0.0, bpY2 = 0.0, noiseX1 = 0.0, pos = 0.0;
	unsigned long long lcg = c_lcg_init;
	int low = kNoLowMark;
	float cur[kNumParams];
#pragma unroll
	for (int i = 0; i < kNumParams; ++i) cur[i] = 0.0f;
	int left = 0;                               // samples left in the current control period
	long long p = 0;                            // next frame
	unsigned long long t = 0;                   // (next output) * inc: position e = t >> 16, phase t & 0xFFFF
	long long k = 0;                            // next output
	const long long nTotal = nInternal + 2 * kSrcZeroCrossings;   // 26 flush zeros (SampleRateConverter.h:462-471)
	const long long nOut = U.n_out;

#pragma unroll 1
	for (long long n = 0; n < nTotal; ++n) {
		double xOut = 0.0;
		if (n < nInternal) {
			// ---- control period start: Controller.cpp:297-300 ----
			if (left == 0) {
				const float4* f0 = reinterpret_cast<const float4*>(frames + p * kNumParams);
				const float4* f1 = (p + 1 < nFrames) ? f0 + 4 : f0;
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const float4 a = f0[q], b = f1[q];
					cur[4 * q + 0] = a.x; cur[4 * q + 1] = a.y; cur[4 * q + 2] = a.z; cur[4 * q + 3] = a.w;
					S->delta[4 * q + 0][tid] = __fmul_rn(__fsub_rn(b.x, a.x), invSteps);
					S->delta[4 * q + 1][tid] = __fmul_rn(__fsub_rn(b.y, a.y), invSteps);
					S->delta[4 * q + 2][tid] = __fmul_rn(__fsub_rn(b.z, a.z), invSteps);
					S->delta[4 * q + 3][tid] = __fmul_rn(__fsub_rn(b.w, a.w), invSteps);
				}
				left = steps;
				p += 1;
			}
			left -= 1;

			// ---- conversions (VocalTractModel0.h:396-404, 484-552, 698-716; VTMUtil.h:50-84) ----
			const double f0hz = 220.0 * gtts_exp2(((double) cur[0] + 3.0) * (1.0 / 12.0));
			const double oscInc = (f0hz / 2.0) * VC(C_BASIC_INC);
			const double ax = amp60((double) cur[1]);
			const double ah1 = amp60((double) cur[2]);
			const double fa = amp60((double) cur[3]);
			double r2[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				double r = (double) cur[7 + i] * VC(C_RC0 + i);
				r = r > 0.01 ? r : 0.01;
				r2[i] = r * r;
			}
			double kk[8];
#pragma unroll
			for (int i = 0; i < 7; ++i) kk[i] = div_fast(r2[i] - r2[i + 1], r2[i] + r2[i + 1]);
			{
				const double ap2 = VC(C_AP2);
				kk[7] = div_fast(r2[7] - ap2, r2[7] + ap2);
			}
			const double vel = (double) cur[15];
			const double v2 = vel * vel;
			const double asum = div_fast(2.0, r2[3] + r2[3] + v2);
			const double aL = asum * r2[3], aU = asum * v2;
			double nk0;
			{
				const double b2 = VC(C_NR1_2);
				nk0 = div_fast(v2 - b2, v2 + b2);
			}
			// frication taps (:524-552)
			const double fpos = (double) cur[4];
			int ip = (int) fpos;
			const double comp = fpos - ip;
			double ta = (1.0 - comp) * fa, tb = comp * fa;
			if (ip < 0 || ip > 7) { ta = 0.0; tb = 0.0; ip = -50; }
			// bandpass coefficients (BandpassFilter.h:91-110)
			double bpA2, bpA1, bpB0;
			{
				const double pi = 3.14159265358979323846;
				const double Ts = VC(C_TS);
#ifndef GTTS_EMU
				double sx, cx, sy, cv;
				gtts_sincos(pi * (double) cur[6] * Ts, sx, cx);
				gtts_sincos(2.0 * pi * (double) cur[5] * Ts, sy, cv);
				bpA2 = div_fast(cx - sx, cx + sx);
#else
				const double tv = tan(pi * (double) cur[6] * Ts);
				const double cv = cos(2.0 * pi * (double) cur[5] * Ts);
				bpA2 = (1.0 - tv) / (1.0 + tv);
#endif
				bpA1 = -(1.0 + bpA2) * cv;
				bpB0 = 0.5 - 0.5 * bpA2;
			}
			// float32 interpolation (Controller.cpp:307-310)
#pragma unroll
			for (int i = 0; i < kNumParams; ++i) cur[i] = __fadd_rn(cur[i], S->delta[i][tid]);

			// ---- noise (NoiseSource.h:40-44, NoiseFilter.h:63-68) ----
			lcg = (lcg * 377ull) & ((1ull << 44) - 1);
			const double noise = (double) lcg * (1.0 / 17592186044416.0) - 0.5;
			const double lp = noise + noiseX1;
			noiseX1 = noise;

			// ---- glottal source (WavetableGlottalSource.h:162-235) ----
			const double div1 = VC(C_DIV1), div2 = VC(C_DIV2);
			double nd2 = div2, inv = VC(C_INV_TNLEN);
			{
				const double tnDelta = VC(C_TN_DELTA);
				if (tnDelta != 0.0 && !sine) {
					nd2 = div2 - rint(ax * tnDelta);
					nd2 = nd2 > 0.0 ? nd2 : 0.0;
					inv = div_fast(1.0, nd2 - div1);             // unused (and not finite) when nd2 <= div1: every entry of the fall segment is then 0
					if (nd2 < div1) { const int c = (int) nd2; low = c < low ? c : low; }
				}
			}
			const double invDiv1 = VC(C_INV_DIV1);
			double vHalf[2];
#pragma unroll
			for (int s = 0; s < 2; ++s) {
				double sp = pos + oscInc;
				pos = (sp > 511.0) ? sp - 512.0 : sp;
				const unsigned lo = __double2uint_rz(pos);
				const unsigned up = (lo + 1 > 511u) ? lo + 1 - 512u : lo + 1;
				double tv[2];
#pragma unroll
				for (int w = 0; w < 2; ++w) {
					const unsigned i = w ? up : lo;
					const double di = (double) (int) i;
					double v;
					if (sine) {
						v = sineTab[i];
					} else if (di < div1) {
						const double x = di * invDiv1;
						const double x2 = x * x;
						v = ((int) i >= low) ? 0.0 : (3.0 * x2) - (2.0 * (x2 * x));
					} else if (di < div2) {
						const double x = (di - div1) * inv;
						v = (di >= nd2) ? 0.0 : 1.0 - (x * x);
					} else {
						v = 0.0;
					}
					tv[w] = v;
				}
				vHalf[s] = tv[0] + ((pos - (double) lo) * (tv[1] - tv[0]));
			}
			const int slot = (int) (n & (kFirRing - 1));
			S->fe[slot][tid] = vHalf[0];
			S->fo[slot][tid] = vHalf[1];
			// 49-tap decimating FIR (WavetableGlottalSourceFIRFilter.h:276-304): y = sum c[i] x2[2n + 1 - i]
			double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
			for (int j = 0; j < 24; j += 2) {
				const int s0 = (slot - j) & (kFirRing - 1), s1 = (slot - j - 1) & (kFirRing - 1);
				a0 += S->fo[s0][tid] * c_fir[2 * j];
				a1 += S->fe[s0][tid] * c_fir[2 * j + 1];
				a2 += S->fo[s1][tid] * c_fir[2 * j + 2];
				a3 += S->fe[s1][tid] * c_fir[2 * j + 3];
			}
			a0 += S->fo[(slot - 24) & (kFirRing - 1)][tid] * c_fir[48];
			double pulse = (a0 + a1) + (a2 + a3);

			// ---- mixing (VocalTractModel0.h:408-438) ----
			const double pn = lp * pulse;
			pulse = ax * ((pulse * VC(C_ONE_MINUS_BREATH)) + (pn * VC(C_BREATH)));
			double sig;
			if (modulation) {
				double cm = ax * VC(C_CROSSMIX);
				cm = (cm < 1.0) ? cm : 1.0;
				sig = (pn * cm) + (lp * (1.0 - cm));
			} else {
				sig = lp;
			}
			// frication bandpass (BandpassFilter.h:114-122)
			const double fr = bpB0 * (sig - bpX2) - bpA1 * bpY1 - bpA2 * bpY2;
			bpX2 = bpX1; bpX1 = sig; bpY2 = bpY1; bpY1 = fr;
			const double pa = ta * fr, pb = tb * fr;      // injected at taps ip and ip + 1
			double tapf[8];
#pragma unroll
			for (int q = 0; q < 8; ++q) tapf[q] = (q == ip) ? pa : ((q == ip + 1) ? pb : 0.0);

			// ---- the tube (VocalTractModel0.h:565-661): new waves from the previous sample's ----
			const double d = VC(C_DAMP);
			const double input = (pulse + (ah1 * sig)) * 0.125;
			double noT[10], noB[10], nnT[6], nnB[6];
			double dl;
			noT[0] = (oB[0] * d) + input;
			dl = kk[0] * (oT[0] - oB[1]);
			noT[1] = (oT[0] + dl) * d;
			noB[0] = (oB[1] + dl) * d;
			dl = kk[1] * (oT[1] - oB[2]);
			noT[2] = ((oT[1] + dl) * d) + tapf[0];
			noB[1] = (oB[2] + dl) * d;
			dl = kk[2] * (oT[2] - oB[3]);
			noT[3] = ((oT[2] + dl) * d) + tapf[1];
			noB[2] = (oB[3] + dl) * d;
			{
				const double jp = (aL * oT[3]) + (aL * oB[4]) + (aU * nB[0]);
				noB[3] = (jp - oT[3]) * d;
				noT[4] = ((jp - oB[4]) * d) + tapf[2];
				nnT[0] = (jp - nB[0]) * d;
			}
			dl = kk[3] * (oT[4] - oB[5]);
			noT[5] = ((oT[4] + dl) * d) + tapf[3];
			noB[4] = (oB[5] + dl) * d;
			noT[6] = (oT[5] * d) + tapf[4];
			noB[5] = oB[6] * d;
			dl = kk[4] * (oT[6] - oB[7]);
			noT[7] = ((oT[6] + dl) * d) + tapf[5];
			noB[6] = (oB[7] + dl) * d;
			dl = kk[5] * (oT[7] - oB[8]);
			noT[8] = ((oT[7] + dl) * d) + tapf[6];
			noB[7] = (oB[8] + dl) * d;
			dl = kk[6] * (oT[8] - oB[9]);
			noT[9] = ((oT[8] + dl) * d) + tapf[7];
			noB[8] = (oB[9] + dl) * d;
			double output;
			{
				// mouth: reflection and radiation (ReflectionFilter.h, RadiationFilter.h:73-79)
				const double y = VC(C_REFL_B0_M) * (kk[7] * oT[9]) - VC(C_REFL_A1_M) * reflY1m;
				reflY1m = y;
				noB[9] = d * y;
				const double a = VC(C_RAD_M);
				const double x = (1.0 + kk[7]) * oT[9];
				const double yr = a * x + (-a) * radX1m - (-a) * radY1m;
				radX1m = x; radY1m = yr;
				output = yr;
			}
			{
				const double nk[5] = {nk0, VC(C_NK1), VC(C_NK2), VC(C_NK3), VC(C_NK4)};
#pragma unroll
				for (int i = 0; i < 5; ++i) {
					dl = nk[i] * (nT[i] - nB[i + 1]);
					nnT[i + 1] = (nT[i] + dl) * d;
					nnB[i] = (nB[i + 1] + dl) * d;
				}
				const double nk5 = VC(C_NK5);
				const double y = VC(C_REFL_B0_N) * (nk5 * nT[5]) - VC(C_REFL_A1_N) * reflY1n;
				reflY1n = y;
				nnB[5] = d * y;
				const double a = VC(C_RAD_N);
				const double x = (1.0 + nk5) * nT[5];
				const double yr = a * x + (-a) * radX1n - (-a) * radY1n;
				radX1n = x; radY1n = yr;
				output += yr;
			}
#pragma unroll
			for (int i = 0; i < 10; ++i) { oT[i] = noT[i]; oB[i] = noB[i]; }
#pragma unroll
			for (int i = 0; i < 6; ++i) { nT[i] = nnT[i]; nB[i] = nnB[i]; }
			// throat (Throat.h:80-85)
			{
				const double y = VC(C_THR_B0) * (pulse * 0.125) - VC(C_THR_A1) * thrY1;
				thrY1 = y;
				xOut = output + (y * VC(C_THR_GAIN));
			}
		}
		S->xring[n & (kRing - 1)][tid] = xOut;

		// ---- sample-rate conversion (SampleRateConverter.h:295-361): the outputs centred on this sample ----
		// Output k is centred on e = (k inc) >> 16 with phase f = (k inc) & 0xFFFF: left wing taps h[(f >> 8) + 256 j]
		// on x[e - 13 - j], right wing taps (from ~f) on x[e - 12 + j], j = 0..12.
#pragma unroll 1
		while ((long long) (t >> 16) <= n && k < nOut) {
			const unsigned f = (unsigned) (t & 0xFFFFu);
			const unsigned g = (~f) & 0xFFFFu;
			const double iL = (double) (f & 0xFFu) / 256, iR = (double) (g & 0xFFu) / 256;
			const double2* pL = S->tab + (f >> 8);
			const double2* pR = S->tab + (g >> 8);
			const int e = (int) (t >> 16);             // == n: up-sampling, the centre moves by at most one sample per output
			double accL = 0.0, accR = 0.0;
#pragma unroll
			for (int j = 0; j < kSrcZeroCrossings; ++j) {
				const double2 cl = pL[256 * j], cr = pR[256 * j];
				accL += S->xring[(e - 13 - j) & (kRing - 1)][tid] * (cl.x + (cl.y * iL));
				accR += S->xring[(e - 12 + j) & (kRing - 1)][tid] * (cr.x + (cr.y * iR));
			}
			out[k] = (float) (accL + accR);
			k += 1;
			t += inc;
		}
	}
#undef VC
}

// CTA body: the SRC table is staged once, then every warp takes groups of 32 utterances from the queue.
GTTS_DEV void tube_v3_cta_body(const KernelParamsV3& P, unsigned char* smem, int tid)
{
	CtaSm* S = reinterpret_cast<CtaSm*>(smem);
	for (int i = tid; i < kSrcFilterLen; i += kThreads) S->tab[i] = P.src_tab[i];
	__syncthreads();
	const int lane = tid & 31;
	for (;;) {
		int g = 0;
		if (lane == 0) g = atomicAdd(P.queue, 1);
		g = __shfl_sync(0xffffffffu, g, 0, 32);
		if (g >= P.n_groups) break;
		const int u = P.order[(size_t) g * 32 + lane];
		if (u >= 0) {
			const UttDesc U = P.utts[u];
			run_thread(S, P, U, tid);
		}
		__syncwarp();
	}
}

#ifndef GTTS_EMU
__global__ void __launch_bounds__(kThreads, 1) tube_kernel_v3(const KernelParamsV3 P)
{
	extern __shared__ __align__(16) unsigned char smem_v3[];
	tube_v3_cta_body(P, smem_v3, (int) threadIdx.x);
}
#endif

} // namespace v3
} // namespace gtts
#endif
