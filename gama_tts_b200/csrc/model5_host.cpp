#include "model5_host.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace gtts {
namespace m5 {

namespace {

// Util::amplitude60dB (VTMUtil.h:48-67)
double amplitude60dB(double db)
{
	if (db <= 0.0) return 0.0;
	if (db == 60.0) return 1.0;
	return std::pow(10.0, (db - 60.0) * (1.0 / 20.0));
}

// PoleZeroRadiationImpedance::update (PoleZeroRadiationImpedance.h:143-176) -> cT1 (= cR1), cT2, cT3, cR2, cR3
void radiationCoefficients(double radius, double samplePeriod, double* out5)
{
	const double rr = radius < 0.5e-2 ? 0.5e-2 : radius;
	const double transFreq = 62.3371 / rr + 320.204;
	const double cosWT = std::cos((2.0 * M_PI) * transFreq * samplePeriod);
	const double qa = 2.0 * cosWT;
	const double qb = -2.0 * (cosWT + 1.0);
	const double qc = cosWT + 1.0;
	const double delta = qb * qb - 4.0 * qa * qc;
	double a = (-qb - std::sqrt(delta)) / (2.0 * qa);
	const double b = 2.0 * a - 1.0;
	if (radius < 0.5e-2) a *= 40391.2 * (radius * radius);
	const double coef = 1.0 / (a + 1.0);
	out5[0] = (a + b) * coef;
	out5[1] = 2.0 * coef;
	out5[2] = -2.0 * b * coef;
	out5[3] = (a - 1.0) * coef;
	out5[4] = (b - a) * coef;
}

// Butterworth1LowPassFilter::update (Butterworth1LowpassFilter.h:73-87); false: cutoff outside [1, 0.48 fs]
bool butterworth1(double fs, double fc, double& b0, double& a1)
{
	if (fc < 1.0 || fc > fs * 0.48) return false;
	const double wcT = 2.0 * std::tan(M_PI * fc / fs);
	const double c1 = 1.0 / (wcT + 2.0);
	b0 = c1 * wcT;
	a1 = c1 * (wcT - 2.0);
	return true;
}

} // namespace

const char* deriveVoice5(const gtts_voice5_config& c, Voice5Dev& v, bool* unsupported)
{
	*unsupported = false;
	v = Voice5Dev{};
	// VocalTractModel5.h:373-425
	double length = c.vocal_tract_length_offset + c.vocal_tract_length;
	if (length < 3.0) length = 3.0; else if (length > 30.0) length = 30.0;
	const double sound = 331.4 + (0.6 * c.temperature);
	v.fs = (sound * 30.0 * 100.0) / length;                     // :464-465
	// PoleZeroRadiationImpedance.h:104-118: the reference throws below 50 kHz
	if (!(v.fs >= 50000.0)) return "internal sample rate below 50 kHz (the reference's radiation impedance refuses it): vocal tract too long";
	v.Ts = 1.0 / v.fs;
	v.output_rate = c.output_rate;
	v.waveform = c.waveform;
	v.modulation = c.noise_modulation;
	v.bypass = c.bypass;
	v.const_mouth = c.constant_radius_mouth_impedance ? 1 : 0;
	v.breath = c.breathiness / 100.0;
	v.one_minus_breath = 1.0 - v.breath;
	const double mix = amplitude60dB(c.mix_offset);
	if (!(mix > 0.0)) return "mix_offset must be above 0 dB";
	v.crossmix = 1.0 / mix;
	v.damping = 1.0 - (c.loss_factor / 100.0);
	for (int i = 0; i < 8; ++i) v.radius_coef[i] = c.radius_coef[i] * c.global_radius_coef;
	double nr[7];
	nr[0] = 0.0;
	for (int i = 0; i < 6; ++i) nr[i + 1] = c.nasal_radius[i] * c.global_nasal_radius_coef;
	v.nr2_2 = nr[1] * nr[1];
	for (int i = 1, j = 1; i < 6; ++i, ++j) {                   // :588-594
		const double r0 = nr[j] * nr[j], r1 = nr[j + 1] * nr[j + 1];
		v.nasal_k[i] = (r0 - r1) / (r0 + r1);
	}
	radiationCoefficients(std::sqrt(0.5 * nr[6] * nr[6]) * static_cast<double>(1.0e-2f), v.Ts, v.rad_n);   // :596-598
	if (v.const_mouth) radiationCoefficients(c.mouth_impedance_radius * static_cast<double>(1.0e-2f), v.Ts, v.rad_m);
	// RosenbergBGlottalSource.h:63-101
	v.tn_min = c.glottal_pulse_tn_min / 100.0;
	v.tn_max = c.glottal_pulse_tn_max / 100.0;
	v.t1 = c.glottal_pulse_tp / 100.0;
	if (v.t1 < 1.0e-2) return "glottal source: tp too small or negative";
	if (v.tn_min < 1.0e-2) return "glottal source: tnMin too small or negative";
	if (v.tn_max < 1.0e-2) return "glottal source: tnMax too small or negative";
	if (v.tn_min > v.tn_max) return "glottal source: tnMin must be <= tnMax";
	if (v.t1 + v.tn_max > 1.0) return "glottal source: tp + tnMax must be <= 1";
	if (!butterworth1(v.fs, c.glottal_noise_cutoff, v.gn_b0, v.gn_a1)) return "glottal_noise_cutoff outside [1 Hz, 0.48 fs]";
	if (!butterworth1(v.fs, c.glottal_lowpass_cutoff, v.gl_b0, v.gl_a1)) return "glottal_lowpass_cutoff outside [1 Hz, 0.48 fs]";
	{
		// Butterworth2LowPassFilter::update (Butterworth2LowpassFilter.h:82-102)
		const double fc = c.frication_noise_cutoff;
		if (fc < 1.0 || fc > v.fs * 0.48) return "frication_noise_cutoff outside [1 Hz, 0.48 fs]";
		const double wcT = 2.0 * std::tan(M_PI * fc / v.fs);
		const double wc2T2 = wcT * wcT;
		const double c1 = 2.0 * std::sqrt(2.0) * wcT;
		const double c2 = 1.0 / (wc2T2 + c1 + 4.0);
		v.fn_b0 = c2 * wc2T2;
		v.fn_b1 = 2.0 * v.fn_b0;
		v.fn_a1 = c2 * (2.0 * wc2T2 - 8.0);
		v.fn_a2 = c2 * (wc2T2 - c1 + 4.0);
	}
	v.min_loss = c.min_glottal_loss / 100.0;
	v.max_loss = c.max_glottal_loss / 100.0;
	v.fric_factor = c.frication_factor;
	// SampleRateConverter.h:136-164
	if (!(c.output_rate > 0.0)) return "output_rate must be positive";
	v.src_ratio = c.output_rate / v.fs;
	v.src_inc = static_cast<uint32_t>(std::rint(std::pow(2.0, 16) / v.src_ratio));
	if (v.src_inc == 0) return "sample rate ratio too large";
	if (v.src_ratio >= 1.0) { *unsupported = true; return "model 5: output rate at or above the internal rate (up-sampling converter) is not implemented"; }
	const double rounded = std::pow(2.0, 16) / v.src_inc;
	v.src_phase_inc = static_cast<uint32_t>(std::rint(v.src_ratio * 65536));
	v.src_pad = static_cast<int>(kSrcZeroCrossings / rounded) + 1;
	if (v.src_pad > kMaxPad) { *unsupported = true; return "model 5: internal rate above ~3.6 x the output rate (converter wing longer than 48 taps) is not implemented"; }
	return nullptr;
}

int32_t controlSteps5(double fs, double controlRate)
{
	return static_cast<int32_t>(static_cast<unsigned int>(std::rint(fs / controlRate)));    // Controller.cpp:286
}

// Outputs after dataFill x n + flushBuffer (SampleRateConverter.h:268-282, 362-416, 462-471): output k sits at ring
// position (k inc) >> 16 and exists while that position is below n + 2 pad.
int64_t outputLength5(const Voice5Dev& v, int64_t nInternal)
{
	const unsigned __int128 limit = static_cast<unsigned __int128>(nInternal + 2 * static_cast<int64_t>(v.src_pad)) << 16;
	return static_cast<int64_t>((limit + v.src_inc - 1) / v.src_inc);
}

std::string planBatch5(const gtts_voice5_config* voices, int32_t nVoices, const int32_t* voiceIndex,
			double controlRate, const int32_t* stepsOverride, const int64_t* frameOffsets,
			int64_t nUtt, BatchPlan5& plan, int* err)
{
	*err = GTTS_ERR_INVALID;
	if (!voices || nVoices <= 0) return "no voices given";
	if (nUtt < 0 || (nUtt > 0 && !frameOffsets)) return "bad utterance count / frame offsets";
	if (!stepsOverride && !(controlRate > 0.0)) return "control_rate must be positive";
	plan.voices.resize(nVoices);
	for (int32_t i = 0; i < nVoices; ++i) {
		bool unsupported = false;
		const char* e = deriveVoice5(voices[i], plan.voices[i], &unsupported);
		if (e) {
			if (unsupported) *err = GTTS_ERR_UNSUPPORTED;
			return std::string("voice ") + std::to_string(i) + ": " + e;
		}
	}
	plan.utts.resize(nUtt);
	plan.out_offsets.assign(nUtt + 1, 0);
	for (int64_t u = 0; u < nUtt; ++u) {
		UttDesc& d = plan.utts[u];
		const int32_t vi = voiceIndex ? voiceIndex[u] : 0;
		if (vi < 0 || vi >= nVoices) return "voice_index out of range";
		const int64_t f0 = frameOffsets[u], f1 = frameOffsets[u + 1];
		if (f1 < f0 || f0 < 0) return "frame_offsets must be non-decreasing";
		const Voice5Dev& v = plan.voices[vi];
		const int32_t steps = (stepsOverride && stepsOverride[u] > 0) ? stepsOverride[u] : controlSteps5(v.fs, controlRate);
		if (steps <= 0) return "control steps must be positive (control_rate above the internal rate?)";
		d = UttDesc{};
		d.frame_begin = f0;
		d.n_frames = f1 - f0;
		d.voice = vi;
		d.steps = steps;
		d.inv_steps = 1.0f / static_cast<float>(static_cast<unsigned int>(steps));   // Controller.cpp:287
		d.n_internal = d.n_frames * steps;
		d.n_out = outputLength5(v, d.n_internal);
		d.out_begin = plan.out_offsets[u];
		d.state_index = -1;
		plan.out_offsets[u + 1] = (plan.out_offsets[u] + d.n_out + 63) & ~int64_t(63);
	}
	plan.n_frames_total = nUtt ? frameOffsets[nUtt] : 0;
	plan.order.resize(nUtt);
	std::iota(plan.order.begin(), plan.order.end(), 0);
	std::stable_sort(plan.order.begin(), plan.order.end(), [&](int32_t a, int32_t b) {
		return plan.utts[a].n_internal > plan.utts[b].n_internal;
	});
	*err = GTTS_OK;
	return std::string();
}

} // namespace m5
} // namespace gtts
