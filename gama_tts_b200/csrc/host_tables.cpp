// Host-side, init-time arithmetic of the tube path: per-voice derived constants, the glottal FIR
// design, the SRC windowed-sinc tables, output-length closed form and the multi-GPU shard plan.
// Everything here is double precision libm work done once per voice / per process (SURVEY.md
// Appendix B); nothing here runs per sample.  Compiled with -ffp-contract=off so that the table
// values do not depend on the host compiler's FMA contraction.
#include "host_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

namespace gtts {

static constexpr double kPiTable = 3.14159265358979323846;

namespace {

constexpr double kPi = 3.14159265358979323846;

// Util::amplitude60dB (reference gama_tts/src/vtm/VTMUtil.h:50-67)
double amplitude60dB(double db)
{
	if (db <= 0.0) return 0.0;
	if (db == 60.0) return 1.0;
	return std::pow(10.0, (db - 60.0) * (1.0 / 20.0));
}

} // namespace

// The glottal source's decimating low-pass (WavetableGlottalSourceFIRFilter.h:74-114) is the same for every voice:
// the reference designs it at start-up with fixed arguments (beta = 0.2, gamma = 0.1, cutoff 1e-8; maximally flat,
// linear phase), which gives 49 symmetric taps.  They are a constant of the path and are shipped as one: the 25
// values from the centre tap outwards (tests/test_host_abi.py checks them bit for bit against
// the reference's own design routine, which is where they would have to be regenerated if the arguments changed).
std::vector<double> designGlottalFir(double beta, double gamma, double cutoff)
{
	static const double kHalf[25] = {
		0.39847427941239039, 0.2965041881371811, 0.087853092781387032, -0.051247637455809028,
		-0.056044214360539933, -0.0013632376466076977, 0.024723951583638423, 0.011204567485223504,
		-0.0059227989476883834, -0.0070911561060204289, -0.00061607377823225985, 0.0023296835220954727,
		0.0011316111860893143, -0.00028644112282104019, -0.00043243790948125553, -8.6968794272942906e-05,
		6.9699584503931649e-05, 4.3261082316484397e-05, 2.2354964101677552e-06, -6.267564867199802e-06,
		-2.4300136304004587e-06, -6.5322947036563118e-08, 2.137841904426972e-07, 7.3133299513665035e-08,
		1.0887157865533967e-08 };
	if (beta != 0.2 || gamma != 0.1 || cutoff != 0.00000001) return std::vector<double>();     // only the shipped design exists
	std::vector<double> taps(49);
	for (int m = 0; m < 25; ++m) taps[24 - m] = taps[24 + m] = kHalf[m];
	return taps;
}

// The SRC's interpolation filter (SampleRateConverter.h:230-255): a sinc low-pass at 11/13 of the Nyquist rate, 13 zero
// crossings of 256 phases each, under a Kaiser window (beta = 5.658), and its first differences (the last one
// against an implied zero).  Closed form; the window uses the library's modified Bessel function.  Agrees with the
// reference's table to 2e-15 relative (tests/test_host_abi.py).
void buildSrcTables(double* h, double* dh)
{
	const double beta = 5.658, corner = 11.0 / 13.0;
	const double i0Beta = std::cyl_bessel_i(0.0, beta);
	for (int i = 0; i < kSrcFilterLen; ++i) {
		const double phase = i * (kPi / kSrcLRange);
		const double sinc = (i == 0) ? corner : std::sin(phase * corner) / phase;
		const double u = static_cast<double>(i) / kSrcFilterLen;
		h[i] = sinc * (std::cyl_bessel_i(0.0, beta * std::sqrt(1.0 - (u * u))) / i0Beta);
	}
	for (int i = 0; i < kSrcFilterLen; ++i) dh[i] = (i + 1 < kSrcFilterLen ? h[i + 1] : 0.0) - h[i];
}

int internalRate(const gtts_voice_config& c)
{
	// VocalTractModel0.h:274-279, 343-344
	double length = c.vocal_tract_length_offset + c.vocal_tract_length;
	if (length < 3.0) length = 3.0; else if (length > 30.0) length = 30.0;
	const double speed = 331.4 + (0.6 * c.temperature);       // VTMUtil.h:107-113
	// models 3 and 4: 30 section delays along the tract instead of 10 (VocalTractModel2.h:419, VocalTractModel4.h)
	return static_cast<int>((speed * (c.tube_model == 0 ? 10 : 30) * 100.0) / length);
}

int controlSteps(int fs, double controlRate)
{
	return static_cast<int>(static_cast<unsigned int>(std::rint(fs / controlRate)));   // Controller.cpp:286
}

const char* deriveVoice(const gtts_voice_config& c, VoiceDev& v)
{
	std::memset(&v, 0, sizeof v);
	if (!(c.output_rate > 0.0)) return "output_rate must be positive";
	if (c.waveform != 0 && c.waveform != 1) return "waveform must be 0 (pulse) or 1 (sine)";
	if (c.tube_model != 0 && c.tube_model != 3 && c.tube_model != 4) return "tube_model must be 0 (models 0 / 2), 3 or 4";
	v.tube_model = c.tube_model;
	v.fs = internalRate(c);
	if (v.fs <= 0) return "internal sample rate is not positive";
	v.waveform = c.waveform;
	v.modulation = c.noise_modulation ? 1 : 0;
	const double nyquist = static_cast<double>(static_cast<float>(v.fs) / 2.0f);   // :345 (int / 2.0f is float)

	v.breath = c.breathiness / 100.0;                       // :349
	v.one_minus_breath = 1.0 - v.breath;                    // :422
	const double mix = amplitude60dB(c.mix_offset);
	if (mix == 0.0) return "mix_offset must be > 0 dB";
	v.crossmix = 1.0 / mix;                                 // :352
	v.damping = 1.0 - (c.loss_factor / 100.0);              // :355

	// Wavetable geometry (WavetableGlottalSource.h:104-109)
	const unsigned d1 = static_cast<unsigned>(std::rint(kTableLen * (c.glottal_pulse_tp / 100.0)));
	const unsigned d2 = static_cast<unsigned>(std::rint(kTableLen * ((c.glottal_pulse_tp + c.glottal_pulse_tn_max) / 100.0)));
	if (c.waveform == 0 && (d1 == 0 || d2 <= d1 || d2 > kTableLen)) return "glottal pulse tp/tn out of range";
	v.div1 = static_cast<int32_t>(d1);
	v.div2 = static_cast<int32_t>(d2);
	v.tn_length = static_cast<double>(d2 - d1);
	v.tn_delta = std::rint(kTableLen * ((c.glottal_pulse_tn_max - c.glottal_pulse_tn_min) / 100.0));
	if (v.tn_delta < 0.0 || v.tn_delta >= v.tn_length) return "glottal_pulse_tn_min/max out of range";
	v.basic_inc = kTableLen / static_cast<double>(v.fs);

	const double am = (nyquist - c.mouth_coefficient) / nyquist;    // :366
	v.rad_m = am; v.refl_b0_m = 1.0 - std::fabs(am); v.refl_a1_m = -am;   // RadiationFilter.h:54-62, ReflectionFilter.h:55-61
	const double an = (nyquist - c.nose_coefficient) / nyquist;     // :371
	v.rad_n = an; v.refl_b0_n = 1.0 - std::fabs(an); v.refl_a1_n = -an;

	const double apr = c.aperture_radius * c.global_radius_coef;   // :290
	double nr[6];
	nr[0] = 0.0;
	for (int i = 0; i < 5; ++i) nr[i + 1] = c.nasal_radius[i] * c.global_nasal_radius_coef;   // :291-296
	for (int i = 1; i < 5; ++i) {                                   // :460-464
		const double a2 = nr[i] * nr[i], b2 = nr[i + 1] * nr[i + 1];
		v.nasal_k[i] = (a2 - b2) / (a2 + b2);
	}
	{
		const double a2 = nr[5] * nr[5], b2 = apr * apr;            // :467-469
		v.nasal_k[5] = (a2 - b2) / (a2 + b2);
	}
	v.ap2 = apr * apr;
	v.nr1_2 = nr[1] * nr[1];
	for (int i = 0; i < 8; ++i) v.radius_coef[i] = c.radius_coef[i] * c.global_radius_coef;   // :297-304

	v.throat_b0 = (c.throat_cutoff * 2.0) / v.fs;                  // Throat.h:52-61
	v.throat_a1 = v.throat_b0 - 1.0;
	v.throat_gain = amplitude60dB(c.throat_volume);
	v.Ts = 1.0 / v.fs;

	// SampleRateConverter.h:136-164
	v.src_ratio = c.output_rate / v.fs;
	v.src_inc = static_cast<uint32_t>(std::rint(std::pow(2.0, 16) / v.src_ratio));
	if (v.src_inc == 0) return "sample rate ratio too large";
	const double rounded = std::pow(2.0, 16) / v.src_inc;
	if (v.src_ratio >= 1.0) {
		v.src_upsample = 1;
		v.src_pad = kSrcZeroCrossings;
		v.src_phase_inc = 0;
	} else {
		v.src_upsample = 0;
		v.src_phase_inc = static_cast<uint32_t>(std::rint(v.src_ratio * 65536));
		v.src_pad = static_cast<int>(kSrcZeroCrossings / rounded) + 1;
		// the kernels' 128-entry ring holds a converter window of 2 pad + 32 inputs (models 3 / 4 with a tract below ~6.2 cm)
		if (v.src_pad > 48) return "internal rate above ~3.6 x the output rate (converter wing longer than 48 taps) is not implemented";
	}
	return nullptr;
}

// Number of outputs after dataFill x n_internal + flushBuffer (SampleRateConverter.h:268-282,
// 295-416, 462-471): output k sits at ring position (k*inc)>>16 and exists while that position is
// below n_internal + 2*pad.
int64_t outputLength(const VoiceDev& v, int64_t nInternal)
{
	const unsigned __int128 limit = static_cast<unsigned __int128>(nInternal + 2 * static_cast<int64_t>(v.src_pad)) << 16;
	return static_cast<int64_t>((limit + v.src_inc - 1) / v.src_inc);
}

// Longest-processing-time-first greedy partition.
void shardPlan(const int64_t* cost, int64_t n, int shards, int32_t* shardOf)
{
	std::vector<int64_t> order(n);
	std::iota(order.begin(), order.end(), 0);
	std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
	std::vector<int64_t> load(shards, 0);
	for (int64_t u : order) {
		int best = 0;
		for (int s = 1; s < shards; ++s) if (load[s] < load[best]) best = s;
		shardOf[u] = best;
		load[best] += cost[u];
	}
}

} // namespace gtts

namespace gtts {

// Glottal wavetable of one voice (WavetableGlottalSource.h:111-136): rise 3x^2 - 2x^3, fall 1 - x^2,
// closed phase 0; or one sine period.  For tn_min != tn_max the fall segment stored here is the
// initial one; the kernels evaluate the amplitude-dependent fall segment analytically.
void buildWavetable(const VoiceDev& v, double* table)
{
	if (v.waveform == 0) {
		for (int i = 0; i < v.div1; ++i) {
			const double x = static_cast<double>(i) / static_cast<double>(v.div1);
			const double x2 = x * x;
			const double x3 = x2 * x;
			table[i] = (3.0 * x2) - (2.0 * x3);
		}
		for (int i = v.div1, j = 0; i < v.div2; ++i, ++j) {
			const double x = static_cast<double>(j) / v.tn_length;
			table[i] = 1.0 - (x * x);
		}
		for (int i = v.div2; i < kTableLen; ++i) table[i] = 0.0;
	} else {
		for (int i = 0; i < kTableLen; ++i)
			table[i] = std::sin((static_cast<double>(i) / kTableLen) * 2.0 * kPiTable);
	}
}

} // namespace gtts
