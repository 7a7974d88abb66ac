// POD types of the model-5 path (VocalTractModel5<double, 1>) shared by the host runtime and the device kernel.
#ifndef GTTS_TUBE5_TYPES_H_
#define GTTS_TUBE5_TYPES_H_

#include <stdint.h>

namespace gtts {
namespace m5 {

enum {
	kOral = 30,                 // oropharynx sections S1..S30 (VocalTractModel5.h:150-194)
	kNasal = 21,                // nasal sections N1..N21 (:112-135)
	kMaxPad = 48,               // widest wing of the down-sampling converter this kernel's 128-entry ring holds
};

// Per-voice constants, derived once on the host in double (VocalTractModel5.h:373-425, 460-525, 588-599;
// RosenbergBGlottalSource.h:70-80; Butterworth{1,2}LowpassFilter.h; PoleZeroRadiationImpedance.h:143-176;
// SampleRateConverter.h:136-164).
struct Voice5Dev {
	double fs;                   // internal sample rate: a double here (60,411.43 Hz for 5_male), not truncated
	double Ts;                   // 1.0f / fs (BandpassFilter.h:101, PoleZeroRadiationImpedance.h:116)
	double output_rate;
	int32_t waveform;            // 0 pulse, 1 sine
	int32_t modulation;
	int32_t bypass;              // 1: glottal waveform straight to the converter, no difference filter
	int32_t const_mouth;         // constant_radius_mouth_impedance
	int32_t src_pad;
	uint32_t src_inc;            // timeRegisterIncrement_
	uint32_t src_phase_inc;
	int32_t pad0_;
	double src_ratio;            // output_rate / fs (< 1: the down-sampling branch)
	double breath, one_minus_breath;
	double crossmix;
	double damping;
	double radius_coef[8];
	double nr2_2;                // (nasal_radius_2 * global_nasal_radius_coef)^2, for the first nasal junction
	double nasal_k[7];           // [1..5] = NJ2..NJ6 fixed; [0] (NJ1) is per sample, [6] unused
	double t1, tn_min, tn_max;   // Rosenberg pulse: rise end, fall time range (fractions of the period)
	double gn_b0, gn_a1;         // glottal noise low-pass (Butterworth 1)
	double fn_b0, fn_b1, fn_a1, fn_a2;   // frication noise low-pass (Butterworth 2)
	double gl_b0, gl_a1;         // glottal wave low-pass (Butterworth 1)
	double min_loss, max_loss;   // min / max glottal loss / 100
	double fric_factor;
	double rad_n[5];             // nose radiation impedance: cT1 (= cR1), cT2, cT3, cR2, cR3
	double rad_m[5];             // mouth, when const_mouth
};

static_assert(sizeof(Voice5Dev) % sizeof(double) == 0, "Voice5Dev is copied as doubles");

} // namespace m5
} // namespace gtts
#endif
