// POD types shared by the host runtime and the device kernels.
#ifndef GTTS_TUBE_TYPES_H_
#define GTTS_TUBE_TYPES_H_

#include <stdint.h>

namespace gtts {

enum {
	kNumParams = 16,
	kFirTaps = 49,             // 49 in double (WavetableGlottalSourceFIRFilter.h:74-114 with beta .2, gamma .1)
	kFirMaxTaps = 64,
	kTableLen = 512,           // WavetableGlottalSource.h:92
	kSrcZeroCrossings = 13,    // SampleRateConverter.h:49
	kSrcLRange = 256,
	kSrcFilterLen = 13 * 256,  // 3328
};

// Per-voice constants, derived once on the host in double (VocalTractModel0.h:266-305, 338-392,
// 457-470; WavetableGlottalSource.h:90-141; SampleRateConverter.h:136-164).
struct VoiceDev {
	int32_t fs;                  // internal sample rate (int, VocalTractModel0.h:344)
	int32_t waveform;            // 0 pulse, 1 sine
	int32_t modulation;
	int32_t div1, div2;          // wavetable rise end / fall end
	int32_t src_upsample;        // 1: ratio >= 1 (SampleRateConverter.h:153)
	int32_t src_pad;
	uint32_t src_inc;            // timeRegisterIncrement_
	uint32_t src_phase_inc;      // downsampling only
	int32_t tube_model;          // 0: models 0 / 2; 3: three-sample section delay; 4: 30 + 18 sections (general kernel only)
	double src_ratio;            // output_rate / fs
	double tn_length;            // div2 - div1
	double tn_delta;             // rint(512 (tnMax - tnMin)/100); != 0 -> fall segment depends on amplitude
	double basic_inc;            // 512 / fs
	double breath, one_minus_breath;
	double crossmix;
	double damping;
	double rad_m, refl_b0_m, refl_a1_m;     // mouth: radiation b0 (b1 = a1 = -b0); reflection b0, a1
	double rad_n, refl_b0_n, refl_a1_n;     // nose
	double nasal_k[6];                      // [1..5] fixed; [0] is per-sample (velum)
	double throat_b0, throat_a1, throat_gain;
	double ap2;                             // (aperture_radius * global_radius_coef)^2
	double nr1_2;                           // (nasal_radius_1 * global_nasal_radius_coef)^2
	double radius_coef[8];
	double Ts;                              // 1.0 / fs  (BandpassFilter.h:101)
};

// One utterance of a batch.
struct UttDesc {
	int64_t frame_begin;         // first frame in the packed track array
	int64_t n_frames;
	int64_t out_begin;           // first output sample in the packed output
	int64_t n_out;               // outputs this utterance produces in total (after flush)
	int64_t n_internal;          // n_frames * steps
	int32_t voice;
	int32_t steps;               // internal samples per control period
	float   inv_steps;           // 1.0f / steps (Controller.cpp:287)
	int32_t flags;               // bit0: resume from state[], bit1: do not flush (streaming chunk)
	int64_t state_index;         // index into the UttState array (-1: none)
	int64_t n_in_base;           // internal samples synthesised before this chunk (streaming; 0 for a whole utterance)
};

// Everything one utterance carries from one internal sample to the next (SURVEY.md section 8a,
// "per-instance dynamic state").  Stored in global memory between launches for streaming.
struct UttState {
	double oral_t[10], oral_b[10];
	double nasal_t[6], nasal_b[6];
	double refl_y1_m, rad_x1_m, rad_y1_m;
	double refl_y1_n, rad_x1_n, rad_y1_n;
	double throat_y1;
	double seed, noise_x1;
	double bp_x1, bp_x2, bp_y1, bp_y2;
	double pos;
	double fir_hist[kFirMaxTaps];   // last 48 values of the 2x stream, oldest first
	double src_hist[64];            // last 64 tube outputs, oldest first (2 pad <= 64)
	int64_t n_in_done;              // internal samples consumed so far
	int64_t n_out_done;             // output samples produced so far
	int32_t started;
	int32_t table_low;              // lowest wavetable closure point below div1 seen so far (1 << 30: none)
};

// The same for the pipelined kernel (tube_kernel_v2.cuh): a chunk is drained through every pipeline stage, so all
// roles stop at the same sample and each saves / restores its own part.
struct UttStateV2 {
	double tube[16][4];             // per cell: T, Bn, nb, last (tube_iteration)
	double pos;                     // oscillator phase
	double bp[4];                   // bandpass x1, x2, y1, y2
	double rad[3][2];               // mouth radiation, nose radiation, throat: x1, y1
	double noise_x1;
	unsigned long long lcg;         // noise generator state on the 2^-44 grid
	double vw[24][2];               // last 24 samples of the 2x oscillator stream {even, odd}
	double xring[128];              // tube-output ring (SRC input), by absolute sample index & 127
	float  walk_cur[16];            // the float32 interpolation accumulators (Controller.cpp:307-310) at the end of the chunk
	int32_t table_low;              // lowest wavetable closure point below div1 seen so far (1 << 30: none)
	int32_t started;
};

} // namespace gtts
#endif
