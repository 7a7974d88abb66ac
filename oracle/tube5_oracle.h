/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's model 5 (see tube5_oracle.c). */
#ifndef TUBE5_ORACLE_H_
#define TUBE5_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Same keys as data/voice/english/5_xxx/vtm.txt + variant/<name>.txt (VocalTractModel5.h:373-425). */
typedef struct oracle5_voice {
	double output_rate;
	int    waveform;                 /* 0 pulse, 1 sine */
	double glottal_pulse_tp;
	double glottal_pulse_tn_min;
	double glottal_pulse_tn_max;
	double breathiness;
	double vocal_tract_length_offset;
	double vocal_tract_length;
	double temperature;
	double loss_factor;
	int    noise_modulation;
	double mix_offset;
	double global_radius_coef;
	double global_nasal_radius_coef;
	double nasal_radius[6];          /* nasal_radius_2 .. 7 */
	double radius_coef[8];           /* radius_1_coef .. 8 */
	double glottal_noise_cutoff;
	double frication_noise_cutoff;
	double frication_factor;
	double min_glottal_loss;
	double max_glottal_loss;
	double glottal_lowpass_cutoff;
	int    bypass;
	int    constant_radius_mouth_impedance;
	double mouth_impedance_radius;
} oracle5_voice;

typedef struct oracle5_model oracle5_model;

oracle5_model* oracle5_create(const oracle5_voice* voice);
void   oracle5_destroy(oracle5_model* m);
double oracle5_internal_rate(const oracle5_model* m);
/* One internal sample: setAllParameters(params[16]) + execSynthesisStep(). */
void   oracle5_step(oracle5_model* m, const float* params16);
/* Controller::synthesize over a packed track + finishSynthesis(); returns n_out, writes at most cap samples.
 * steps_override > 0 replaces rint(internal rate / control rate). */
long oracle5_synthesize(const oracle5_voice* voice, double control_rate, int steps_override, const float* frames, long n_frames,
			float* out, long cap, double* internal_rate);

#ifdef __cplusplus
}
#endif
#endif
