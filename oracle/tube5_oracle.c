/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's model 5 (VocalTractModel5<double, 1>), the voice the
 * reference's documentation uses by default (data/voice/english/5_*), driven by the Controller::synthesize
 * interpolation loop.  Plain-C restatement; every function cites the reference lines it follows
 * (/root/reference/gama_tts/src/vtm/...).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may build,
 * load or call this; the product never links it.
 *
 * Parity status: PINNED against the reference itself (oracle/_ref/libgtts_ref_nofma.so, model = 5, built by
 * oracle/Makefile from the unmodified sources): tests/test_oracle.py::test_model5_oracle_vs_reference, bit-identical
 * in -ffp-contract=off builds.  The sample-rate converter is the one of tube_oracle.c (oracle_src_run: both
 * branches pinned there; model 5 runs the down-sampling one, 60,411 Hz -> 48 kHz).
 */
#define _GNU_SOURCE 1
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tube5_oracle.h"

long oracle_src_run(double input_rate, double output_rate, const double* x, long n_in, float* out, long cap);

enum { N_ORAL = 30, N_NASAL = 21, S6 = 5, S12 = 11, S13 = 12, S28 = 27 };

/* PoleZeroRadiationImpedance.h:93-189 */
typedef struct {
	double sample_period, in1, outT1, outR1;
	double cT1, cT2, cT3, cR1, cR2, cR3;
	double prev_radius;
} radiation;

static void rad_init(radiation* r, double fs)
{
	memset(r, 0, sizeof *r);
	r->prev_radius = -1.0;
	r->sample_period = 1.0f / fs;
}

static void rad_update(radiation* r, double radius)           /* :143-176 */
{
	if (radius == r->prev_radius) return;
	r->prev_radius = radius;
	double rr = radius;
	if (rr < 0.5e-2) rr = 0.5e-2;                             /* transitionFrequency, :134-140 */
	const double trans_freq = 62.3371 / rr + 320.204;
	const double cos_wt = cos((2.0 * M_PI) * trans_freq * r->sample_period);
	const double qa = 2.0f * cos_wt;
	const double qb = -2.0f * (cos_wt + 1.0f);
	const double qc = cos_wt + 1.0f;
	const double delta = qb * qb - 4.0f * qa * qc;
	double a = (-qb - sqrt(delta)) / (2.0f * qa);
	const double b = 2.0f * a - 1.0f;
	if (radius < 0.5e-2) a *= 40391.2 * (radius * radius);
	const double coef = 1.0f / (a + 1.0f);
	const double a_plus_b = a + b;
	r->cT1 = a_plus_b * coef;
	r->cT2 = 2.0f * coef;
	r->cT3 = -2.0f * b * coef;
	r->cR1 = a_plus_b * coef;
	r->cR2 = (a - 1.0f) * coef;
	r->cR3 = (b - a) * coef;
}

static void rad_process(radiation* r, double in, double* outT, double* outR)    /* :178-189 */
{
	*outT = r->cT1 * r->outT1 + r->cT2 * in + r->cT3 * r->in1;
	*outR = r->cR1 * r->outR1 + r->cR2 * in + r->cR3 * r->in1;
	r->in1 = in;
	r->outT1 = *outT;
	r->outR1 = *outR;
}

struct oracle5_model {
	oracle5_voice v;
	double fs, breath, crossmix, damping;
	double nasal_radius[7];              /* NR1 (unused, 0) .. NR7, already * global_nasal_radius_coef */
	double radius_coef[8];
	double nasal_k[7];                   /* NJ1 is per sample */
	radiation rad_m, rad_n;
	/* RosenbergBGlottalSource.h */
	double tn_min, tn_max, t1, t2, next_t2, prev_amp, t;
	/* Butterworth filters: glottal noise (1st order), frication noise (2nd), glottal wave (1st) */
	double gn_b0, gn_a1, gn_x1, gn_y1;
	double fn_b0, fn_b1, fn_a1, fn_a2, fn_x1, fn_x2, fn_y1, fn_y2;
	double gl_b0, gl_a1, gl_x1, gl_y1;
	/* BandpassFilter.h */
	double bp_x1, bp_x2, bp_y1, bp_y2, bp_a1, bp_a2, bp_b0, bp_prev_bw, bp_prev_cf;
	double seed;
	double oT[N_ORAL], oB[N_ORAL], nT[N_NASAL], nB[N_NASAL];   /* what the sections hold at their out pointer */
	double* x; long n_x, cap_x;          /* stream fed to the sample-rate converter */
};

static double amp60(double db)                               /* VTMUtil.h:48-67 */
{
	if (db <= 0.0) return 0.0;
	if (db == 60.0) return 1.0;
	db -= 60.0;
	return pow(10.0, db * (1.0 / 20.0));
}

static void butter1(double fs, double fc, double* b0, double* a1)   /* Butterworth1LowpassFilter.h:73-87 */
{
	const double wcT = 2.0f * tan(M_PI * fc / fs);
	const double c1 = 1.0f / (wcT + 2.0f);
	*b0 = c1 * wcT;
	*a1 = c1 * (wcT - 2.0f);
}

oracle5_model* oracle5_create(const oracle5_voice* voice)
{
	oracle5_model* m = (oracle5_model*) calloc(1, sizeof *m);
	if (!m) return NULL;
	m->v = *voice;
	const oracle5_voice* v = &m->v;
	/* VocalTractModel5.h:373-425 loadConfiguration */
	double length = v->vocal_tract_length_offset + v->vocal_tract_length;
	if (length < 3.0) length = 3.0; else if (length > 30.0) length = 30.0;
	m->nasal_radius[0] = 0.0;
	for (int i = 0; i < 6; i++) m->nasal_radius[i + 1] = v->nasal_radius[i] * v->global_nasal_radius_coef;
	for (int i = 0; i < 8; i++) m->radius_coef[i] = v->radius_coef[i] * v->global_radius_coef;
	/* :460-525 initializeSynthesizer */
	const double c = 331.4 + (0.6 * v->temperature);          /* VTMUtil.h:107-113 */
	m->fs = (c * (30u * 1u) * 100.0f) / length;
	m->breath = v->breathiness / 100.0f;
	m->crossmix = 1.0f / amp60(v->mix_offset);
	m->damping = 1.0f - (v->loss_factor / 100.0f);
	m->tn_min = v->glottal_pulse_tn_min / 100.0f;             /* RosenbergBGlottalSource.h:70-80 */
	m->tn_max = v->glottal_pulse_tn_max / 100.0f;
	m->t1 = v->glottal_pulse_tp / 100.0f;
	m->t2 = m->t1 + m->tn_max;
	m->next_t2 = m->t2;
	m->prev_amp = -1.0;
	m->t = 0.0;
	rad_init(&m->rad_m, m->fs);
	if (v->constant_radius_mouth_impedance) rad_update(&m->rad_m, v->mouth_impedance_radius * 1.0e-2f);
	rad_init(&m->rad_n, m->fs);
	/* :588-599 initializeNasalCavity */
	for (int i = 1, j = 1; i < 6; ++i, ++j) {
		const double r0_2 = m->nasal_radius[j] * m->nasal_radius[j], r1_2 = m->nasal_radius[j + 1] * m->nasal_radius[j + 1];
		m->nasal_k[i] = (r0_2 - r1_2) / (r0_2 + r1_2);
	}
	rad_update(&m->rad_n, sqrt(0.5f * m->nasal_radius[6] * m->nasal_radius[6]) * 1.0e-2f);
	butter1(m->fs, v->glottal_noise_cutoff, &m->gn_b0, &m->gn_a1);
	{
		/* Butterworth2LowpassFilter.h:82-102 */
		const double wcT = 2.0f * tan(M_PI * v->frication_noise_cutoff / m->fs);
		const double wc2T2 = wcT * wcT;
		const double c1 = 2.0f * sqrt(2.0) * wcT;
		const double c2 = 1.0f / (wc2T2 + c1 + 4.0f);
		m->fn_b0 = c2 * wc2T2;
		m->fn_b1 = 2.0f * m->fn_b0;
		m->fn_a1 = c2 * (2.0f * wc2T2 - 8.0f);
		m->fn_a2 = c2 * (wc2T2 - c1 + 4.0f);
	}
	butter1(m->fs, v->glottal_lowpass_cutoff, &m->gl_b0, &m->gl_a1);
	m->bp_prev_bw = m->bp_prev_cf = -1.0;
	m->seed = 0.7892347;
	return m;
}

void oracle5_destroy(oracle5_model* m)
{
	if (!m) return;
	free(m->x);
	free(m);
}

double oracle5_internal_rate(const oracle5_model* m) { return m->fs; }

/* VocalTractModel5.h:646-730 vocalTract.  With SectionDelay = 1 the in / out pointers alternate between the two
 * slots of every section: each call reads what the previous call wrote, and every slot read is written once per
 * call -- old arrays in, new arrays out. */
static double tube5(oracle5_model* m, const double* P, const double* junction, const double* velum, double nj1,
			double input, double frication, double glottal_loss)
{
	const double d = m->damping;
	const double *oT = m->oT, *oB = m->oB, *nT = m->nT, *nB = m->nB;
	double ooT[N_ORAL], ooB[N_ORAL], nnT[N_NASAL], nnB[N_NASAL];
	ooT[0] = oB[0] * glottal_loss + input;
	/* region boundaries: junction J(i) sits between sections jl[i] and jl[i] + 1 (S3|S4, S5|S6, S9|S10, S15|S16,
	 * S21|S22, S25|S26, S27|S28); the 3-way junction between S12 and S13 */
	static const int jl[7] = {2, 4, 8, 14, 20, 24, 26};
	int nextj = 0;
	for (int i = 0; i < N_ORAL - 1; ++i) {
		if (nextj < 7 && jl[nextj] == i) {                     /* propagateJunction, :320-325 */
			const double delta = junction[nextj] * (oT[i] + oB[i + 1]);
			ooT[i + 1] = (oT[i] - delta) * d;
			ooB[i] = (oB[i + 1] + delta) * d;
			++nextj;
		} else if (i == S12) {                                 /* 3-way junction, :326-332 */
			const double partial = oT[i] + oB[i + 1] + nB[0];
			ooB[i] = (oB[i + 1] + nB[0] + velum[0] * partial) * d;
			ooT[i + 1] = (oT[i] + nB[0] + velum[1] * partial) * d;
			nnT[0] = (oT[i] + oB[i + 1] + velum[2] * partial) * d;
		} else {                                               /* propagate, :316-319 */
			ooT[i + 1] = oT[i] * d;
			ooB[i] = oB[i + 1] * d;
		}
	}
	double mouth_flow, refl;
	rad_process(&m->rad_m, oT[N_ORAL - 1], &mouth_flow, &refl);
	ooB[N_ORAL - 1] = refl * d;
	for (int i = 0; i < N_NASAL - 1; ++i) {
		if (i % 3 == 2) {                                      /* N3|N4 (NJ1, per sample), N6|N7 (NJ2) ... N18|N19 (NJ6) */
			const int j = i / 3;
			const double k = j == 0 ? nj1 : m->nasal_k[j];
			const double delta = k * (nT[i] + nB[i + 1]);
			nnT[i + 1] = (nT[i] - delta) * d;
			nnB[i] = (nB[i + 1] + delta) * d;
		} else {
			nnT[i + 1] = nT[i] * d;
			nnB[i] = nB[i + 1] * d;
		}
	}
	double nose_flow;
	rad_process(&m->rad_n, nT[N_NASAL - 1], &nose_flow, &refl);
	nnB[N_NASAL - 1] = refl * d;
	/* frication, :711-723 */
	const double fric_offset = (S28 - S6) * (P[4] / 7.0);
	const int fo = (int) fric_offset;
	const double fric_right = fric_offset - fo;
	const double fric_left = 1.0f - fric_right;
	const double fric_value = amp60(P[3]) * frication;
	if (fo >= 0 && S6 + fo <= S28) {                          /* outside: undefined in the reference (array bounds) */
		ooT[S6 + fo] += fric_value * fric_left;
		if (S6 + fo < S28) ooT[S6 + fo + 1] += fric_value * fric_right;
	}
	memcpy(m->oT, ooT, sizeof ooT); memcpy(m->oB, ooB, sizeof ooB);
	memcpy(m->nT, nnT, sizeof nnT); memcpy(m->nB, nnB, sizeof nnB);
	return mouth_flow + nose_flow;
}

/* :776-792 setAllParameters + :527-582 execSynthesisStep */
void oracle5_step(oracle5_model* m, const float* p)
{
	const oracle5_voice* v = &m->v;
	double P[16];
	for (int i = 0; i <= 6; i++) P[i] = p[i];
	for (int i = 7; i <= 14; i++) {
		const double r = p[i] * m->radius_coef[i - 7];
		P[i] = r > 0.01 ? r : 0.01;
	}
	P[15] = p[15];
	const double f0 = 220.0 * pow(2.0, (P[0] + 3.0) * (1.0 / 12.0));    /* VTMUtil.h:76-84 */
	const double glot_amp = amp60(P[1]);
	const double asp_amp = amp60(P[2]);
	/* :610-630 calculateTubeCoefficients */
	double junction[7], velum[3], nj1;
	for (int i = 0; i < 7; i++) {
		const double r0_2 = P[7 + i] * P[7 + i], r1_2 = P[8 + i] * P[8 + i];
		junction[i] = (r0_2 - r1_2) / (r0_2 + r1_2);
	}
	if (!v->constant_radius_mouth_impedance) rad_update(&m->rad_m, P[14] * 1.0e-2f);
	{
		const double r0_2 = P[10] * P[10], r1_2 = r0_2, r2_2 = P[15] * P[15];
		const double c = 1.0f / (r0_2 + r1_2 + r2_2);
		velum[0] = c * (r0_2 - r1_2 - r2_2);
		velum[1] = c * (r1_2 - r0_2 - r2_2);
		velum[2] = c * (r2_2 - r0_2 - r1_2);
	}
	{
		const double r0_2 = P[15] * P[15], r1_2 = m->nasal_radius[1] * m->nasal_radius[1];
		nj1 = (r0_2 - r1_2) / (r0_2 + r1_2);
	}
	/* BandpassFilter.h:88-110 update(sampleRate, bandwidth = FRIC_BW, centerFreq = FRIC_CF) */
	if (!(P[6] == m->bp_prev_bw && P[5] == m->bp_prev_cf)) {
		m->bp_prev_bw = P[6]; m->bp_prev_cf = P[5];
		const double T = 1.0f / m->fs;
		const double tv = tan(M_PI * P[6] * T);
		const double cv = cos(2.0f * M_PI * P[5] * T);
		m->bp_a2 = (1.0f - tv) / (1.0f + tv);
		m->bp_a1 = -(1.0f + m->bp_a2) * cv;
		m->bp_b0 = 0.5f - 0.5f * m->bp_a2;
	}
	/* NoiseSource.h:40-44 */
	const double product = m->seed * 377.0;
	m->seed = product - (int) product;
	const double noise = m->seed - 0.5;
	/* glottal noise: Butterworth1LowpassFilter.h:89-96 */
	const double glottal_noise = m->gn_b0 * (noise + m->gn_x1) - m->gn_a1 * m->gn_y1;
	m->gn_x1 = noise; m->gn_y1 = glottal_noise;
	/* RosenbergBGlottalSource.h:112-150 */
	if (v->waveform == 0 && !(m->tn_min == m->tn_max || glot_amp == m->prev_amp)) {
		m->next_t2 = m->t1 + m->tn_max - glot_amp * (m->tn_max - m->tn_min);
		m->prev_amp = glot_amp;
	}
	double value;
	if (v->waveform == 0) {
		if (m->t < m->t1) {
			const double x = m->t / m->t1;
			value = (x * x) * (3.0f - 2.0f * x);
		} else if (m->t < m->t2) {
			const double x = (m->t - m->t1) / (m->t2 - m->t1);
			value = 1.0f - x * x;
		} else {
			value = 0.0;
		}
	} else {
		value = sin(m->t * (2.0 * M_PI));
	}
	m->t += f0 / m->fs;
	if (m->t > 1.0f) { m->t -= 1.0f; m->t2 = m->next_t2; }
	const double pulse = m->gl_b0 * (value + m->gl_x1) - m->gl_a1 * m->gl_y1;
	m->gl_x1 = value; m->gl_y1 = pulse;
	const double pulsed_noise = glottal_noise * pulse;
	const double noisy_pulse = glot_amp * (pulse * (1.0f - m->breath) + pulsed_noise * m->breath);
	/* frication noise: Butterworth2LowpassFilter.h:104-113 */
	double fric_noise = m->fn_b0 * (noise + m->fn_x2) + m->fn_b1 * m->fn_x1 - m->fn_a1 * m->fn_y1 - m->fn_a2 * m->fn_y2;
	m->fn_x2 = m->fn_x1; m->fn_x1 = noise; m->fn_y2 = m->fn_y1; m->fn_y1 = fric_noise;
	if (v->noise_modulation) {
		double crossmix = glot_amp * m->crossmix;
		crossmix = (crossmix < 1.0f) ? crossmix : 1.0f;
		fric_noise = fric_noise * (noisy_pulse * crossmix + (1.0f - crossmix));
	}
	double signal;
	if (v->bypass == 1) {
		signal = noisy_pulse + asp_amp * fric_noise;
	} else {
		const double min_gl = 1.0f - glot_amp * (v->min_glottal_loss / 100.0f);
		const double max_gl = 1.0f - glot_amp * (v->max_glottal_loss / 100.0f);
		const double gl = min_gl + (max_gl - min_gl) * pulse;
		const double fr = m->bp_b0 * (fric_noise - m->bp_x2) - m->bp_a1 * m->bp_y1 - m->bp_a2 * m->bp_y2;
		m->bp_x2 = m->bp_x1; m->bp_x1 = fric_noise; m->bp_y2 = m->bp_y1; m->bp_y1 = fr;
		signal = tube5(m, P, junction, velum, nj1, noisy_pulse + asp_amp * fric_noise, v->frication_factor * fr, gl);
	}
	if (m->n_x == m->cap_x) {
		m->cap_x = m->cap_x ? m->cap_x * 2 : 4096;
		m->x = (double*) realloc(m->x, sizeof(double) * m->cap_x);
	}
	m->x[m->n_x++] = signal;
}

/* Controller.cpp:277-313 + finishSynthesis (:794-797) + the output callback (:497-513). */
long oracle5_synthesize(const oracle5_voice* voice, double control_rate, int steps_override, const float* frames, long n_frames,
			float* out, long cap, double* internal_rate)
{
	oracle5_model* m = oracle5_create(voice);
	if (!m) return -1;
	if (internal_rate) *internal_rate = m->fs;
	const unsigned steps = steps_override > 0 ? (unsigned) steps_override : (unsigned) rint(m->fs / control_rate);
	const float coef = 1.0f / steps;
	float cur[16], delta[16];
	for (long i = 1; i <= n_frames; ++i) {
		const float* prev = frames + (i - 1) * 16;
		const float* next = (i < n_frames) ? frames + i * 16 : prev;
		for (int j = 0; j < 16; ++j) {
			cur[j] = prev[j];
			delta[j] = (next[j] - cur[j]) * coef;
		}
		for (unsigned j = 0; j < steps; ++j) {
			oracle5_step(m, cur);
			for (int k = 0; k < 16; ++k) cur[k] += delta[k];
		}
	}
	const long bound = (long) ((double) (m->n_x + 200) * (voice->output_rate / m->fs > 1.0 ? voice->output_rate / m->fs : 1.0)) + 256;
	float* buf = (float*) malloc(sizeof(float) * bound);
	long n = oracle_src_run(m->fs, voice->output_rate, m->x, m->n_x, buf, bound);
	if (n > bound) n = bound;
	if (voice->bypass != 1) {
		/* DifferenceFilter<float> (DifferenceFilter.h:62-69) * outputRate, "does not use the 0.5 factor" */
		float x1 = 0.0f, x2 = 0.0f;
		for (long k = 0; k < n; ++k) {
			const float x = buf[k];
			const float y = x - x2;
			x2 = x1; x1 = x;
			buf[k] = (float) (y * voice->output_rate);
		}
	}
	memcpy(out, buf, sizeof(float) * (n < cap ? n : cap));
	free(buf);
	oracle5_destroy(m);
	return n;
}
