/* TEST INFRASTRUCTURE ONLY -- see tube_oracle.h.  Parity status: PINNED against the compiled reference
 * (oracle/_ref) and tests/golden/.  Built with -ffp-contract=off: every double operation below is a
 * single IEEE-754 rounding in the order written, i.e. the reference's arithmetic without FMA
 * contraction (the reference's own FMA-on / FMA-off builds differ by <= 7e-8 of peak).
 *
 * Each function cites the reference lines it restates (paths relative to gama_tts/src/).
 */
#include "tube_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { N_PARAM = 16, N_ORAL = 10, N_NASAL = 6, TABLE_LEN = 512, FIR_LIMIT = 200,
       SRC_ZC = 13, SRC_LRANGE = 256, SRC_FLEN = SRC_ZC * SRC_LRANGE };

struct oracle_model {
	oracle_voice v;
	/* derived (VocalTractModel0.h:338-392) */
	int    fs;
	double breath, crossmix, damping;
	double rad_m, refl_b0_m, refl_a1_m;      /* mouth: radiation b0 (b1 = a1 = -b0), reflection b0, a1 */
	double rad_n, refl_b0_n, refl_a1_n;      /* nose */
	double nasal_k[N_NASAL];                 /* [0] is recomputed per sample from the velum */
	double throat_b0, throat_a1, throat_gain;
	double aperture_radius, nasal_r1;
	double radius_coef[8];
	/* wavetable source (WavetableGlottalSource.h) */
	unsigned div1, div2;
	double tn_length, tn_delta, basic_inc;
	double table[TABLE_LEN];
	double pos, prev_amp;
	/* FIR (WavetableGlottalSourceFIRFilter.h) */
	int    n_taps;
	double fir_coef[2 * FIR_LIMIT + 2];
	double fir_data[2 * FIR_LIMIT + 2];
	int    fir_ptr;
	/* noise */
	double seed, noise_x1;
	/* bandpass */
	double bp_b0, bp_a1, bp_a2, bp_x1, bp_x2, bp_y1, bp_y2, bp_prev_bw, bp_prev_cf, bp_prev_fs;
	/* tube state: [section][0 = top, 1 = bottom], previous-sample values */
	double oral[N_ORAL][2], nasal[N_NASAL][2];
	double refl_y1_m, rad_x1_m, rad_y1_m, refl_y1_n, rad_x1_n, rad_y1_n, throat_y1;
	/* model 3: the two other phases of the three-sample section delay (the waves written at step n are read at step
	   n + 3: three interleaved copies of the wave state, VocalTractModel2.h:234-250); model 4: its 30 + 18 sections */
	double oral_d[2][N_ORAL][2], nasal_d[2][N_NASAL][2];
	long   step;
	double oral4[30][2], nasal4[18][2];
	/* current parameters in double (setAllParameters) */
	double P[N_PARAM];
	/* SRC: all tube output is kept, conversion is done by the closed form at finish */
	double* x; long n_x, cap_x;
	float* out; long n_out;
};

/* ---- vtm/VTMUtil.h:50-67, 76-84 ---- */
static double amp60(double db)
{
	if (db <= 0.0) return 0.0;
	if (db == 60.0) return 1.0;
	db -= 60.0;
	return pow(10.0, db * (1.0 / 20.0));
}

static double pitch_to_freq(double pitch)
{
	return 220.0 * pow(2.0, (pitch + 3.0) * (1.0 / 12.0));
}

/* ---- vtm/WavetableGlottalSourceFIRFilter.h:336-382 (rationalApproximation) ---- */
static void rational_approx(double number, int* order, int* numerator, int* denominator)
{
	if (*order <= 0) { *numerator = 0; *denominator = 0; *order = -1; return; }
	const double frac = fabs(number - (int) number);
	int order_max = 2 * (*order);
	if (order_max > FIR_LIMIT) order_max = FIR_LIMIT;
	double min_err = 1.0;
	int modulus = 0;
	for (int i = *order; i <= order_max; i++) {
		const double ps = i * frac;
		const int ip = (int) (ps + 0.5);
		const double err = fabs((ps - (double) ip) / i);
		if (err < min_err) { min_err = err; modulus = ip; *denominator = i; }
	}
	*numerator = (int) fabs(number) * (*denominator) + modulus;
	if (number < 0.0) *numerator *= -1;
	*order = *denominator - 1;
	if (*numerator == *denominator) {
		*denominator = order_max;
		*order = *numerator = *denominator - 1;
	}
}

/* ---- :137-215 (maximallyFlat), :228-237 (trim), :74-114 (tap layout) ---- */
static int design_fir(double beta, double gamma, double cutoff, double* taps)
{
	double a[FIR_LIMIT + 1], c[FIR_LIMIT + 1], coef[FIR_LIMIT + 1];
	int np = 0, numerator;
	int nt = (int) (1.0 / (4.0 * gamma * gamma));
	const double ac = (1.0 + cos((2.0 * M_PI) * beta)) / 2.0;
	rational_approx(ac, &nt, &numerator, &np);
	const int n = 2 * np - 1;
	if (numerator == 0) numerator = 1;
	c[1] = a[1] = 1.0;
	const int ll = nt - numerator;
	for (int i = 2; i <= np; i++) {
		c[i] = cos((2.0 * M_PI) * ((double) (i - 1) / n));
		const double x = (1.0 - c[i]) / 2.0;
		double y = x;
		if (numerator == nt) continue;
		double sum = 1.0;
		for (int j = 1; j <= ll; j++) {
			double z = y;
			if (numerator != 1) {
				for (int jj = 1; jj <= numerator - 1; jj++) z *= 1.0 + ((double) j / jj);
			}
			y *= x;
			sum += z;
		}
		a[i] = sum * pow(1.0 - x, numerator);
	}
	for (int i = 1; i <= np; i++) {
		coef[i] = a[1] / 2.0;
		for (int j = 2; j <= np; j++) {
			int m = ((i - 1) * (j - 1)) % n;
			if (m > nt) m = n - m;
			coef[i] += c[m + 1] * a[j];
		}
		coef[i] *= 2.0 / (double) n;
	}
	int ncoef = np;
	for (int i = ncoef; i > 0; i--) {
		if (fabs(coef[i]) >= fabs(cutoff)) { ncoef = i; break; }
	}
	const int n_taps = ncoef * 2 - 1;
	int inc = -1, ptr = ncoef;
	for (int i = 0; i < n_taps; i++) {
		taps[i] = coef[ptr];
		ptr += inc;
		if (ptr <= 0) { ptr = 2; inc = 1; }
	}
	return n_taps;
}

/* ---- vtm/SampleRateConverter.h:175-194 (Izero), 230-255 (initializeFilter) ---- */
static double izero(double x)
{
	double sum = 1.0, u = 1.0;
	int n = 1;
	const double halfx = x / 2.0;
	do {
		double t = halfx / n;
		n += 1;
		t *= t;
		u *= t;
		sum += u;
	} while (u >= 1e-21 * sum);
	return sum;
}

static double g_h[SRC_FLEN], g_dh[SRC_FLEN];
static int g_src_ready = 0;

static void src_init_tables(void)
{
	if (g_src_ready) return;
	const double beta = 5.658, lp = 11.0 / 13.0;
	g_h[0] = lp;
	const double x = M_PI / SRC_LRANGE;
	for (int i = 1; i < SRC_FLEN; i++) {
		const double y = i * x;
		g_h[i] = sin(y * lp) / y;
	}
	const double ibeta = 1.0 / izero(beta);
	for (int i = 0; i < SRC_FLEN; i++) {
		const double t = (double) i / SRC_FLEN;
		g_h[i] *= izero(beta * sqrt(1.0 - (t * t))) * ibeta;
	}
	for (int i = 0; i < SRC_FLEN - 1; i++) g_dh[i] = g_h[i + 1] - g_h[i];
	g_dh[SRC_FLEN - 1] = 0.0 - g_h[SRC_FLEN - 1];
	g_src_ready = 1;
}

/* ---- :136-164 (initializeConversion) ---- */
static void src_params(double in_rate, double out_rate, double* ratio, unsigned* inc, unsigned* phase_inc, int* pad)
{
	*ratio = out_rate / in_rate;
	*inc = (unsigned) rint(pow(2.0, 16) / *ratio);
	const double rounded = pow(2.0, 16) / *inc;
	*phase_inc = 0;
	if (*ratio >= 1.0) {
		*pad = SRC_ZC;
	} else {
		*phase_inc = (unsigned) rint(*ratio * 65536);
		*pad = (int) (SRC_ZC / rounded) + 1;
	}
}

/* Whole-signal form of dataFill/dataEmpty/flushBuffer (:268-282, 295-416, 462-471).  The ring-buffer
 * bookkeeping of the reference reduces to: output k is centred on ring position e = (k*inc)>>16 with
 * fraction f = (k*inc)&0xFFFF; input i sits at ring position pad+i; positions outside the fed data
 * read 0; 2*pad zeros are appended by the flush; outputs exist while e < n_in + 2*pad.
 * (SURVEY.md section 4 verified this closed form bit-identical; tests re-verify it against oracle/_ref.) */
static long src_convert(double in_rate, double out_rate, const double* x, long n_in, float** out_p)
{
	src_init_tables();
	double ratio; unsigned inc, phase_inc; int pad;
	src_params(in_rate, out_rate, &ratio, &inc, &phase_inc, &pad);
	const long n_pos = n_in + 2L * pad;              /* endPtr after the flush */
	const unsigned long long lim = (unsigned long long) n_pos << 16;
	const long n_out = (long) ((lim + inc - 1) / inc);
	float* out = (float*) malloc(sizeof(float) * (n_out > 0 ? n_out : 1));
#define XAT(pos) (((pos) - pad) >= 0 && ((pos) - pad) < n_in ? x[(pos) - pad] : 0.0)
	for (long k = 0; k < n_out; k++) {
		const unsigned long long t = (unsigned long long) k * inc;
		const long e = (long) (t >> 16);
		const unsigned f = (unsigned) (t & 0xFFFF);
		double acc = 0.0;
		if (ratio >= 1.0) {
			double interp = (double) (f & 0xFF) / 256;
			long idx = e;
			for (unsigned fi = (f >> 8) & 0xFF; fi < SRC_FLEN; fi += SRC_LRANGE, idx--)
				acc += (XAT(idx) * (g_h[fi] + (g_dh[fi] * interp)));
			const unsigned g = (~f) & 0xFFFF;
			interp = (double) (g & 0xFF) / 256;
			idx = e + 1;
			for (unsigned fi = (g >> 8) & 0xFF; fi < SRC_FLEN; fi += SRC_LRANGE, idx++)
				acc += (XAT(idx) * (g_h[fi] + (g_dh[fi] * interp)));
		} else {
			unsigned ph = (unsigned) rint(f * ratio);
			long idx = e;
			unsigned ii;
			while ((ii = (ph >> 8)) < SRC_FLEN) {
				const double imp = g_h[ii] + (g_dh[ii] * ((double) (ph & 0xFF) / 256));
				acc += XAT(idx) * imp;
				idx--;
				ph += phase_inc;
			}
			ph = (unsigned) rint((double) ((~f) & 0xFFFF) * ratio);
			idx = e + 1;
			while ((ii = (ph >> 8)) < SRC_FLEN) {
				const double imp = g_h[ii] + (g_dh[ii] * ((double) (ph & 0xFF) / 256));
				acc += XAT(idx) * imp;
				idx++;
				ph += phase_inc;
			}
		}
		out[k] = (float) acc;
	}
#undef XAT
	*out_p = out;
	return n_out;
}

/* ---- vtm/WavetableGlottalSource.h:90-141 ---- */
static void wavetable_init(oracle_model* m)
{
	const oracle_voice* v = &m->v;
	m->div1 = (unsigned) rint(TABLE_LEN * (v->glottal_pulse_tp / 100.0));
	m->div2 = (unsigned) rint(TABLE_LEN * ((v->glottal_pulse_tp + v->glottal_pulse_tn_max) / 100.0));
	m->tn_length = (double) (m->div2 - m->div1);
	m->tn_delta = rint(TABLE_LEN * ((v->glottal_pulse_tn_max - v->glottal_pulse_tn_min) / 100.0));
	m->basic_inc = TABLE_LEN / (double) m->fs;
	if (v->waveform == 0) {
		for (unsigned i = 0; i < m->div1; i++) {
			const double x = (double) i / m->div1;
			const double x2 = x * x;
			const double x3 = x2 * x;
			m->table[i] = (3.0 * x2) - (2.0 * x3);
		}
		for (unsigned i = m->div1, j = 0; i < m->div2; i++, j++) {
			const double x = (double) j / m->tn_length;
			m->table[i] = 1.0 - (x * x);
		}
		for (unsigned i = m->div2; i < TABLE_LEN; i++) m->table[i] = 0.0;
	} else {
		for (unsigned i = 0; i < TABLE_LEN; i++)
			m->table[i] = sin(((double) i / TABLE_LEN) * 2.0 * M_PI);
	}
}

/* ---- :162-184 (setup) ---- */
static void wavetable_setup(oracle_model* m, double amplitude)
{
	if (m->tn_delta == 0.0 || amplitude == m->prev_amp) return;
	m->prev_amp = amplitude;
	double new_div2 = m->div2 - rint(amplitude * m->tn_delta);
	if (!(new_div2 > 0.0)) new_div2 = 0.0;
	const double inv = 1.0 / (new_div2 - m->div1);
	double x = 0.0;
	const unsigned end = (unsigned) new_div2;
	for (unsigned i = m->div1; i < end; ++i, x += inv) m->table[i] = 1.0 - (x * x);
	for (unsigned i = end; i < m->div2; i++) m->table[i] = 0.0;
}

static double mod0(double v)
{
	if (v > TABLE_LEN - 1) v -= TABLE_LEN;
	return v;
}

/* ---- WavetableGlottalSourceFIRFilter.h:276-304 ---- */
static double fir_push(oracle_model* m, double in, int need_output)
{
	if (need_output) {
		double out = 0.0;
		m->fir_data[m->fir_ptr] = in;
		for (int i = 0; i < m->n_taps; i++) {
			out += m->fir_data[m->fir_ptr] * m->fir_coef[i];
			if (++m->fir_ptr >= m->n_taps) m->fir_ptr = 0;
		}
		if (--m->fir_ptr < 0) m->fir_ptr = m->n_taps - 1;
		return out;
	}
	m->fir_data[m->fir_ptr] = in;
	if (--m->fir_ptr < 0) m->fir_ptr = m->n_taps - 1;
	return 0.0;
}

/* ---- WavetableGlottalSource.h:196-235 (2x oversampling oscillator) ---- */
static double glottal_sample(oracle_model* m, double f0)
{
	double out = 0.0;
	for (int i = 0; i < 2; i++) {
		m->pos = mod0(m->pos + ((f0 / 2.0) * m->basic_inc));
		const unsigned lo = (unsigned) (long long) m->pos;   /* (-1,0) truncates to 0 as on x86-64 */
		const unsigned up = (unsigned) mod0((double) (lo + 1));
		const double v = m->table[lo] + ((m->pos - lo) * (m->table[up] - m->table[lo]));
		out = fir_push(m, v, i);
	}
	return out;
}

/* ---- vtm/VocalTractModel0.h:309-326 and the sub-objects' reset() ---- */
void oracle_reset(oracle_model* m)
{
	memset(m->oral, 0, sizeof m->oral);
	memset(m->nasal, 0, sizeof m->nasal);
	memset(m->oral_d, 0, sizeof m->oral_d);
	memset(m->nasal_d, 0, sizeof m->nasal_d);
	memset(m->oral4, 0, sizeof m->oral4);
	memset(m->nasal4, 0, sizeof m->nasal4);
	m->step = 0;
	m->refl_y1_m = m->rad_x1_m = m->rad_y1_m = 0.0;
	m->refl_y1_n = m->rad_x1_n = m->rad_y1_n = 0.0;
	m->throat_y1 = 0.0;
	m->pos = 0.0;
	m->prev_amp = -1.0;
	memset(m->fir_data, 0, sizeof m->fir_data);
	m->fir_ptr = 0;
	m->bp_x1 = m->bp_x2 = m->bp_y1 = m->bp_y2 = 0.0;
	m->bp_prev_bw = m->bp_prev_cf = m->bp_prev_fs = -1.0;
	m->noise_x1 = 0.0;
	m->seed = 0.7892347;
	m->n_x = 0;
	free(m->out); m->out = NULL; m->n_out = 0;
	/* note: like the reference, reset() leaves P[] (currentParameter_) and the wavetable untouched */
}

/* ---- :266-305 (loadConfiguration), :338-392 (initializeSynthesizer), :457-470 ---- */
oracle_model* oracle_create(const oracle_voice* voice)
{
	oracle_model* m = (oracle_model*) calloc(1, sizeof *m);
	if (!m) return NULL;
	m->v = *voice;
	const oracle_voice* v = &m->v;
	double length = v->vocal_tract_length_offset + v->vocal_tract_length;
	if (length < 3.0) length = 3.0; else if (length > 30.0) length = 30.0;
	m->aperture_radius = v->aperture_radius * v->global_radius_coef;
	double nr[N_NASAL];
	nr[0] = 0.0;
	for (int i = 0; i < 5; i++) nr[i + 1] = v->nasal_radius[i] * v->global_nasal_radius_coef;
	for (int i = 0; i < 8; i++) m->radius_coef[i] = v->radius_coef[i] * v->global_radius_coef;
	m->nasal_r1 = nr[1];

	const double c = 331.4 + (0.6 * v->temperature);
	/* VocalTractModel0.h:344; VocalTractModel2.h:419 with SectionDelay = 3 and VocalTractModel4.h (30 sections): 30 */
	m->fs = (int) ((c * (v->tube_model == 0 ? N_ORAL : 30) * 100.0) / length);
	const double nyquist = (double) ((float) m->fs / 2.0f);
	m->breath = v->breathiness / 100.0;
	m->crossmix = 1.0 / amp60(v->mix_offset);
	m->damping = 1.0 - (v->loss_factor / 100.0);
	wavetable_init(m);
	const double am = (nyquist - v->mouth_coefficient) / nyquist;
	m->rad_m = am; m->refl_b0_m = 1.0 - fabs(am); m->refl_a1_m = -am;
	const double an = (nyquist - v->nose_coefficient) / nyquist;
	m->rad_n = an; m->refl_b0_n = 1.0 - fabs(an); m->refl_a1_n = -an;
	for (int i = 1; i < 5; i++) {
		const double a2 = nr[i] * nr[i], b2 = nr[i + 1] * nr[i + 1];
		m->nasal_k[i] = (a2 - b2) / (a2 + b2);
	}
	{
		const double a2 = nr[5] * nr[5], b2 = m->aperture_radius * m->aperture_radius;
		m->nasal_k[5] = (a2 - b2) / (a2 + b2);
	}
	m->throat_b0 = (v->throat_cutoff * 2.0) / m->fs;
	m->throat_a1 = m->throat_b0 - 1.0;
	m->throat_gain = amp60(v->throat_volume);
	m->n_taps = design_fir(0.2, 0.1, 0.00000001, m->fir_coef);
	src_init_tables();
	oracle_reset(m);
	return m;
}

void oracle_destroy(oracle_model* m)
{
	if (!m) return;
	free(m->x);
	free(m->out);
	free(m);
}

double oracle_internal_rate(const oracle_model* m) { return (double) m->fs; }

static double refl(double b0, double a1, double* y1, double x)
{
	const double y = b0 * x - a1 * *y1;
	*y1 = y;
	return y;
}

static double rad(double a, double* x1, double* y1, double x)
{
	const double y = a * x + (-a) * *x1 - (-a) * *y1;
	*x1 = x;
	*y1 = y;
	return y;
}

/* ---- :565-661 (vocalTract); o/n hold the previous sample, results are written to fresh arrays ---- */
static double tube(oracle_model* m, double input, double fric, const double* k, const double* alpha, const double* tap)
{
	enum { T = 0, B = 1 };
	const double d = m->damping;
	double (*o)[2] = m->oral, (*n)[2] = m->nasal;
	double oo[N_ORAL][2], nn[N_NASAL][2];
	double dl;

	oo[0][T] = (o[0][B] * d) + input;
	dl = k[0] * (o[0][T] - o[1][B]);
	oo[1][T] = (o[0][T] + dl) * d;
	oo[0][B] = (o[1][B] + dl) * d;
	for (int i = 1, j = 1, t = 0; i < 3; i++, j++, t++) {
		dl = k[j] * (o[i][T] - o[i + 1][B]);
		oo[i + 1][T] = ((o[i][T] + dl) * d) + (tap[t] * fric);
		oo[i][B] = (o[i + 1][B] + dl) * d;
	}
	const double jp = (alpha[0] * o[3][T]) + (alpha[1] * o[4][B]) + (alpha[2] * n[0][B]);
	oo[3][B] = (jp - o[3][T]) * d;
	oo[4][T] = ((jp - o[4][B]) * d) + (tap[2] * fric);
	nn[0][T] = (jp - n[0][B]) * d;
	dl = k[3] * (o[4][T] - o[5][B]);
	oo[5][T] = ((o[4][T] + dl) * d) + (tap[3] * fric);
	oo[4][B] = (o[5][B] + dl) * d;
	oo[6][T] = (o[5][T] * d) + (tap[4] * fric);
	oo[5][B] = o[6][B] * d;
	for (int i = 6, j = 4, t = 5; i < 9; i++, j++, t++) {
		dl = k[j] * (o[i][T] - o[i + 1][B]);
		oo[i + 1][T] = ((o[i][T] + dl) * d) + (tap[t] * fric);
		oo[i][B] = (o[i + 1][B] + dl) * d;
	}
	oo[9][B] = d * refl(m->refl_b0_m, m->refl_a1_m, &m->refl_y1_m, k[7] * o[9][T]);
	double output = rad(m->rad_m, &m->rad_x1_m, &m->rad_y1_m, (1.0 + k[7]) * o[9][T]);
	for (int i = 0; i < 5; i++) {
		dl = m->nasal_k[i] * (n[i][T] - n[i + 1][B]);
		nn[i + 1][T] = (n[i][T] + dl) * d;
		nn[i][B] = (n[i + 1][B] + dl) * d;
	}
	nn[5][B] = d * refl(m->refl_b0_n, m->refl_a1_n, &m->refl_y1_n, m->nasal_k[5] * n[5][T]);
	output += rad(m->rad_n, &m->rad_x1_n, &m->rad_y1_n, (1.0 + m->nasal_k[5]) * n[5][T]);
	memcpy(m->oral, oo, sizeof oo);
	memcpy(m->nasal, nn, sizeof nn);
	return output;
}

/* ---- model 3: VocalTractModel2<double, 3>::vocalTract (VocalTractModel2.h:626-670) ----
 * The same junctions as model 0, every section a delay line of three samples (in / out pointers over four slots,
 * :234-250): what a step writes is read three steps later, so the wave state is three interleaved copies, each
 * advanced every third step.  The end filters (reflection, radiation) run on every sample. */
static double tube3(oracle_model* m, double input, double fric, const double* k, const double* alpha, const double* tap)
{
	const int ph = (int) (m->step % 3);
	m->step += 1;
	if (ph != 0) {
		/* bring phase ph's copy into the working arrays, run the model-0 step, put it back */
		double so[N_ORAL][2], sn[N_NASAL][2];
		memcpy(so, m->oral, sizeof so); memcpy(sn, m->nasal, sizeof sn);
		memcpy(m->oral, m->oral_d[ph - 1], sizeof so); memcpy(m->nasal, m->nasal_d[ph - 1], sizeof sn);
		const double out = tube(m, input, fric, k, alpha, tap);
		memcpy(m->oral_d[ph - 1], m->oral, sizeof so); memcpy(m->nasal_d[ph - 1], m->nasal, sizeof sn);
		memcpy(m->oral, so, sizeof so); memcpy(m->nasal, sn, sizeof sn);
		return out;
	}
	return tube(m, input, fric, k, alpha, tap);
}

/* ---- model 4: VocalTractModel4<double, 1>::vocalTract (VocalTractModel4.h:671-744) ----
 * 30 oropharynx + 18 nasal sections of one sample; junctions J1..J7 between S3|S4, S5|S6, S9|S10, S15|S16, S21|S22,
 * S25|S26, S27|S28, the 3-way junction between S12 and S13, nasal junctions every three sections; sections inside a
 * region are plain copies (no damping, :303-307) except S18 -> S19, which carries frication tap FC5 (:308-312). */
static double tube4(oracle_model* m, double input, double fric, const double* k, const double* alpha, const double* tap)
{
	enum { T = 0, B = 1, NO = 30, NN = 18 };
	const double d = m->damping;
	double (*o)[2] = m->oral4, (*n)[2] = m->nasal4;
	double oo[NO][2], nn[NN][2];
	/* junction index at the boundary i | i + 1 (or -1), frication tap injected into section i + 1 (or -1) */
	static const int jn[NO - 1] = {-1, -1, 0, -1, 1, -1, -1, -1, 2, -1, -1, /*S12|S13: 3-way*/ -2, -1, -1, 3, -1, -1, /*S18|S19*/ -3, -1, -1, 4, -1, -1, -1, 5, -1, 6, -1, -1};
	static const int ft[NO - 1] = {-1, -1, -1, -1, 0, -1, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, 5, -1, -1, -1, 6, -1, 7, -1, -1};
	oo[0][T] = o[0][B] * d + input;
	for (int i = 0; i < NO - 1; ++i) {
		const double fn = ft[i] >= 0 ? tap[ft[i]] * fric : 0.0;
		if (jn[i] >= 0) {
			const double delta = k[jn[i]] * (o[i][T] - o[i + 1][B]);
			oo[i + 1][T] = (o[i][T] + delta) * d + fn;
			oo[i][B] = (o[i + 1][B] + delta) * d;
		} else if (jn[i] == -2) {
			const double jp = alpha[0] * o[i][T] + alpha[1] * o[i + 1][B] + alpha[2] * n[0][B];
			oo[i][B] = (jp - o[i][T]) * d;
			oo[i + 1][T] = (jp - o[i + 1][B]) * d + fn;
			nn[0][T] = (jp - n[0][B]) * d;
		} else if (jn[i] == -3) {
			oo[i + 1][T] = o[i][T] * d + fn;
			oo[i][B] = o[i + 1][B] * d;
		} else {
			oo[i + 1][T] = o[i][T];
			oo[i][B] = o[i + 1][B];
		}
	}
	oo[NO - 1][B] = d * refl(m->refl_b0_m, m->refl_a1_m, &m->refl_y1_m, k[7] * o[NO - 1][T]);
	double output = rad(m->rad_m, &m->rad_x1_m, &m->rad_y1_m, (1.0 + k[7]) * o[NO - 1][T]);
	for (int i = 0; i < NN - 1; ++i) {
		if (i % 3 == 2) {
			const double delta = m->nasal_k[i / 3] * (n[i][T] - n[i + 1][B]);
			nn[i + 1][T] = (n[i][T] + delta) * d + 0.0;
			nn[i][B] = (n[i + 1][B] + delta) * d;
		} else {
			nn[i + 1][T] = n[i][T];
			nn[i][B] = n[i + 1][B];
		}
	}
	nn[NN - 1][B] = d * refl(m->refl_b0_n, m->refl_a1_n, &m->refl_y1_n, m->nasal_k[5] * n[NN - 1][T]);
	output += rad(m->rad_n, &m->rad_x1_n, &m->rad_y1_n, (1.0 + m->nasal_k[5]) * n[NN - 1][T]);
	memcpy(m->oral4, oo, sizeof oo);
	memcpy(m->nasal4, nn, sizeof nn);
	return output;
}

/* ---- :698-716 (setAllParameters) + :396-445 (execSynthesisStep) ---- */
void oracle_step(oracle_model* m, const float* p)
{
	double* P = m->P;
	for (int i = 0; i <= 6; i++) P[i] = p[i];
	for (int i = 7; i <= 14; i++) {
		const double r = p[i] * m->radius_coef[i - 7];
		P[i] = r > 0.01 ? r : 0.01;          /* std::max(r, 0.01) */
	}
	P[15] = p[15];

	const double f0 = pitch_to_freq(P[0]);
	const double ax = amp60(P[1]);
	const double ah1 = amp60(P[2]);

	/* :484-512 calculateTubeCoefficients */
	double k[8], alpha[3];
	for (int i = 0; i < 7; i++) {
		const double a2 = P[7 + i] * P[7 + i], b2 = P[8 + i] * P[8 + i];
		k[i] = (a2 - b2) / (a2 + b2);
	}
	{
		const double a2 = P[14] * P[14], b2 = m->aperture_radius * m->aperture_radius;
		k[7] = (a2 - b2) / (a2 + b2);
	}
	const double r1_2 = P[10] * P[10], r0_2 = r1_2, r2_2 = P[15] * P[15];
	const double sum = 2.0 / (r0_2 + r1_2 + r2_2);
	alpha[0] = sum * r0_2; alpha[1] = sum * r1_2; alpha[2] = sum * r2_2;
	{
		const double b2 = m->nasal_r1 * m->nasal_r1;
		m->nasal_k[0] = (r2_2 - b2) / (r2_2 + b2);
	}

	/* :524-552 setFricationTaps */
	double tap[8];
	const double fa = amp60(P[3]);
	const int ip = (int) P[4];
	const double complement = P[4] - ip, remainder = 1.0 - complement;
	for (int i = 0; i < 8; i++) {
		if (i == ip) {
			tap[i] = remainder * fa;
			if (i + 1 < 8) tap[++i] = complement * fa;
		} else {
			tap[i] = 0.0;
		}
	}

	/* BandpassFilter.h:91-110 update */
	if (!((double) m->fs == m->bp_prev_fs && P[6] == m->bp_prev_bw && P[5] == m->bp_prev_cf)) {
		m->bp_prev_fs = m->fs; m->bp_prev_bw = P[6]; m->bp_prev_cf = P[5];
		const double Ts = 1.0 / m->fs;
		const double tv = tan(M_PI * P[6] * Ts);
		const double cv = cos(2.0 * M_PI * P[5] * Ts);
		m->bp_a2 = (1.0 - tv) / (1.0 + tv);
		m->bp_a1 = -(1.0 + m->bp_a2) * cv;
		m->bp_b0 = 0.5 - 0.5 * m->bp_a2;
	}

	/* NoiseSource.h:40-44, NoiseFilter.h:63-68 */
	const double product = m->seed * 377.0;
	m->seed = product - (int) product;
	const double noise = m->seed - 0.5;
	const double lp = noise + m->noise_x1;
	m->noise_x1 = noise;

	if (m->v.waveform == 0) wavetable_setup(m, ax);
	double pulse = glottal_sample(m, f0);
	const double pn = lp * pulse;
	pulse = ax * ((pulse * (1.0 - m->breath)) + (pn * m->breath));
	double sig;
	if (m->v.noise_modulation) {
		double cm = ax * m->crossmix;
		cm = (cm < 1.0) ? cm : 1.0;
		sig = (pn * cm) + (lp * (1.0 - cm));
	} else {
		sig = lp;
	}
	/* BandpassFilter.h:114-122 */
	const double fr = m->bp_b0 * (sig - m->bp_x2) - m->bp_a1 * m->bp_y1 - m->bp_a2 * m->bp_y2;
	m->bp_x2 = m->bp_x1; m->bp_x1 = sig; m->bp_y2 = m->bp_y1; m->bp_y1 = fr;

	const double tube_in = (pulse + (ah1 * sig)) * 0.125;
	double s = m->v.tube_model == 3 ? tube3(m, tube_in, fr, k, alpha, tap)
	         : (m->v.tube_model == 4 ? tube4(m, tube_in, fr, k, alpha, tap) : tube(m, tube_in, fr, k, alpha, tap));
	/* Throat.h:80-85 */
	{
		const double y = m->throat_b0 * (pulse * 0.125) - m->throat_a1 * m->throat_y1;
		m->throat_y1 = y;
		s += y * m->throat_gain;
	}
	if (m->n_x == m->cap_x) {
		m->cap_x = m->cap_x ? m->cap_x * 2 : 4096;
		m->x = (double*) realloc(m->x, sizeof(double) * m->cap_x);
	}
	m->x[m->n_x++] = s;
}

/* ---- vtm_control_model/Controller.cpp:277-313 ---- */
void oracle_run_track(oracle_model* m, double control_rate, const float* frames, long n_frames)
{
	if (n_frames <= 0) { oracle_finish(m); return; }   /* Controller.cpp:232-233 */
	const unsigned steps = (unsigned) rint((double) m->fs / control_rate);
	const float coef = 1.0f / steps;
	float cur[N_PARAM], delta[N_PARAM];
	for (long i = 1; i <= n_frames; i++) {
		const float* prev = frames + (i - 1) * N_PARAM;
		const float* next = (i < n_frames) ? frames + i * N_PARAM : prev;
		for (int j = 0; j < N_PARAM; j++) {
			cur[j] = prev[j];
			delta[j] = (next[j] - cur[j]) * coef;
		}
		for (unsigned j = 0; j < steps; j++) {
			oracle_step(m, cur);
			for (int k = 0; k < N_PARAM; k++) cur[k] += delta[k];
		}
	}
	oracle_finish(m);
}

void oracle_finish(oracle_model* m)
{
	free(m->out);
	m->n_out = src_convert((double) m->fs, m->v.output_rate, m->x, m->n_x, &m->out);
}

long oracle_output_size(const oracle_model* m) { return m->n_out; }
const float* oracle_output(const oracle_model* m) { return m->out; }
long oracle_internal_size(const oracle_model* m) { return m->n_x; }
const double* oracle_internal(const oracle_model* m) { return m->x; }

long oracle_synthesize(const oracle_voice* voice, double control_rate, const float* frames, long n_frames,
			float* out, long cap)
{
	oracle_model* m = oracle_create(voice);
	if (!m) return -1;
	oracle_run_track(m, control_rate, frames, n_frames);
	const long n = m->n_out;
	if (out) memcpy(out, m->out, sizeof(float) * (n < cap ? n : cap));
	oracle_destroy(m);
	return n;
}

/* ---- probes ---- */
void oracle_noise(double* out, long n)
{
	double seed = 0.7892347;
	for (long i = 0; i < n; i++) {
		const double product = seed * 377.0;
		seed = product - (int) product;
		out[i] = seed - 0.5;
	}
}

int oracle_fir_taps(double* out, int cap)
{
	double taps[2 * FIR_LIMIT + 2];
	const int n = design_fir(0.2, 0.1, 0.00000001, taps);
	for (int i = 0; i < n && i < cap; i++) out[i] = taps[i];
	return n;
}

void oracle_src_tables(double* h, double* dh)
{
	src_init_tables();
	memcpy(h, g_h, sizeof g_h);
	memcpy(dh, g_dh, sizeof g_dh);
}

void oracle_src_params(double input_rate, double output_rate, unsigned* inc, unsigned* phase_inc, int* pad)
{
	double ratio;
	src_params(input_rate, output_rate, &ratio, inc, phase_inc, pad);
}

long oracle_src_run(double input_rate, double output_rate, const double* x, long n_in, float* out, long cap)
{
	float* buf = NULL;
	const long n = src_convert(input_rate, output_rate, x, n_in, &buf);
	memcpy(out, buf, sizeof(float) * (n < cap ? n : cap));
	free(buf);
	return n;
}

void oracle_wavetable(const oracle_voice* voice, double amplitude, double* table512, double* scalars5)
{
	oracle_model* m = oracle_create(voice);
	if (amplitude >= 0.0) wavetable_setup(m, amplitude);
	memcpy(table512, m->table, sizeof m->table);
	scalars5[0] = m->div1; scalars5[1] = m->div2; scalars5[2] = m->tn_length;
	scalars5[3] = m->tn_delta; scalars5[4] = m->basic_inc;
	oracle_destroy(m);
}

int oracle_constants(const oracle_voice* voice, double* out)
{
	oracle_model* m = oracle_create(voice);
	int k = 0;
	out[k++] = m->fs; out[k++] = m->breath; out[k++] = m->crossmix; out[k++] = m->damping;
	out[k++] = m->rad_m; out[k++] = m->refl_b0_m; out[k++] = m->refl_a1_m;
	out[k++] = m->rad_n; out[k++] = m->refl_b0_n; out[k++] = m->refl_a1_n;
	out[k++] = m->throat_b0; out[k++] = m->throat_a1; out[k++] = m->throat_gain;
	for (int i = 1; i < 6; i++) out[k++] = m->nasal_k[i];
	out[k++] = m->aperture_radius; out[k++] = m->nasal_r1; out[k++] = m->basic_inc;
	out[k++] = m->div1; out[k++] = m->div2; out[k++] = m->tn_delta;
	oracle_destroy(m);
	return k;
}

/* ---- output stage: peak normalisation + 16-bit PCM ------------------------------------------------------------
 * VTM::Util::maximumAbsoluteValue / calculateOutputScale (vtm/VTMUtil.h:115-127, vtm/VTMUtil.cpp:20-21, 48-57),
 * Controller::writeOutputToFile (vtm_control_model/Controller.cpp:315-328: sample * scale in float),
 * WAVEFileWriter::writeSample (WAVEFileWriter.cpp:36-37, 122-126: round(sample * 32767.0f) as int, low 16 bits). */
float oracle_output_scale(const float* x, long n)
{
	float max_value = 0.0f;
	for (long i = 0; i < n; i++) {
		const float a = fabsf(x[i]);
		if (a > max_value) max_value = a;
	}
	if (max_value < 1.0e-30f) return 0.0f;
	return 0.95f / max_value;
}

float oracle_pcm16(const float* x, long n, short* out)
{
	const float scale = oracle_output_scale(x, n);
	const float sample_scale = 32767.0f;           /* sampleScale_(INT16_MAX) */
	for (long i = 0; i < n; i++) {
		const float sample = x[i] * scale;
		const int v = (int) roundf(sample * sample_scale);
		out[i] = (short) (v & 0xffff);
	}
	return scale;
}
