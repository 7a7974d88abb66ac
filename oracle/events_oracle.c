/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's control-frame generation, EventList::generateOutput
 * (vtm_control_model/EventList.cpp:929-1091) with DriftGenerator::drift (vtm_control_model/DriftGenerator.cpp:72-84)
 * and its Butterworth2LowPassFilter (vtm/Butterworth2LowpassFilter.h:104-113): from an utterance's event list to its
 * control frames.  Plain-C restatement.
 *
 * Parity status: PINNED against the reference itself: tests/golden/events_v1.npz holds event lists and the frames the
 * unmodified reference produced from them (oracle/ref_events.cpp around the reference's own front end,
 * tools/make_golden_events.py); tests/test_oracle.py::test_events_oracle_bit_exact demands bit equality.
 */
#include <math.h>
#include <string.h>

#include "events_oracle.h"

static int is_empty(double v) { return isinf(v) && v > 0; }

static double drift(oracle_event_config* c)
{
	const double temp = c->drift_seed * 377.0;
	c->drift_seed = temp - (int) temp;
	const double x = (c->drift_seed * c->drift_deviation2) - c->drift_offset;
	const double y = c->drift_b0 * (x + c->drift_x2) + c->drift_b1 * c->drift_x1 - c->drift_a1 * c->drift_y1 - c->drift_a2 * c->drift_y2;
	c->drift_x2 = c->drift_x1; c->drift_x1 = x; c->drift_y2 = c->drift_y1; c->drift_y1 = y;
	return y;
}

long oracle_events_generate(oracle_event_config* c, const oracle_event* ev, int n, float* frames, long cap)
{
	if (n < 2) return 0;
	double cur[16], dlt[16], scur[16], sdlt[16];
	memset(cur, 0, sizeof cur); memset(dlt, 0, sizeof dlt); memset(scur, 0, sizeof scur); memset(sdlt, 0, sizeof sdlt);
	const int period = c->control_period;
	for (int i = 0; i < 16; ++i) {                      /* :942-953 */
		cur[i] = ev[0].param[i];
		int j = 1;
		double value;
		while (is_empty(value = ev[j].param[i])) { if (++j >= n) break; }
		if (j < n) dlt[i] = ((value - cur[i]) / ev[j].time) * period;
	}
	double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
	if (c->macro_intonation) {                          /* :957-978 */
		int j = 0;
		for (; j < n; ++j) if (ev[j].has_interp) break;
		if (j < n) {
			const double y1 = c->initial_pitch;
			const double x2 = ev[j].time;
			if (c->smooth_intonation) {
				const double y2 = x2 * (x2 * (x2 * ev[j].a + ev[j].b) + ev[j].c) + ev[j].d;
				pc = (y2 - y1) / x2;
				pd = y1;
			} else {
				const double y2 = x2 * ev[j].a + ev[j].b;
				pa = (y2 - y1) / x2;
				pb = y1;
			}
		}
	}
	int target = 1;
	int target_time = ev[target].time;
	int now = 0;
	long n_frames = 0;
	while (target < n) {
		float p[16];
		for (int j = 0; j < 16; ++j) p[j] = (float) (cur[j] + scur[j]);
		if (!c->micro_intonation) p[0] = 0.0f;
		if (c->intonation_drift) p[0] += (float) drift(c);
		if (c->macro_intonation) {
			const double x = now;
			const double intonation = c->smooth_intonation ? x * (x * (x * pa + pb) + pc) + pd : x * pa + pb;
			p[0] += (float) intonation;
		}
		p[0] += (float) c->mean_pitch;
		if (n_frames < cap) memcpy(frames + 16 * n_frames, p, sizeof p);
		n_frames++;
		for (int j = 0; j < 16; ++j) if (dlt[j]) cur[j] += dlt[j];
		for (int j = 0; j < 16; ++j) if (sdlt[j]) scur[j] += sdlt[j];
		now += period;
		if (now >= target_time) {
			if (++target == n) break;
			target_time = ev[target].time;
			for (int j = 0; j < 16; ++j) {                  /* :1036-1054 */
				if (!is_empty(ev[target - 1].param[j])) {
					int k = target;
					double value;
					while (is_empty(value = ev[k].param[j])) { if (++k >= n) break; }
					if (!is_empty(value)) dlt[j] = ((value - cur[j]) / (ev[k].time - now)) * period;
					else dlt[j] = 0.0;
				}
			}
			for (int j = 0; j < 16; ++j) {                  /* :1056-1072 */
				if (!is_empty(ev[target - 1].special[j])) {
					int k = target;
					double value;
					while (is_empty(value = ev[k].special[j])) { if (++k >= n) break; }
					if (!is_empty(value)) sdlt[j] = ((value - scur[j]) / (ev[k].time - now)) * period;
					else sdlt[j] = 0.0;
				}
			}
			if (c->macro_intonation && ev[target - 1].has_interp) {   /* :1074-1086 */
				pa = ev[target - 1].a;
				pb = ev[target - 1].b;
				if (c->smooth_intonation) { pc = ev[target - 1].c; pd = ev[target - 1].d; }
			}
		}
	}
	return n_frames;
}
