/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the control-frame generation (see events_oracle.c). */
#ifndef EVENTS_ORACLE_H_
#define EVENTS_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* One event of the reference's event list (vtm_control_model/EventList.h:125-161); EMPTY = +infinity. */
typedef struct oracle_event {
	int    time;                 /* ms */
	int    has_interp;           /* macro-intonation polynomial present */
	double param[16];
	double special[16];
	double a, b, c, d;
} oracle_event;

/* What EventList::generateOutput reads besides the events, and the drift generator it steps. */
typedef struct oracle_event_config {
	int    control_period;       /* ms */
	int    macro_intonation, micro_intonation, intonation_drift, smooth_intonation;
	double initial_pitch, mean_pitch;
	double drift_deviation2, drift_offset;        /* DriftGenerator: pitchDeviation_ (= 2 deviation), pitchOffset_ */
	double drift_seed;
	double drift_b0, drift_b1, drift_a1, drift_a2;   /* its Butterworth-2 low-pass */
	double drift_x1, drift_x2, drift_y1, drift_y2;
} oracle_event_config;

/* Returns the number of control frames (16 float32 each); writes at most cap of them.  The drift generator's state in
 * *cfg is advanced (the reference keeps it across the chunks of an utterance). */
long oracle_events_generate(oracle_event_config* cfg, const oracle_event* events, int n_events, float* frames, long cap);

#ifdef __cplusplus
}
#endif
#endif
