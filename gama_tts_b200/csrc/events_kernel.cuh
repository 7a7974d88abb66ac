// events_kernel -- control-frame generation on the device: from the event list of a chunk
// of an utterance to its control frames, the float32 [frame][16] array the tube kernels read, without the frames ever
// crossing PCIe.  Replaces EventList::generateOutput (vtm_control_model/EventList.cpp:929-1091) with the drift
// generator it steps (DriftGenerator.cpp:72-84, Butterworth2LowpassFilter.h:104-113).
//
// One warp per utterance: the chunks of an utterance share one drift generator (the reference keeps its state from one
// chunk to the next), so they are walked in order; utterances are independent and are dealt from a queue, longest first.
//
//   lane & 15        the parameter: every lane carries the value and the per-period delta of one regular and of one
//                    special parameter (EventList.cpp:942-953, 1021-1072).  The accumulation `value += delta` once a
//                    control period is the reference's own and is kept (a closed form would round differently).
//   lane >> 4        the frame of a pair: lanes 0..15 emit frame 2i of a segment, lanes 16..31 frame 2i + 1 from the
//                    value one delta on, so a pair of frames leaves the warp as one 128-byte row.
//   parameter 0      pitch: micro-intonation switch, drift, macro-intonation polynomial at the frame's time, mean pitch,
//                    added in float32 in the reference's order (:992-1006).  The drift recurrence and the polynomial are
//                    evaluated by all lanes (uniform values: same cost as one lane, no broadcast).
//
// Segment boundaries (:1026-1087): the next non-empty target of a parameter is found by a scan of its column of the
// event array, four rows in flight; a column is scanned at most once over the chunk.
//
// Arithmetic: IEEE double / float32 in the reference's order, no FMA contraction: bit-identical to the frames of the
// reference front end (tests/golden/events_v1.npz) and to the tests' plain-C restatement of generateOutput.
//
// Also compiled for the host by tests/simt_emu (GTTS_EMU): test infrastructure, never linked into the product.
#ifndef GTTS_EVENTS_KERNEL_CUH_
#define GTTS_EVENTS_KERNEL_CUH_

#include <cstdint>

#include "tube_kernel.cuh"
#include "events_types.h"

namespace gtts {
namespace evt {

struct DriftState { double seed, x1, x2, y1, y2; };

struct EventsParams {
	const double* events;           // gtts_event [n_events_total], viewed as rows of 37 doubles
	const gtts_event_config* cfgs;  // per chunk
	gtts_event_config* cfgs_out;    // per chunk: cfgs with the drift generator's state as the chunk left it (may be null)
	const ChunkDesc* chunks;
	const ChainDesc* chains;
	const int32_t* order;           // chains, longest first
	float* frames;
	int32_t* queue;                 // [0] next entry of order, [1] error flag (frame count mismatch)
	int32_t n_chains;
};

GTTS_DEV bool is_empty(double v) { return v > 1.7976931348623157e308; }      // Event::EMPTY_PARAMETER = +infinity

// One step of DriftGenerator::drift(); c = {deviation * 2, deviation, b0, b1, a1, a2}.
GTTS_DEV double drift_step(DriftState& s, double dev2, double off, double b0, double b1, double a1, double a2)
{
	const double temp = __dmul_rn(s.seed, 377.0);
	s.seed = __dsub_rn(temp, (double) (int) temp);
	const double x = __dsub_rn(__dmul_rn(s.seed, dev2), off);
	const double y = __dsub_rn(__dsub_rn(__dadd_rn(__dmul_rn(b0, __dadd_rn(x, s.x2)), __dmul_rn(b1, s.x1)), __dmul_rn(a1, s.y1)),
			__dmul_rn(a2, s.y2));
	s.x2 = s.x1; s.x1 = x; s.y2 = s.y1; s.y1 = y;
	return y;
}

// First row k in [from, n) whose entry in column `col` is not empty (n if none), with the entry in *value (the last
// entry read if none).  Four rows in flight; rows past the end are read as row n - 1 and ignored.
GTTS_DEV int scan_column(const double* col, int from, int n, double* value)
{
	int k = from;
	double v = *value;
	while (k < n) {
		double r[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) r[i] = col[(int64_t) (k + i < n ? k + i : n - 1) * kEventDoubles];
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (k + i < n) {
				v = r[i];
				if (!is_empty(v)) { *value = v; return k + i; }
			}
		}
		k += 4;
	}
	*value = v;
	return n;
}

GTTS_DEV int event_time(const double* events, int k)
{
	return reinterpret_cast<const int32_t*>(events + (int64_t) k * kEventDoubles)[0];
}

GTTS_DEV int event_has_interp(const double* events, int k)
{
	return reinterpret_cast<const int32_t*>(events + (int64_t) k * kEventDoubles)[1];
}

// One chunk on one warp.  Returns the number of frames written.
GTTS_DEV int chunk_frames(const gtts_event_config& c, const double* ev, int n, float* frames, DriftState& ds, int lane)
{
	if (n < 2) return 0;                                    // :931-933
	const int j = lane & 15, half = lane >> 4;
	const int period = c.control_period;
	const double dperiod = (double) period;
	const bool macro = c.macro_intonation != 0, smooth = c.smooth_intonation != 0;
	const double* colP = ev + 1 + j;                        // parameters[j]
	const double* colS = ev + 17 + j;                       // specialParameters[j]
	const double* poly = ev + 33;

	// ---- initial values and deltas (:942-953) ----
	double cur = colP[0], dlt = 0.0, scur = 0.0, sdlt = 0.0;
	{
		double value = 0.0;
		const int k = scan_column(colP, 1, n, &value);
		if (k < n) dlt = __dmul_rn(__dsub_rn(value, cur) / (double) event_time(ev, k), dperiod);
	}
	// ---- first segment of the macro-intonation curve (:957-978) ----
	double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
	if (macro) {
		int first = n;
		for (int base = 0; base < n; base += 32) {
			const unsigned b = __ballot_sync(0xffffffffu, base + lane < n && event_has_interp(ev, base + lane));
			if (b) { first = base + __ffs((int) b) - 1; break; }
		}
		if (first < n) {
			const double* q = poly + (int64_t) first * kEventDoubles;
			const double y1 = c.initial_pitch;
			const double x2 = (double) event_time(ev, first);
			if (smooth) {
				const double y2 = __dadd_rn(__dmul_rn(x2, __dadd_rn(__dmul_rn(x2, __dadd_rn(__dmul_rn(x2, q[0]), q[1])), q[2])), q[3]);
				pc = __dsub_rn(y2, y1) / x2;
				pd = y1;
			} else {
				const double y2 = __dadd_rn(__dmul_rn(x2, q[0]), q[1]);
				pa = __dsub_rn(y2, y1) / x2;
				pb = y1;
			}
		}
	}
	const bool micro = c.micro_intonation != 0, drift = c.intonation_drift != 0;
	const double dev2 = c.drift_deviation2, off = c.drift_offset, b0 = c.drift_b0, b1 = c.drift_b1, a1 = c.drift_a1, a2 = c.drift_a2;
	const float meanPitch = (float) c.mean_pitch;

	int target = 1, now = 0, nFrames = 0;
	double rowP = colP[kEventDoubles], rowS = colS[kEventDoubles];   // entries of row `target` (the "previous event" of the next boundary)
	int targetTime = event_time(ev, 1);
	for (;;) {
		// ---- the frames of this segment: one while now < targetTime, at least one (:985-1024) ----
		const int k = targetTime > now ? (targetTime - now + period - 1) / period : 1;
		for (int i = 0; i < k; i += 2) {
			const bool two = i + 1 < k;
			const double c1 = dlt != 0.0 ? __dadd_rn(cur, dlt) : cur;
			const double s1 = sdlt != 0.0 ? __dadd_rn(scur, sdlt) : scur;
			float p = (float) __dadd_rn(half ? c1 : cur, half ? s1 : scur);
			double d0 = 0.0, d1 = 0.0;
			if (drift) {
				d0 = drift_step(ds, dev2, off, b0, b1, a1, a2);
				if (two) d1 = drift_step(ds, dev2, off, b0, b1, a1, a2);
			}
			if (j == 0) {
				if (!micro) p = 0.0f;
				if (drift) p = __fadd_rn(p, (float) (half ? d1 : d0));
				if (macro) {
					const double x = (double) (now + half * period);
					const double intonation = smooth
						? __dadd_rn(__dmul_rn(x, __dadd_rn(__dmul_rn(x, __dadd_rn(__dmul_rn(x, pa), pb)), pc)), pd)
						: __dadd_rn(__dmul_rn(x, pa), pb);
					p = __fadd_rn(p, (float) intonation);
				}
				p = __fadd_rn(p, meanPitch);
			}
			if (!half || two) frames[(int64_t) (nFrames + half) * 16 + j] = p;
			cur = two && dlt != 0.0 ? __dadd_rn(c1, dlt) : c1;
			scur = two && sdlt != 0.0 ? __dadd_rn(s1, sdlt) : s1;
			now += two ? 2 * period : period;
			nFrames += two ? 2 : 1;
		}
		// ---- segment boundary (:1026-1087) ----
		if (++target == n) break;
		targetTime = event_time(ev, target);
		const double prevP = rowP, prevS = rowS;
		rowP = colP[(int64_t) target * kEventDoubles];
		rowS = colS[(int64_t) target * kEventDoubles];
		if (!is_empty(prevP)) {
			double value = rowP;
			const int kk = is_empty(value) ? scan_column(colP, target + 1, n, &value) : target;
			dlt = !is_empty(value) ? __dmul_rn(__dsub_rn(value, cur) / (double) (event_time(ev, kk) - now), dperiod) : 0.0;
		}
		if (!is_empty(prevS)) {
			double value = rowS;
			const int kk = is_empty(value) ? scan_column(colS, target + 1, n, &value) : target;
			sdlt = !is_empty(value) ? __dmul_rn(__dsub_rn(value, scur) / (double) (event_time(ev, kk) - now), dperiod) : 0.0;
		}
		if (macro && event_has_interp(ev, target - 1)) {
			const double* q = poly + (int64_t) (target - 1) * kEventDoubles;
			pa = q[0];
			pb = q[1];
			if (smooth) { pc = q[2]; pd = q[3]; }
		}
	}
	return nFrames;
}

// The body of a CTA: its warps take utterances from the queue until it is empty.
GTTS_DEV void events_cta_body(const EventsParams& P, int tid)
{
	const int lane = tid & 31;
	for (;;) {
		int slot = 0;
		if (lane == 0) slot = atomicAdd(P.queue, 1);
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (slot >= P.n_chains) return;
		const ChainDesc chain = P.chains[P.order[slot]];
		DriftState ds = {0.0, 0.0, 0.0, 0.0, 0.0};
		for (int ci = chain.first; ci < chain.first + chain.count; ++ci) {
			const gtts_event_config c = P.cfgs[ci];
			const ChunkDesc d = P.chunks[ci];
			if (ci == chain.first) { ds.seed = c.drift_seed; ds.x1 = c.drift_x1; ds.x2 = c.drift_x2; ds.y1 = c.drift_y1; ds.y2 = c.drift_y2; }
			const int made = chunk_frames(c, P.events + d.event_offset * kEventDoubles, d.n_events, P.frames + d.frame_offset * 16, ds, lane);
			if (lane == 0) {
				if (made != d.n_frames) P.queue[1] = 1;
				if (P.cfgs_out) {
					gtts_event_config o = c;
					o.drift_seed = ds.seed; o.drift_x1 = ds.x1; o.drift_x2 = ds.x2; o.drift_y1 = ds.y1; o.drift_y2 = ds.y2;
					P.cfgs_out[ci] = o;
				}
			}
		}
	}
}

#ifndef GTTS_EMU
constexpr int kEventsWarps = 8;

__global__ void __launch_bounds__(kEventsWarps * 32) events_kernel(const EventsParams P)
{
	events_cta_body(P, (int) threadIdx.x);
}
#endif

} // namespace evt
} // namespace gtts

#endif
