// events_kernel -- control-frame generation on the device: from the event list of a chunk
// of an utterance to its control frames, the float32 [frame][16] array the tube kernels read, without the frames ever
// crossing PCIe.  Replaces EventList::generateOutput (vtm_control_model/EventList.cpp:929-1091) with the drift
// generator it steps (DriftGenerator.cpp:72-84, Butterworth2LowpassFilter.h:104-113).
//
// Two passes.  The drift generator (one per utterance: the reference keeps its state from one chunk to the next) is a
// scalar sequential recurrence, the same for all parameters: one THREAD per utterance writes its output per frame into a
// scratch array (events_drift_kernel).  Then one WARP per chunk, dealt from a queue, longest first (events_kernel):
//
//   lane & 15        the parameter: every lane carries the value and the per-period delta of one regular and of one
//                    special parameter (EventList.cpp:942-953, 1021-1072).  The accumulation `value += delta` once a
//                    control period is the reference's own and is kept (a closed form would round differently).
//   lane >> 4        the frame of a pair: lanes 0..15 emit frame 2i of a segment, lanes 16..31 frame 2i + 1 from the
//                    value one delta on, so a pair of frames leaves the warp as one 128-byte row.
//   parameter 0      pitch: micro-intonation switch, drift, macro-intonation polynomial at the frame's time, mean pitch,
//                    added in float32 in the reference's order (:992-1006).
//
// Segment boundaries (:1026-1087): the next non-empty target of a parameter is found by a scan of its column of the
// event array -- in the ring of rows each warp stages ahead of itself in shared memory (below), beyond it in global
// memory; a column is scanned at most once over the chunk.
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
	const int32_t* order;           // chains, longest first (drift pass: one thread each)
	const int32_t* chunk_order;     // chunks, longest first (frame pass: one warp each, from a queue)
	float* frames;
	float* drift;                   // scratch [n_frames_total]: the drift generator's output per frame, already float32
	int32_t* queue;                 // [0] next entry of chunk_order, [1] error flag (frame count mismatch)
	int32_t n_chains;
	int32_t n_chunks;
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

// ---- the ring: the next rows of the event array, staged in shared memory by asynchronous copies -------------------------
// Every boundary needs the row of the new target, the row before it and, for the parameters whose next target lies
// further on, the rows after it: read from global memory these are two to three dependent round trips per segment
// (measured: 880 cycles per frame, all of it waiting).  Each warp therefore keeps rows [target - 1, target + 14] of its
// chunk in a ring of 16 slots; a boundary retires one row and starts the copy of the next (cp.async, 8 bytes per lane),
// twelve or more rows ahead of its first use.
enum { kRingRows = 16, kRingPending = 4, kMaskRows = 512, kMaskWords = kMaskRows / 32 };

#ifndef GTTS_EMU
GTTS_DEV void ring_copy8(double* dst, const double* src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
GTTS_DEV void ring_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> GTTS_DEV void ring_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#else
GTTS_DEV void ring_copy8(double* dst, const double* src) { *dst = *src; }
GTTS_DEV void ring_commit() {}
template<int N> GTTS_DEV void ring_wait() {}
#endif

struct Ring {
	double* rows;                   // [kRingRows][kEventDoubles], this warp's
	const double* ev;               // the chunk's rows in global memory
	int n;                          // rows of the chunk
	int issued;                     // rows [0, issued) have been requested; rows [issued - 16, issued) are in the ring
	int safe;                       // rows [issued - 16, safe) have arrived and are visible to every lane
	unsigned* mask;                 // [kMaskWords][32]: bit r of mask[w][c] = entry c of row maskBase + 32 w + r is not empty
	int maskBase, maskEnd;          // the rows the masks cover
};

// ---- the occupancy masks ---------------------------------------------------------------------------------------------
// Most columns are sparse (special parameters: one entry in a hundred and more), and "the next row with an entry in my
// column" read row by row from global memory is a chain of round trips that dominated the first version (95 such scans
// of about 50 rows per 10 s utterance: two thirds of its time).  One coalesced pass over the rows of a window of 512
// leaves every column's occupancy as 16 words in shared memory; a scan is then a find-first-set over at most 16 words.
GTTS_DEV_NOINLINE int build_masks(const double* ev, int n, unsigned* mask, int base, int lane)
{
	// (out of line, with scalar arguments: the caller's Ring stays in registers and the eight loads in flight here do not
	// count against the frame loop's register budget.)  Returns the end of the window.
	const int end = base + kMaskRows < n ? base + kMaskRows : n;
	__syncwarp();                                           // the readers of the previous window are done
	const double* col = ev + 1 + lane;
	for (int w = 0; base + 32 * w < end; ++w) {
		unsigned word = 0;
		const int k0 = base + 32 * w;
#pragma unroll 8
		for (int r = 0; r < 32; ++r) {
			const int k = k0 + r < end ? k0 + r : end - 1;
			const double v = col[(int64_t) k * kEventDoubles];
			if (k0 + r < end && !is_empty(v)) word |= 1u << r;
		}
		mask[w * 32 + lane] = word;
	}
	__syncwarp();
	return end;
}

GTTS_DEV void ring_issue(Ring& R, int lane)
{
	if (R.issued < R.n) {
		double* dst = R.rows + (R.issued & (kRingRows - 1)) * kEventDoubles;
		const double* src = R.ev + (int64_t) R.issued * kEventDoubles;
		ring_copy8(dst + lane, src + lane);
		if (lane < kEventDoubles - 32) ring_copy8(dst + 32 + lane, src + 32 + lane);
	}
	ring_commit();                  // one group per call, empty past the end: the count of pending groups stays uniform
	R.issued++;
}

GTTS_DEV const double* ring_row(const Ring& R, int k) { return R.rows + (k & (kRingRows - 1)) * kEventDoubles; }
GTTS_DEV int row_time(const double* row) { return reinterpret_cast<const int32_t*>(row)[0]; }
GTTS_DEV int row_has_interp(const double* row) { return reinterpret_cast<const int32_t*>(row)[1]; }

// First row k in [from, n) whose entry in column `col` (1 + j: parameter j, 17 + j: special parameter j) is not empty, with
// the entry in *value and the row's time in *time; n if there is none (*value then is EMPTY).  Within the mask window a
// find-first-set; beyond it row by row from global memory, four rows in flight.
GTTS_DEV int scan_column(const Ring& R, int col, int from, double* value, int* time)
{
	const int n = R.n;
	int k = from;
	bool found = false;
	if (from >= R.maskBase && from < R.maskEnd) {
		const unsigned* m = R.mask + (col - 1);
		const int rel = from - R.maskBase, words = (R.maskEnd - R.maskBase + 31) >> 5;
		int w = rel >> 5;
		unsigned word = m[w * 32] & (0xffffffffu << (rel & 31));
		while (!word && ++w < words) word = m[w * 32];
		if (word) {
			k = R.maskBase + 32 * w + __ffs((int) word) - 1;
			found = true;
		} else {
			k = R.maskEnd;
		}
	}
	if (!found) {
		const double* g = R.ev + col;
		while (k < n && !found) {
			double r[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) r[i] = g[(int64_t) (k + i < n ? k + i : n - 1) * kEventDoubles];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				if (!found && k + i < n && !is_empty(r[i])) { k += i; found = true; }
			}
			if (!found) k += 4;
		}
	}
	if (!found) {
		*value = 1.0 / 0.0;
		return n;
	}
	const double* row = (k < R.safe && k >= R.issued - kRingRows) ? ring_row(R, k) : R.ev + (int64_t) k * kEventDoubles;
	*value = row[col];
	*time = row_time(row);
	return k;
}

// One chunk on one warp.  Returns the number of frames written.
GTTS_DEV int chunk_frames(const gtts_event_config& c, const double* ev, int n, float* frames, const float* driftRow, double* ringRows, unsigned* maskWords, float* pitch, int lane)
{
	if (n < 2) return 0;                                    // :931-933
	const int j = lane & 15, half = lane >> 4;
	const int period = c.control_period;
	const double dperiod = (double) period;
	const bool macro = c.macro_intonation != 0, smooth = c.smooth_intonation != 0;
	const int colP = 1 + j, colS = 17 + j;                 // parameters[j], specialParameters[j] within a row

	Ring R = {ringRows, ev, n, 0, 0, maskWords, 0, 0};
	__syncwarp();                                           // the previous chunk's readers are done with the ring
	for (int r = 0; r < kRingRows; ++r) ring_issue(R, lane);
	R.maskBase = 0;
	R.maskEnd = build_masks(ev, n, maskWords, 0, lane);
	ring_wait<0>();
	__syncwarp();
	R.safe = kRingRows;

	// ---- initial values and deltas (:942-953) ----
	double cur = ring_row(R, 0)[colP], dlt = 0.0, scur = 0.0, sdlt = 0.0;
	{
		double value = 0.0;
		int time = 0;
		const int k = scan_column(R, colP, 1, &value, &time);
		if (k < n) dlt = __dmul_rn(__dsub_rn(value, cur) / (double) time, dperiod);
	}
	// ---- first segment of the macro-intonation curve (:957-978) ----
	double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
	if (macro) {
		int first = n;
		for (int base = 0; base < n; base += 32) {
			const unsigned b = __ballot_sync(0xffffffffu, base + lane < n && row_has_interp(ev + (int64_t) (base + lane) * kEventDoubles));
			if (b) { first = base + __ffs((int) b) - 1; break; }
		}
		if (first < n) {
			const double* q = ev + (int64_t) first * kEventDoubles;
			const double y1 = c.initial_pitch;
			const double x2 = (double) row_time(q);
			if (smooth) {
				const double y2 = __dadd_rn(__dmul_rn(x2, __dadd_rn(__dmul_rn(x2, __dadd_rn(__dmul_rn(x2, q[33]), q[34])), q[35])), q[36]);
				pc = __dsub_rn(y2, y1) / x2;
				pd = y1;
			} else {
				const double y2 = __dadd_rn(__dmul_rn(x2, q[33]), q[34]);
				pa = __dsub_rn(y2, y1) / x2;
				pb = y1;
			}
		}
	}
	const bool micro = c.micro_intonation != 0, drift = c.intonation_drift != 0;
	const float meanPitch = (float) c.mean_pitch;
	const int periodShift = (period & (period - 1)) == 0 ? 31 - __clz(period) : -1;      // control periods of 1, 2, 4, 8 ms: a shift

	int target = 1, now = 0, nFrames = 0;
	double rowP = ring_row(R, 1)[colP], rowS = ring_row(R, 1)[colS];   // entries of row `target` (the "previous event" of the next boundary)
	int targetTime = row_time(ring_row(R, 1));
	for (;;) {
		// ---- the frames of this segment: one while now < targetTime, at least one (:985-1024) ----
		// In blocks of up to 32 frames: first the pitch terms of the block's frames, one frame per lane (the drift
		// value of the scratch array, the polynomial at the frame's time: neither depends on the accumulated values),
		// into the warp's 64 floats of shared memory; then the frames pair by pair.  `value += delta` is applied
		// unconditionally: the reference skips a zero delta, which differs only in the sign of a zero value, and
		// that sign never reaches a frame (the special value it is added to is never -0).
		const int span = targetTime - now;
		int k = span > 0 ? (periodShift >= 0 ? (span + period - 1) >> periodShift : (span + period - 1) / period) : 1;
		while (k > 0) {
			const int blk = k < 32 ? k : 32;
			if (drift || macro) {
				float addD = 0.0f, addI = 0.0f;
				if (lane < blk) {
					if (drift) addD = driftRow[nFrames + lane];
					if (macro) {
						const double x = (double) (now + lane * period);
						addI = (float) (smooth
							? __dadd_rn(__dmul_rn(x, __dadd_rn(__dmul_rn(x, __dadd_rn(__dmul_rn(x, pa), pb)), pc)), pd)
							: __dadd_rn(__dmul_rn(x, pa), pb));
					}
				}
				__syncwarp();                               // the previous block's readers are done
				pitch[lane] = addD;
				pitch[32 + lane] = addI;
				__syncwarp();
			}
			float* o = frames + (int64_t) (nFrames + half) * 16 + j;
			const float* pt = pitch + half;
			// lanes 16..31 emit the value one delta on: cur + (half ? dlt : 0) is the same addition as the c1 of the next
			// line for them and leaves the value as it is for lanes 0..15 (a -0 becomes +0: see above), without selects.
			// The pitch terms are evaluated by every lane (uniform conditions, broadcast reads) and kept by parameter 0: no
			// divergent branch in the loop.
			const double stepP = half ? dlt : 0.0, stepS = half ? sdlt : 0.0;
#pragma unroll 2
			for (int i = blk >> 1; i > 0; --i) {
				float p = (float) __dadd_rn(__dadd_rn(cur, stepP), __dadd_rn(scur, stepS));
				float q = micro ? p : 0.0f;
				if (drift) q = __fadd_rn(q, pt[0]);
				if (macro) q = __fadd_rn(q, pt[32]);
				q = __fadd_rn(q, meanPitch);
				*o = j == 0 ? q : p;
				o += 32;
				pt += 2;
				cur = __dadd_rn(__dadd_rn(cur, dlt), dlt);
				scur = __dadd_rn(__dadd_rn(scur, sdlt), sdlt);
			}
			if (blk & 1) {                                   // the odd frame: lanes 0..15
				float p = (float) __dadd_rn(cur, scur);
				if (j == 0) {
					if (!micro) p = 0.0f;
					if (drift) p = __fadd_rn(p, pitch[blk - 1]);
					if (macro) p = __fadd_rn(p, pitch[32 + blk - 1]);
					p = __fadd_rn(p, meanPitch);
				}
				if (!half) *o = p;
				cur = __dadd_rn(cur, dlt);
				scur = __dadd_rn(scur, sdlt);
			}
			now += blk * period;
			nFrames += blk;
			k -= blk;
		}
		// ---- segment boundary (:1026-1087) ----
		if (++target == n) break;
		// row target - 2 is dead: its slot takes row target + 14; all but the newest copies have landed
		__syncwarp();
		ring_issue(R, lane);
		ring_wait<kRingPending>();
		__syncwarp();
		R.safe = R.issued - kRingPending;
		if (R.maskEnd < n && target + kMaskRows / 4 > R.maskEnd) {
			R.maskBase = target;
			R.maskEnd = build_masks(ev, n, maskWords, target, lane);
		}
		const double* rowT = ring_row(R, target);
		targetTime = row_time(rowT);
		const double prevP = rowP, prevS = rowS;
		rowP = rowT[colP];
		rowS = rowT[colS];
		if (!is_empty(prevP)) {
			double value = rowP;
			int time = targetTime;
			if (is_empty(value)) scan_column(R, colP, target + 1, &value, &time);
			dlt = !is_empty(value) ? __dmul_rn(__dsub_rn(value, cur) / (double) (time - now), dperiod) : 0.0;
		}
		if (!is_empty(prevS)) {
			double value = rowS;
			int time = targetTime;
			if (is_empty(value)) scan_column(R, colS, target + 1, &value, &time);
			sdlt = !is_empty(value) ? __dmul_rn(__dsub_rn(value, scur) / (double) (time - now), dperiod) : 0.0;
		}
		const double* rowB = ring_row(R, target - 1);
		if (macro && row_has_interp(rowB)) {
			pa = rowB[33];
			pb = rowB[34];
			if (smooth) { pc = rowB[35]; pd = rowB[36]; }
		}
	}
	ring_wait<0>();                 // nothing of this chunk may land in the ring once the next chunk fills it
	return nFrames;
}

// Drift pass: thread t walks the drift generator of utterance order[t] through the frames of its chunks -- a chaotic seed
// map and a recursive low-pass, sequential by nature and the same for every parameter, so it costs a thread, not a warp,
// and 12 FP64 instructions per frame of one lane instead of 32 -- and leaves (float) drift per frame in the scratch array
// (what generateOutput adds to the pitch, :995) and the state after every chunk in cfgs_out.
GTTS_DEV void drift_body(const EventsParams& P, int t)
{
	if (t >= P.n_chains) return;
	const ChainDesc chain = P.chains[P.order[t]];
	DriftState ds = {0.0, 0.0, 0.0, 0.0, 0.0};
	for (int ci = chain.first; ci < chain.first + chain.count; ++ci) {
		const gtts_event_config c = P.cfgs[ci];
		const ChunkDesc d = P.chunks[ci];
		if (ci == chain.first) { ds.seed = c.drift_seed; ds.x1 = c.drift_x1; ds.x2 = c.drift_x2; ds.y1 = c.drift_y1; ds.y2 = c.drift_y2; }
		if (c.intonation_drift) {
			float* out = P.drift + d.frame_offset;
			for (int f = 0; f < d.n_frames; ++f)
				out[f] = (float) drift_step(ds, c.drift_deviation2, c.drift_offset, c.drift_b0, c.drift_b1, c.drift_a1, c.drift_a2);
		}
		if (P.cfgs_out) {
			gtts_event_config o = c;
			o.drift_seed = ds.seed; o.drift_x1 = ds.x1; o.drift_x2 = ds.x2; o.drift_y1 = ds.y1; o.drift_y2 = ds.y2;
			P.cfgs_out[ci] = o;
		}
	}
}

// Frame pass: the warps of a CTA take chunks from the queue until it is empty.
GTTS_DEV void events_cta_body(const EventsParams& P, double* ringBase, unsigned* maskBase, float* pitchBase, int tid)
{
	const int lane = tid & 31;
	double* ringRows = ringBase + (tid >> 5) * (kRingRows * kEventDoubles);
	unsigned* maskWords = maskBase + (tid >> 5) * (kMaskWords * 32);
	float* pitch = pitchBase + (tid >> 5) * 64;
	for (;;) {
		int slot = 0;
		if (lane == 0) slot = atomicAdd(P.queue, 1);
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (slot >= P.n_chunks) return;
		const int ci = P.chunk_order[slot];
		const gtts_event_config c = P.cfgs[ci];
		const ChunkDesc d = P.chunks[ci];
		const int made = chunk_frames(c, P.events + d.event_offset * kEventDoubles, d.n_events, P.frames + d.frame_offset * 16,
				P.drift + d.frame_offset, ringRows, maskWords, pitch, lane);
		if (lane == 0 && made != d.n_frames) P.queue[1] = 1;
	}
}

#ifndef GTTS_EMU
constexpr int kEventsWarps = 8;
constexpr int kEventsSmem = kEventsWarps * (kRingRows * kEventDoubles * 8 + kMaskWords * 32 * 4 + 64 * 4);   // 56,320 bytes

// Two builds of the frame pass: 3 resident CTAs per SM at 80 registers -- the faster warp, for batches that do not fill
// the GPU (1,024 chunks: 0.72 ms against 0.78) -- and 4 at 64 registers -- more warps to cover each other's waits, for
// batches that do (37,888 chunks: 1.58 ms against 1.72; 2 CTAs per SM: 2.20 ms).
template<int CTAS>
__global__ void __launch_bounds__(kEventsWarps * 32, CTAS) events_kernel(const EventsParams P)
{
	extern __shared__ double smem_events[];
	double* ring = smem_events;
	unsigned* masks = reinterpret_cast<unsigned*>(ring + kEventsWarps * kRingRows * kEventDoubles);
	float* pitch = reinterpret_cast<float*>(masks + kEventsWarps * kMaskWords * 32);
	events_cta_body(P, ring, masks, pitch, (int) threadIdx.x);
}

constexpr int kDriftThreads = 32;

__global__ void __launch_bounds__(kDriftThreads) events_drift_kernel(const EventsParams P)
{
	drift_body(P, (int) (blockIdx.x * kDriftThreads + threadIdx.x));
}
#endif

} // namespace evt
} // namespace gtts

#endif
