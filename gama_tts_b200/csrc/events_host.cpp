// Control-frame generation: host-side planning (frame layout from the event times, utterance chains, work order).
#include "events_types.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

namespace gtts {
namespace evt {

int64_t countFrames(int32_t control_period, const gtts_event* events, int64_t n_events)
{
	if (n_events < 2) return 0;
	int64_t now = 0, frames = 0;
	for (int64_t target = 1; target < n_events; ++target) {
		const int64_t t = events[target].time;
		const int64_t k = t > now ? (t - now + control_period - 1) / control_period : 1;
		frames += k;
		now += k * control_period;
	}
	return frames;
}

bool driftSetup(double deviation, double sampleRate, double lowpassCutoff, gtts_event_config& cfg)
{
	if (!(lowpassCutoff >= 1.0) || !(lowpassCutoff <= sampleRate * 0.48)) return false;
	cfg.drift_deviation2 = deviation * 2.0;
	cfg.drift_offset = deviation;
	cfg.drift_seed = 0.7892347;
	const double wcT = 2.0 * std::tan(M_PI * lowpassCutoff / sampleRate);
	const double wc2T2 = wcT * wcT;
	const double c1 = 2.0 * std::sqrt(2.0) * wcT;
	const double c2 = 1.0 / (wc2T2 + c1 + 4.0);
	cfg.drift_b0 = c2 * wc2T2;
	cfg.drift_b1 = 2.0 * cfg.drift_b0;
	cfg.drift_a1 = c2 * (2.0 * wc2T2 - 8.0);
	cfg.drift_a2 = c2 * (wc2T2 - c1 + 4.0);
	cfg.drift_x1 = cfg.drift_x2 = cfg.drift_y1 = cfg.drift_y2 = 0.0;
	return true;
}

std::string planEvents(const gtts_event_config* configs, const int32_t* continues_previous, const gtts_event* events,
		const int64_t* event_offsets, int64_t n_chunks, EventsPlan& plan, int* err)
{
	*err = GTTS_ERR_INVALID;
	if (n_chunks < 0 || (n_chunks > 0 && (!configs || !event_offsets))) return "null configs / event_offsets";
	if (n_chunks > std::numeric_limits<int32_t>::max()) return "too many chunks";
	if (n_chunks > 0 && event_offsets[0] != 0) return "event_offsets[0] must be 0";
	plan.cfgs.assign(configs, configs + n_chunks);
	plan.chunks.resize(n_chunks);
	plan.frame_offsets.assign(n_chunks + 1, 0);
	plan.chains.clear();
	std::vector<int64_t> chainFrames;
	for (int64_t c = 0; c < n_chunks; ++c) {
		const int64_t n = event_offsets[c + 1] - event_offsets[c];
		if (n < 0 || n > std::numeric_limits<int32_t>::max()) return "event_offsets must not decrease";
		if (n > 0 && !events) return "null events";
		if (configs[c].control_period <= 0) return "control_period must be positive";
		for (int64_t k = 0; k < n; ++k) {
			const int64_t t = events[event_offsets[c] + k].time;
			if (t < 0 || t > (int64_t(1) << 30)) return "event time out of range";
		}
		const int64_t frames = countFrames(configs[c].control_period, events + event_offsets[c], n);
		// the kernel keeps the chunk's time in milliseconds and its frame count in 32-bit integers
		if (frames > (int64_t(1) << 30) || frames * configs[c].control_period > (int64_t(1) << 30)) return "chunk too long";
		plan.chunks[c].event_offset = event_offsets[c];
		plan.chunks[c].frame_offset = plan.frame_offsets[c];
		plan.chunks[c].n_events = static_cast<int32_t>(n);
		plan.chunks[c].n_frames = static_cast<int32_t>(frames);
		plan.frame_offsets[c + 1] = plan.frame_offsets[c] + frames;
		if (c > 0 && continues_previous && continues_previous[c]) {
			plan.chains.back().count++;
			chainFrames.back() += frames;
		} else {
			plan.chains.push_back(ChainDesc{static_cast<int32_t>(c), 1});
			chainFrames.push_back(frames);
		}
	}
	plan.n_events_total = n_chunks > 0 ? event_offsets[n_chunks] : 0;
	plan.order.resize(plan.chains.size());
	std::iota(plan.order.begin(), plan.order.end(), 0);
	std::stable_sort(plan.order.begin(), plan.order.end(), [&](int32_t a, int32_t b) { return chainFrames[a] > chainFrames[b]; });
	plan.chunk_order.resize(n_chunks);
	std::iota(plan.chunk_order.begin(), plan.chunk_order.end(), 0);
	std::stable_sort(plan.chunk_order.begin(), plan.chunk_order.end(),
			[&](int32_t a, int32_t b) { return plan.chunks[a].n_frames > plan.chunks[b].n_frames; });
	*err = GTTS_OK;
	return std::string();
}

} // namespace evt
} // namespace gtts
