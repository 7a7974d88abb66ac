// Control-frame generation (events_kernel.cuh): the descriptors the host plans and the kernel reads.
#ifndef GTTS_EVENTS_TYPES_H_
#define GTTS_EVENTS_TYPES_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gtts_b200.h"

namespace gtts {
namespace evt {

enum { kEventDoubles = 37 };        // sizeof(gtts_event) / 8: time + flag, 16 + 16 values, a b c d

struct ChunkDesc {
	int64_t event_offset;           // first event of the chunk in the packed event array
	int64_t frame_offset;           // first frame of the chunk in the packed frame array
	int32_t n_events;
	int32_t n_frames;               // what the host's count says the chunk produces (checked)
};

struct ChainDesc {                  // the chunks of one utterance: [first, first + count)
	int32_t first;
	int32_t count;
};

static_assert(sizeof(gtts_event) == 8 * kEventDoubles, "gtts_event is read as rows of 37 doubles");
static_assert(sizeof(gtts_event_config) == 128, "gtts_event_config layout");

struct EventsPlan {
	std::vector<gtts_event_config> cfgs;
	std::vector<ChunkDesc> chunks;
	std::vector<ChainDesc> chains;
	std::vector<int32_t> order;             // chains by frames, longest first
	std::vector<int32_t> chunk_order;       // chunks by frames, longest first
	std::vector<int64_t> frame_offsets;     // [n_chunks + 1]
	int64_t n_events_total = 0;
};

// Frames EventList::generateOutput makes of a list with these times (EventList.cpp:931-933, 985-1030): every target from the
// second event on takes the control periods up to its time, at least one.
int64_t countFrames(int32_t control_period, const gtts_event* events, int64_t n_events);

// A fresh drift generator (DriftGenerator.cpp:23-34, 55-63; Butterworth2LowpassFilter.h:80-100); false: cutoff out of range.
bool driftSetup(double deviation, double sampleRate, double lowpassCutoff, gtts_event_config& cfg);

// Returns an empty string or the reason the batch is refused (*err = GTTS_ERR_*).
std::string planEvents(const gtts_event_config* configs, const int32_t* continues_previous, const gtts_event* events,
		const int64_t* event_offsets, int64_t n_chunks, EventsPlan& plan, int* err);

} // namespace evt
} // namespace gtts

#endif
