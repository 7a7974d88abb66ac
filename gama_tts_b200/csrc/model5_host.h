// Host side of the model-5 path: per-voice constants and batch planning (pure host code, no CUDA; compiled without
// FMA contraction like host_tables.cpp, shared with the emulated tests).
#ifndef GTTS_MODEL5_HOST_H_
#define GTTS_MODEL5_HOST_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gtts_b200.h"
#include "tube5_types.h"
#include "tube_types.h"

namespace gtts {
namespace m5 {

// Returns nullptr on success, else a static error text (*unsupported set when the voice is valid in the reference but
// outside what the kernel implements).
const char* deriveVoice5(const gtts_voice5_config& c, Voice5Dev& v, bool* unsupported);
int32_t controlSteps5(double fs, double controlRate);
int64_t outputLength5(const Voice5Dev& v, int64_t nInternal);

struct BatchPlan5 {
	std::vector<Voice5Dev> voices;
	std::vector<UttDesc> utts;
	std::vector<int32_t> order;          // longest first
	std::vector<int64_t> out_offsets;    // n_utt + 1
	int64_t n_frames_total = 0;
};

std::string planBatch5(const gtts_voice5_config* voices, int32_t nVoices, const int32_t* voiceIndex,
			double controlRate, const int32_t* stepsOverride, const int64_t* frameOffsets,
			int64_t nUtt, BatchPlan5& plan, int* err);

} // namespace m5
} // namespace gtts
#endif
