// Host-side planning of a batch: per-utterance descriptors, output layout, processing order.
// Pure C++ (no CUDA) so that it can be unit-tested without a GPU.
#ifndef GTTS_BATCH_PLAN_H_
#define GTTS_BATCH_PLAN_H_

#include <cstdint>
#include <string>
#include <vector>

#include "host_tables.h"

namespace gtts {

struct BatchPlan {
	std::vector<VoiceDev> voices;
	std::vector<UttDesc> utts;
	std::vector<int32_t> order;          // longest first (persistent warps pop from the front)
	std::vector<int64_t> out_offsets;    // n_utt + 1
	int64_t n_frames_total = 0;
	int64_t n_internal_total = 0;
};

// Returns an empty string on success, else the error text (code in *err: GTTS_ERR_*).
std::string planBatch(const gtts_voice_config* voices, int32_t nVoices, const int32_t* voiceIndex,
			double controlRate, const int32_t* stepsOverride, const int64_t* frameOffsets,
			int64_t nUtt, BatchPlan& plan, int* err);

// ---- wide-batch kernel (tube_kernel_v3.cuh: one thread per utterance, 32 utterances of a warp in lockstep) ----------
// Groups the utterances `utts` (indices into plan.utts) into warps: utterances of similar internal rate together
// (the SRC loop of a warp runs as often as its fastest-converting lane needs), longest first inside a rate class
// (a warp runs until its longest utterance ends), the groups themselves longest first (the persistent warps pop
// them in this order).  Returns n_groups x 32 indices, -1 for the empty lanes of the last group of a class.
std::vector<int32_t> wideGroups(const BatchPlan& plan, std::vector<int32_t> utts);

// ---- streaming on the pipelined kernel: chunk planning (pure host logic, shared with the emulated tests) ----------
// A stream is synthesised in chunks of whole 32-sample blocks; the frame window handed to the kernel starts at the
// control period `period0` that contains the first sample of the chunk.
// Samples that can be synthesised now: the control periods whose end frame is known (`have` frames in the window),
// in whole blocks, leaving at least one sample for the final chunk.
int64_t streamSamplesReady(int64_t period0, int64_t have, int32_t steps, int64_t nInDone, int32_t block);
// Outputs complete after nIn internal samples (flush: the whole utterance, SampleRateConverter.h:462-471).
int64_t streamOutputsAfter(const VoiceDev& v, int64_t nIn, bool flush);
// The descriptor of a chunk of nSamples internal samples over a window of nAvail frames.
UttDesc streamChunkDesc(const UttDesc& base, int64_t nAvail, int64_t nSamples, int64_t nInDone, int64_t nOutDone,
			int64_t outputsAfter, bool flush);

// 377^(j+1) mod 2^44, j = 0..31: jump-ahead multipliers of the noise generator
// (reference NoiseSource.h:40-44 is exactly this LCG on the 2^-44 grid).
void lcgMultipliers(unsigned long long* out32);
// The state s0 with 377 * s0 = s1 (mod 2^44), s1 = frac(0.7892347 * 377) * 2^44 being the reference's first
// seed after reset (exactly on the grid): lane j of the first block then gets s0 * 377^(j+1) like any other block.
unsigned long long lcgInitialState();

} // namespace gtts
#endif
