#include "batch_plan.h"

#include <algorithm>
#include <numeric>

namespace gtts {

std::string planBatch(const gtts_voice_config* voices, int32_t nVoices, const int32_t* voiceIndex,
			double controlRate, const int32_t* stepsOverride, const int64_t* frameOffsets,
			int64_t nUtt, BatchPlan& plan, int* err)
{
	*err = GTTS_ERR_INVALID;
	if (!voices || nVoices <= 0) return "no voices given";
	if (nUtt < 0 || (nUtt > 0 && !frameOffsets)) return "bad utterance count / frame offsets";
	if (!stepsOverride && !(controlRate > 0.0)) return "control_rate must be positive";
	plan.voices.resize(nVoices);
	for (int32_t i = 0; i < nVoices; ++i) {
		const char* e = deriveVoice(voices[i], plan.voices[i]);
		if (e) {
			if (std::string(e).find("not implemented") != std::string::npos) *err = GTTS_ERR_UNSUPPORTED;
			return std::string("voice ") + std::to_string(i) + ": " + e;
		}
	}
	plan.utts.resize(nUtt);
	plan.out_offsets.assign(nUtt + 1, 0);
	plan.n_internal_total = 0;
	for (int64_t u = 0; u < nUtt; ++u) {
		UttDesc& d = plan.utts[u];
		const int32_t vi = voiceIndex ? voiceIndex[u] : 0;
		if (vi < 0 || vi >= nVoices) return "voice_index out of range";
		const int64_t f0 = frameOffsets[u], f1 = frameOffsets[u + 1];
		if (f1 < f0 || f0 < 0) return "frame_offsets must be non-decreasing";
		const VoiceDev& v = plan.voices[vi];
		int32_t steps = (stepsOverride && stepsOverride[u] > 0) ? stepsOverride[u] : controlSteps(v.fs, controlRate);
		if (steps <= 0) return "control steps must be positive (control_rate above the internal rate?)";
		d.frame_begin = f0;
		d.n_frames = f1 - f0;
		d.voice = vi;
		d.steps = steps;
		d.inv_steps = 1.0f / static_cast<float>(static_cast<unsigned int>(steps));   // Controller.cpp:287
		d.n_internal = d.n_frames * steps;
		d.n_out = outputLength(v, d.n_internal);
		d.out_begin = plan.out_offsets[u];
		d.flags = 0;
		d.state_index = -1;
		d.n_in_base = 0;
		// every utterance starts on a 64-sample (256-byte) boundary of the output buffer: the kernel then
		// writes whole aligned row pairs (see src_rows in tube_kernel_v2.cuh)
		plan.out_offsets[u + 1] = (plan.out_offsets[u] + d.n_out + 63) & ~int64_t(63);
		plan.n_internal_total += d.n_internal;
	}
	plan.n_frames_total = nUtt ? frameOffsets[nUtt] : 0;
	plan.order.resize(nUtt);
	std::iota(plan.order.begin(), plan.order.end(), 0);
	std::stable_sort(plan.order.begin(), plan.order.end(), [&](int32_t a, int32_t b) {
		return plan.utts[a].n_internal > plan.utts[b].n_internal;
	});
	*err = GTTS_OK;
	return std::string();
}

std::vector<int32_t> wideGroups(const BatchPlan& plan, std::vector<int32_t> utts)
{
	std::vector<int32_t> out;
	if (utts.empty()) return out;
	int32_t lo = plan.utts[utts[0]].steps, hi = lo;
	for (int32_t u : utts) { lo = std::min(lo, plan.utts[u].steps); hi = std::max(hi, plan.utts[u].steps); }
	const int kClasses = 8;
	auto cls = [&](int32_t u) { return hi == lo ? 0 : static_cast<int>((static_cast<int64_t>(plan.utts[u].steps - lo) * kClasses) / (hi - lo + 1)); };
	std::stable_sort(utts.begin(), utts.end(), [&](int32_t a, int32_t b) {
		const int ca = cls(a), cb = cls(b);
		if (ca != cb) return ca < cb;
		return plan.utts[a].n_internal > plan.utts[b].n_internal;
	});
	// whole groups inside a class; what is left of every class (fewer than 32 each) is grouped across classes
	std::vector<std::vector<int32_t>> groups;
	std::vector<int32_t> rest;
	size_t i = 0;
	while (i < utts.size()) {
		size_t j = i;
		while (j < utts.size() && cls(utts[j]) == cls(utts[i])) ++j;
		size_t k = i;
		for (; k + 32 <= j; k += 32) groups.emplace_back(utts.begin() + k, utts.begin() + k + 32);
		rest.insert(rest.end(), utts.begin() + k, utts.begin() + j);
		i = j;
	}
	std::stable_sort(rest.begin(), rest.end(), [&](int32_t a, int32_t b) { return plan.utts[a].n_internal > plan.utts[b].n_internal; });
	for (size_t k = 0; k < rest.size(); k += 32) groups.emplace_back(rest.begin() + k, rest.begin() + std::min(rest.size(), k + 32));
	std::stable_sort(groups.begin(), groups.end(), [&](const std::vector<int32_t>& a, const std::vector<int32_t>& b) {
		return plan.utts[a[0]].n_internal > plan.utts[b[0]].n_internal;
	});
	out.assign(groups.size() * 32, -1);
	for (size_t g = 0; g < groups.size(); ++g) std::copy(groups[g].begin(), groups[g].end(), out.begin() + g * 32);
	return out;
}

int64_t streamSamplesReady(int64_t period0, int64_t have, int32_t steps, int64_t nInDone, int32_t block)
{
	if (have < 2) return 0;
	const int64_t avail = (period0 + have - 1) * steps - nInDone;
	return avail > 0 ? ((avail - 1) / block) * block : 0;
}

int64_t streamOutputsAfter(const VoiceDev& v, int64_t nIn, bool flush)
{
	if (flush) return outputLength(v, nIn);
	if (nIn == 0) return 0;
	return static_cast<int64_t>(((static_cast<unsigned __int128>(nIn) << 16) + v.src_inc - 1) / v.src_inc);
}

UttDesc streamChunkDesc(const UttDesc& base, int64_t nAvail, int64_t nSamples, int64_t nInDone, int64_t nOutDone,
			int64_t outputsAfter, bool flush)
{
	UttDesc d = base;
	d.frame_begin = 0;
	d.n_frames = nAvail;
	d.n_internal = nSamples;
	d.n_in_base = nInDone;
	d.out_begin = -nOutDone;              // the kernel indexes outputs absolutely; staging index = k - nOutDone
	d.n_out = outputsAfter;
	d.flags = 1 | (flush ? 0 : 2);
	d.state_index = 0;
	return d;
}

void lcgMultipliers(unsigned long long* out32)
{
	const unsigned long long mask = (1ull << 44) - 1;
	unsigned long long p = 1;
	for (int j = 0; j < 32; ++j) {
		p = (p * 377ull) & mask;
		out32[j] = p;
	}
}

unsigned long long lcgInitialState()
{
	const unsigned long long mask = (1ull << 44) - 1;
	const double seed0 = 0.7892347;
	const double product = seed0 * 377.0;                                  // one IEEE multiplication, as in the reference
	const double s1d = product - static_cast<double>(static_cast<int>(product));
	const unsigned long long s1 = static_cast<unsigned long long>(s1d * 17592186044416.0);    // exact: s1d is on the 2^-44 grid
	// inverse of 377 modulo 2^44 by Newton iteration (377 is odd)
	unsigned long long inv = 377;
	for (int i = 0; i < 6; ++i) inv = (inv * (2 - 377 * inv)) & mask;
	return (s1 * inv) & mask;
}

} // namespace gtts
