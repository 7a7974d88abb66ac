// TEST INFRASTRUCTURE ONLY: runs the device source of the tube kernel under the SIMT emulator
// (see simt_emu.h).  Built by tests/simt_emu/Makefile into tests/simt_emu/libemu_tube.so.
#include "simt_emu.h"

#define GTTS_EMU 1
namespace gtts {
double c_fir[64];
unsigned long long c_lcg[32];
unsigned long long c_lcg_init;
}
#include "../../gama_tts_b200/csrc/tube_kernel.cuh"
#include "../../gama_tts_b200/csrc/batch_plan.h"

namespace simt {
Cta* g_cta = nullptr;
void fiber_entry(int tid)
{
	g_cta->body(tid);
	g_cta->fibers[tid].state = 2;
}
}

static std::string g_err;

extern "C" const char* emu_last_error() { return g_err.c_str(); }

// Same planning code as the product host runtime, kernel body run CTA by CTA on fibers.
extern "C" int emu_batch(const gtts_voice_config* voices, int n_voices, const int* voice_index, double control_rate,
			const int* steps_override, const float* frames, const long long* frame_offsets, long long n_utt,
			float* out, long long* out_offsets, long long* out_lengths, int warps_per_cta, int n_ctas)
{
	using namespace gtts;
	BatchPlan plan;
	int err = 0;
	g_err = planBatch(voices, n_voices, voice_index, control_rate, steps_override,
			reinterpret_cast<const int64_t*>(frame_offsets), n_utt, plan, &err);
	if (err) return err;
	for (long long u = 0; u <= n_utt; ++u) out_offsets[u] = plan.out_offsets[u];
	for (long long u = 0; u < n_utt; ++u) out_lengths[u] = plan.utts[u].n_out;
	if (!out) return 0;

	std::vector<double> taps = designGlottalFir();
	std::memset(c_fir, 0, sizeof c_fir);
	for (size_t i = 0; i < taps.size(); ++i) c_fir[i] = taps[i];
	lcgMultipliers(c_lcg);
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }

	int queue = 0;
	KernelParams P;
	P.voices = plan.voices.data();
	P.utts = plan.utts.data();
	P.order = plan.order.data();
	P.frames = frames;
	P.out = out;
	P.states = nullptr;
	P.src_tab = tab.data();
	P.queue = &queue;
	P.n_utt = static_cast<int32_t>(n_utt);

	const int nthreads = warps_per_cta * 32;
	std::vector<unsigned char> smem(tube_smem_bytes(warps_per_cta) + 64);
	unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	for (int b = 0; b < n_ctas; ++b) {
		simt::run_cta(nthreads, [&](int tid) { tube_cta_body(P, base, tid, nthreads); });
	}
	return 0;
}

// ---- v1 (pipelined, warp-specialised kernel) -----------------------------------------------------------
#include "../../gama_tts_b200/csrc/tube_kernel_v1.cuh"

extern "C" int emu_batch_v1(const gtts_voice_config* voices, int n_voices, const int* voice_index, double control_rate,
			const int* steps_override, const float* frames, const long long* frame_offsets, long long n_utt,
			float* out, long long* out_offsets, long long* out_lengths, int n_ctas)
{
	using namespace gtts;
	BatchPlan plan;
	int err = 0;
	g_err = planBatch(voices, n_voices, voice_index, control_rate, steps_override,
			reinterpret_cast<const int64_t*>(frame_offsets), n_utt, plan, &err);
	if (err) return err;
	// EMU_OUT_SHIFT=s (tests only): utterance u starts s + 7 u samples (mod 32) off its 32-sample row, which the
	// planner never produces -- exercises the partial first row and the row phases of the SRC stage.
	if (const char* sh = std::getenv("EMU_OUT_SHIFT")) {
		const long long shift = std::atoll(sh);
		for (long long u = 0; u < n_utt; ++u) {
			plan.out_offsets[u] += 64 * u + ((7 * u + shift) & 31);
			plan.utts[u].out_begin = plan.out_offsets[u];
		}
		plan.out_offsets[n_utt] += 64 * n_utt + 32;
	}
	for (long long u = 0; u <= n_utt; ++u) out_offsets[u] = plan.out_offsets[u];
	for (long long u = 0; u < n_utt; ++u) out_lengths[u] = plan.utts[u].n_out;
	if (!out) return 0;
	for (const UttDesc& d : plan.utts) {
		if (d.steps < kBlock && d.steps != 1) { g_err = "v1 needs control periods of at least one block (or of one sample)"; return GTTS_ERR_UNSUPPORTED; }
	}
	std::vector<double> taps = designGlottalFir();
	std::memset(c_fir, 0, sizeof c_fir);
	for (size_t i = 0; i < taps.size(); ++i) c_fir[i] = taps[i];
	lcgMultipliers(c_lcg);
	c_lcg_init = lcgInitialState();
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }
	std::vector<double> tables(static_cast<size_t>(n_voices) * kTableLen);
	for (int v = 0; v < n_voices; ++v) buildWavetable(plan.voices[v], tables.data() + static_cast<size_t>(v) * kTableLen);

	int queue = 0;
	v1::KernelParamsV1 P;
	P.voices = plan.voices.data();
	P.tables = tables.data();
	P.utts = plan.utts.data();
	P.order = plan.order.data();
	P.frames = frames;
	P.out = out;
	P.src_tab = tab.data();
	P.queue = &queue;
	P.n_utt = static_cast<int32_t>(n_utt);
	P.prof = nullptr;
	P.debug_skip = 0;

	std::vector<unsigned char> smem(v1::smem_bytes() + 64);
	unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	for (int b = 0; b < n_ctas; ++b) {
		simt::run_cta(v1::kThreads, [&](int tid) { v1::tube_v1_cta_body(P, base, tid); });
	}
	return 0;
}

// ---- v2 (decoupled roles, walker warps, two-output SRC) ---------------------------------------------------------
#include "../../gama_tts_b200/csrc/tube_kernel_v2.cuh"

extern "C" int emu_batch_v2(const gtts_voice_config* voices, int n_voices, const int* voice_index, double control_rate,
			const int* steps_override, const float* frames, const long long* frame_offsets, long long n_utt,
			float* out, long long* out_offsets, long long* out_lengths, int n_ctas)
{
	using namespace gtts;
	BatchPlan plan;
	int err = 0;
	g_err = planBatch(voices, n_voices, voice_index, control_rate, steps_override,
			reinterpret_cast<const int64_t*>(frame_offsets), n_utt, plan, &err);
	if (err) return err;
	// EMU_OUT_SHIFT=s (tests only): utterance u starts s + 7 u samples (mod 32) off its row, which the planner never
	// produces -- exercises the partial first row and the row phases of the SRC stage.
	if (const char* sh = std::getenv("EMU_OUT_SHIFT")) {
		const long long shift = std::atoll(sh);
		for (long long u = 0; u < n_utt; ++u) {
			plan.out_offsets[u] += 128 * u + ((7 * u + shift) & 31);
			plan.utts[u].out_begin = plan.out_offsets[u];
		}
		plan.out_offsets[n_utt] += 128 * n_utt + 64;
	}
	for (long long u = 0; u <= n_utt; ++u) out_offsets[u] = plan.out_offsets[u];
	for (long long u = 0; u < n_utt; ++u) out_lengths[u] = plan.utts[u].n_out;
	if (!out) return 0;
	for (const UttDesc& d : plan.utts) {
		if (d.steps < kBlock && d.steps != 1) { g_err = "v2 needs control periods of at least one block (or of one sample)"; return GTTS_ERR_UNSUPPORTED; }
	}
	std::vector<double> taps = designGlottalFir();
	std::memset(c_fir, 0, sizeof c_fir);
	for (size_t i = 0; i < taps.size(); ++i) c_fir[i] = taps[i];
	lcgMultipliers(c_lcg);
	c_lcg_init = lcgInitialState();
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }
	std::vector<double> tables(static_cast<size_t>(n_voices) * kTableLen);
	for (int v = 0; v < n_voices; ++v) buildWavetable(plan.voices[v], tables.data() + static_cast<size_t>(v) * kTableLen);

	int queue = 0;
	v2::KernelParamsV2 P;
	P.voices = plan.voices.data();
	P.tables = tables.data();
	P.utts = plan.utts.data();
	P.order = plan.order.data();
	P.frames = frames;
	P.out = out;
	P.src_tab = tab.data();
	P.queue = &queue;
	P.n_utt = static_cast<int32_t>(n_utt);
	P.prof = nullptr;
	P.debug_skip = 0;
	P.states = nullptr;

	std::vector<unsigned char> smem(v2::smem_bytes() + 64);
	unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	for (int b = 0; b < n_ctas; ++b) {
		simt::run_cta(v2::kThreads, [&](int tid) { v2::tube_v2_cta_body<false>(P, base, tid); });
	}
	return 0;
}

// A stream on the pipelined kernel: frames are pushed in pieces of push_sizes[i] frames, every chunk is one emulated
// launch of the streaming instantiation with the roles' state carried in UttStateV2 (host logic of runtime.cu:
// gtts_stream_push_frames / gtts_stream_finish, through the same planning functions).
extern "C" long long emu_stream_v2(const gtts_voice_config* voice, double control_rate, const float* frames, long long n_frames,
			const int* push_sizes, int n_pushes, float* out, long long cap)
{
	using namespace gtts;
	BatchPlan plan;
	int err = 0;
	const int64_t fo[2] = {0, 0};
	g_err = planBatch(voice, 1, nullptr, control_rate, nullptr, fo, 1, plan, &err);
	if (err) return -err;
	const UttDesc base = plan.utts[0];
	const int32_t steps = base.steps;
	if (steps < kBlock) { g_err = "control period shorter than one block"; return -GTTS_ERR_UNSUPPORTED; }
	std::vector<double> taps = designGlottalFir();
	std::memset(c_fir, 0, sizeof c_fir);
	for (size_t i = 0; i < taps.size(); ++i) c_fir[i] = taps[i];
	lcgMultipliers(c_lcg);
	c_lcg_init = lcgInitialState();
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }
	std::vector<double> table(kTableLen);
	buildWavetable(plan.voices[0], table.data());
	UttStateV2 state;
	std::memset(&state, 0, sizeof state);
	std::vector<float> pending;
	int64_t period0 = 0, nInDone = 0, nOutDone = 0, fed = 0;
	const int32_t order0 = 0;
	std::vector<unsigned char> smem(v2::smem_bytes() + 64);
	unsigned char* sbase = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	auto run = [&](int64_t nSamples, bool flush) -> int {
		const int64_t nAvail = static_cast<int64_t>(pending.size()) / kNumParams;
		const int64_t kAfter = streamOutputsAfter(plan.voices[0], nInDone + nSamples, flush);
		if (kAfter > cap) { g_err = "output buffer too small"; return 1; }
		UttDesc d = streamChunkDesc(base, nAvail, nSamples, nInDone, nOutDone, kAfter, flush);
		int queue = 0;
		v2::KernelParamsV2 P;
		P.voices = plan.voices.data();
		P.tables = table.data();
		P.utts = &d;
		P.order = &order0;
		P.frames = pending.data();
		P.out = out + nOutDone;               // out_begin = -nOutDone: absolute output k lands at out[k]
		P.src_tab = tab.data();
		P.queue = &queue;
		P.n_utt = 1;
		P.states = &state;
		P.prof = nullptr;
		P.debug_skip = 0;
		simt::run_cta(v2::kThreads, [&](int tid) { v2::tube_v2_cta_body<true>(P, sbase, tid); });
		nInDone += nSamples;
		nOutDone = kAfter;
		return 0;
	};
	for (int i = 0; i < n_pushes && fed < n_frames; ++i) {
		const int64_t n = std::min<int64_t>(push_sizes[i], n_frames - fed);
		pending.insert(pending.end(), frames + fed * kNumParams, frames + (fed + n) * kNumParams);
		fed += n;
		const int64_t have = static_cast<int64_t>(pending.size()) / kNumParams;
		const int64_t nSamples = streamSamplesReady(period0, have, steps, nInDone, kBlock);
		if (nSamples == 0) continue;
		if (run(nSamples, false)) return -1;
		const int64_t p0 = nInDone / steps;
		pending.erase(pending.begin(), pending.begin() + (p0 - period0) * kNumParams);
		period0 = p0;
	}
	if (fed < n_frames) {
		pending.insert(pending.end(), frames + fed * kNumParams, frames + n_frames * kNumParams);
		fed = n_frames;
	}
	const int64_t have = static_cast<int64_t>(pending.size()) / kNumParams;
	if (run(have > 0 ? (period0 + have) * steps - nInDone : 0, true)) return -1;
	return nOutDone;
}

// ---- v3 (wide batch: one thread per utterance) --------------------------------------------------------------------
#include "../../gama_tts_b200/csrc/tube_kernel_v3.cuh"

extern "C" int emu_batch_v3(const gtts_voice_config* voices, int n_voices, const int* voice_index, double control_rate,
			const int* steps_override, const float* frames, const long long* frame_offsets, long long n_utt,
			float* out, long long* out_offsets, long long* out_lengths, int n_ctas)
{
	using namespace gtts;
	BatchPlan plan;
	int err = 0;
	g_err = planBatch(voices, n_voices, voice_index, control_rate, steps_override,
			reinterpret_cast<const int64_t*>(frame_offsets), n_utt, plan, &err);
	if (err) return err;
	for (long long u = 0; u <= n_utt; ++u) out_offsets[u] = plan.out_offsets[u];
	for (long long u = 0; u < n_utt; ++u) out_lengths[u] = plan.utts[u].n_out;
	if (!out) return 0;
	for (const UttDesc& d : plan.utts) {
		if (!plan.voices[d.voice].src_upsample) { g_err = "v3 takes up-sampling voices only"; return GTTS_ERR_UNSUPPORTED; }
	}
	std::vector<double> taps = designGlottalFir();
	std::memset(c_fir, 0, sizeof c_fir);
	for (size_t i = 0; i < taps.size(); ++i) c_fir[i] = taps[i];
	c_lcg_init = lcgInitialState();
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }
	std::vector<double> tables(static_cast<size_t>(n_voices) * kTableLen);
	for (int v = 0; v < n_voices; ++v) buildWavetable(plan.voices[v], tables.data() + static_cast<size_t>(v) * kTableLen);
	const std::vector<int32_t> groups = wideGroups(plan, plan.order);

	int queue = 0;
	v3::KernelParamsV3 P;
	P.voices = plan.voices.data();
	P.tables = tables.data();
	P.utts = plan.utts.data();
	P.order = groups.data();
	P.frames = frames;
	P.out = out;
	P.src_tab = tab.data();
	P.queue = &queue;
	P.n_groups = static_cast<int32_t>(groups.size() / 32);

	std::vector<unsigned char> smem(v3::smem_bytes() + 64);
	unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	for (int b = 0; b < n_ctas; ++b) {
		simt::run_cta(v3::kThreads, [&](int tid) { v3::tube_v3_cta_body(P, base, tid); });
	}
	return 0;
}

// ---- model 5 (tube5_kernel.cuh: one warp per utterance) -----------------------------------------------------------
#include "../../gama_tts_b200/csrc/tube5_kernel.cuh"
#include "../../gama_tts_b200/csrc/model5_host.h"

extern "C" int emu_batch_m5(const gtts_voice5_config* voices, int n_voices, const int* voice_index, double control_rate,
			const int* steps_override, const float* frames, const long long* frame_offsets, long long n_utt,
			float* out, long long* out_offsets, long long* out_lengths, int warps_per_cta)
{
	using namespace gtts;
	m5::BatchPlan5 plan;
	int err = 0;
	g_err = m5::planBatch5(voices, n_voices, voice_index, control_rate, steps_override,
			reinterpret_cast<const int64_t*>(frame_offsets), n_utt, plan, &err);
	if (err) return err;
	for (long long u = 0; u <= n_utt; ++u) out_offsets[u] = plan.out_offsets[u];
	for (long long u = 0; u < n_utt; ++u) out_lengths[u] = plan.utts[u].n_out;
	if (!out) return 0;
	lcgMultipliers(c_lcg);
	c_lcg_init = lcgInitialState();
	std::vector<double> h(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(h.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) { tab[i].x = h[i]; tab[i].y = dh[i]; }
	int queue = 0;
	m5::KernelParams5 P;
	P.voices = plan.voices.data();
	P.utts = plan.utts.data();
	P.order = plan.order.data();
	P.frames = frames;
	P.out = out;
	P.src_tab = tab.data();
	P.queue = &queue;
	P.n_utt = static_cast<int32_t>(n_utt);
	const int nthreads = warps_per_cta * 32;
	std::vector<unsigned char> smem(m5::smem_bytes(warps_per_cta) + 64);
	unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
	simt::run_cta(nthreads, [&](int tid) { m5::tube5_cta_body(P, base, tid, nthreads); });
	return 0;
}

// ---- control-frame generation (events_kernel.cuh: one warp per utterance) -------------------------------------------
#include "../../gama_tts_b200/csrc/events_kernel.cuh"

extern "C" int emu_events(const gtts_event_config* configs, const int* continues_previous, const gtts_event* events,
			const long long* event_offsets, long long n_chunks, float* frames, long long* frame_offsets,
			gtts_event_config* configs_out, int warps_per_cta)
{
	using namespace gtts;
	evt::EventsPlan plan;
	int err = 0;
	g_err = evt::planEvents(configs, continues_previous, events, reinterpret_cast<const int64_t*>(event_offsets), n_chunks, plan, &err);
	if (err) return err;
	for (long long c = 0; c <= n_chunks; ++c) frame_offsets[c] = plan.frame_offsets[c];
	if (!frames) return 0;
	int queue[2] = {0, 0};
	evt::EventsParams P;
	P.events = reinterpret_cast<const double*>(events);
	P.cfgs = plan.cfgs.data();
	P.cfgs_out = configs_out;
	P.chunks = plan.chunks.data();
	P.chains = plan.chains.data();
	P.order = plan.order.data();
	P.chunk_order = plan.chunk_order.data();
	P.frames = frames;
	std::vector<float> drift(static_cast<size_t>(plan.frame_offsets.back()) + 1);
	P.drift = drift.data();
	P.queue = queue;
	P.n_chains = static_cast<int32_t>(plan.chains.size());
	P.n_chunks = static_cast<int32_t>(plan.chunks.size());
	for (int t = 0; t < P.n_chains; ++t) evt::drift_body(P, t);
	std::vector<double> ring(static_cast<size_t>(warps_per_cta) * evt::kRingRows * evt::kEventDoubles);
	std::vector<unsigned> masks(static_cast<size_t>(warps_per_cta) * evt::kMaskWords * 32);
	std::vector<float> pitch(static_cast<size_t>(warps_per_cta) * 64);
	simt::run_cta(warps_per_cta * 32, [&](int tid) { evt::events_cta_body(P, ring.data(), masks.data(), pitch.data(), tid); });
	if (queue[1]) { g_err = "frame count mismatch"; return GTTS_ERR_CUDA; }
	return 0;
}
