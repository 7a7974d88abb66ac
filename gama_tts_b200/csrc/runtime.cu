// Host runtime + C ABI (include/gtts_b200.h) of the B200 tube path.
//
// One gtts_handle per GPU.  A gtts_batch owns the plan of U utterances (descriptors, voice
// constants, processing order) in device memory; running it is one persistent kernel launch whose
// warps pull utterances longest-first from an atomic queue.  No CPU fallback exists: every device
// entry point fails with GTTS_ERR_NO_DEVICE / GTTS_ERR_CUDA if CUDA is unusable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>
#include <new>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gtts_b200.h"
#include "batch_plan.h"
#include "host_tables.h"
#include "tube_kernel.cuh"
#include "tube_kernel_v1.cuh"
#include "tube_kernel_v2.cuh"
#include "tube_kernel_v3.cuh"
#include "tube5_kernel.cuh"
#include "model5_host.h"
#include "events_kernel.cuh"

using namespace gtts;

namespace {

thread_local std::string t_error;

int fail(int code, const std::string& text)
{
	t_error = text;
	return code;
}

int failCuda(cudaError_t e, const char* what)
{
	return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? GTTS_ERR_NO_DEVICE : GTTS_ERR_CUDA,
			std::string(what) + ": " + cudaGetErrorString(e));
}

#define GTTS_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return failCuda(e_, #call); } while (0)

constexpr int kWarpsPerCta = 8;
constexpr int kWarps5 = 12;              // model-5 kernel: warps (utterances) per CTA, one CTA per SM (12 x 12 KB + the 53 KB converter table)

// FP64 FMA-pipe peak probe: 8 independent register-resident DFMA chains per thread.
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b)
{
	double x[8];
#pragma unroll
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- output stage on the device: peak normalisation + 16-bit PCM ------------------------------------------------
// The reference's Controller::writeOutputToFile (vtm_control_model/Controller.cpp:315-328): scale =
// VTM::Util::calculateOutputScale(buffer) = 0.95f / max|x| (0 below 1e-30f; vtm/VTMUtil.cpp:20-21, 48-57), then
// WAVEFileWriter::writeSample(x * scale) = (int) round((x * scale) * 32767.0f), low 16 bits (WAVEFileWriter.cpp:36-37,
// 122-126).  All in float32, every product rounded on its own (no contraction), round half away from zero: the PCM
// payload is bit-identical to the reference's for the same float32 input.  One CTA per utterance, HBM-bound.
__global__ void utterance_peak_kernel(const float* audio, const UttDesc* utts, unsigned* peakBits)
{
	const UttDesc U = utts[blockIdx.x];
	const float* x = audio + U.out_begin;
	float m = 0.0f;
	const long long n4 = U.n_out >> 2;
	const float4* x4 = reinterpret_cast<const float4*>(x);          // utterances start on 256-byte boundaries
	for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
		const float4 v = x4[i];
		m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
	}
	for (long long i = (n4 << 2) + threadIdx.x; i < U.n_out; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
	for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
	__shared__ float warpMax[32];
	if ((threadIdx.x & 31) == 0) warpMax[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x < 32) {
		m = threadIdx.x < (blockDim.x >> 5) ? warpMax[threadIdx.x] : 0.0f;
		for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
		if (threadIdx.x == 0) peakBits[blockIdx.x] = __float_as_uint(m);     // NaN-free maxima of non-negative floats
	}
}

// Order-independent 64-bit checksum of one utterance's float32 audio (sum of bits_i * (2 i + 1) mod 2^64): what the
// multi-GPU determinism checks compare instead of the audio itself (BASELINE config 4).
__global__ void utterance_checksum_kernel(const float* audio, const UttDesc* utts, unsigned long long* sums)
{
	const UttDesc U = utts[blockIdx.x];
	const unsigned* x = reinterpret_cast<const unsigned*>(audio + U.out_begin);
	unsigned long long h = 0;
	for (long long i = threadIdx.x; i < U.n_out; i += blockDim.x) h += (unsigned long long) x[i] * (unsigned long long) (2 * i + 1);
	for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
	__shared__ unsigned long long warpSum[32];
	if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = h;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long t = 0;
		for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += warpSum[w];
		sums[blockIdx.x] = t + (unsigned long long) U.n_out;
	}
}

__device__ __forceinline__ short pcm16_of(float x, float scale)
{
	const float sample = __fmul_rn(x, scale);
	const int v = (int) roundf(__fmul_rn(sample, 32767.0f));
	return (short) (v & 0xffff);
}

// `utts` gives where the float32 audio of an utterance lies, `dst` where its payload goes (the same layout for a
// batch on one GPU; a shard of a multi-GPU batch keeps its audio compact on the device and writes the payload at the
// utterance's place in the caller's buffer).
__global__ void utterance_pcm16_kernel(const float* audio, const UttDesc* utts, const UttDesc* dst, const unsigned* peakBits, short* pcm, float* scaleOut)
{
	const UttDesc U = utts[blockIdx.x];
	const float peak = __uint_as_float(peakBits[blockIdx.x]);
	const float scale = (peak < 1.0e-30f) ? 0.0f : __fdiv_rn(0.95f, peak);
	if (scaleOut != nullptr && threadIdx.x == 0) scaleOut[blockIdx.x] = scale;
	const float* x = audio + U.out_begin;
	short* o = pcm + dst[blockIdx.x].out_begin;
	const long long n8 = U.n_out >> 3;
	const float4* x4 = reinterpret_cast<const float4*>(x);
	for (long long i = threadIdx.x; i < n8; i += blockDim.x) {
		const float4 a = x4[2 * i], b = x4[2 * i + 1];
		short4 lo, hi;
		lo.x = pcm16_of(a.x, scale); lo.y = pcm16_of(a.y, scale); lo.z = pcm16_of(a.z, scale); lo.w = pcm16_of(a.w, scale);
		hi.x = pcm16_of(b.x, scale); hi.y = pcm16_of(b.y, scale); hi.z = pcm16_of(b.z, scale); hi.w = pcm16_of(b.w, scale);
		int4 packed;
		packed.x = (unsigned short) lo.x | ((unsigned) (unsigned short) lo.y << 16);
		packed.y = (unsigned short) lo.z | ((unsigned) (unsigned short) lo.w << 16);
		packed.z = (unsigned short) hi.x | ((unsigned) (unsigned short) hi.y << 16);
		packed.w = (unsigned short) hi.z | ((unsigned) (unsigned short) hi.w << 16);
		reinterpret_cast<int4*>(o)[i] = packed;                         // 16 bytes per thread, utterances start on 128-byte boundaries
	}
	for (long long i = (n8 << 3) + threadIdx.x; i < U.n_out; i += blockDim.x) o[i] = pcm16_of(x[i], scale);
}

} // namespace

struct gtts_handle {
	int device = 0;
	int sms = 0;
	std::mutex order_lock;
	cudaEvent_t last_output_stage = nullptr;   // end of the output stage of the batch submitted last (PCM path)
	double2* d_src_tab = nullptr;
	std::string description;
};

struct gtts_batch {
	gtts_handle* h = nullptr;
	BatchPlan plan;
	VoiceDev* d_voices = nullptr;
	UttDesc* d_utts = nullptr;
	int32_t* d_order = nullptr;
	int32_t* d_queue = nullptr;
	UttState* d_states = nullptr;       // streaming only
	float* d_frames = nullptr;          // staging for run_host
	float* d_out = nullptr;
	short* d_pcm = nullptr;             // run_host_pcm16 staging
	UttDesc* d_utts_local = nullptr;    // shard of a multi-GPU batch: the same utterances laid out compactly (device-side audio of the PCM path)
	int64_t local_out_total = 0;
	unsigned* d_peak = nullptr;         // per-utterance max |sample| (float bits)
	float* d_scale = nullptr;           // per-utterance normalisation scale
	int64_t cap_frames = 0, cap_out = 0, cap_pcm = 0;
	cudaStream_t stream = nullptr;      // used by run_host
	cudaStream_t stream_out = nullptr;  // highest priority: output stage + payload copy of the PCM path (see submit_host_pcm16)
	cudaEvent_t ev_synth = nullptr;
	int32_t last_launches = 0;
	double* d_tables = nullptr;         // per-voice glottal wavetables (v1 kernel)
	int kernel_hint = -1;               // shard of a multi-GPU batch: 1 = the whole batch is uniform (tube_kernel_v1), 0 = not
	bool legacy_v1 = false;             // tube_kernel_v1 (one CTA barrier per iteration) instead of v2: uniform batches, or GTTS_KERNEL=v1
	bool streaming = false;             // one-utterance batch of a gtts_stream (set before the plan is uploaded)
	int32_t n_fast = 0;                 // the first n_fast entries of the order list run on the pipelined kernel
	int32_t* d_order_wide = nullptr;    // wide-batch kernel (one thread per utterance): n_wide_groups x 32 indices, -1 = empty lane
	int32_t n_wide_groups = 0;
	int32_t n_wide = 0;                 // utterances in those groups (they are NOT in the order list)
	const char* last_kernel = "none";
};

struct gtts_stream {
	gtts_handle* h = nullptr;
	gtts_voice_config voice;
	VoiceDev vdev;
	int32_t steps = 0;
	gtts_batch* batch = nullptr;        // one-utterance batch re-used for every chunk
	std::vector<float> pending;         // frames not yet consumed as a period start (<= 1 after a push)
	int64_t n_in_done = 0, n_out_done = 0;
	bool finished = false;
	// fast path (pipelined kernel, control periods of at least one block): whole 32-sample blocks per chunk, every
	// role's state in d_state between chunks, frames / descriptor / audio through mapped pinned staging buffers,
	// one graph launch (counter reset + kernel) per chunk
	bool fast = false;
	int64_t period0 = 0;                // control period of pending[0]
	UttStateV2* d_state = nullptr;
	float* m_frames = nullptr;          // mapped pinned: the frame window the kernel reads
	UttDesc* m_utt = nullptr;           // mapped pinned: the chunk's descriptor
	float* m_out = nullptr;             // mapped pinned: the chunk's audio
	int64_t cap_frames = 0, cap_out = 0;
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t graph_exec = nullptr;
};

namespace {

int launchBatch(gtts_batch* b, const float* dFrames, float* dOut, cudaStream_t stream, const UttDesc* dUtts = nullptr)
{
	if (!dUtts) dUtts = b->d_utts;
	const int64_t nUtt = static_cast<int64_t>(b->plan.utts.size());
	b->last_launches = 0;
	if (nUtt == 0) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaMemsetAsync(b->d_queue, 0, 4 * sizeof(int32_t), stream));
	// The processing order is [pipelined-kernel utterances | general-kernel utterances], each part longest first:
	// one launch per non-empty part, each with its own work counter (uploadPlan).
	const int32_t nFast = b->n_fast;
	b->last_kernel = "none";
	int32_t nWide = 0;
	if (b->n_wide_groups > 0) {
		// wide-batch kernel: one thread per utterance, one CTA of 128 per SM, warps pop groups of 32 (uploadPlan)
		v3::KernelParamsV3 W;
		W.voices = b->d_voices;
		W.tables = b->d_tables;
		W.utts = dUtts;
		W.order = b->d_order_wide;
		W.frames = dFrames;
		W.out = dOut;
		W.src_tab = b->h->d_src_tab;
		W.queue = b->d_queue + 2;
		W.n_groups = b->n_wide_groups;
		const int64_t ctasWanted = (static_cast<int64_t>(b->n_wide_groups) + v3::kWarps - 1) / v3::kWarps;
		const int grid = static_cast<int>(std::min<int64_t>(ctasWanted, b->h->sms));
		v3::tube_kernel_v3<<<grid, v3::kThreads, v3::smem_bytes(), stream>>>(W);
		GTTS_CUDA(cudaGetLastError());
		b->last_kernel = "tube_kernel_v3";
		b->last_launches += 1;
		nWide = b->n_wide;
	}
	const int32_t nGeneral = static_cast<int32_t>(nUtt) - nFast - nWide;
	if (nFast > 0 && !b->legacy_v1) {
		// decoupled warp-specialised kernel: one persistent CTA per SM, 7 utterance slots each
		v2::KernelParamsV2 Q;
		Q.voices = b->d_voices;
		Q.tables = b->d_tables;
		Q.utts = dUtts;
		Q.order = b->d_order;
		Q.frames = dFrames;
		Q.out = dOut;
		Q.src_tab = b->h->d_src_tab;
		Q.queue = b->d_queue;
		Q.n_utt = nFast;
		Q.prof = nullptr;
		Q.debug_skip = 0;
		Q.states = nullptr;
		const int64_t ctasWanted = (static_cast<int64_t>(nFast) + v2::kSlots - 1) / v2::kSlots;
		const int grid = static_cast<int>(std::min<int64_t>(ctasWanted, b->h->sms));
#ifdef GTTS_EXPERIMENTS
		// development builds only (tools/ab_build.sh NAME -DGTTS_EXPERIMENTS): GTTS_DEBUG_SKIP=<bits> leaves pipeline
		// roles out (wrong audio, for isolating a role under ncu); the shipped library has no such switch
		if (const char* dbg = std::getenv("GTTS_DEBUG_SKIP")) Q.debug_skip = std::atoi(dbg);
#endif
#ifdef GTTS_ROLE_PROFILE
		// development builds only: busy and waiting cycles per role (GTTS_PROFILE=1)
		const bool profile = std::getenv("GTTS_PROFILE") != nullptr;
		long long* dProf = nullptr;
		const size_t rowLen = 2 * v2::kWarps + 1;
		if (profile) {
			GTTS_CUDA(cudaMalloc(&dProf, sizeof(long long) * grid * rowLen));
			if (cudaMemsetAsync(dProf, 0, sizeof(long long) * grid * rowLen, stream) != cudaSuccess) {
				cudaFree(dProf);
				return failCuda(cudaGetLastError(), "profile buffer");
			}
			Q.prof = dProf;
		}
#endif
		v2::tube_kernel_v2<<<grid, v2::kThreads, v2::smem_bytes(), stream>>>(Q);
		GTTS_CUDA(cudaGetLastError());
#ifdef GTTS_ROLE_PROFILE
		if (profile) {
			std::vector<long long> hp(static_cast<size_t>(grid) * rowLen);
			cudaError_t pe = cudaStreamSynchronize(stream);
			if (pe == cudaSuccess) pe = cudaMemcpy(hp.data(), dProf, sizeof(long long) * hp.size(), cudaMemcpyDeviceToHost);
			cudaFree(dProf);
			if (pe != cudaSuccess) return failCuda(pe, "profile readback");
			double busy[v2::kWarps] = {0}, wait[v2::kWarps] = {0};
			double iters = 0;
			for (int c = 0; c < grid; ++c) {
				for (int w = 0; w < v2::kWarps; ++w) {
					busy[w] += static_cast<double>(hp[static_cast<size_t>(c) * rowLen + w]);
					wait[w] += static_cast<double>(hp[static_cast<size_t>(c) * rowLen + v2::kWarps + 1 + w]);
				}
				iters += static_cast<double>(hp[static_cast<size_t>(c) * rowLen + v2::kWarps]);
			}
			std::fprintf(stderr, "[gtts profile] grid %d, iterations per CTA %.0f; busy cycles per iteration by role:", grid, iters / grid);
			for (int w = 0; w < v2::kWarps; ++w) std::fprintf(stderr, " %d:%.0f", w, busy[w] / (iters > 0 ? iters : 1));
			std::fprintf(stderr, "\n[gtts profile] waiting cycles per iteration by role:");
			for (int w = 0; w < v2::kWarps; ++w) std::fprintf(stderr, " %d:%.0f", w, wait[w] / (iters > 0 ? iters : 1));
			std::fprintf(stderr, "\n");
		}
#endif
		b->last_kernel = nWide > 0 ? "tube_kernel_v3+tube_kernel_v2" : "tube_kernel_v2";
		b->last_launches += 1;
	} else if (nFast > 0) {
		// the barrier-per-iteration kernel: batches whose utterances all have one voice and one length (prepare), or GTTS_KERNEL=v1
		v1::KernelParamsV1 Q;
		Q.voices = b->d_voices;
		Q.tables = b->d_tables;
		Q.utts = dUtts;
		Q.order = b->d_order;
		Q.frames = dFrames;
		Q.out = dOut;
		Q.src_tab = b->h->d_src_tab;
		Q.queue = b->d_queue;
		Q.n_utt = nFast;
		Q.prof = nullptr;
		Q.debug_skip = 0;
		const int64_t ctasWanted = (static_cast<int64_t>(nFast) + v1::kSlots - 1) / v1::kSlots;
		const int grid = static_cast<int>(std::min<int64_t>(ctasWanted, b->h->sms));
		v1::tube_kernel_v1<<<grid, v1::kThreads, v1::smem_bytes(), stream>>>(Q);
		GTTS_CUDA(cudaGetLastError());
		b->last_kernel = "tube_kernel_v1";
		b->last_launches += 1;
	}
	if (nGeneral > 0) {
		KernelParams P;
		P.voices = b->d_voices;
		P.utts = dUtts;
		P.order = b->d_order + nFast;
		P.frames = dFrames;
		P.out = dOut;
		P.states = b->d_states;
		P.src_tab = b->h->d_src_tab;
		P.queue = b->d_queue + 1;
		P.n_utt = nGeneral;
		const int64_t ctasWanted = (static_cast<int64_t>(nGeneral) + kWarpsPerCta - 1) / kWarpsPerCta;
		const int grid = static_cast<int>(std::min<int64_t>(ctasWanted, b->h->sms));
		const size_t smem = tube_smem_bytes(kWarpsPerCta);
		tube_kernel_v0<kWarpsPerCta><<<grid, kWarpsPerCta * 32, smem, stream>>>(P);
		GTTS_CUDA(cudaGetLastError());
		b->last_kernel = (nFast > 0 || nWide > 0) ? "pipelined+tube_kernel_v0" : "tube_kernel_v0";
		b->last_launches += 1;
	}
	return GTTS_OK;
}

int uploadPlan(gtts_batch* b)
{
	BatchPlan& p = b->plan;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	// Every upload of the plan goes through the batch's own stream and is synchronised before the call returns:
	// the kernels run on that stream or on the caller's, neither of which is ordered against the legacy default
	// stream a plain cudaMemcpy / cudaMemset would use.
	if (!b->stream) GTTS_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	GTTS_CUDA(cudaMalloc(&b->d_voices, sizeof(VoiceDev) * std::max<size_t>(p.voices.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_utts, sizeof(UttDesc) * std::max<size_t>(p.utts.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_order, sizeof(int32_t) * std::max<size_t>(p.order.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_queue, 4 * sizeof(int32_t)));
	// Kernel choice per utterance: v1 (pipelined) needs control periods of at least one 32-sample block, or of
	// exactly one sample (the plugin shim's mode: every internal sample has its own frame); the general v0 kernel
	// takes everything else (and all streaming / resumed work).  The order list is partitioned accordingly, each
	// part longest first (stable partition of the planner's order), so one odd utterance no longer moves the whole
	// batch to the slow kernel.  GTTS_KERNEL=v0 forces v0 for all.
	bool forceGeneral = b->streaming;     // resumed utterances (UttState) exist in the general kernel only
	if (const char* env = std::getenv("GTTS_KERNEL")) {
		if (std::strcmp(env, "v0") == 0) forceGeneral = true;
		b->legacy_v1 = std::strcmp(env, "v1") == 0;
	}
	// The wide-batch kernel (one thread per utterance, tube_kernel_v3.cuh) is kept as a measured alternative, not a
	// default: a thread advances one sample per ~2.5 us (one warp per SM sub-partition is what its shared-memory
	// arrays leave room for), against 0.25 us per sample of an utterance on the pipelined kernel.  Measured on B200:
	// BASELINE config 3 (lengths up to 20 s: the run time is bounded by the longest utterance) 4.5x slower, 37,888
	// utterances of 0.5-1.5 s -- its best case -- 194 ms against 184 ms.  GTTS_KERNEL=v3 selects it (tests,
	// tools/v3_probe.py, tools/v3_short_probe.py); it takes the up-sampling utterances of the batch.
	{
		bool wideForced = false;
		if (const char* env = std::getenv("GTTS_KERNEL")) wideForced = std::strcmp(env, "v3") == 0;
		auto wide = [&](int32_t u) { const VoiceDev& v = p.voices[p.utts[u].voice]; return v.src_upsample != 0 && v.tube_model == 0 && (p.utts[u].flags & 3) == 0; };
		std::vector<int32_t> cand;
		if (wideForced && !forceGeneral && !b->streaming) for (int32_t u : p.order) if (wide(u)) cand.push_back(u);
		if (!cand.empty()) {
			const std::vector<int32_t> groups = wideGroups(p, cand);
			p.order.erase(std::remove_if(p.order.begin(), p.order.end(), wide), p.order.end());
			b->n_wide = static_cast<int32_t>(cand.size());
			b->n_wide_groups = static_cast<int32_t>(groups.size() / 32);
			GTTS_CUDA(cudaMalloc(&b->d_order_wide, sizeof(int32_t) * groups.size()));
			GTTS_CUDA(cudaMemcpyAsync(b->d_order_wide, groups.data(), sizeof(int32_t) * groups.size(), cudaMemcpyHostToDevice, b->stream));
			GTTS_CUDA(cudaStreamSynchronize(b->stream));      // `groups` goes out of scope
		}
	}
	// (models 3 and 4 -- VoiceDev::tube_model -- exist in the general kernel only)
	auto fast = [&](int32_t u) { const UttDesc& d = p.utts[u]; return !forceGeneral && p.voices[d.voice].tube_model == 0 && (d.steps >= kBlock || d.steps == 1); };
	const auto mid = std::stable_partition(p.order.begin(), p.order.end(), fast);
	b->n_fast = static_cast<int32_t>(mid - p.order.begin());
	// Which pipelined kernel: when every utterance of the part has the same voice and length (BASELINE config 2), all
	// slots of a CTA stay aligned, the CTA-wide barrier of tube_kernel_v1 costs next to nothing and its smaller code wins:
	// measured on B200 25.7 ms against 27.4 ms on tube_kernel_v2 for 1,024 x 10 s.  Ragged or mixed-voice batches
	// (config 3) are where v2's decoupled roles pay: 208 k against 169 k audio-s/s.  GTTS_KERNEL=v1 / v2 forces one.
	// A shard of a multi-GPU batch takes the choice made for the whole batch (kernel_hint), so that its audio stays bit
	// for bit what one GPU produces.
	if (b->n_fast > 0 && !std::getenv("GTTS_KERNEL")) {
		if (b->kernel_hint >= 0) {
			b->legacy_v1 = b->kernel_hint == 1;
		} else {
			const UttDesc& first = p.utts[p.order[0]];
			bool uniform = true;
			for (int32_t i = 1; i < b->n_fast && uniform; ++i) {
				const UttDesc& d = p.utts[p.order[i]];
				uniform = d.voice == first.voice && d.n_frames == first.n_frames && d.steps == first.steps;
			}
			b->legacy_v1 = uniform && b->n_fast > 1;
		}
	}
	if (b->n_fast > 0 || b->n_wide > 0) {
		std::vector<double> tables(p.voices.size() * kTableLen);
		for (size_t v = 0; v < p.voices.size(); ++v) buildWavetable(p.voices[v], tables.data() + v * kTableLen);
		GTTS_CUDA(cudaMalloc(&b->d_tables, sizeof(double) * std::max<size_t>(tables.size(), 1)));
		GTTS_CUDA(cudaMemcpyAsync(b->d_tables, tables.data(), sizeof(double) * tables.size(), cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaStreamSynchronize(b->stream));      // `tables` goes out of scope
	}
	GTTS_CUDA(cudaMemcpyAsync(b->d_voices, p.voices.data(), sizeof(VoiceDev) * p.voices.size(), cudaMemcpyHostToDevice, b->stream));
	if (!p.utts.empty()) {
		GTTS_CUDA(cudaMemcpyAsync(b->d_utts, p.utts.data(), sizeof(UttDesc) * p.utts.size(), cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_order, p.order.data(), sizeof(int32_t) * p.order.size(), cudaMemcpyHostToDevice, b->stream));
	}
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}

} // namespace

extern "C" {

const char* gtts_last_error(void) { return t_error.c_str(); }
int gtts_abi_version(void) { return GTTS_ABI_VERSION; }

int gtts_voice_internal_rate(const gtts_voice_config* voice, int32_t* fs_out)
{
	if (!voice || !fs_out) return fail(GTTS_ERR_INVALID, "null argument");
	*fs_out = internalRate(*voice);
	return GTTS_OK;
}

int gtts_voice_control_steps(const gtts_voice_config* voice, double control_rate, int32_t* steps_out)
{
	if (!voice || !steps_out) return fail(GTTS_ERR_INVALID, "null argument");
	if (!(control_rate > 0.0)) return fail(GTTS_ERR_INVALID, "control_rate must be positive");
	*steps_out = controlSteps(internalRate(*voice), control_rate);
	return GTTS_OK;
}

int gtts_output_length(const gtts_voice_config* voice, int32_t steps, int64_t n_frames,
			int64_t* n_internal_out, int64_t* n_output_out)
{
	if (!voice) return fail(GTTS_ERR_INVALID, "null argument");
	if (steps <= 0 || n_frames < 0) return fail(GTTS_ERR_INVALID, "steps must be > 0 and n_frames >= 0");
	VoiceDev v;
	if (const char* e = deriveVoice(*voice, v)) return fail(GTTS_ERR_INVALID, e);
	const int64_t nInternal = n_frames * steps;
	if (n_internal_out) *n_internal_out = nInternal;
	if (n_output_out) *n_output_out = outputLength(v, nInternal);
	return GTTS_OK;
}

int gtts_shard_plan(const int64_t* cost, int64_t n_utt, int32_t n_shards, int32_t* shard_of)
{
	if (n_utt < 0 || n_shards <= 0 || (n_utt > 0 && (!cost || !shard_of))) return fail(GTTS_ERR_INVALID, "bad shard plan arguments");
	shardPlan(cost, n_utt, n_shards, shard_of);
	return GTTS_OK;
}

int gtts_probe_fir_taps(double* taps, int32_t cap, int32_t* n_taps_out)
{
	const std::vector<double> t = designGlottalFir();
	if (n_taps_out) *n_taps_out = static_cast<int32_t>(t.size());
	for (size_t i = 0; i < t.size() && static_cast<int32_t>(i) < cap; ++i) taps[i] = t[i];
	return GTTS_OK;
}

int gtts_probe_src_tables(double* h3328, double* dh3328)
{
	if (!h3328 || !dh3328) return fail(GTTS_ERR_INVALID, "null argument");
	buildSrcTables(h3328, dh3328);
	return GTTS_OK;
}

int gtts_probe_voice_constants(const gtts_voice_config* voice, double* out, int32_t cap, int32_t* n_out)
{
	if (!voice || !out) return fail(GTTS_ERR_INVALID, "null argument");
	VoiceDev v;
	if (const char* e = deriveVoice(*voice, v)) return fail(GTTS_ERR_INVALID, e);
	const double vals[] = {
		(double) v.fs, v.breath, v.crossmix, v.damping, v.rad_m, v.refl_b0_m, v.refl_a1_m, v.rad_n, v.refl_b0_n,
		v.refl_a1_n, v.throat_b0, v.throat_a1, v.throat_gain, v.nasal_k[1], v.nasal_k[2], v.nasal_k[3], v.nasal_k[4],
		v.nasal_k[5], std::sqrt(v.ap2), std::sqrt(v.nr1_2), v.basic_inc, (double) v.div1, (double) v.div2, v.tn_delta,
		(double) v.src_inc, (double) v.src_pad };
	const int32_t n = static_cast<int32_t>(sizeof vals / sizeof vals[0]);
	if (n_out) *n_out = n;
	for (int32_t i = 0; i < n && i < cap; ++i) out[i] = vals[i];
	return GTTS_OK;
}

namespace { int createHandle(int32_t device, gtts_handle** handle_out); }

int gtts_create(int32_t device, gtts_handle** handle_out)
{
	// no exception may cross the C ABI (std::vector / std::string allocations inside)
	try {
		return createHandle(device, handle_out);
	} catch (const std::bad_alloc&) {
		return fail(GTTS_ERR_NOMEM, "out of host memory in gtts_create");
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_INVALID, e.what());
	}
}

namespace {
int createHandle(int32_t device, gtts_handle** handle_out)
{
	if (!handle_out) return fail(GTTS_ERR_INVALID, "null handle_out");
	*handle_out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) return failCuda(e, "cudaGetDeviceCount");
	if (count == 0) return fail(GTTS_ERR_NO_DEVICE, "no CUDA device");
	if (device < 0 || device >= count) return fail(GTTS_ERR_INVALID, "device index out of range");
	GTTS_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	GTTS_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		return fail(GTTS_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
				"; this library carries sm_100a code only");
	}
	gtts_handle* h = new (std::nothrow) gtts_handle;
	if (!h) return fail(GTTS_ERR_NOMEM, "out of memory");
	h->device = device;
	h->sms = prop.multiProcessorCount;

	// constants: FIR taps, LCG jump multipliers, SRC tables
	const std::vector<double> taps = designGlottalFir();
	if (taps.size() != kFirTaps) { delete h; return fail(GTTS_ERR_INVALID, "unexpected glottal FIR length"); }
	double fir[kFirMaxTaps] = {0};
	std::copy(taps.begin(), taps.end(), fir);
	unsigned long long lcg[kBlock];
	lcgMultipliers(lcg);
	const unsigned long long lcgInit = lcgInitialState();
	std::vector<double> hh(kSrcFilterLen), dh(kSrcFilterLen);
	buildSrcTables(hh.data(), dh.data());
	std::vector<double2> tab(kSrcFilterLen);
	for (int i = 0; i < kSrcFilterLen; ++i) tab[i] = make_double2(hh[i], dh[i]);
	cudaError_t ce;
	if ((ce = cudaMemcpyToSymbol(c_fir, fir, sizeof fir)) != cudaSuccess ||
	    (ce = cudaMemcpyToSymbol(c_lcg, lcg, sizeof lcg)) != cudaSuccess ||
	    (ce = cudaMemcpyToSymbol(c_lcg_init, &lcgInit, sizeof lcgInit)) != cudaSuccess ||
	    (ce = cudaMalloc(&h->d_src_tab, sizeof(double2) * kSrcFilterLen)) != cudaSuccess ||
	    (ce = cudaMemcpy(h->d_src_tab, tab.data(), sizeof(double2) * kSrcFilterLen, cudaMemcpyHostToDevice)) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(tube_kernel_v0<kWarpsPerCta>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) tube_smem_bytes(kWarpsPerCta))) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(v1::tube_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) v1::smem_bytes())) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(v2::tube_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) v2::smem_bytes())) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(v2::tube_kernel_v2_stream, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) v2::smem_bytes())) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(v3::tube_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) v3::smem_bytes())) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(m5::tube5_kernel<kWarps5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               (int) m5::smem_bytes(kWarps5))) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(evt::events_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               evt::kEventsSmem)) != cudaSuccess ||
	    (ce = cudaFuncSetAttribute(evt::events_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                               evt::kEventsSmem)) != cudaSuccess) {
		if (h->d_src_tab) cudaFree(h->d_src_tab);
		delete h;
		return failCuda(ce, "gtts_create: device setup");
	}
	// the uploads above went through the legacy default stream from pageable memory: wait for them, the kernels
	// run on non-blocking streams that are not ordered against it
	if ((ce = cudaDeviceSynchronize()) != cudaSuccess) { cudaFree(h->d_src_tab); delete h; return failCuda(ce, "gtts_create: device setup"); }
	char buf[1024];
	std::snprintf(buf, sizeof buf,
			"{\"device\": %d, \"name\": \"%s\", \"sm\": \"%d.%d\", \"sms\": %d, "
			"\"kernels\": {\"tube_kernel_v2\": {\"warps_per_cta\": %d, \"utterance_slots_per_cta\": %d, \"smem_per_cta\": %zu}, "
			"\"tube_kernel_v1\": {\"warps_per_cta\": %d, \"utterance_slots_per_cta\": %d, \"smem_per_cta\": %zu, "
			"\"used_for\": \"batches of one voice and one length\"}, "
			"\"tube_kernel_v0\": {\"warps_per_cta\": %d, \"utterances_per_warp\": 1, \"smem_per_cta\": %zu}}}",
			device, prop.name, prop.major, prop.minor, h->sms, (int) v2::kWarps, (int) v2::kSlots, v2::smem_bytes(),
			(int) (v1::kThreads / 32), (int) v1::kSlots, v1::smem_bytes(), kWarpsPerCta, tube_smem_bytes(kWarpsPerCta));
	h->description = buf;
	*handle_out = h;
	return GTTS_OK;
}
} // namespace

void gtts_destroy(gtts_handle* h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->d_src_tab) cudaFree(h->d_src_tab);
	if (h->last_output_stage) cudaEventDestroy(h->last_output_stage);
	delete h;
}

const char* gtts_describe(gtts_handle* h) { return h ? h->description.c_str() : "{}"; }

namespace {
__global__ void probe_exp_kernel(const double* x, int n, double* out2, double* out10)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { out2[i] = gtts_exp2(x[i]); out10[i] = gtts_exp10(x[i]); }
}
} // namespace

int gtts_probe_exp(gtts_handle* h, const double* x, int32_t n, double* exp2_out, double* exp10_out)
{
	if (!h || !x || !exp2_out || !exp10_out || n < 0) return fail(GTTS_ERR_INVALID, "bad argument");
	if (n == 0) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(h->device));
	double* d = nullptr;
	GTTS_CUDA(cudaMalloc(&d, sizeof(double) * 3 * n));
	cudaError_t e = cudaMemcpy(d, x, sizeof(double) * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) {
		probe_exp_kernel<<<(n + 255) / 256, 256>>>(d, n, d + n, d + 2 * (size_t) n);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpy(exp2_out, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(exp10_out, d + 2 * (size_t) n, sizeof(double) * n, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return failCuda(e, "gtts_probe_exp");
	return GTTS_OK;
}

int gtts_probe_fp64_peak(gtts_handle* h, double* tflops_out)
{
	if (!h || !tflops_out) return fail(GTTS_ERR_INVALID, "null argument");
	GTTS_CUDA(cudaSetDevice(h->device));
	const int blocks = h->sms * 8, threads = 512, iters = 8192;
	double* d = nullptr;
	GTTS_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
	cudaEvent_t e0, e1;
	GTTS_CUDA(cudaEventCreate(&e0));
	GTTS_CUDA(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 6; ++rep) {
		GTTS_CUDA(cudaEventRecord(e0));
		fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
		GTTS_CUDA(cudaEventRecord(e1));
		GTTS_CUDA(cudaEventSynchronize(e1));
		float ms = 0.f;
		GTTS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		const double tf = 2.0 * 8 * iters * (double) blocks * threads / (ms * 1e-3) * 1e-12;
		if (rep >= 2 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	*tflops_out = best;
	return GTTS_OK;
}

namespace {
int prepareBatch(gtts_handle* h, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, bool streaming, gtts_batch** batch_out)
{
	if (!h || !batch_out) return fail(GTTS_ERR_INVALID, "null argument");
	*batch_out = nullptr;
	if (n_utt > std::numeric_limits<int32_t>::max()) return fail(GTTS_ERR_INVALID, "too many utterances");
	gtts_batch* b = new (std::nothrow) gtts_batch;
	if (!b) return fail(GTTS_ERR_NOMEM, "out of memory");
	b->h = h;
	b->streaming = streaming;
	// no exception may cross the C ABI: host-side allocation failures become GTTS_ERR_NOMEM
	try {
		int err = GTTS_OK;
		const std::string msg = planBatch(voices, n_voices, voice_index, control_rate, steps_override, frame_offsets, n_utt, b->plan, &err);
		if (err != GTTS_OK) { delete b; return fail(err, msg); }
		const int rc = uploadPlan(b);
		if (rc != GTTS_OK) { gtts_batch_free(b); return rc; }
	} catch (const std::bad_alloc&) {
		gtts_batch_free(b);
		return fail(GTTS_ERR_NOMEM, "out of host memory while planning the batch");
	} catch (const std::exception& e) {
		gtts_batch_free(b);
		return fail(GTTS_ERR_INVALID, e.what());
	}
	*batch_out = b;
	return GTTS_OK;
}
} // namespace

int gtts_batch_prepare(gtts_handle* h, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts_batch** batch_out)
{
	return prepareBatch(h, voices, n_voices, voice_index, control_rate, steps_override, frame_offsets, n_utt, false, batch_out);
}

int gtts_batch_layout(const gtts_batch* b, int64_t* out_offsets, int64_t* n_internal)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (out_offsets) std::copy(b->plan.out_offsets.begin(), b->plan.out_offsets.end(), out_offsets);
	if (n_internal) for (size_t u = 0; u < b->plan.utts.size(); ++u) n_internal[u] = b->plan.utts[u].n_internal;
	return GTTS_OK;
}

int gtts_batch_run_device(gtts_batch* b, const float* d_frames, float* d_out, void* cuda_stream)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (!b->plan.utts.empty() && (!d_out || (b->plan.n_frames_total > 0 && !d_frames))) return fail(GTTS_ERR_INVALID, "null device buffer");
	return launchBatch(b, d_frames, d_out, static_cast<cudaStream_t>(cuda_stream));
}

int gtts_batch_lengths(const gtts_batch* b, int64_t* n_out)
{
	if (!b || !n_out) return fail(GTTS_ERR_INVALID, "null argument");
	for (size_t u = 0; u < b->plan.utts.size(); ++u) n_out[u] = b->plan.utts[u].n_out;
	return GTTS_OK;
}

int gtts_batch_run_host(gtts_batch* b, const float* h_frames, float* h_out)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nFrames = b->plan.n_frames_total;
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	if ((nFrames > 0 && !h_frames) || (nOut > 0 && !h_out)) return fail(GTTS_ERR_INVALID, "null host buffer");
	GTTS_CUDA(cudaSetDevice(b->h->device));
	if (!b->stream) GTTS_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	// Input path.  The kernels read every track value exactly once, one control period before it is needed
	// (walk_block), so pinned host frames are read in place over PCIe (measured: no slowdown against
	// device-resident frames, and the serial 3 ms host->device copy of BASELINE config 2 disappears).
	// Pageable frames are copied to the device first.  GTTS_HOST_INPUT=staged forces the copy.
	const float* dFrames = nullptr;
	if (nFrames > 0) {
		cudaPointerAttributes attr;
		const char* env = std::getenv("GTTS_HOST_INPUT");
		const bool allow = !(env && std::strcmp(env, "staged") == 0);
		if (allow && cudaPointerGetAttributes(&attr, h_frames) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
			dFrames = static_cast<const float*>(attr.devicePointer);
		} else {
			cudaGetLastError();
			if (nFrames > b->cap_frames) {
				if (b->d_frames) cudaFree(b->d_frames);
				b->d_frames = nullptr; b->cap_frames = 0;
				GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * nFrames));
				b->cap_frames = nFrames;
			}
			GTTS_CUDA(cudaMemcpyAsync(b->d_frames, h_frames, sizeof(float) * kNumParams * nFrames, cudaMemcpyHostToDevice, b->stream));
			dFrames = b->d_frames;
		}
	}
	// Output path.  Pinned (page-locked) host memory is device-accessible under UVA: the kernel then
	// stores the float32 audio straight into the caller's buffer over PCIe while it computes, so the
	// device->host transfer overlaps the synthesis instead of following it.  Pageable buffers are
	// staged through device memory and copied afterwards.  GTTS_HOST_OUTPUT=staged forces staging.
	float* dOut = nullptr;
	bool zeroCopy = false;
	if (nOut > 0) {
		cudaPointerAttributes attr;
		const char* env = std::getenv("GTTS_HOST_OUTPUT");
		const bool allow = !(env && std::strcmp(env, "staged") == 0);
		if (allow && cudaPointerGetAttributes(&attr, h_out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
			dOut = static_cast<float*>(attr.devicePointer);
			zeroCopy = true;
		} else {
			cudaGetLastError();     // clear the error a pageable pointer may have left
		}
	}
	if (!zeroCopy && nOut > b->cap_out) {
		if (b->d_out) cudaFree(b->d_out);
		b->d_out = nullptr; b->cap_out = 0;
		GTTS_CUDA(cudaMalloc(&b->d_out, sizeof(float) * nOut));
		b->cap_out = nOut;
	}
	if (!zeroCopy) dOut = b->d_out;
	const int rc = launchBatch(b, dFrames, dOut, b->stream);
	if (rc != GTTS_OK) return rc;
	if (!zeroCopy && nOut > 0) GTTS_CUDA(cudaMemcpyAsync(h_out, b->d_out, sizeof(float) * nOut, cudaMemcpyDeviceToHost, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}

namespace {
// the output stage after the synthesis kernels, same stream
int launchPcm16(gtts_batch* b, const float* dAudio, short* dPcm, float* dScale, cudaStream_t stream, const UttDesc* dUttsAudio = nullptr)
{
	const int nUtt = static_cast<int>(b->plan.utts.size());
	if (nUtt == 0) return GTTS_OK;
	if (!dUttsAudio) dUttsAudio = b->d_utts;
	if (!b->d_peak) GTTS_CUDA(cudaMalloc(&b->d_peak, sizeof(unsigned) * nUtt));
	utterance_peak_kernel<<<nUtt, 256, 0, stream>>>(dAudio, dUttsAudio, b->d_peak);
	GTTS_CUDA(cudaGetLastError());
	utterance_pcm16_kernel<<<nUtt, 256, 0, stream>>>(dAudio, dUttsAudio, b->d_utts, b->d_peak, dPcm, dScale);
	GTTS_CUDA(cudaGetLastError());
	b->last_launches += 2;
	return GTTS_OK;
}
} // namespace

int gtts_batch_run_device_pcm16(gtts_batch* b, const float* d_frames, float* d_audio, int16_t* d_pcm, float* d_scale, void* cuda_stream)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (!b->plan.utts.empty() && (!d_audio || !d_pcm || (b->plan.n_frames_total > 0 && !d_frames))) return fail(GTTS_ERR_INVALID, "null device buffer");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	const int rc = launchBatch(b, d_frames, d_audio, stream);
	if (rc != GTTS_OK) return rc;
	return launchPcm16(b, d_audio, reinterpret_cast<short*>(d_pcm), d_scale, stream);
}

int gtts_batch_checksum_device(gtts_batch* b, const float* d_audio, uint64_t* d_sums, void* cuda_stream)
{
	if (!b || (!b->plan.utts.empty() && (!d_audio || !d_sums))) return fail(GTTS_ERR_INVALID, "null argument");
	if (b->plan.utts.empty()) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	utterance_checksum_kernel<<<static_cast<int>(b->plan.utts.size()), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
			d_audio, b->d_utts, reinterpret_cast<unsigned long long*>(d_sums));
	GTTS_CUDA(cudaGetLastError());
	return GTTS_OK;
}

int gtts_batch_wait(gtts_batch* b)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (!b->stream) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	if (b->stream_out) GTTS_CUDA(cudaStreamSynchronize(b->stream_out));
	return GTTS_OK;
}

int gtts_batch_run_host_pcm16(gtts_batch* b, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	const int rc = gtts_batch_submit_host_pcm16(b, h_frames, h_pcm, h_scale);
	if (rc != GTTS_OK) return rc;
	return gtts_batch_wait(b);
}

int gtts_batch_submit_host_pcm16(gtts_batch* b, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nFrames = b->plan.n_frames_total;
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	const int64_t nUtt = static_cast<int64_t>(b->plan.utts.size());
	if ((nFrames > 0 && !h_frames) || (nOut > 0 && !h_pcm)) return fail(GTTS_ERR_INVALID, "null host buffer");
	if (nUtt == 0) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	// frames: always copied to the device first (copy engine, 3 ms for BASELINE config 2).  Reading pinned frames in
	// place, as gtts_batch_run_host does, would compete with the payload of another batch travelling the other way
	// at the same time: the kernel's reads then starve behind the DMA stream (measured: kernel 26 -> 50 ms).
	const float* dFrames = nullptr;
	if (nFrames > 0) {
		if (nFrames > b->cap_frames) {
			if (b->d_frames) cudaFree(b->d_frames);
			b->d_frames = nullptr; b->cap_frames = 0;
			GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * nFrames));
			b->cap_frames = nFrames;
		}
		GTTS_CUDA(cudaMemcpyAsync(b->d_frames, h_frames, sizeof(float) * kNumParams * nFrames, cudaMemcpyHostToDevice, b->stream));
		dFrames = b->d_frames;
	}
	// The float32 audio stays in device memory: the scale of an utterance needs all of it.  Only the 16-bit payload
	// (half the bytes of the float32 path) and, if asked for, the scales travel to the host.
	if (nOut > b->cap_out) {
		if (b->d_out) cudaFree(b->d_out);
		b->d_out = nullptr; b->cap_out = 0;
		GTTS_CUDA(cudaMalloc(&b->d_out, sizeof(float) * nOut));
		b->cap_out = nOut;
	}
	if (nOut > b->cap_pcm) {
		if (b->d_pcm) cudaFree(b->d_pcm);
		b->d_pcm = nullptr; b->cap_pcm = 0;
		GTTS_CUDA(cudaMalloc(&b->d_pcm, sizeof(short) * nOut));
		b->cap_pcm = nOut;
	}
	if (h_scale && !b->d_scale) GTTS_CUDA(cudaMalloc(&b->d_scale, sizeof(float) * nUtt));
	// The output stage and the payload copy run on a stream of the highest priority: when the synthesis kernel of
	// this batch ends, they get the SMs before the (already queued) synthesis kernel of another batch does -- its
	// persistent CTAs fill the register file and would otherwise keep the output stage, and with it the copy, waiting
	// until they have finished.
	if (!b->stream_out) {
		int lo = 0, hi = 0;
		GTTS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		GTTS_CUDA(cudaStreamCreateWithPriority(&b->stream_out, cudaStreamNonBlocking, hi));
		GTTS_CUDA(cudaEventCreateWithFlags(&b->ev_synth, cudaEventDisableTiming));
	}
	// ... and a synthesis kernel is ordered behind the output stage (not the copy) of the batch submitted before it:
	// its CTAs become resident the moment the previous synthesis kernel's CTAs exit, i.e. before that batch's output
	// stage is even runnable, and priorities do not preempt resident CTAs.
	std::lock_guard<std::mutex> order(b->h->order_lock);
	if (b->h->last_output_stage) GTTS_CUDA(cudaStreamWaitEvent(b->stream, b->h->last_output_stage, 0));
	else GTTS_CUDA(cudaEventCreateWithFlags(&b->h->last_output_stage, cudaEventDisableTiming));
	int rc = launchBatch(b, dFrames, b->d_out, b->stream);
	if (rc != GTTS_OK) return rc;
	GTTS_CUDA(cudaEventRecord(b->ev_synth, b->stream));
	GTTS_CUDA(cudaStreamWaitEvent(b->stream_out, b->ev_synth, 0));
	rc = launchPcm16(b, b->d_out, b->d_pcm, h_scale ? b->d_scale : nullptr, b->stream_out);
	if (rc != GTTS_OK) return rc;
	GTTS_CUDA(cudaEventRecord(b->h->last_output_stage, b->stream_out));
	GTTS_CUDA(cudaMemcpyAsync(h_pcm, b->d_pcm, sizeof(short) * nOut, cudaMemcpyDeviceToHost, b->stream_out));
	if (h_scale) GTTS_CUDA(cudaMemcpyAsync(h_scale, b->d_scale, sizeof(float) * nUtt, cudaMemcpyDeviceToHost, b->stream_out));
	return GTTS_OK;
}

int gtts_batch_last_launches(const gtts_batch* b, int32_t* n_out)
{
	if (!b || !n_out) return fail(GTTS_ERR_INVALID, "null argument");
	*n_out = b->last_launches;
	return GTTS_OK;
}

const char* gtts_batch_last_kernel(const gtts_batch* b) { return b ? b->last_kernel : "none"; }

void gtts_batch_free(gtts_batch* b)
{
	if (!b) return;
	cudaSetDevice(b->h->device);
	if (b->stream_out) { cudaStreamSynchronize(b->stream_out); cudaStreamDestroy(b->stream_out); }
	if (b->ev_synth) cudaEventDestroy(b->ev_synth);
	if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
	cudaFree(b->d_voices); cudaFree(b->d_utts); cudaFree(b->d_order); cudaFree(b->d_order_wide); cudaFree(b->d_queue);
	cudaFree(b->d_states); cudaFree(b->d_frames); cudaFree(b->d_out); cudaFree(b->d_tables);
	cudaFree(b->d_pcm); cudaFree(b->d_peak); cudaFree(b->d_scale); cudaFree(b->d_utts_local);
	delete b;
}

int gtts_batch_synthesize(gtts_handle* h, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const float* frames,
			const int64_t* frame_offsets, int64_t n_utt, float* out, int64_t out_capacity,
			int64_t* out_offsets)
{
	gtts_batch* b = nullptr;
	int rc = gtts_batch_prepare(h, voices, n_voices, voice_index, control_rate, nullptr, frame_offsets, n_utt, &b);
	if (rc != GTTS_OK) return rc;
	if (out_offsets) std::copy(b->plan.out_offsets.begin(), b->plan.out_offsets.end(), out_offsets);
	if (b->plan.out_offsets.back() > out_capacity) {
		gtts_batch_free(b);
		return fail(GTTS_ERR_INVALID, "output buffer too small");
	}
	rc = gtts_batch_run_host(b, frames, out);
	gtts_batch_free(b);
	return rc;
}

// ---- streaming ----------------------------------------------------------------------------------------

namespace { void streamFreeFast(gtts_stream* s); }

int gtts_stream_open(gtts_handle* h, const gtts_voice_config* voice, double control_rate,
			int32_t steps_override, gtts_stream** stream_out)
{
	if (!h || !voice || !stream_out) return fail(GTTS_ERR_INVALID, "null argument");
	*stream_out = nullptr;
	gtts_stream* s = new (std::nothrow) gtts_stream;
	if (!s) return fail(GTTS_ERR_NOMEM, "out of memory");
	s->h = h;
	s->voice = *voice;
	if (voice->tube_model != 0) { delete s; return fail(GTTS_ERR_UNSUPPORTED, "streaming is implemented for models 0 / 2 only (tube_model 0)"); }
	// a one-utterance batch whose descriptor is rewritten for every chunk
	const int64_t fo[2] = {0, 0};
	const int32_t so[1] = {steps_override};
	int rc = prepareBatch(h, voice, 1, nullptr, control_rate, steps_override > 0 ? so : nullptr, fo, 1, true, &s->batch);
	if (rc != GTTS_OK) { delete s; return rc; }
	s->vdev = s->batch->plan.voices[0];
	s->steps = s->batch->plan.utts[0].steps;
	// control periods of at least one block run on the pipelined kernel (whole blocks per chunk), shorter ones on the
	// general kernel (whole control periods per chunk); GTTS_KERNEL=v0 forces the latter
	s->fast = s->steps >= kBlock;
	if (const char* env = std::getenv("GTTS_KERNEL")) { if (std::strcmp(env, "v0") == 0 || std::strcmp(env, "v1") == 0) s->fast = false; }
	// the state is cleared on the stream every chunk of this utterance runs on (stream-ordered before the first kernel)
	cudaError_t e;
	if (s->fast) {
		std::vector<double> table(kTableLen);
		buildWavetable(s->vdev, table.data());
		e = cudaMalloc(&s->d_state, sizeof(UttStateV2));
		if (e == cudaSuccess) e = cudaMemsetAsync(s->d_state, 0, sizeof(UttStateV2), s->batch->stream);
		if (e == cudaSuccess && !s->batch->d_tables) e = cudaMalloc(&s->batch->d_tables, sizeof(double) * kTableLen);
		if (e == cudaSuccess) e = cudaMemcpyAsync(s->batch->d_tables, table.data(), sizeof(double) * kTableLen, cudaMemcpyHostToDevice, s->batch->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(s->batch->stream);
	} else {
		e = cudaMalloc(&s->batch->d_states, sizeof(UttState));
		if (e == cudaSuccess) e = cudaMemsetAsync(s->batch->d_states, 0, sizeof(UttState), s->batch->stream);
	}
	if (e != cudaSuccess) { streamFreeFast(s); gtts_batch_free(s->batch); delete s; return failCuda(e, "gtts_stream_open"); }
	*stream_out = s;
	return GTTS_OK;
}

namespace {

// Runs `periods` control periods over the first `periods` (+1 if lookahead) pending frames.
void streamFreeFast(gtts_stream* s)
{
	if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
	if (s->graph) cudaGraphDestroy(s->graph);
	s->graph_exec = nullptr; s->graph = nullptr;
	cudaFree(s->d_state); s->d_state = nullptr;
	cudaFreeHost(s->m_frames); cudaFreeHost(s->m_utt); cudaFreeHost(s->m_out);
	s->m_frames = nullptr; s->m_utt = nullptr; s->m_out = nullptr;
	s->cap_frames = s->cap_out = 0;
}

// (Re)captures the per-chunk work -- work-counter reset + kernel on the staging buffers -- into a graph.
int streamBuildGraph(gtts_stream* s)
{
	gtts_batch* b = s->batch;
	if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
	if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
	v2::KernelParamsV2 Q;
	Q.voices = b->d_voices;
	Q.tables = b->d_tables;
	GTTS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(const_cast<UttDesc**>(&Q.utts)), s->m_utt, 0));
	Q.order = b->d_order;
	GTTS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(const_cast<float**>(&Q.frames)), s->m_frames, 0));
	GTTS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&Q.out), s->m_out, 0));
	Q.src_tab = s->h->d_src_tab;
	Q.queue = b->d_queue;
	Q.n_utt = 1;
	Q.states = s->d_state;
	Q.prof = nullptr;
	Q.debug_skip = 0;
	GTTS_CUDA(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
	cudaError_t e = cudaMemsetAsync(b->d_queue, 0, 4 * sizeof(int32_t), b->stream);
	if (e == cudaSuccess) {
		v2::tube_kernel_v2_stream<<<1, v2::kThreads, v2::smem_bytes(), b->stream>>>(Q);
		e = cudaGetLastError();
	}
	cudaError_t e2 = cudaStreamEndCapture(b->stream, &s->graph);
	if (e != cudaSuccess) return failCuda(e, "stream graph capture");
	GTTS_CUDA(e2);
	GTTS_CUDA(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
	return GTTS_OK;
}

// Fast path: synthesises `nSamples` internal samples (whole blocks unless `flush`) from the frame window in
// s->pending, which starts at control period s->period0.
int streamChunkFast(gtts_stream* s, int64_t nSamples, bool flush, float* out, int64_t cap, int64_t* written)
{
	gtts_batch* b = s->batch;
	const int64_t nAvail = static_cast<int64_t>(s->pending.size()) / kNumParams;
	const int64_t nInAfter = s->n_in_done + nSamples;
	const int64_t kAfter = streamOutputsAfter(s->vdev, nInAfter, flush);
	const int64_t produced = kAfter - s->n_out_done;
	if (produced > cap) return fail(GTTS_ERR_INVALID, "stream output buffer too small");
	GTTS_CUDA(cudaSetDevice(s->h->device));
	bool rebuild = s->graph_exec == nullptr;
	if (nAvail > s->cap_frames) {
		cudaFreeHost(s->m_frames); s->m_frames = nullptr; s->cap_frames = 0;
		const int64_t want = std::max<int64_t>(2 * nAvail, 256);
		GTTS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->m_frames), sizeof(float) * kNumParams * want, cudaHostAllocMapped));
		s->cap_frames = want;
		rebuild = true;
	}
	if (produced + 64 > s->cap_out) {
		cudaFreeHost(s->m_out); s->m_out = nullptr; s->cap_out = 0;
		const int64_t want = std::max<int64_t>(2 * (produced + 64), 16384);
		GTTS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->m_out), sizeof(float) * want, cudaHostAllocMapped));
		s->cap_out = want;
		rebuild = true;
	}
	if (!s->m_utt) GTTS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->m_utt), sizeof(UttDesc), cudaHostAllocMapped));
	if (rebuild) { const int rc = streamBuildGraph(s); if (rc != GTTS_OK) return rc; }
	if (nSamples == 0 && !flush) { if (written) *written = 0; return GTTS_OK; }
	std::memcpy(s->m_frames, s->pending.data(), sizeof(float) * kNumParams * nAvail);
	*s->m_utt = streamChunkDesc(b->plan.utts[0], nAvail, nSamples, s->n_in_done, s->n_out_done, kAfter, flush);
	GTTS_CUDA(cudaGraphLaunch(s->graph_exec, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	if (produced > 0) std::memcpy(out, s->m_out, sizeof(float) * produced);
	s->n_in_done = nInAfter;
	s->n_out_done = kAfter;
	if (written) *written = produced;
	return GTTS_OK;
}

int streamChunk(gtts_stream* s, int64_t periods, bool lookahead, bool flush, float* out, int64_t cap, int64_t* written)
{
	gtts_batch* b = s->batch;
	UttDesc& d = b->plan.utts[0];
	const int64_t nInAfter = s->n_in_done + periods * s->steps;
	int64_t kAfter;
	if (flush) {
		kAfter = outputLength(s->vdev, nInAfter);
	} else {
		kAfter = static_cast<int64_t>(((static_cast<unsigned __int128>(nInAfter) << 16) + s->vdev.src_inc - 1) / s->vdev.src_inc);
		// nothing is final before the first input exists
		if (nInAfter == 0) kAfter = 0;
	}
	const int64_t produced = kAfter - s->n_out_done;
	if (produced > cap) return fail(GTTS_ERR_INVALID, "stream output buffer too small");
	d.frame_begin = 0;
	d.n_frames = periods;
	d.n_internal = periods * s->steps;
	d.out_begin = -s->n_out_done;            // the kernel indexes outputs absolutely
	d.n_out = flush ? kAfter : std::numeric_limits<int64_t>::max() / 4;
	d.flags = 1 | (flush ? 0 : 2) | (lookahead ? 4 : 0);
	d.state_index = 0;
	const int64_t nFrames = periods + (lookahead ? 1 : 0);
	GTTS_CUDA(cudaSetDevice(s->h->device));
	if (!b->stream) GTTS_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	if (nFrames > b->cap_frames) {
		if (b->d_frames) cudaFree(b->d_frames);
		b->d_frames = nullptr; b->cap_frames = 0;
		const int64_t want = std::max<int64_t>(nFrames, 64);
		GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * want));
		b->cap_frames = want;
	}
	if (produced > b->cap_out) {
		if (b->d_out) cudaFree(b->d_out);
		b->d_out = nullptr; b->cap_out = 0;
		const int64_t want = std::max<int64_t>(produced, 16384);
		GTTS_CUDA(cudaMalloc(&b->d_out, sizeof(float) * want));
		b->cap_out = want;
	}
	GTTS_CUDA(cudaMemcpyAsync(b->d_utts, &d, sizeof d, cudaMemcpyHostToDevice, b->stream));
	if (nFrames > 0) GTTS_CUDA(cudaMemcpyAsync(b->d_frames, s->pending.data(), sizeof(float) * kNumParams * nFrames, cudaMemcpyHostToDevice, b->stream));
	const int rc = launchBatch(b, b->d_frames, b->d_out, b->stream);
	if (rc != GTTS_OK) return rc;
	if (produced > 0) GTTS_CUDA(cudaMemcpyAsync(out, b->d_out, sizeof(float) * produced, cudaMemcpyDeviceToHost, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	s->n_in_done = nInAfter;
	s->n_out_done = kAfter;
	if (written) *written = produced;
	return GTTS_OK;
}

} // namespace

int gtts_stream_push_frames(gtts_stream* s, const float* frames, int64_t n_frames,
			float* out, int64_t out_capacity, int64_t* n_written)
{
	if (!s || (n_frames > 0 && !frames)) return fail(GTTS_ERR_INVALID, "null argument");
	if (s->finished) return fail(GTTS_ERR_INVALID, "stream already finished; call gtts_stream_reset");
	if (n_written) *n_written = 0;
	if (n_frames <= 0) return GTTS_OK;
	// A failed push leaves the stream as it was (the frames are not queued), so the caller may retry, e.g. with a
	// larger output buffer.
	const size_t before = s->pending.size();
	try {
		s->pending.insert(s->pending.end(), frames, frames + n_frames * kNumParams);
	} catch (const std::bad_alloc&) {
		s->pending.resize(before);
		return fail(GTTS_ERR_NOMEM, "out of host memory while queueing frames");
	}
	const int64_t have = static_cast<int64_t>(s->pending.size()) / kNumParams;
	if (have < 2) return GTTS_OK;
	if (s->fast) {
		// the control periods whose end frame is known, in whole blocks; at least one sample is always left for finish()
		const int64_t nSamples = streamSamplesReady(s->period0, have, s->steps, s->n_in_done, kBlock);
		if (nSamples == 0) return GTTS_OK;
		const int rc = streamChunkFast(s, nSamples, false, out, out_capacity, n_written);
		if (rc != GTTS_OK) { s->pending.resize(before); return rc; }
		const int64_t p0 = s->n_in_done / s->steps;                // the period the next chunk starts in
		s->pending.erase(s->pending.begin(), s->pending.begin() + (p0 - s->period0) * kNumParams);
		s->period0 = p0;
		return GTTS_OK;
	}
	const int rc = streamChunk(s, have - 1, true, false, out, out_capacity, n_written);
	if (rc != GTTS_OK) { s->pending.resize(before); return rc; }
	s->pending.erase(s->pending.begin(), s->pending.begin() + (have - 1) * kNumParams);
	return GTTS_OK;
}

int gtts_stream_finish(gtts_stream* s, float* out, int64_t out_capacity, int64_t* n_written)
{
	if (!s) return fail(GTTS_ERR_INVALID, "null argument");
	if (s->finished) return fail(GTTS_ERR_INVALID, "stream already finished; call gtts_stream_reset");
	if (n_written) *n_written = 0;
	const int64_t have = static_cast<int64_t>(s->pending.size()) / kNumParams;
	int rc;
	if (s->fast) {
		// everything that is left, the last control period (the duplicated last frame, Controller.cpp:283) included
		const int64_t total = (s->period0 + have) * s->steps;
		rc = streamChunkFast(s, have > 0 ? total - s->n_in_done : 0, true, out, out_capacity, n_written);
	} else {
		rc = streamChunk(s, have, false, true, out, out_capacity, n_written);
	}
	if (rc != GTTS_OK) return rc;
	s->pending.clear();
	s->finished = true;
	return GTTS_OK;
}

int gtts_stream_reset(gtts_stream* s)
{
	if (!s) return fail(GTTS_ERR_INVALID, "null argument");
	GTTS_CUDA(cudaSetDevice(s->h->device));
	if (s->fast) GTTS_CUDA(cudaMemsetAsync(s->d_state, 0, sizeof(UttStateV2), s->batch->stream));
	else GTTS_CUDA(cudaMemsetAsync(s->batch->d_states, 0, sizeof(UttState), s->batch->stream));
	s->pending.clear();
	s->period0 = 0;
	s->n_in_done = 0;
	s->n_out_done = 0;
	s->finished = false;
	return GTTS_OK;
}

void gtts_stream_close(gtts_stream* s)
{
	if (!s) return;
	cudaSetDevice(s->h->device);
	if (s->batch && s->batch->stream) cudaStreamSynchronize(s->batch->stream);
	streamFreeFast(s);
	gtts_batch_free(s->batch);
	delete s;
}

// ---- one batch over several GPUs (BASELINE config 4) ------------------------------------------
// Utterances are independent, so the batch is partitioned by utterance (gtts_shard_plan on internal + output samples),
// every GPU gets its own gtts_batch over ITS utterances -- whose frames and audio stay where they are in the caller's
// packed arrays -- and gtts_multi_batch_run_host drives the GPUs from one host thread each.  No collective, no
// inter-GPU traffic: the result is, utterance by utterance, what one GPU produces.

struct gtts_multi {
	std::vector<gtts_handle*> handles;
};

struct gtts_multi_batch {
	gtts_multi* m = nullptr;
	BatchPlan plan;                       // the whole batch: layout, lengths
	std::vector<gtts_batch*> shards;      // one per GPU (nullptr: no utterance)
	std::vector<int32_t> shard_of;
	std::string error;
};

int gtts_multi_create(const int32_t* devices, int32_t n_devices, gtts_multi** multi_out)
{
	if (!devices || n_devices <= 0 || !multi_out) return fail(GTTS_ERR_INVALID, "bad device list");
	*multi_out = nullptr;
	try {
		gtts_multi* m = new gtts_multi;
		for (int32_t i = 0; i < n_devices; ++i) {
			gtts_handle* h = nullptr;
			const int rc = gtts_create(devices[i], &h);
			if (rc != GTTS_OK) {
				for (gtts_handle* g : m->handles) gtts_destroy(g);
				delete m;
				return rc;
			}
			m->handles.push_back(h);
		}
		*multi_out = m;
		return GTTS_OK;
	} catch (const std::bad_alloc&) {
		return fail(GTTS_ERR_NOMEM, "out of host memory in gtts_multi_create");
	}
}

void gtts_multi_destroy(gtts_multi* m)
{
	if (!m) return;
	for (gtts_handle* h : m->handles) gtts_destroy(h);
	delete m;
}

int32_t gtts_multi_device_count(const gtts_multi* m) { return m ? static_cast<int32_t>(m->handles.size()) : 0; }

void gtts_multi_batch_free(gtts_multi_batch* b)
{
	if (!b) return;
	for (gtts_batch* s : b->shards) gtts_batch_free(s);
	delete b;
}

int gtts_multi_batch_prepare(gtts_multi* m, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts_multi_batch** batch_out)
{
	if (!m || !batch_out) return fail(GTTS_ERR_INVALID, "null argument");
	*batch_out = nullptr;
	if (n_utt > std::numeric_limits<int32_t>::max()) return fail(GTTS_ERR_INVALID, "too many utterances");
	gtts_multi_batch* b = nullptr;
	try {
		b = new gtts_multi_batch;
		b->m = m;
		int err = GTTS_OK;
		const std::string msg = planBatch(voices, n_voices, voice_index, control_rate, steps_override, frame_offsets, n_utt, b->plan, &err);
		if (err != GTTS_OK) { delete b; return fail(err, msg); }
		const int nDev = static_cast<int>(m->handles.size());
		std::vector<int64_t> cost(static_cast<size_t>(n_utt));
		for (int64_t u = 0; u < n_utt; ++u) cost[u] = b->plan.utts[u].n_internal + b->plan.utts[u].n_out + 1;
		b->shard_of.assign(static_cast<size_t>(n_utt), 0);
		if (n_utt > 0) shardPlan(cost.data(), n_utt, nDev, b->shard_of.data());
		b->shards.assign(nDev, nullptr);
		bool uniform = n_utt > 1;
		for (int64_t u = 1; u < n_utt && uniform; ++u) {
			const UttDesc& d = b->plan.utts[u], & first = b->plan.utts[0];
			uniform = d.voice == first.voice && d.n_frames == first.n_frames && d.steps == first.steps;
		}
		for (int g = 0; g < nDev; ++g) {
			gtts_batch* sb = new gtts_batch;
			sb->h = m->handles[g];
			sb->kernel_hint = uniform ? 1 : 0;
			sb->plan.voices = b->plan.voices;
			for (int64_t u = 0; u < n_utt; ++u) if (b->shard_of[u] == g) sb->plan.utts.push_back(b->plan.utts[u]);   // global frame_begin / out_begin kept
			sb->plan.out_offsets = b->plan.out_offsets;              // sizes of the caller's buffers
			sb->plan.n_frames_total = b->plan.n_frames_total;
			sb->plan.order.resize(sb->plan.utts.size());
			for (size_t i = 0; i < sb->plan.order.size(); ++i) sb->plan.order[i] = static_cast<int32_t>(i);
			std::stable_sort(sb->plan.order.begin(), sb->plan.order.end(), [&](int32_t x, int32_t y) {
				return sb->plan.utts[x].n_internal > sb->plan.utts[y].n_internal;
			});
			b->shards[g] = sb;
			int rc = uploadPlan(sb);
			if (rc != GTTS_OK) { gtts_multi_batch_free(b); return rc; }
			// the same utterances laid out compactly: where the shard keeps its float32 audio on the device (PCM path)
			std::vector<UttDesc> local = sb->plan.utts;
			int64_t off = 0;
			for (UttDesc& d : local) { d.out_begin = off; off = (off + d.n_out + 63) & ~int64_t(63); }
			sb->local_out_total = off;
			if (!local.empty()) {
				cudaError_t ce = cudaMalloc(&sb->d_utts_local, sizeof(UttDesc) * local.size());
				if (ce == cudaSuccess) ce = cudaMemcpyAsync(sb->d_utts_local, local.data(), sizeof(UttDesc) * local.size(), cudaMemcpyHostToDevice, sb->stream);
				if (ce == cudaSuccess) ce = cudaStreamSynchronize(sb->stream);
				if (ce != cudaSuccess) { gtts_multi_batch_free(b); return failCuda(ce, "shard layout"); }
			}
		}
	} catch (const std::bad_alloc&) {
		gtts_multi_batch_free(b);
		return fail(GTTS_ERR_NOMEM, "out of host memory while planning the batch");
	} catch (const std::exception& e) {
		gtts_multi_batch_free(b);
		return fail(GTTS_ERR_INVALID, e.what());
	}
	*batch_out = b;
	return GTTS_OK;
}

int gtts_multi_batch_layout(const gtts_multi_batch* b, int64_t* out_offsets, int64_t* n_out, int32_t* shard_of)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (out_offsets) std::copy(b->plan.out_offsets.begin(), b->plan.out_offsets.end(), out_offsets);
	if (n_out) for (size_t u = 0; u < b->plan.utts.size(); ++u) n_out[u] = b->plan.utts[u].n_out;
	if (shard_of) std::copy(b->shard_of.begin(), b->shard_of.end(), shard_of);
	return GTTS_OK;
}

namespace {
// pins a pageable host range for the duration of a call (the shards read frames / write audio in place)
struct ScopedPin {
	void* p = nullptr;
	bool mine = false;
	int pin(const void* ptr, size_t bytes)
	{
		if (!ptr || bytes == 0) return GTTS_OK;
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost) return GTTS_OK;
		cudaGetLastError();
		const cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
		if (e != cudaSuccess) return failCuda(e, "cudaHostRegister");
		p = const_cast<void*>(ptr);
		mine = true;
		return GTTS_OK;
	}
	~ScopedPin() { if (mine) cudaHostUnregister(p); }
};
} // namespace

// One shard's PCM path: synthesis into a compact device buffer, then the output stage stores the payload straight
// into the caller's (pinned, mapped) buffer at the utterances' own places.
static int shardRunPcm16(gtts_batch* sb, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	GTTS_CUDA(cudaSetDevice(sb->h->device));
	cudaPointerAttributes fa, pa;
	if (cudaPointerGetAttributes(&fa, h_frames) != cudaSuccess || fa.type != cudaMemoryTypeHost || !fa.devicePointer ||
	    cudaPointerGetAttributes(&pa, h_pcm) != cudaSuccess || pa.type != cudaMemoryTypeHost || !pa.devicePointer) {
		cudaGetLastError();
		return fail(GTTS_ERR_CUDA, "host buffers are not mapped into the device address space");
	}
	const int64_t nUtt = static_cast<int64_t>(sb->plan.utts.size());
	if (sb->local_out_total > sb->cap_out) {
		if (sb->d_out) cudaFree(sb->d_out);
		sb->d_out = nullptr; sb->cap_out = 0;
		GTTS_CUDA(cudaMalloc(&sb->d_out, sizeof(float) * sb->local_out_total));
		sb->cap_out = sb->local_out_total;
	}
	if (h_scale && !sb->d_scale) GTTS_CUDA(cudaMalloc(&sb->d_scale, sizeof(float) * nUtt));
	int rc = launchBatch(sb, static_cast<const float*>(fa.devicePointer), sb->d_out, sb->stream, sb->d_utts_local);
	if (rc != GTTS_OK) return rc;
	rc = launchPcm16(sb, sb->d_out, static_cast<short*>(pa.devicePointer), h_scale ? sb->d_scale : nullptr, sb->stream, sb->d_utts_local);
	if (rc != GTTS_OK) return rc;
	if (h_scale) GTTS_CUDA(cudaMemcpyAsync(h_scale, sb->d_scale, sizeof(float) * nUtt, cudaMemcpyDeviceToHost, sb->stream));
	GTTS_CUDA(cudaStreamSynchronize(sb->stream));
	return GTTS_OK;
}

// pcm == nullptr: float32 audio into h_out; else the 16-bit payload into pcm (and the scales into h_scale if not null)
static int multiRun(gtts_multi_batch* b, const float* h_frames, float* h_out, int16_t* h_pcm, float* h_scale)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nFrames = b->plan.n_frames_total;
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	if ((nFrames > 0 && !h_frames) || (nOut > 0 && !h_out && !h_pcm)) return fail(GTTS_ERR_INVALID, "null host buffer");
	ScopedPin pinFrames, pinOut;
	int rc = pinFrames.pin(h_frames, sizeof(float) * kNumParams * static_cast<size_t>(nFrames));
	if (rc != GTTS_OK) return rc;
	if (h_pcm) rc = pinOut.pin(h_pcm, sizeof(int16_t) * static_cast<size_t>(nOut));
	else rc = pinOut.pin(h_out, sizeof(float) * static_cast<size_t>(nOut));
	if (rc != GTTS_OK) return rc;
	const size_t nDev = b->shards.size();
	std::vector<int> codes(nDev, GTTS_OK);
	std::vector<std::string> texts(nDev);
	std::vector<std::thread> pool;
	// per-utterance scales come back in shard order: gather them into the caller's order afterwards
	std::vector<std::vector<float>> shardScale(nDev);
	for (size_t g = 0; g < nDev; ++g) {
		gtts_batch* sb = b->shards[g];
		if (!sb || sb->plan.utts.empty()) continue;
		if (h_pcm && h_scale) shardScale[g].resize(sb->plan.utts.size());
		pool.emplace_back([=, &codes, &texts, &shardScale]() {
			int r;
			if (h_pcm) r = shardRunPcm16(sb, h_frames, h_pcm, shardScale[g].empty() ? nullptr : shardScale[g].data());
			else r = gtts_batch_run_host(sb, h_frames, h_out);
			codes[g] = r;
			if (r != GTTS_OK) texts[g] = gtts_last_error();
		});
	}
	for (std::thread& t : pool) t.join();
	for (size_t g = 0; g < nDev; ++g) if (codes[g] != GTTS_OK) return fail(codes[g], "GPU " + std::to_string(g) + ": " + texts[g]);
	if (h_pcm && h_scale) {
		std::vector<size_t> next(nDev, 0);
		for (size_t u = 0; u < b->shard_of.size(); ++u) {
			const size_t g = static_cast<size_t>(b->shard_of[u]);
			h_scale[u] = shardScale[g][next[g]++];
		}
	}
	return GTTS_OK;
}

int gtts_multi_batch_run_host(gtts_multi_batch* b, const float* h_frames, float* h_out)
{
	try {
		return multiRun(b, h_frames, h_out, nullptr, nullptr);
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

int gtts_multi_batch_run_host_pcm16(gtts_multi_batch* b, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	try {
		return multiRun(b, h_frames, nullptr, h_pcm, h_scale);
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

} // extern "C"

// ---- model 5 (VocalTractModel5<double, 1>): tube5_kernel.cuh, one warp per utterance ---------------------------------

struct gtts5_batch {
	gtts_handle* h = nullptr;
	m5::BatchPlan5 plan;
	m5::Voice5Dev* d_voices = nullptr;
	UttDesc* d_utts = nullptr;
	int32_t* d_order = nullptr;
	int32_t* d_queue = nullptr;
	float* d_frames = nullptr;
	float* d_out = nullptr;
	short* d_pcm = nullptr;
	unsigned* d_peak = nullptr;
	float* d_scale = nullptr;
	cudaStream_t stream = nullptr;
};

extern "C" {

int gtts5_voice_internal_rate(const gtts_voice5_config* voice, double* fs_out)
{
	if (!voice || !fs_out) return fail(GTTS_ERR_INVALID, "null argument");
	m5::Voice5Dev v;
	bool unsupported = false;
	const char* e = m5::deriveVoice5(*voice, v, &unsupported);
	if (e && !unsupported) return fail(GTTS_ERR_INVALID, e);
	*fs_out = v.fs;
	return GTTS_OK;
}

int gtts5_output_length(const gtts_voice5_config* voice, double control_rate, int32_t steps, int64_t n_frames,
			int32_t* steps_out, int64_t* n_internal_out, int64_t* n_output_out)
{
	if (!voice || n_frames < 0) return fail(GTTS_ERR_INVALID, "bad argument");
	m5::Voice5Dev v;
	bool unsupported = false;
	const char* e = m5::deriveVoice5(*voice, v, &unsupported);
	if (e) return fail(unsupported ? GTTS_ERR_UNSUPPORTED : GTTS_ERR_INVALID, e);
	if (steps <= 0) {
		if (!(control_rate > 0.0)) return fail(GTTS_ERR_INVALID, "control_rate must be positive");
		steps = m5::controlSteps5(v.fs, control_rate);
	}
	if (steps <= 0) return fail(GTTS_ERR_INVALID, "control steps must be positive");
	if (steps_out) *steps_out = steps;
	if (n_internal_out) *n_internal_out = n_frames * steps;
	if (n_output_out) *n_output_out = m5::outputLength5(v, n_frames * steps);
	return GTTS_OK;
}

void gtts5_batch_free(gtts5_batch* b)
{
	if (!b) return;
	cudaSetDevice(b->h->device);
	cudaFree(b->d_voices); cudaFree(b->d_utts); cudaFree(b->d_order); cudaFree(b->d_queue);
	cudaFree(b->d_frames); cudaFree(b->d_out); cudaFree(b->d_pcm); cudaFree(b->d_peak); cudaFree(b->d_scale);
	if (b->stream) cudaStreamDestroy(b->stream);
	delete b;
}

static int prepare5(gtts5_batch* b)
{
	m5::BatchPlan5& p = b->plan;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	GTTS_CUDA(cudaMalloc(&b->d_voices, sizeof(m5::Voice5Dev) * std::max<size_t>(p.voices.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_utts, sizeof(UttDesc) * std::max<size_t>(p.utts.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_order, sizeof(int32_t) * std::max<size_t>(p.order.size(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_queue, sizeof(int32_t)));
	GTTS_CUDA(cudaMemcpyAsync(b->d_voices, p.voices.data(), sizeof(m5::Voice5Dev) * p.voices.size(), cudaMemcpyHostToDevice, b->stream));
	if (!p.utts.empty()) {
		GTTS_CUDA(cudaMemcpyAsync(b->d_utts, p.utts.data(), sizeof(UttDesc) * p.utts.size(), cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_order, p.order.data(), sizeof(int32_t) * p.order.size(), cudaMemcpyHostToDevice, b->stream));
	}
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}

int gtts5_batch_prepare(gtts_handle* h, const gtts_voice5_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts5_batch** batch_out)
{
	if (!h || !batch_out) return fail(GTTS_ERR_INVALID, "null handle / output pointer");
	*batch_out = nullptr;
	try {
		if (n_utt > std::numeric_limits<int32_t>::max()) return fail(GTTS_ERR_INVALID, "too many utterances");
		gtts5_batch* b = new gtts5_batch;
		b->h = h;
		int err = GTTS_OK;
		const std::string text = m5::planBatch5(voices, n_voices, voice_index, control_rate, steps_override, frame_offsets, n_utt, b->plan, &err);
		if (err != GTTS_OK) { delete b; return fail(err, text); }
		const int rc = prepare5(b);
		if (rc != GTTS_OK) { gtts5_batch_free(b); return rc; }
		*batch_out = b;
		return GTTS_OK;
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

int gtts5_batch_layout(const gtts5_batch* b, int64_t* out_offsets, int64_t* n_out, int64_t* n_internal)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (out_offsets) std::copy(b->plan.out_offsets.begin(), b->plan.out_offsets.end(), out_offsets);
	for (size_t u = 0; u < b->plan.utts.size(); ++u) {
		if (n_out) n_out[u] = b->plan.utts[u].n_out;
		if (n_internal) n_internal[u] = b->plan.utts[u].n_internal;
	}
	return GTTS_OK;
}

int gtts5_batch_run_device(gtts5_batch* b, const float* d_frames, float* d_out, void* cuda_stream)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int32_t nUtt = static_cast<int32_t>(b->plan.utts.size());
	if (nUtt == 0) return GTTS_OK;
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaMemsetAsync(b->d_queue, 0, sizeof(int32_t), stream));
	m5::KernelParams5 P;
	P.voices = b->d_voices;
	P.utts = b->d_utts;
	P.order = b->d_order;
	P.frames = d_frames;
	P.out = d_out;
	P.src_tab = b->h->d_src_tab;
	P.queue = b->d_queue;
	P.n_utt = nUtt;
	// few utterances: one warp per CTA spreads them over the SMs; many: kWarps5 per SM
	const int64_t ctasWanted = std::max<int64_t>(std::min<int64_t>(nUtt, b->h->sms), (static_cast<int64_t>(nUtt) + kWarps5 - 1) / kWarps5);
	const int grid = static_cast<int>(std::min<int64_t>(ctasWanted, b->h->sms));
	m5::tube5_kernel<kWarps5><<<grid, kWarps5 * 32, m5::smem_bytes(kWarps5), stream>>>(P);
	GTTS_CUDA(cudaGetLastError());
	return GTTS_OK;
}

int gtts5_batch_run_host(gtts5_batch* b, const float* h_frames, float* h_out)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nFrames = b->plan.n_frames_total;
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	if ((nFrames > 0 && !h_frames) || (nOut > 0 && !h_out)) return fail(GTTS_ERR_INVALID, "null host buffer");
	GTTS_CUDA(cudaSetDevice(b->h->device));
	if (!b->d_frames && nFrames > 0) GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * nFrames));
	if (!b->d_out && nOut > 0) GTTS_CUDA(cudaMalloc(&b->d_out, sizeof(float) * nOut));
	if (nFrames > 0) GTTS_CUDA(cudaMemcpyAsync(b->d_frames, h_frames, sizeof(float) * kNumParams * nFrames, cudaMemcpyHostToDevice, b->stream));
	const int rc = gtts5_batch_run_device(b, b->d_frames, b->d_out, b->stream);
	if (rc != GTTS_OK) return rc;
	if (nOut > 0) GTTS_CUDA(cudaMemcpyAsync(h_out, b->d_out, sizeof(float) * nOut, cudaMemcpyDeviceToHost, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}


// The reference's output stage for model 5 (Controller::writeOutputToFile + WAVEFileWriter, as for model 0: see
// gtts_batch_run_device_pcm16): per-utterance scale 0.95f / max|x| and the 16-bit payload, same kernels.
int gtts5_batch_run_device_pcm16(gtts5_batch* b, const float* d_frames, float* d_audio, int16_t* d_pcm, float* d_scale, void* cuda_stream)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int nUtt = static_cast<int>(b->plan.utts.size());
	if (nUtt == 0) return GTTS_OK;
	if (!d_audio || !d_pcm) return fail(GTTS_ERR_INVALID, "null device buffer");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	const int rc = gtts5_batch_run_device(b, d_frames, d_audio, stream);
	if (rc != GTTS_OK) return rc;
	if (!b->d_peak) GTTS_CUDA(cudaMalloc(&b->d_peak, sizeof(unsigned) * nUtt));
	utterance_peak_kernel<<<nUtt, 256, 0, stream>>>(d_audio, b->d_utts, b->d_peak);
	GTTS_CUDA(cudaGetLastError());
	utterance_pcm16_kernel<<<nUtt, 256, 0, stream>>>(d_audio, b->d_utts, b->d_utts, b->d_peak, reinterpret_cast<short*>(d_pcm), d_scale);
	GTTS_CUDA(cudaGetLastError());
	return GTTS_OK;
}

int gtts5_batch_run_host_pcm16(gtts5_batch* b, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nFrames = b->plan.n_frames_total;
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	const size_t nUtt = b->plan.utts.size();
	if ((nFrames > 0 && !h_frames) || (nOut > 0 && !h_pcm)) return fail(GTTS_ERR_INVALID, "null host buffer");
	if (nUtt == 0) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	if (!b->d_frames && nFrames > 0) GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * nFrames));
	if (!b->d_out && nOut > 0) GTTS_CUDA(cudaMalloc(&b->d_out, sizeof(float) * nOut));
	if (!b->d_pcm && nOut > 0) GTTS_CUDA(cudaMalloc(&b->d_pcm, sizeof(short) * nOut));
	if (!b->d_scale) GTTS_CUDA(cudaMalloc(&b->d_scale, sizeof(float) * nUtt));
	if (nFrames > 0) GTTS_CUDA(cudaMemcpyAsync(b->d_frames, h_frames, sizeof(float) * kNumParams * nFrames, cudaMemcpyHostToDevice, b->stream));
	const int rc = gtts5_batch_run_device_pcm16(b, b->d_frames, b->d_out, reinterpret_cast<int16_t*>(b->d_pcm), b->d_scale, b->stream);
	if (rc != GTTS_OK) return rc;
	if (nOut > 0) GTTS_CUDA(cudaMemcpyAsync(h_pcm, b->d_pcm, sizeof(short) * nOut, cudaMemcpyDeviceToHost, b->stream));
	if (h_scale) GTTS_CUDA(cudaMemcpyAsync(h_scale, b->d_scale, sizeof(float) * nUtt, cudaMemcpyDeviceToHost, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}

// ---- model 5 over the GPUs of a box -----------------------------------------------------------------------------------
// Same partition as gtts_multi_batch (by utterance, no inter-GPU traffic).  The model-5 kernel takes device buffers, so
// every GPU's host thread packs its utterances' frames, runs its own gtts5_batch and puts the audio (or the 16-bit
// payload) at the utterances' offsets of the caller's buffer.  One warp synthesises one utterance whatever else is in
// the batch: the result is bit for bit that of one GPU.

struct gtts5_multi_batch {
	gtts_multi* m = nullptr;
	m5::BatchPlan5 plan;                         // the whole batch: layout, lengths
	std::vector<int32_t> shard_of;
	std::vector<gtts5_batch*> shards;            // one per GPU (nullptr: no utterance)
	std::vector<std::vector<int64_t>> members;   // utterances of each shard, ascending
	std::vector<int64_t> frame_offsets;          // the caller's
};

void gtts5_multi_batch_free(gtts5_multi_batch* b)
{
	if (!b) return;
	for (gtts5_batch* sb : b->shards) gtts5_batch_free(sb);
	delete b;
}

int gtts5_multi_batch_prepare(gtts_multi* m, const gtts_voice5_config* voices, int32_t n_voices, const int32_t* voice_index,
			double control_rate, const int32_t* steps_override, const int64_t* frame_offsets, int64_t n_utt,
			gtts5_multi_batch** batch_out)
{
	if (!m || !batch_out) return fail(GTTS_ERR_INVALID, "null multi handle / output pointer");
	*batch_out = nullptr;
	gtts5_multi_batch* b = nullptr;
	try {
		if (n_utt > std::numeric_limits<int32_t>::max()) return fail(GTTS_ERR_INVALID, "too many utterances");
		b = new gtts5_multi_batch;
		b->m = m;
		int err = GTTS_OK;
		const std::string msg = m5::planBatch5(voices, n_voices, voice_index, control_rate, steps_override, frame_offsets, n_utt, b->plan, &err);
		if (err != GTTS_OK) { delete b; return fail(err, msg); }
		b->frame_offsets.assign(frame_offsets, frame_offsets + n_utt + 1);
		const int nDev = static_cast<int>(m->handles.size());
		std::vector<int64_t> cost(static_cast<size_t>(n_utt));
		for (int64_t u = 0; u < n_utt; ++u) cost[u] = b->plan.utts[u].n_internal + b->plan.utts[u].n_out + 1;
		b->shard_of.assign(static_cast<size_t>(n_utt), 0);
		if (n_utt > 0) shardPlan(cost.data(), n_utt, nDev, b->shard_of.data());
		b->shards.assign(nDev, nullptr);
		b->members.assign(nDev, {});
		for (int64_t u = 0; u < n_utt; ++u) b->members[b->shard_of[u]].push_back(u);
		for (int g = 0; g < nDev; ++g) {
			const std::vector<int64_t>& mem = b->members[g];
			if (mem.empty()) continue;
			std::vector<int64_t> fo(mem.size() + 1, 0);
			std::vector<int32_t> vi(mem.size(), 0), so(mem.size(), 0);
			for (size_t i = 0; i < mem.size(); ++i) {
				const int64_t u = mem[i];
				fo[i + 1] = fo[i] + (frame_offsets[u + 1] - frame_offsets[u]);
				vi[i] = voice_index ? voice_index[u] : 0;
				so[i] = steps_override ? steps_override[u] : 0;
			}
			const int rc = gtts5_batch_prepare(m->handles[g], voices, n_voices, vi.data(), control_rate, steps_override ? so.data() : nullptr,
					fo.data(), static_cast<int64_t>(mem.size()), &b->shards[g]);
			if (rc != GTTS_OK) { gtts5_multi_batch_free(b); return rc; }
		}
	} catch (const std::exception& e) {
		gtts5_multi_batch_free(b);
		return fail(GTTS_ERR_NOMEM, e.what());
	}
	*batch_out = b;
	return GTTS_OK;
}

int gtts5_multi_batch_layout(const gtts5_multi_batch* b, int64_t* out_offsets, int64_t* n_out, int32_t* shard_of)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	if (out_offsets) std::copy(b->plan.out_offsets.begin(), b->plan.out_offsets.end(), out_offsets);
	if (n_out) for (size_t u = 0; u < b->plan.utts.size(); ++u) n_out[u] = b->plan.utts[u].n_out;
	if (shard_of) std::copy(b->shard_of.begin(), b->shard_of.end(), shard_of);
	return GTTS_OK;
}

static int multiRun5(gtts5_multi_batch* b, const float* h_frames, float* h_out, int16_t* h_pcm, float* h_scale)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nOut = b->plan.out_offsets.empty() ? 0 : b->plan.out_offsets.back();
	if ((b->plan.n_frames_total > 0 && !h_frames) || (nOut > 0 && !h_out && !h_pcm)) return fail(GTTS_ERR_INVALID, "null host buffer");
	const size_t nDev = b->shards.size();
	std::vector<int> codes(nDev, GTTS_OK);
	std::vector<std::string> texts(nDev);
	std::vector<std::thread> pool;
	for (size_t g = 0; g < nDev; ++g) {
		gtts5_batch* sb = b->shards[g];
		if (!sb) continue;
		pool.emplace_back([=, &codes, &texts]() {
			try {
				const std::vector<int64_t>& mem = b->members[g];
				std::vector<float> frames(static_cast<size_t>(sb->plan.n_frames_total) * kNumParams + 1);
				size_t at = 0;
				for (int64_t u : mem) {
					const size_t n = static_cast<size_t>(b->frame_offsets[u + 1] - b->frame_offsets[u]) * kNumParams;
					std::memcpy(frames.data() + at, h_frames + b->frame_offsets[u] * kNumParams, n * sizeof(float));
					at += n;
				}
				const size_t subOut = static_cast<size_t>(sb->plan.out_offsets.back());
				int r;
				if (h_pcm) {
					std::vector<int16_t> pcm(subOut + 1);
					std::vector<float> scale(mem.size());
					r = gtts5_batch_run_host_pcm16(sb, frames.data(), pcm.data(), scale.data());
					for (size_t i = 0; r == GTTS_OK && i < mem.size(); ++i) {
						const UttDesc& d = b->plan.utts[mem[i]];
						std::memcpy(h_pcm + d.out_begin, pcm.data() + sb->plan.utts[i].out_begin, sizeof(int16_t) * static_cast<size_t>(d.n_out));
						if (h_scale) h_scale[mem[i]] = scale[i];
					}
				} else {
					std::vector<float> out(subOut + 1);
					r = gtts5_batch_run_host(sb, frames.data(), out.data());
					for (size_t i = 0; r == GTTS_OK && i < mem.size(); ++i) {
						const UttDesc& d = b->plan.utts[mem[i]];
						std::memcpy(h_out + d.out_begin, out.data() + sb->plan.utts[i].out_begin, sizeof(float) * static_cast<size_t>(d.n_out));
					}
				}
				codes[g] = r;
				if (r != GTTS_OK) texts[g] = gtts_last_error();
			} catch (const std::exception& e) {
				codes[g] = GTTS_ERR_NOMEM;
				texts[g] = e.what();
			}
		});
	}
	for (std::thread& t : pool) t.join();
	for (size_t g = 0; g < nDev; ++g) if (codes[g] != GTTS_OK) return fail(codes[g], "GPU " + std::to_string(g) + ": " + texts[g]);
	return GTTS_OK;
}

int gtts5_multi_batch_run_host(gtts5_multi_batch* b, const float* h_frames, float* h_out)
{
	try {
		return multiRun5(b, h_frames, h_out, nullptr, nullptr);
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

int gtts5_multi_batch_run_host_pcm16(gtts5_multi_batch* b, const float* h_frames, int16_t* h_pcm, float* h_scale)
{
	if (!h_pcm) return fail(GTTS_ERR_INVALID, "null host buffer");
	try {
		return multiRun5(b, h_frames, nullptr, h_pcm, h_scale);
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

// ---- control-frame generation on the device (events_kernel.cuh) ---------------------------------------------------

struct gtts_events_batch {
	gtts_handle* h = nullptr;
	evt::EventsPlan plan;
	gtts_event_config* d_cfgs = nullptr;
	gtts_event_config* d_cfgs_out = nullptr;   // run_host
	evt::ChunkDesc* d_chunks = nullptr;
	evt::ChainDesc* d_chains = nullptr;
	int32_t* d_order = nullptr;
	int32_t* d_chunk_order = nullptr;
	float* d_drift = nullptr;                  // scratch: the drift generator's output per frame
	int32_t* d_queue = nullptr;
	gtts_event* d_events = nullptr;            // run_host staging
	float* d_frames = nullptr;
	cudaStream_t stream = nullptr;
};

void gtts_events_free(gtts_events_batch* b)
{
	if (!b) return;
	cudaSetDevice(b->h->device);
	cudaFree(b->d_cfgs); cudaFree(b->d_cfgs_out); cudaFree(b->d_chunks); cudaFree(b->d_chains); cudaFree(b->d_order);
	cudaFree(b->d_queue); cudaFree(b->d_events); cudaFree(b->d_frames); cudaFree(b->d_chunk_order); cudaFree(b->d_drift);
	if (b->stream) cudaStreamDestroy(b->stream);
	delete b;
}

int gtts_events_drift_setup(double deviation, double sample_rate, double lowpass_cutoff, gtts_event_config* config)
{
	if (!config) return fail(GTTS_ERR_INVALID, "null config");
	if (!evt::driftSetup(deviation, sample_rate, lowpass_cutoff, *config))
		return fail(GTTS_ERR_INVALID, "drift low-pass cutoff must lie between 1 Hz and 0.48 of the control rate");
	return GTTS_OK;
}

int gtts_events_frame_count(const gtts_event_config* config, const gtts_event* events, int64_t n_events, int64_t* n_frames_out)
{
	if (!config || !n_frames_out || (n_events > 0 && !events) || n_events < 0) return fail(GTTS_ERR_INVALID, "null argument");
	if (config->control_period <= 0) return fail(GTTS_ERR_INVALID, "control_period must be positive");
	*n_frames_out = evt::countFrames(config->control_period, events, n_events);
	return GTTS_OK;
}

static int prepareEvents(gtts_events_batch* b)
{
	const evt::EventsPlan& p = b->plan;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	const size_t nChunks = p.chunks.size(), nChains = p.chains.size();
	GTTS_CUDA(cudaMalloc(&b->d_cfgs, sizeof(gtts_event_config) * std::max<size_t>(nChunks, 1)));
	GTTS_CUDA(cudaMalloc(&b->d_chunks, sizeof(evt::ChunkDesc) * std::max<size_t>(nChunks, 1)));
	GTTS_CUDA(cudaMalloc(&b->d_chains, sizeof(evt::ChainDesc) * std::max<size_t>(nChains, 1)));
	GTTS_CUDA(cudaMalloc(&b->d_order, sizeof(int32_t) * std::max<size_t>(nChains, 1)));
	GTTS_CUDA(cudaMalloc(&b->d_chunk_order, sizeof(int32_t) * std::max<size_t>(nChunks, 1)));
	GTTS_CUDA(cudaMalloc(&b->d_drift, sizeof(float) * std::max<int64_t>(p.frame_offsets.empty() ? 0 : p.frame_offsets.back(), 1)));
	GTTS_CUDA(cudaMalloc(&b->d_queue, 2 * sizeof(int32_t)));
	if (nChunks > 0) {
		GTTS_CUDA(cudaMemcpyAsync(b->d_chunk_order, p.chunk_order.data(), sizeof(int32_t) * nChunks, cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_cfgs, p.cfgs.data(), sizeof(gtts_event_config) * nChunks, cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_chunks, p.chunks.data(), sizeof(evt::ChunkDesc) * nChunks, cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_chains, p.chains.data(), sizeof(evt::ChainDesc) * nChains, cudaMemcpyHostToDevice, b->stream));
		GTTS_CUDA(cudaMemcpyAsync(b->d_order, p.order.data(), sizeof(int32_t) * nChains, cudaMemcpyHostToDevice, b->stream));
	}
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	return GTTS_OK;
}

int gtts_events_prepare(gtts_handle* h, const gtts_event_config* configs, const int32_t* continues_previous,
			const gtts_event* events, const int64_t* event_offsets, int64_t n_chunks, gtts_events_batch** batch_out)
{
	if (!h || !batch_out) return fail(GTTS_ERR_INVALID, "null handle / output pointer");
	*batch_out = nullptr;
	try {
		gtts_events_batch* b = new gtts_events_batch;
		b->h = h;
		int err = GTTS_OK;
		const std::string text = evt::planEvents(configs, continues_previous, events, event_offsets, n_chunks, b->plan, &err);
		if (err != GTTS_OK) { delete b; return fail(err, text); }
		const int rc = prepareEvents(b);
		if (rc != GTTS_OK) { gtts_events_free(b); return rc; }
		*batch_out = b;
		return GTTS_OK;
	} catch (const std::exception& e) {
		return fail(GTTS_ERR_NOMEM, e.what());
	}
}

int gtts_events_layout(const gtts_events_batch* b, int64_t* frame_offsets)
{
	if (!b || !frame_offsets) return fail(GTTS_ERR_INVALID, "null argument");
	std::copy(b->plan.frame_offsets.begin(), b->plan.frame_offsets.end(), frame_offsets);
	return GTTS_OK;
}

int gtts_events_run_device(gtts_events_batch* b, const gtts_event* d_events, float* d_frames, gtts_event_config* d_configs_out,
			void* cuda_stream)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int nChains = static_cast<int>(b->plan.chains.size());
	if (nChains == 0) return GTTS_OK;
	if ((b->plan.n_events_total > 0 && !d_events) || (b->plan.frame_offsets.back() > 0 && !d_frames))
		return fail(GTTS_ERR_INVALID, "null device buffer");
	if (reinterpret_cast<uintptr_t>(d_events) % 8 != 0 || reinterpret_cast<uintptr_t>(d_frames) % 4 != 0)
		return fail(GTTS_ERR_INVALID, "d_events must be 8-byte aligned and d_frames 4-byte aligned");
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	GTTS_CUDA(cudaSetDevice(b->h->device));
	GTTS_CUDA(cudaMemsetAsync(b->d_queue, 0, 2 * sizeof(int32_t), stream));
	evt::EventsParams P;
	P.events = reinterpret_cast<const double*>(d_events);
	P.cfgs = b->d_cfgs;
	P.cfgs_out = d_configs_out;
	P.chunks = b->d_chunks;
	P.chains = b->d_chains;
	P.order = b->d_order;
	P.chunk_order = b->d_chunk_order;
	P.frames = d_frames;
	P.drift = b->d_drift;
	P.queue = b->d_queue;
	P.n_chains = nChains;
	P.n_chunks = static_cast<int>(b->plan.chunks.size());
	// drift pass: one thread per utterance
	evt::events_drift_kernel<<<(nChains + evt::kDriftThreads - 1) / evt::kDriftThreads, evt::kDriftThreads, 0, stream>>>(P);
	GTTS_CUDA(cudaGetLastError());
	// frame pass: one warp per chunk from a queue; CTAs of eight warps, three per SM resident, or four (64 registers) when
	// the batch has more chunks than three per SM keep busy
	const int ctasWanted = (P.n_chunks + evt::kEventsWarps - 1) / evt::kEventsWarps;
	if (ctasWanted > b->h->sms * 3) {
		evt::events_kernel<4><<<std::min(ctasWanted, b->h->sms * 4), evt::kEventsWarps * 32, evt::kEventsSmem, stream>>>(P);
	} else {
		evt::events_kernel<3><<<ctasWanted, evt::kEventsWarps * 32, evt::kEventsSmem, stream>>>(P);
	}
	GTTS_CUDA(cudaGetLastError());
	return GTTS_OK;
}

int gtts_events_run_host(gtts_events_batch* b, const gtts_event* h_events, float* h_frames, gtts_event_config* configs_out)
{
	if (!b) return fail(GTTS_ERR_INVALID, "null batch");
	const int64_t nEvents = b->plan.n_events_total, nFrames = b->plan.frame_offsets.empty() ? 0 : b->plan.frame_offsets.back();
	const size_t nChunks = b->plan.chunks.size();
	if ((nEvents > 0 && !h_events) || (nFrames > 0 && !h_frames)) return fail(GTTS_ERR_INVALID, "null host buffer");
	if (nChunks == 0) return GTTS_OK;
	GTTS_CUDA(cudaSetDevice(b->h->device));
	if (!b->d_events && nEvents > 0) GTTS_CUDA(cudaMalloc(&b->d_events, sizeof(gtts_event) * nEvents));
	if (!b->d_frames && nFrames > 0) GTTS_CUDA(cudaMalloc(&b->d_frames, sizeof(float) * kNumParams * nFrames));
	if (!b->d_cfgs_out) GTTS_CUDA(cudaMalloc(&b->d_cfgs_out, sizeof(gtts_event_config) * nChunks));
	if (nEvents > 0) GTTS_CUDA(cudaMemcpyAsync(b->d_events, h_events, sizeof(gtts_event) * nEvents, cudaMemcpyHostToDevice, b->stream));
	const int rc = gtts_events_run_device(b, b->d_events, b->d_frames, b->d_cfgs_out, b->stream);
	if (rc != GTTS_OK) return rc;
	int32_t flags[2] = {0, 0};
	GTTS_CUDA(cudaMemcpyAsync(flags, b->d_queue, sizeof flags, cudaMemcpyDeviceToHost, b->stream));
	if (nFrames > 0) GTTS_CUDA(cudaMemcpyAsync(h_frames, b->d_frames, sizeof(float) * kNumParams * nFrames, cudaMemcpyDeviceToHost, b->stream));
	if (configs_out) GTTS_CUDA(cudaMemcpyAsync(configs_out, b->d_cfgs_out, sizeof(gtts_event_config) * nChunks, cudaMemcpyDeviceToHost, b->stream));
	GTTS_CUDA(cudaStreamSynchronize(b->stream));
	if (flags[1]) return fail(GTTS_ERR_CUDA, "control-frame kernel: a chunk produced a frame count other than planned");
	return GTTS_OK;
}

} // extern "C"
