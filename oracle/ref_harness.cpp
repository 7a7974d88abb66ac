// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Thin C-ABI harness around the UNMODIFIED reference tube model.  It is compiled by oracle/Makefile
// against the reference sources where they lie (-I/root/reference/gama_tts/src{,/vtm}); no reference
// source is copied into this repository.  Output goes to oracle/_ref/libgtts_ref.so (git-ignored).
//
// What it drives (reference file:line):
//   * GS::VTM::VocalTractModel::getInstance          gama_tts/src/vtm/VocalTractModel.cpp:35-59
//   * the Controller::synthesize interpolation loop   gama_tts/src/vtm_control_model/Controller.cpp:277-313
//     (restated below because Controller drags in the whole articulatory database; the loop is the
//      same float32 sequence: cur = frame[i-1]; delta = (frame[i]-cur)*coef; steps x {set, exec, cur+=delta})
//   * finishSynthesis / outputBuffer                  gama_tts/src/vtm/VocalTractModel0.h:720-723, :73
//
// It is used (a) to pin oracle/tube_oracle.c, (b) to generate tests/golden/*.npz, (c) as the
// "reference" CPU baseline of bench.py (kind = "reference").

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <array>
#include <algorithm>
#include <unistd.h>

// Probe access to the reference's private members (KAT extraction only).
#define private public
#define protected public
#include "ConfigurationData.h"
#include "VocalTractModel.h"
#include "VocalTractModel0.h"
#include "NoiseSource.h"
#include "SampleRateConverter.h"
#include "WavetableGlottalSource.h"
#include "WavetableGlottalSourceFIRFilter.h"
#include "VTMUtil.h"
#include "WAVEFileWriter.h"
#undef private
#undef protected

namespace {

thread_local std::string g_err;

std::string writeTempConfig(const char* text)
{
	char path[] = "/tmp/gtts_ref_cfg_XXXXXX";
	int fd = mkstemp(path);
	if (fd < 0) throw std::runtime_error("mkstemp failed");
	size_t len = std::strlen(text);
	if (write(fd, text, len) != (ssize_t) len) { close(fd); throw std::runtime_error("write failed"); }
	close(fd);
	return path;
}

std::unique_ptr<GS::VTM::VocalTractModel> makeModel(const char* configText)
{
	const std::string path = writeTempConfig(configText);
	std::unique_ptr<GS::VTM::VocalTractModel> vtm;
	try {
		GS::ConfigurationData data(path);
		vtm = GS::VTM::VocalTractModel::getInstance(data, false);
	} catch (...) {
		unlink(path.c_str());
		throw;
	}
	unlink(path.c_str());
	return vtm;
}

// The loop of Controller::synthesize (Controller.cpp:277-313) on a packed float32 track, followed by
// finishSynthesis() as in Controller::synthesizeToFile/ToBuffer (Controller.cpp:231-234).
void runTrack(GS::VTM::VocalTractModel& vtm, double controlRate, const float* frames, long nFrames, int nParam)
{
	if (nFrames <= 0) { vtm.finishSynthesis(); return; } // synthesize() returns early, finishSynthesis() still runs (Controller.cpp:232-233)
	const unsigned int controlSteps = static_cast<unsigned int>(std::rint(vtm.internalSampleRate() / controlRate));
	const float coef = 1.0f / controlSteps;
	std::vector<float> cur(nParam), delta(nParam);
	for (long i = 1; i <= nFrames; ++i) {
		const float* prev = frames + (i - 1) * nParam;
		const float* next = (i < nFrames) ? frames + i * nParam : prev; // duplicated last frame (:283)
		for (int j = 0; j < nParam; ++j) {
			cur[j] = prev[j];
			delta[j] = (next[j] - cur[j]) * coef;
		}
		for (unsigned int j = 0; j < controlSteps; ++j) {
			vtm.setAllParameters(cur);
			vtm.execSynthesisStep();
			for (int k = 0; k < nParam; ++k) cur[k] += delta[k];
		}
	}
	vtm.finishSynthesis();
}

} // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_free(void* p) { std::free(p); }

// Synthesises one control track with the reference model selected by "model = N" in config_text.
// Returns 0 and a malloc'ed float32 buffer (raw outputBuffer(), before peak normalisation).
int ref_synthesize(const char* config_text, double control_rate, const float* frames, long n_frames,
			float** out, long* n_out, double* internal_rate)
{
	try {
		auto vtm = makeModel(config_text);
		if (internal_rate) *internal_rate = vtm->internalSampleRate();
		runTrack(*vtm, control_rate, frames, n_frames, 16);
		const std::vector<float>& buf = vtm->outputBuffer();
		*n_out = static_cast<long>(buf.size());
		*out = static_cast<float*>(std::malloc(sizeof(float) * std::max<size_t>(buf.size(), 1)));
		std::memcpy(*out, buf.data(), sizeof(float) * buf.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// The editor's interactive call pattern (gama_tts_editor/src/interactive/InteractiveAudio.cpp:131-186, the JACK process
// callback) without JACK: the model is constructed with interactive = true; every callback wants `callback_frames`
// output samples, takes what outputBuffer() holds with VTM::Util::getSamples (which clears the buffer once it is
// used up, VTMUtil.cpp:30-46) and, if that is not enough, steps the model -- setParameter x 16 + execSynthesisStep,
// one row of `params` per step -- until it is.  Runs until the rows are used up; returns every sample handed out.
int ref_interactive(const char* config_text, const float* params, long n_steps, long callback_frames, float** out, long* n_out)
{
	try {
		const std::string path = writeTempConfig(config_text);
		std::unique_ptr<GS::VTM::VocalTractModel> vtm;
		try {
			GS::ConfigurationData data(path);
			vtm = GS::VTM::VocalTractModel::getInstance(data, true);
		} catch (...) {
			unlink(path.c_str());
			throw;
		}
		unlink(path.c_str());
		std::vector<float> all, cb(static_cast<size_t>(callback_frames));
		std::vector<float>& buf = vtm->outputBuffer();
		std::size_t pos = 0;
		long step = 0;
		while (step < n_steps) {
			const std::size_t n = GS::VTM::Util::getSamples(buf, pos, cb.data(), callback_frames, 1.0f);
			all.insert(all.end(), cb.begin(), cb.begin() + n);
			if (n == static_cast<std::size_t>(callback_frames)) continue;
			const std::size_t target = callback_frames - n;
			while (buf.size() < target && step < n_steps) {
				for (int i = 0; i < 16; ++i) vtm->setParameter(i, params[step * 16 + i]);
				vtm->execSynthesisStep();
				++step;
			}
			const std::size_t n2 = GS::VTM::Util::getSamples(buf, pos, cb.data(), target, 1.0f);
			all.insert(all.end(), cb.begin(), cb.begin() + n2);
		}
		*n_out = static_cast<long>(all.size());
		*out = static_cast<float*>(std::malloc(sizeof(float) * std::max<size_t>(all.size(), 1)));
		std::memcpy(*out, all.data(), sizeof(float) * all.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// Same, but through the per-sample entry points with caller-supplied per-sample parameters (no
// interpolation): what the plugin seam sees (setAllParameters + execSynthesisStep per internal sample).
int ref_synthesize_samples(const char* config_text, const float* params, long n_samples, float** out, long* n_out)
{
	try {
		auto vtm = makeModel(config_text);
		std::vector<float> cur(16);
		for (long i = 0; i < n_samples; ++i) {
			std::memcpy(cur.data(), params + i * 16, sizeof(float) * 16);
			vtm->setAllParameters(cur);
			vtm->execSynthesisStep();
		}
		vtm->finishSynthesis();
		const std::vector<float>& buf = vtm->outputBuffer();
		*n_out = static_cast<long>(buf.size());
		*out = static_cast<float*>(std::malloc(sizeof(float) * std::max<size_t>(buf.size(), 1)));
		std::memcpy(*out, buf.data(), sizeof(float) * buf.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// Batch over U tracks (packed frames, frame_offsets[U+1], one config per utterance or a single shared
// one when n_configs == 1), n_threads std::threads each owning its own model instance (the reference
// has no shared mutable state besides Log::debugEnabled).  out may be NULL (timing only);
// otherwise out_offsets[U+1] gives where each utterance's float32 samples go.
// Returns wall seconds of the synthesis span (model construction excluded) in *seconds.
int ref_batch(const char* const* config_texts, int n_configs, double control_rate,
		const float* frames, const long* frame_offsets, long n_utt, int n_threads,
		float* out, const long* out_offsets, long* n_out_each, double* seconds)
{
	try {
		if (n_threads < 1) n_threads = 1;
		std::atomic<long> next{0};
		std::atomic<int> failed{0};
		std::vector<double> busy(n_threads, 0.0);
		auto worker = [&](int tid) {
			try {
				std::unique_ptr<GS::VTM::VocalTractModel> shared;
				if (n_configs == 1) shared = makeModel(config_texts[0]);
				for (;;) {
					const long u = next.fetch_add(1);
					if (u >= n_utt) break;
					std::unique_ptr<GS::VTM::VocalTractModel> own;
					GS::VTM::VocalTractModel* vtm;
					if (n_configs == 1) {
						vtm = shared.get();
						vtm->reset(); // Controller.cpp:231
					} else {
						own = makeModel(config_texts[u]);
						vtm = own.get();
					}
					auto t0 = std::chrono::steady_clock::now();
					runTrack(*vtm, control_rate, frames + frame_offsets[u] * 16,
							frame_offsets[u + 1] - frame_offsets[u], 16);
					auto t1 = std::chrono::steady_clock::now();
					busy[tid] += std::chrono::duration<double>(t1 - t0).count();
					const std::vector<float>& buf = vtm->outputBuffer();
					if (n_out_each) n_out_each[u] = static_cast<long>(buf.size());
					if (out) {
						const long cap = out_offsets[u + 1] - out_offsets[u];
						const long n = std::min<long>(cap, static_cast<long>(buf.size()));
						std::memcpy(out + out_offsets[u], buf.data(), sizeof(float) * n);
					}
				}
			} catch (const std::exception& e) {
				failed = 1;
				g_err = e.what();
			}
		};
		auto w0 = std::chrono::steady_clock::now();
		std::vector<std::thread> pool;
		for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
		worker(0);
		for (auto& th : pool) th.join();
		auto w1 = std::chrono::steady_clock::now();
		if (seconds) *seconds = std::chrono::duration<double>(w1 - w0).count();
		return failed ? 1 : 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// The reference's own output stage on a raw output buffer: Controller::writeOutputToFile (Controller.cpp:315-328)
// -- VTM::Util::calculateOutputScale + WAVEFileWriter::writeSample(sample * scale) -- into a temporary WAVE file,
// whose 16-bit payload (after the 44-byte header, WAVEFileWriter.cpp:62-105) is read back.  Returns the scale.
int ref_pcm16(const float* x, long n, float rate, short* out, float* scale_out)
{
	try {
		char path[] = "/tmp/gtts_ref_wav_XXXXXX";
		int fd = mkstemp(path);
		if (fd < 0) throw std::runtime_error("mkstemp failed");
		close(fd);
		const std::vector<float> audioData(x, x + n);
		float scale;
		{
			GS::WAVEFileWriter fileWriter(path, 1, audioData.size(), rate);
			scale = GS::VTM::Util::calculateOutputScale(audioData);
			for (std::size_t i = 0, end = audioData.size(); i < end; ++i) fileWriter.writeSample(audioData[i] * scale);
		}
		if (scale_out) *scale_out = scale;
		FILE* f = std::fopen(path, "rb");
		if (!f) { unlink(path); throw std::runtime_error("cannot reopen the WAVE file"); }
		std::fseek(f, 44, SEEK_SET);
		const size_t got = std::fread(out, sizeof(short), static_cast<size_t>(n), f);
		std::fclose(f);
		unlink(path);
		if (got != static_cast<size_t>(n)) throw std::runtime_error("short WAVE payload");
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// ---- KAT probes (private members of the reference) -------------------------

// First n values of NoiseSource::getSample() (NoiseSource.h:40-44).
void ref_noise(double* out, long n)
{
	GS::VTM::NoiseSource src;
	for (long i = 0; i < n; ++i) out[i] = src.getSample();
}

// Glottal FIR taps in double (WavetableGlottalSourceFIRFilter.h:74-114). Returns the tap count.
int ref_fir_taps(double* out, int cap)
{
	GS::VTM::WavetableGlottalSourceFIRFilter<double> fir(0.2, 0.1, 0.00000001);
	const int n = fir.numberTaps_;
	for (int i = 0; i < n && i < cap; ++i) out[i] = fir.coef_[i];
	return n;
}

// SRC tables h_, deltaH_ (3328 each) and the per-rate integers (SampleRateConverter.h:136-164, 230-255).
void ref_src_tables(double input_rate, double output_rate, double* h, double* dh, unsigned int* incs, int* pad)
{
	GS::VTM::SampleRateConverter<double> src(input_rate, output_rate, [](float) {});
	for (size_t i = 0; i < src.h_.size(); ++i) { h[i] = src.h_[i]; dh[i] = src.deltaH_[i]; }
	incs[0] = src.timeRegisterIncrement_;
	incs[1] = src.filterIncrement_;
	incs[2] = src.phaseIncrement_;
	*pad = src.padSize_;
}

// Runs the reference SampleRateConverter alone on n_in doubles (dataFill + flushBuffer).
int ref_src_run(double input_rate, double output_rate, const double* x, long n_in, float** out, long* n_out)
{
	try {
		std::vector<float> buf;
		GS::VTM::SampleRateConverter<double> src(input_rate, output_rate, [&](float s) { buf.push_back(s); });
		for (long i = 0; i < n_in; ++i) src.dataFill(x[i]);
		src.flushBuffer();
		*n_out = static_cast<long>(buf.size());
		*out = static_cast<float*>(std::malloc(sizeof(float) * std::max<size_t>(buf.size(), 1)));
		std::memcpy(*out, buf.data(), sizeof(float) * buf.size());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

// Wavetable (512 doubles) and its integers after construction and an optional setup(amplitude)
// (WavetableGlottalSource.h:90-141, 162-184). amplitude < 0 skips the setup call.
void ref_wavetable(int sine, double sample_rate, double tp, double tn_min, double tn_max, double amplitude,
			double* table, double* scalars)
{
	using Src = GS::VTM::WavetableGlottalSource<double>;
	Src src(sine ? Src::Type::sine : Src::Type::pulse, sample_rate, tp, tn_min, tn_max);
	if (amplitude >= 0.0) src.setup(amplitude);
	for (int i = 0; i < 512; ++i) table[i] = src.wavetable_[i];
	scalars[0] = src.tableDiv1_;
	scalars[1] = src.tableDiv2_;
	scalars[2] = src.tnLength_;
	scalars[3] = src.tnDelta_;
	scalars[4] = src.basicIncrement_;
}

// Derived per-voice constants of VocalTractModel0<double> (VocalTractModel0.h:338-392, 457-470).
int ref_model0_constants(const char* config_text, double* out /* >= 24 */)
{
	try {
		const std::string path = writeTempConfig(config_text);
		GS::ConfigurationData data(path);
		unlink(path.c_str());
		GS::VTM::VocalTractModel0<double> m(data, false);
		int k = 0;
		out[k++] = m.sampleRate_;
		out[k++] = m.breathinessFactor_;
		out[k++] = m.crossmixFactor_;
		out[k++] = m.dampingFactor_;
		out[k++] = m.mouthRadiationFilter_->b0_;
		out[k++] = m.mouthReflectionFilter_->b0_;
		out[k++] = m.mouthReflectionFilter_->a1_;
		out[k++] = m.nasalRadiationFilter_->b0_;
		out[k++] = m.nasalReflectionFilter_->b0_;
		out[k++] = m.nasalReflectionFilter_->a1_;
		out[k++] = m.throat_->b0_;
		out[k++] = m.throat_->a1_;
		out[k++] = m.throat_->throatGain_;
		for (int i = 1; i < 6; ++i) out[k++] = m.nasalCoeff_[i];
		out[k++] = m.config_.apertureRadius;
		out[k++] = m.config_.nasalRadius[1];
		out[k++] = m.glottalSource_->basicIncrement_;
		out[k++] = m.glottalSource_->tableDiv1_;
		out[k++] = m.glottalSource_->tableDiv2_;
		out[k++] = m.glottalSource_->tnDelta_;
		return k;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

} // extern "C"
