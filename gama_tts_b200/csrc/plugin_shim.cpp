// Plugin shim: makes the B200 tube path a drop-in GS::VTM::VocalTractModel for the UNMODIFIED
// reference through its own plugin seam.
//
//   reference side (not rebuilt here)                      this file
//   ---------------------------------                      ---------
//   vtm.txt: model = 2000, dll_path = .../libgtts_plugin.so
//   VocalTractModel::getInstance case 2000                 (gama_tts/src/vtm/VocalTractModel.cpp:53-54)
//   VocalTractModelPlugin: dlopen + dlsym of               (gama_tts/src/vtm/VocalTractModelPlugin.cpp:57-91)
//     GAMA_TTS_construct_vocal_tract_model(const void* config_data, int is_interactive)   -> exported below
//     GAMA_TTS_destruct_vocal_tract_model(void* vtm)                                       -> exported below
//   every virtual forwarded to the returned object         (VocalTractModelPlugin.cpp:104-150)
//
// The returned object is used through the Itanium vtable of GS::VTM::VocalTractModel
// (gama_tts/src/vtm/VocalTractModel.h:43-71) and receives a GS::ConfigurationData, so this one file is
// compiled against the reference headers where they lie (-I$(REF)/gama_tts/src{,/vtm}, plus the
// reference's ConfigurationData.cpp for the non-inline convertString<> specialisations) with the same
// g++ / libstdc++ as the host binary.  Nothing of the reference is copied.  Everything below the seam
// goes through the C ABI of include/gtts_b200.h.
//
// Call pattern honoured (Controller.cpp:231-234, 277-313):
//   [reset] -> (setAllParameters -> execSynthesisStep) x N -> finishSynthesis -> outputBuffer()
// execSynthesisStep() records the 16 float parameters of that internal sample; finishSynthesis() runs
// the whole recording as ONE utterance with steps = 1 (every recorded row is used verbatim for one
// internal sample, so the host's own float32 interpolation is reproduced exactly) and fills the
// output buffer.
//
// Interactive callers (gama_tts_editor/src/interactive/InteractiveAudio.cpp:131-186: the editor steps the model until
// outputBuffer() holds what the audio callback needs, drains it, and steps on) are served in chunks: construction with
// is_interactive != 0 opens a stream (gtts_stream_*), execSynthesisStep() queues the step's parameters and every
// kInteractiveChunk steps the queue is synthesised and its audio appended to outputBuffer() -- the same samples as the
// reference's model, available up to kInteractiveChunk - 1 internal samples (3 ms) later.  Model-0 voices only; an
// interactive model-5 voice is refused (NULL: "Could not construct the vocal tract model",
// VocalTractModelPlugin.cpp:88-90).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "ConfigurationData.h"
#include "VocalTractModel.h"

#include "../../include/gtts_b200.h"

namespace {

class B200VocalTractModel final : public GS::VTM::VocalTractModel {
public:
	B200VocalTractModel(const GS::ConfigurationData& data, int device, bool interactive);
	~B200VocalTractModel() noexcept override;

	void reset() noexcept override;
	double internalSampleRate() const noexcept override { return model5_ ? internalRate5_ : internalRate_; }
	double outputSampleRate() const noexcept override { return model5_ ? voice5_.output_rate : voice_.output_rate; }
	void setParameter(int parameter, float value) noexcept override;
	void setAllParameters(const std::vector<float>& parameters) noexcept override;
	void execSynthesisStep() noexcept override;
	void finishSynthesis() noexcept override;
	std::vector<float>& outputBuffer() noexcept override { return outputBuffer_; }

private:
	static gtts_handle* sharedHandle(int device);

	void loadModel5(const GS::ConfigurationData& data);

	gtts_voice_config voice_;
	gtts_voice5_config voice5_;           // model-5 voice (the configuration has model 5's keys: see the constructor)
	bool model5_ = false;
	double internalRate5_ = 0.0;
	gtts_handle* handle_ = nullptr;       // process-wide, not owned
	int32_t internalRate_ = 0;
	float current_[GTTS_NUM_PARAMS];
	std::vector<float> recorded_;          // one row of 16 per execSynthesisStep()
	std::vector<float> outputBuffer_;
	bool failed_ = false;                  // recording abandoned: nothing more is recorded until finishSynthesis() / reset()

	// interactive mode: a stream, fed kInteractiveChunk internal samples at a time
	static constexpr int kInteractiveChunk = 64;
	bool interactive_ = false;
	gtts_stream* stream_ = nullptr;
	std::vector<float> chunkOut_;
	bool pushChunk(bool finish) noexcept;

	void fail(const char* what) noexcept;
};

int g_lastStatus = GTTS_OK;                // status of the last finishSynthesis() in this process (GTTS_plugin_last_status)

// Same keys, same order as VocalTractModel0::loadConfiguration (VocalTractModel0.h:266-305).
B200VocalTractModel::B200VocalTractModel(const GS::ConfigurationData& data, int device, bool interactive)
		: interactive_(interactive)
{
	std::memset(&voice_, 0, sizeof voice_);
	std::memset(&voice5_, 0, sizeof voice5_);
	std::memset(current_, 0, sizeof current_);
	// Which of the reference's models this object stands in for: the configuration of a 5_xxx voice directory carries
	// model 5's keys (VocalTractModel5.h:373-425), that of a 0_xxx directory model 0's; "model" itself says 2000 (the
	// plugin) in both.
	try {
		(void) data.value<double>("glottal_noise_cutoff");
		model5_ = true;
	} catch (...) {
		model5_ = false;
	}
	if (model5_) {
		if (interactive_) throw std::runtime_error("interactive mode is implemented for model-0 voices only");
		loadModel5(data);
		handle_ = sharedHandle(device);
		outputBuffer_.reserve(OUTPUT_BUFFER_RESERVE);
		return;
	}
	voice_.output_rate = data.value<double>("output_rate");
	voice_.waveform = data.value<int>("waveform");
	voice_.glottal_pulse_tp = data.value<double>("glottal_pulse_tp");
	voice_.glottal_pulse_tn_min = data.value<double>("glottal_pulse_tn_min");
	voice_.glottal_pulse_tn_max = data.value<double>("glottal_pulse_tn_max");
	voice_.breathiness = data.value<double>("breathiness");
	voice_.vocal_tract_length_offset = data.value<double>("vocal_tract_length_offset");
	voice_.vocal_tract_length = data.value<double>("vocal_tract_length");
	voice_.temperature = data.value<double>("temperature");
	voice_.loss_factor = data.value<double>("loss_factor");
	voice_.mouth_coefficient = data.value<double>("mouth_coefficient");
	voice_.nose_coefficient = data.value<double>("nose_coefficient");
	voice_.throat_cutoff = data.value<double>("throat_cutoff");
	voice_.throat_volume = data.value<double>("throat_volume");
	voice_.noise_modulation = data.value<int>("noise_modulation");
	voice_.mix_offset = data.value<double>("mix_offset");
	voice_.global_radius_coef = data.value<double>("global_radius_coef");
	voice_.global_nasal_radius_coef = data.value<double>("global_nasal_radius_coef");
	voice_.aperture_radius = data.value<double>("aperture_radius");
	for (int i = 0; i < 5; ++i) voice_.nasal_radius[i] = data.value<double>("nasal_radius_" + std::to_string(i + 1));
	for (int i = 0; i < 8; ++i) voice_.radius_coef[i] = data.value<double>("radius_" + std::to_string(i + 1) + "_coef");
	// "model" says 2000 here; a host that wants the plugin to stand in for model 3 or 4 (same keys as model 0) adds
	// plugin_tube_model = 3 | 4 to the configuration (absent: models 0 / 2)
	try {
		voice_.tube_model = data.value<int>("plugin_tube_model");
	} catch (...) {
		voice_.tube_model = 0;
	}
	if (interactive_ && voice_.tube_model != 0) throw std::runtime_error("interactive mode is implemented for model-0 voices only");

	if (gtts_voice_internal_rate(&voice_, &internalRate_) != GTTS_OK) throw std::runtime_error(gtts_last_error());
	int64_t nInternal = 0, nOut = 0;
	if (gtts_output_length(&voice_, 1, 0, &nInternal, &nOut) != GTTS_OK) throw std::runtime_error(gtts_last_error());
	handle_ = sharedHandle(device);
	outputBuffer_.reserve(OUTPUT_BUFFER_RESERVE);
	if (interactive_) {
		// every queued row is one internal sample (steps = 1), as in finishSynthesis() below
		if (gtts_stream_open(handle_, &voice_, 250.0, 1, &stream_) != GTTS_OK) throw std::runtime_error(gtts_last_error());
		chunkOut_.resize(static_cast<size_t>(kInteractiveChunk) * 8 + 256);
	}
}

// Same keys as VocalTractModel5::loadConfiguration (VocalTractModel5.h:373-425).
void B200VocalTractModel::loadModel5(const GS::ConfigurationData& data)
{
	gtts_voice5_config& v = voice5_;
	v.output_rate = data.value<double>("output_rate");
	v.waveform = data.value<int>("waveform");
	v.glottal_pulse_tp = data.value<double>("glottal_pulse_tp");
	v.glottal_pulse_tn_min = data.value<double>("glottal_pulse_tn_min");
	v.glottal_pulse_tn_max = data.value<double>("glottal_pulse_tn_max");
	v.breathiness = data.value<double>("breathiness");
	v.vocal_tract_length_offset = data.value<double>("vocal_tract_length_offset");
	v.vocal_tract_length = data.value<double>("vocal_tract_length");
	v.temperature = data.value<double>("temperature");
	v.loss_factor = data.value<double>("loss_factor");
	v.noise_modulation = data.value<int>("noise_modulation");
	v.mix_offset = data.value<double>("mix_offset");
	v.global_radius_coef = data.value<double>("global_radius_coef");
	v.global_nasal_radius_coef = data.value<double>("global_nasal_radius_coef");
	for (int i = 0; i < 6; ++i) v.nasal_radius[i] = data.value<double>("nasal_radius_" + std::to_string(i + 2));
	for (int i = 0; i < 8; ++i) v.radius_coef[i] = data.value<double>("radius_" + std::to_string(i + 1) + "_coef");
	v.glottal_noise_cutoff = data.value<double>("glottal_noise_cutoff");
	v.frication_noise_cutoff = data.value<double>("frication_noise_cutoff");
	v.frication_factor = data.value<double>("frication_factor");
	v.min_glottal_loss = data.value<double>("min_glottal_loss");
	v.max_glottal_loss = data.value<double>("max_glottal_loss");
	v.glottal_lowpass_cutoff = data.value<double>("glottal_lowpass_cutoff");
	v.bypass = data.value<int>("bypass");
	v.constant_radius_mouth_impedance = data.value<bool>("constant_radius_mouth_impedance") ? 1 : 0;
	if (v.constant_radius_mouth_impedance) v.mouth_impedance_radius = data.value<double>("mouth_impedance_radius");
	int64_t nInternal = 0, nOut = 0;
	if (gtts5_output_length(&v, 250.0, 1, 0, nullptr, &nInternal, &nOut) != GTTS_OK) throw std::runtime_error(gtts_last_error());
	if (gtts5_voice_internal_rate(&v, &internalRate5_) != GTTS_OK) throw std::runtime_error(gtts_last_error());
}

// The host constructs one model per synthesis and dlcloses the plugin after it (VocalTractModelPlugin.cpp:95-98).
// The device handle (CUDA context, tables, loaded kernels: ~0.4 s to set up) is therefore kept for the life of
// the process -- one per device -- and the libraries are linked -z nodelete so that it survives the dlclose.
gtts_handle* B200VocalTractModel::sharedHandle(int device)
{
	static std::mutex lock;
	static std::map<int, gtts_handle*> handles;
	std::lock_guard<std::mutex> g(lock);
	auto it = handles.find(device);
	if (it != handles.end()) return it->second;
	gtts_handle* h = nullptr;
	if (gtts_create(device, &h) != GTTS_OK) throw std::runtime_error(gtts_last_error());
	handles[device] = h;
	return h;
}

B200VocalTractModel::~B200VocalTractModel() noexcept
{
	if (stream_) gtts_stream_close(stream_);
}

// Interactive mode: synthesises the queued steps and appends their audio to the output buffer.
bool B200VocalTractModel::pushChunk(bool finish) noexcept
{
	try {
		const int64_t n = static_cast<int64_t>(recorded_.size() / GTTS_NUM_PARAMS);
		int64_t written = 0;
		int rc = GTTS_OK;
		if (n > 0) {
			const size_t need = static_cast<size_t>(n) * 8 + 256;       // output rate / internal rate is below 8 for every tract length
			if (chunkOut_.size() < need) chunkOut_.resize(need);
			rc = gtts_stream_push_frames(stream_, recorded_.data(), n, chunkOut_.data(), static_cast<int64_t>(chunkOut_.size()), &written);
			if (rc == GTTS_OK) outputBuffer_.insert(outputBuffer_.end(), chunkOut_.begin(), chunkOut_.begin() + written);
			recorded_.clear();
		}
		if (rc == GTTS_OK && finish) {
			rc = gtts_stream_finish(stream_, chunkOut_.data(), static_cast<int64_t>(chunkOut_.size()), &written);
			if (rc == GTTS_OK) outputBuffer_.insert(outputBuffer_.end(), chunkOut_.begin(), chunkOut_.begin() + written);
		}
		if (rc != GTTS_OK) { g_lastStatus = rc; fail(gtts_last_error()); return false; }
		return true;
	} catch (...) {
		g_lastStatus = GTTS_ERR_NOMEM;
		fail("out of memory in the interactive chunk");
		return false;
	}
}

// VocalTractModel0::reset (VocalTractModel0.h:309-326): all dynamic state back to zero, output cleared;
// like the reference, the current parameters are kept.
void B200VocalTractModel::reset() noexcept
{
	recorded_.clear();
	outputBuffer_.clear();
	failed_ = false;
	if (stream_ && gtts_stream_reset(stream_) != GTTS_OK) { g_lastStatus = GTTS_ERR_CUDA; fail(gtts_last_error()); }
}

// The VocalTractModel interface has no error channel (all eight virtuals are noexcept void, VocalTractModel.h:43-71)
// and the host would write a silent, empty WAV with exit code 0 from an empty outputBuffer() (Controller.cpp:315-328).
// A failed synthesis therefore ends the process the way an exception escaping a noexcept function of the reference
// would: message on stderr, std::terminate().  Hosts that must survive (a speech server) set
// GTTS_PLUGIN_ON_ERROR=continue: the output stays empty, the recording is dropped, further steps are ignored until
// reset(), and GTTS_plugin_last_status() reports the error code.
void B200VocalTractModel::fail(const char* what) noexcept
{
	std::fprintf(stderr, "[gtts_plugin] synthesis failed: %s\n", what);
	recorded_.clear();
	outputBuffer_.clear();
	failed_ = true;
	if (g_lastStatus == GTTS_OK) g_lastStatus = GTTS_ERR_CUDA;
	const char* mode = std::getenv("GTTS_PLUGIN_ON_ERROR");
	if (!(mode && std::strcmp(mode, "continue") == 0)) std::terminate();
}

// VocalTractModel0.h:665-694: an invalid index is ignored silently.  The radius scaling
// (max(r * radiusCoef, 0.01)) is applied on the device from the raw float.
void B200VocalTractModel::setParameter(int parameter, float value) noexcept
{
	if (parameter < 0 || parameter >= GTTS_NUM_PARAMS) return;
	current_[parameter] = value;
}

// VocalTractModel0.h:698-716: a vector of the wrong size is ignored silently.
void B200VocalTractModel::setAllParameters(const std::vector<float>& parameters) noexcept
{
	if (parameters.size() != GTTS_NUM_PARAMS) return;
	std::memcpy(current_, parameters.data(), sizeof current_);
}

void B200VocalTractModel::execSynthesisStep() noexcept
{
	if (failed_) return;
	try {
		recorded_.insert(recorded_.end(), current_, current_ + GTTS_NUM_PARAMS);
	} catch (...) {
		g_lastStatus = GTTS_ERR_NOMEM;
		fail("out of memory while recording parameters");
		return;
	}
	if (interactive_ && recorded_.size() >= static_cast<size_t>(kInteractiveChunk) * GTTS_NUM_PARAMS) pushChunk(false);
}

// VocalTractModel0.h:720-723 (SRC flush) -- here: the deferred synthesis of everything recorded.
void B200VocalTractModel::finishSynthesis() noexcept
{
	if (failed_) {
		// the recording of this utterance was abandoned (GTTS_PLUGIN_ON_ERROR=continue): nothing to synthesise; the
		// host does not call reset() on an empty output (Controller.cpp:231), so the next utterance starts here
		failed_ = false;
		return;
	}
	g_lastStatus = GTTS_OK;
	if (interactive_) {
		pushChunk(true);
		failed_ = false;
		return;
	}
	if (model5_) {
		gtts5_batch* b5 = nullptr;
		try {
			const int64_t nSamples = static_cast<int64_t>(recorded_.size() / GTTS_NUM_PARAMS);
			const int64_t frameOffsets[2] = {0, nSamples};
			const int32_t steps[1] = {1};
			int rc = gtts5_batch_prepare(handle_, &voice5_, 1, nullptr, 250.0, steps, frameOffsets, 1, &b5);
			if (rc != GTTS_OK) { g_lastStatus = rc; fail(gtts_last_error()); failed_ = false; return; }
			int64_t outOffsets[2] = {0, 0};
			int64_t nOut = 0;
			gtts5_batch_layout(b5, outOffsets, &nOut, nullptr);
			const size_t base = outputBuffer_.size();
			outputBuffer_.resize(base + static_cast<size_t>(outOffsets[1]));
			rc = gtts5_batch_run_host(b5, recorded_.data(), outputBuffer_.data() + base);
			gtts5_batch_free(b5);
			b5 = nullptr;
			if (rc != GTTS_OK) { g_lastStatus = rc; fail(gtts_last_error()); failed_ = false; return; }
			outputBuffer_.resize(base + static_cast<size_t>(nOut));
			recorded_.clear();
		} catch (...) {
			if (b5) gtts5_batch_free(b5);
			g_lastStatus = GTTS_ERR_NOMEM;
			fail("out of memory in finishSynthesis");
			failed_ = false;
		}
		return;
	}
	gtts_batch* batch = nullptr;
	try {
		const int64_t nSamples = static_cast<int64_t>(recorded_.size() / GTTS_NUM_PARAMS);
		const int64_t frameOffsets[2] = {0, nSamples};
		const int32_t steps[1] = {1};
		int rc = gtts_batch_prepare(handle_, &voice_, 1, nullptr, 250.0, steps, frameOffsets, 1, &batch);
		if (rc != GTTS_OK) { g_lastStatus = rc; fail(gtts_last_error()); failed_ = false; return; }
		int64_t outOffsets[2] = {0, 0};
		int64_t nOut = 0;
		gtts_batch_layout(batch, outOffsets, nullptr);
		gtts_batch_lengths(batch, &nOut);
		const size_t base = outputBuffer_.size();
		outputBuffer_.resize(base + static_cast<size_t>(outOffsets[1]));     // the layout is padded to whole rows
		rc = gtts_batch_run_host(batch, recorded_.data(), outputBuffer_.data() + base);
		gtts_batch_free(batch);
		batch = nullptr;
		if (rc != GTTS_OK) { g_lastStatus = rc; fail(gtts_last_error()); failed_ = false; return; }
		outputBuffer_.resize(base + static_cast<size_t>(nOut));
		recorded_.clear();
	} catch (...) {
		if (batch) gtts_batch_free(batch);
		g_lastStatus = GTTS_ERR_NOMEM;
		fail("out of memory in finishSynthesis");
		failed_ = false;
	}
}

} // namespace

extern "C" {

// Replaces nothing in the reference: these are the two symbols VocalTractModelPlugin looks up
// (gama_tts/src/vtm/VocalTractModelPlugin.cpp:37-38, 77, 82).
void* GAMA_TTS_construct_vocal_tract_model(const void* config_data, int is_interactive)
{
	if (!config_data) return nullptr;
	try {
		const char* dev = std::getenv("GTTS_DEVICE");
		const GS::ConfigurationData& data = *static_cast<const GS::ConfigurationData*>(config_data);
		GS::VTM::VocalTractModel* vtm = new B200VocalTractModel(data, dev ? std::atoi(dev) : 0, is_interactive != 0);
		return vtm;
	} catch (const std::exception& e) {
		std::fprintf(stderr, "[gtts_plugin] %s\n", e.what());
		return nullptr;
	} catch (...) {
		return nullptr;
	}
}

void GAMA_TTS_destruct_vocal_tract_model(void* vtm)
{
	delete static_cast<GS::VTM::VocalTractModel*>(vtm);
}

// Not looked up by the reference: GTTS_OK or the GTTS_ERR_* code of the last finishSynthesis() in this process, for
// hosts that run with GTTS_PLUGIN_ON_ERROR=continue.
int GTTS_plugin_last_status(void) { return g_lastStatus; }

} // extern "C"
