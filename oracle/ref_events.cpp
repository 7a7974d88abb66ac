// TEST INFRASTRUCTURE ONLY -- fixture generator for the control-frame generation path.
//
// Runs the UNMODIFIED reference front end (text parser, rule engine: every source of gama_tts/src but main.cpp,
// compiled where they lie by oracle/Makefile) on a text and, for every chunk of the phonetic string, dumps
//   * the event list EventList::generateOutput is about to consume (vtm_control_model/EventList.cpp:929-1091):
//     time, 16 parameters, 16 special parameters (EMPTY = +inf), macro-intonation polynomial where present,
//   * the settings it reads (control period, pitches, intonation flags) and the state of the drift generator
//     (vtm_control_model/DriftGenerator.cpp:55-84: seed, scaling, Butterworth-2 coefficients and history),
//   * the control frames it produces (float32 x 16 per control period),
// as one binary file.  tools/make_golden_events.py turns it into tests/golden/events_v1.npz.  Needs /root/reference
// (sources and voice data): container only.
//
// With GTTS_LIB=<path to libgtts_b200.so> it is also the SEAM TEST of the control-frame path: the binding INTEGRATION.md
// describes -- EventList::list_ copied into gtts_event records, gtts_events_prepare / gtts_events_run_host in place of
// generateOutput -- runs inside the reference's own process on the chunks of the utterance (one batch, later chunks
// chained to the first through continues_previous) and its frames are compared bit for bit with the reference's
// (tests/test_plugin_seam.py, on a GPU box with the prebuilt binary and the voice directory copied next to it).
//
// usage: [REF_EVENTS_FLAGS=macro,micro,drift,smooth] [GTTS_LIB=...] ref_events <voice dir> <out.bin> <text...>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "Controller.h"
#include "EventList.h"
#include "DriftGenerator.h"
#include "Butterworth2LowpassFilter.h"
#undef private
#undef protected
#include "Index.h"
#include "Model.h"
#include "TextParser.h"
#include "PhoneticStringParser.h"

#include <dlfcn.h>
#include "../include/gtts_b200.h"

namespace {

template<typename T> void put(std::ofstream& f, const T& v) { f.write(reinterpret_cast<const char*>(&v), sizeof v); }

} // namespace

int main(int argc, char* argv[])
{
	if (argc < 4) { std::cerr << "usage: ref_events <voice dir> <out.bin> <text...>" << std::endl; return 2; }
	std::string text;
	for (int i = 3; i < argc; ++i) { text += argv[i]; text += ' '; }
	try {
		using namespace GS;
		const Index index{argv[1]};
		auto model = std::make_unique<VTMControlModel::Model>();
		model->load(index);
		auto controller = std::make_unique<VTMControlModel::Controller>(index, *model);
		auto textParser = TextParser::TextParser::getInstance(index, controller->vtmControlModelConfiguration().phoStrFormat);
		const std::string phoneticString = textParser->parse(text.c_str());

		std::ofstream f(argv[2], std::ios::binary);
		VTMControlModel::Controller& c = *controller;
		VTMControlModel::EventList& el = c.eventList_;
		// Controller::getParametersFromPhoneticString (Controller.cpp:119-156), gnuspeech format branch, with the dumps
		c.vtmParamList_.clear();
		c.initUtterance();
		if (const char* flags = std::getenv("REF_EVENTS_FLAGS")) {   // "macro,micro,drift,smooth": the setters of EventList.h:182-192
			int m = 1, u = 1, d = 1, s = 1;
			std::sscanf(flags, "%d,%d,%d,%d", &m, &u, &d, &s);
			el.setMacroIntonation(m); el.setMicroIntonation(u); el.setIntonationDrift(d); el.setSmoothIntonation(s);
		}
		c.phoneticStringParser_ = std::make_unique<VTMControlModel::PhoneticStringParser>(c.index_, c.model_, el);
		std::size_t pos = 0, size = 0;
		int chunks = 0;
		std::vector<gtts_event_config> gCfgs;
		std::vector<int32_t> gCont;
		std::vector<gtts_event> gEvents;
		std::vector<int64_t> gOffsets{0};
		std::vector<float> refFrames;
		while (pos < phoneticString.size()) {
			if (c.nextChunk(phoneticString, pos, size)) {
				el.setUp();
				c.phoneticStringParser_->parse(&phoneticString[pos], size);
				el.generateEventList();
				el.applyIntonation();
				// ---- dump the inputs ----
				const int nEvents = static_cast<int>(el.list_.size());
				put(f, int(0x45564E54));     // 'EVNT'
				put(f, nEvents);
				put(f, int(el.controlPeriod_));
				put(f, int(el.macroIntonation_)); put(f, int(el.microIntonation_)); put(f, int(el.intonationDrift_)); put(f, int(el.smoothIntonation_));
				put(f, double(el.initialPitch_)); put(f, double(el.meanPitch_));
				const VTMControlModel::DriftGenerator& d = el.driftGenerator_;
				put(f, double(d.pitchDeviation_)); put(f, double(d.pitchOffset_)); put(f, double(d.seed_));
				put(f, double(d.filter_.b0_)); put(f, double(d.filter_.b1_)); put(f, double(d.filter_.a1_)); put(f, double(d.filter_.a2_));
				put(f, double(d.filter_.x1_)); put(f, double(d.filter_.x2_)); put(f, double(d.filter_.y1_)); put(f, double(d.filter_.y2_));
				for (const auto& e : el.list_) {
					put(f, int(e->time));
					put(f, int(e->interpData ? 1 : 0));
					for (int j = 0; j < 16; ++j) put(f, double(e->parameters[j]));
					for (int j = 0; j < 16; ++j) put(f, double(e->specialParameters[j]));
					const double z = 0.0;
					put(f, e->interpData ? e->interpData->a : z); put(f, e->interpData ? e->interpData->b : z);
					put(f, e->interpData ? e->interpData->c : z); put(f, e->interpData ? e->interpData->d : z);
				}
				// ---- the binding of INTEGRATION.md: the same inputs as C-ABI records ----
				{
					gtts_event_config g = {};
					g.control_period = el.controlPeriod_;
					g.macro_intonation = el.macroIntonation_; g.micro_intonation = el.microIntonation_;
					g.intonation_drift = el.intonationDrift_; g.smooth_intonation = el.smoothIntonation_;
					g.initial_pitch = el.initialPitch_; g.mean_pitch = el.meanPitch_;
					g.drift_deviation2 = d.pitchDeviation_; g.drift_offset = d.pitchOffset_;
					g.drift_b0 = d.filter_.b0_; g.drift_b1 = d.filter_.b1_; g.drift_a1 = d.filter_.a1_; g.drift_a2 = d.filter_.a2_;
					if (gCfgs.empty()) {     // later chunks: the state is carried on the device (continues_previous)
						g.drift_seed = d.seed_;
						g.drift_x1 = d.filter_.x1_; g.drift_x2 = d.filter_.x2_; g.drift_y1 = d.filter_.y1_; g.drift_y2 = d.filter_.y2_;
					}
					gCont.push_back(gCfgs.empty() ? 0 : 1);
					gCfgs.push_back(g);
					for (const auto& e : el.list_) {
						gtts_event ev = {};
						ev.time = e->time;
						ev.has_interp = e->interpData ? 1 : 0;
						for (int j = 0; j < 16; ++j) { ev.param[j] = e->parameters[j]; ev.special[j] = e->specialParameters[j]; }
						if (e->interpData) { ev.a = e->interpData->a; ev.b = e->interpData->b; ev.c = e->interpData->c; ev.d = e->interpData->d; }
						gEvents.push_back(ev);
					}
					gOffsets.push_back(static_cast<int64_t>(gEvents.size()));
				}
				// ---- run it, dump the frames of this chunk ----
				const std::size_t before = c.vtmParamList_.size();
				el.generateOutput(c.vtmParamList_);
				const int nFrames = static_cast<int>(c.vtmParamList_.size() - before);
				put(f, nFrames);
				for (std::size_t i = before; i < c.vtmParamList_.size(); ++i) {
					for (int j = 0; j < 16; ++j) { put(f, float(c.vtmParamList_[i][j])); refFrames.push_back(c.vtmParamList_[i][j]); }
				}
				++chunks;
			}
			pos += size;
		}
		std::cout << "chunks " << chunks << " frames " << c.vtmParamList_.size() << std::endl;
		if (const char* libPath = std::getenv("GTTS_LIB")) {
			void* lib = dlopen(libPath, RTLD_NOW);
			if (!lib) { std::cerr << "dlopen: " << dlerror() << std::endl; return 3; }
			auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p) { std::cerr << "missing " << name << std::endl; std::exit(3); } return p; };
			auto create = reinterpret_cast<decltype(&gtts_create)>(sym("gtts_create"));
			auto lastError = reinterpret_cast<decltype(&gtts_last_error)>(sym("gtts_last_error"));
			auto prepare = reinterpret_cast<decltype(&gtts_events_prepare)>(sym("gtts_events_prepare"));
			auto layout = reinterpret_cast<decltype(&gtts_events_layout)>(sym("gtts_events_layout"));
			auto runHost = reinterpret_cast<decltype(&gtts_events_run_host)>(sym("gtts_events_run_host"));
			auto freeBatch = reinterpret_cast<decltype(&gtts_events_free)>(sym("gtts_events_free"));
			auto destroy = reinterpret_cast<decltype(&gtts_destroy)>(sym("gtts_destroy"));
			gtts_handle* h = nullptr;
			if (create(0, &h) != GTTS_OK) { std::cerr << "gtts_create: " << lastError() << std::endl; return 4; }
			gtts_events_batch* b = nullptr;
			if (prepare(h, gCfgs.data(), gCont.data(), gEvents.data(), gOffsets.data(), static_cast<int64_t>(gCfgs.size()), &b) != GTTS_OK) {
				std::cerr << "gtts_events_prepare: " << lastError() << std::endl; return 4;
			}
			std::vector<int64_t> fo(gCfgs.size() + 1);
			layout(b, fo.data());
			std::vector<float> got(static_cast<std::size_t>(fo.back()) * 16 + 1);
			if (runHost(b, gEvents.data(), got.data(), nullptr) != GTTS_OK) { std::cerr << "gtts_events_run_host: " << lastError() << std::endl; return 4; }
			long mismatches = fo.back() * 16 == static_cast<int64_t>(refFrames.size()) ? 0 : -1;
			if (mismatches == 0) {
				for (std::size_t i = 0; i < refFrames.size(); ++i) {
					const bool bothNaN = refFrames[i] != refFrames[i] && got[i] != got[i];
					if (!bothNaN && std::memcmp(&refFrames[i], &got[i], 4) != 0) ++mismatches;
				}
			}
			std::cout << "gpu_check chunks " << gCfgs.size() << " frames " << fo.back() << " mismatches " << mismatches << std::endl;
			freeBatch(b);
			destroy(h);
			if (mismatches != 0) return 5;
		}
	} catch (std::exception& e) {
		std::cerr << "Exception: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
