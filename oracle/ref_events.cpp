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
// usage: [REF_EVENTS_FLAGS=macro,micro,drift,smooth] ref_events <voice dir> <out.bin> <text...>
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
				// ---- run it, dump the frames of this chunk ----
				const std::size_t before = c.vtmParamList_.size();
				el.generateOutput(c.vtmParamList_);
				const int nFrames = static_cast<int>(c.vtmParamList_.size() - before);
				put(f, nFrames);
				for (std::size_t i = before; i < c.vtmParamList_.size(); ++i) {
					for (int j = 0; j < 16; ++j) put(f, float(c.vtmParamList_[i][j]));
				}
				++chunks;
			}
			pos += size;
		}
		std::cout << "chunks " << chunks << " frames " << c.vtmParamList_.size() << std::endl;
	} catch (std::exception& e) {
		std::cerr << "Exception: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
