// BASELINE config 5 through the C ABI: one utterance pushed control frame by control frame (and in chunks).
// g++ -O2 -std=c++17 -I include -o tools/microbench/stream_push tools/microbench/stream_push.cpp -L gama_tts_b200/csrc -lgtts_b200 -Wl,-rpath,'$ORIGIN/../../gama_tts_b200/csrc'
// usage: stream_push frames.f32 n_frames   (frames: float32 [n_frames][16]; voice: 0_male/male, hard-coded below)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gtts_b200.h"

int main(int argc, char** argv)
{
	if (argc < 3) { std::fprintf(stderr, "usage: stream_push frames.f32 n_frames\n"); return 2; }
	const long nFrames = std::atol(argv[2]);
	std::vector<float> frames(static_cast<size_t>(nFrames) * 16);
	FILE* f = std::fopen(argv[1], "rb");
	if (!f || std::fread(frames.data(), sizeof(float), frames.size(), f) != frames.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
	std::fclose(f);
	gtts_voice_config v = {};      // data/voice/english/0_male: vtm.txt + variant/male.txt
	v.output_rate = 48000; v.waveform = 0; v.noise_modulation = 1; v.glottal_pulse_tp = 40; v.glottal_pulse_tn_min = 24;
	v.glottal_pulse_tn_max = 24; v.breathiness = 0.5; v.vocal_tract_length_offset = 0; v.vocal_tract_length = 17.5;
	v.temperature = 32; v.loss_factor = 0.8; v.mouth_coefficient = 5000; v.nose_coefficient = 5000; v.throat_cutoff = 1500;
	v.throat_volume = 6; v.mix_offset = 48; v.global_radius_coef = 1; v.global_nasal_radius_coef = 1; v.aperture_radius = 3.05;
	const double nr[5] = {1.35, 1.96, 1.91, 1.3, 0.73};
	for (int i = 0; i < 5; ++i) v.nasal_radius[i] = nr[i];
	for (int i = 0; i < 8; ++i) v.radius_coef[i] = 1.0;
	gtts_handle* h = nullptr;
	if (gtts_create(0, &h) != GTTS_OK) { std::fprintf(stderr, "%s\n", gtts_last_error()); return 1; }
	std::vector<float> out(1 << 20);
	std::printf("{");
	const int chunks[] = {1, 4, 25, 250};
	for (int ci = 0; ci < 4; ++ci) {
		const int chunk = chunks[ci];
		gtts_stream* s = nullptr;
		if (gtts_stream_open(h, &v, 250.0, 0, &s) != GTTS_OK) { std::fprintf(stderr, "%s\n", gtts_last_error()); return 1; }
		long long total = 0;
		int64_t n = 0;
		// warm-up: the first pushes build the graph and the staging buffers
		for (long i = 0; i < 8 * chunk && i < nFrames; i += chunk) gtts_stream_push_frames(s, &frames[i * 16], chunk, out.data(), (int64_t) out.size(), &n);
		gtts_stream_reset(s);
		const auto t0 = std::chrono::steady_clock::now();
		for (long i = 0; i + chunk <= nFrames; i += chunk) {
			if (gtts_stream_push_frames(s, &frames[i * 16], chunk, out.data(), (int64_t) out.size(), &n) != GTTS_OK) { std::fprintf(stderr, "%s\n", gtts_last_error()); return 1; }
			total += n;
		}
		gtts_stream_finish(s, out.data(), (int64_t) out.size(), &n);
		total += n;
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		const long pushes = nFrames / chunk;
		std::printf("%s\"chunk_%d\": {\"frames\": %ld, \"audio_seconds\": %.3f, \"wall_s\": %.4f, \"audio_s_per_s\": %.1f, \"us_per_push\": %.2f}",
				ci ? ", " : "", chunk, nFrames, total / 48000.0, sec, total / 48000.0 / sec, sec / pushes * 1e6);
		gtts_stream_close(s);
	}
	std::printf("}\n");
	gtts_destroy(h);
	return 0;
}
