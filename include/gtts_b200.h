/*
 * gtts_b200.h -- C ABI of the B200-native GamaTTS tube-model path.
 *
 * One drop-in boundary: batched synthesis of the reference's model-0 vocal-tract tube model
 * (GS::VTM::VocalTractModel0<double>) driven by control-parameter tracks, i.e. the work of
 *
 *     Controller::synthesize()                 gama_tts/src/vtm_control_model/Controller.cpp:277-313
 *       VocalTractModel::setAllParameters()    gama_tts/src/vtm/VocalTractModel.h:54, VocalTractModel0.h:698-716
 *       VocalTractModel::execSynthesisStep()   gama_tts/src/vtm/VocalTractModel.h:56, VocalTractModel0.h:396-445
 *     VocalTractModel::finishSynthesis()       gama_tts/src/vtm/VocalTractModel.h:57, VocalTractModel0.h:720-723
 *     VocalTractModel::outputBuffer()          gama_tts/src/vtm/VocalTractModel.h:59
 *
 * for U independent utterances at once on one GPU -- and, either side of it: the same for models 3 / 4 / 5 (tube_model,
 * gtts5_*), the output stage (peak normalisation + 16-bit PCM, *_pcm16), one batch over the GPUs of a box (gtts_multi_*,
 * gtts5_multi_*), frame-by-frame streaming (gtts_stream_*), and the control frames themselves from the rule engine's event
 * lists (EventList::generateOutput, gtts_events_*).  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary.  The C++ plugin shim that the unmodified reference loads through
 * VocalTractModelPlugin (gama_tts/src/vtm/VocalTractModelPlugin.cpp:40-48, model = 2000) is built on
 * top of these entry points (gama_tts_b200/csrc/plugin_shim.cpp); INTEGRATION.md shows the bindings.
 *
 * All functions return GTTS_OK (0) or a GTTS_ERR_* code; gtts_last_error() gives the text of the
 * last failure on the calling thread.  There is no CPU fallback: device entry points fail with
 * GTTS_ERR_NO_DEVICE / GTTS_ERR_CUDA when no sm_100 GPU is usable.
 */
#ifndef GTTS_B200_H_
#define GTTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTTS_NUM_PARAMS 16          /* control parameters per frame (artic.xml:38-53 order) */
#define GTTS_ABI_VERSION 2          /* 2: gtts_voice_config::tube_model */

enum {
	GTTS_OK = 0,
	GTTS_ERR_INVALID = 1,           /* bad argument / configuration value */
	GTTS_ERR_CUDA = 2,              /* a CUDA runtime call failed */
	GTTS_ERR_NO_DEVICE = 3,         /* no usable sm_100 device */
	GTTS_ERR_NOMEM = 4,
	GTTS_ERR_UNSUPPORTED = 5        /* valid in the reference but outside what this path implements */
};

/* Parameter indices of one control frame (VocalTractModel0.h:160-178). */
enum {
	GTTS_PARAM_GLOT_PITCH = 0, GTTS_PARAM_GLOT_VOL = 1, GTTS_PARAM_ASP_VOL = 2, GTTS_PARAM_FRIC_VOL = 3,
	GTTS_PARAM_FRIC_POS = 4, GTTS_PARAM_FRIC_CF = 5, GTTS_PARAM_FRIC_BW = 6, GTTS_PARAM_R1 = 7,
	GTTS_PARAM_R8 = 14, GTTS_PARAM_VELUM = 15
};

/* The configuration keys VocalTractModel0 reads (VocalTractModel0.h:266-305), same names, same units.
 * It is the merged vtm.txt + variant/<name>.txt map the reference's Controller builds
 * (Controller.cpp:48-49). */
typedef struct gtts_voice_config {
	double output_rate;                 /* Hz */
	int32_t waveform;                   /* 0 = glottal pulse, 1 = sine */
	int32_t noise_modulation;           /* 0 = off, 1 = on */
	double glottal_pulse_tp;            /* % */
	double glottal_pulse_tn_min;        /* % */
	double glottal_pulse_tn_max;        /* % */
	double breathiness;                 /* % */
	double vocal_tract_length_offset;   /* cm */
	double vocal_tract_length;          /* cm */
	double temperature;                 /* deg C */
	double loss_factor;                 /* % */
	double mouth_coefficient;
	double nose_coefficient;
	double throat_cutoff;               /* Hz */
	double throat_volume;               /* dB */
	double mix_offset;                  /* dB */
	double global_radius_coef;
	double global_nasal_radius_coef;
	double aperture_radius;             /* cm */
	double nasal_radius[5];             /* nasal_radius_1 .. nasal_radius_5, cm */
	double radius_coef[8];              /* radius_1_coef .. radius_8_coef */
	/* Which of the reference's models that read THIS key set the voice runs on (VocalTractModel.cpp:38-49):
	 *   0  models 0 and 2: VocalTractModel0<double> / VocalTractModel2<double, 1> -- the same arithmetic
	 *   3  model 3: VocalTractModel2<double, 3>, three samples of delay per section (VocalTractModel2.h:234-268, 626-670)
	 *   4  model 4: VocalTractModel4<double, 1>, 30 + 18 sections (VocalTractModel4.h:671-744)
	 * Models 3 and 4 run at three times the internal rate (60,102 Hz for 0_male) on the general kernel: batch entry
	 * points and the plugin seam; no streaming. */
	int32_t tube_model;
	int32_t reserved_;                  /* 0 */
} gtts_voice_config;

typedef struct gtts_handle gtts_handle;    /* one per GPU */
typedef struct gtts_batch gtts_batch;      /* a prepared batch (plan + device metadata) */
typedef struct gtts_stream gtts_stream;    /* one utterance fed control frame by control frame */

const char* gtts_last_error(void);
int gtts_abi_version(void);

/* ---- host-only helpers (no GPU touched) ----------------------------------------------------------- */

/* VocalTractModel::internalSampleRate() right after construction (VocalTractModel0.h:343-344). */
int gtts_voice_internal_rate(const gtts_voice_config* voice, int32_t* fs_out);
/* controlSteps of Controller::synthesize (Controller.cpp:286): rint(fs_int / control_rate). */
int gtts_voice_control_steps(const gtts_voice_config* voice, double control_rate, int32_t* steps_out);
/* Number of internal samples and of output samples (outputBuffer().size() after finishSynthesis())
 * for a track of n_frames control frames stepped `steps` times each
 * (SampleRateConverter.h:268-282, 295-416, 462-471 reduced to a closed form). */
int gtts_output_length(const gtts_voice_config* voice, int32_t steps, int64_t n_frames,
			int64_t* n_internal_out, int64_t* n_output_out);
/* Greedy longest-first partition of utterances over n_shards GPUs (no collective: SURVEY.md s.8e).
 * cost[u] is any positive work estimate (e.g. internal samples); shard_of[u] receives 0..n_shards-1. */
int gtts_shard_plan(const int64_t* cost, int64_t n_utt, int32_t n_shards, int32_t* shard_of);
/* Known-answer probes of the host-side table builders (used by the tests; cheap). */
int gtts_probe_fir_taps(double* taps, int32_t cap, int32_t* n_taps_out);
int gtts_probe_src_tables(double* h3328, double* dh3328);
int gtts_probe_voice_constants(const gtts_voice_config* voice, double* out, int32_t cap, int32_t* n_out);

/* ---- device ---------------------------------------------------------------------------------------- */

int gtts_create(int32_t device, gtts_handle** handle_out);
void gtts_destroy(gtts_handle* handle);
/* Device and kernel facts for logs (SM count, kernel variant, shared memory per CTA ...), JSON text. */
const char* gtts_describe(gtts_handle* handle);

/* Measures this GPU's FP64 FMA-pipe peak (TFLOP/s) with a register-resident DFMA kernel (~30 ms):
 * the roofline denominator of this path (MEASURED_PEAKS.json only carries HBM and bf16 figures). */
int gtts_probe_fp64_peak(gtts_handle* handle, double* tflops_out);
/* Test hook: the kernels' 2^x and 10^x (the reference's pow(2, .) / pow(10, .) in VTMUtil.h:50-84, valid for |x| < 16)
 * evaluated on the device for x[n] (host arrays). */
int gtts_probe_exp(gtts_handle* handle, const double* x, int32_t n, double* exp2_out, double* exp10_out);

/* Plans a batch of n_utt utterances.
 *   voices[n_voices]        voice table; voice_index[u] selects one (NULL: every utterance uses voices[0])
 *   control_rate            Hz (1000 / control_period of vtm_control_model.txt); steps = rint(fs_int/rate)
 *   steps_override          NULL, or per-utterance controlSteps (> 0 overrides; 1 = "every frame is one
 *                           internal sample", the mode the plugin shim uses)
 *   frame_offsets[n_utt+1]  utterance u owns frames [frame_offsets[u], frame_offsets[u+1]) of the packed
 *                           float32 [n_frames_total][16] track array
 * All arrays are host memory and are copied. */
int gtts_batch_prepare(gtts_handle* handle, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts_batch** batch_out);
/* Layout of the output buffer, in float32 samples: utterance u occupies [out_offsets[u], out_offsets[u] +
 * n_out[u]); out_offsets[n_utt] is the size of the whole buffer.  Every utterance starts on a multiple of
 * 64 samples (256 bytes) so that the kernel writes whole aligned rows -- this is what lets it store straight
 * into pinned host memory at PCIe rate; the up to 63 samples between two utterances are never written.
 * n_internal[n_utt] (samples at the tube's internal rate) may be NULL. */
int gtts_batch_layout(const gtts_batch* batch, int64_t* out_offsets, int64_t* n_internal);
/* n_out[n_utt]: output samples of each utterance (== gtts_output_length of its voice, steps and frames). */
int gtts_batch_lengths(const gtts_batch* batch, int64_t* n_out);
/* Runs the batch on DEVICE buffers, asynchronously on `cuda_stream` (a cudaStream_t, may be NULL):
 * d_frames float32 [n_frames_total][16], d_out float32 [out_offsets[n_utt]]. */
int gtts_batch_run_device(gtts_batch* batch, const float* d_frames, float* d_out, void* cuda_stream);
/* Same with HOST buffers, synchronised on return.  Pageable buffers are staged: host->device copy of the frames,
 * kernel, device->host copy of the audio.  Pinned (page-locked, e.g. cudaHostAlloc) buffers are used in place: the
 * kernel reads each frame over PCIe one control period ahead of its use and stores the audio straight into h_out
 * in whole 128-byte rows while it computes, so the transfers overlap the synthesis (h_out should be 128-byte
 * aligned, which pinned allocations are).  The samples between two utterances (see gtts_batch_layout) are not
 * written. */
int gtts_batch_run_host(gtts_batch* batch, const float* h_frames, float* h_out);
/* The reference's output stage on the device (BASELINE next row 2): per utterance the peak-normalisation scale
 * 0.95f / max|x| (0 below 1e-30f: VTM::Util::calculateOutputScale, gama_tts/src/vtm/VTMUtil.cpp:20-21, 48-57) and the
 * 16-bit PCM payload Controller::writeOutputToFile + WAVEFileWriter::writeSample would put into the WAVE file
 * (gama_tts/src/vtm_control_model/Controller.cpp:315-328, gama_tts/src/WAVEFileWriter.cpp:36-37, 122-126): bit-identical
 * to the reference's for the same float32 audio.  pcm has the layout of the float32 output (gtts_batch_layout, in
 * samples); scale[n_utt] may be NULL.  The float32 audio is kept in device memory (d_audio: a scratch buffer of
 * out_offsets[n_utt] floats for the device call), only the 16-bit payload -- half the bytes -- goes to the host. */
int gtts_batch_run_device_pcm16(gtts_batch* batch, const float* d_frames, float* d_audio, int16_t* d_pcm,
			float* d_scale, void* cuda_stream);
int gtts_batch_run_host_pcm16(gtts_batch* batch, const float* h_frames, int16_t* h_pcm, float* h_scale);
/* The same without waiting: everything (frame copy if the frames are pageable, synthesis, output stage, device->host
 * copy of the payload) is queued on the batch's own stream and the call returns; gtts_batch_wait() blocks until it
 * is done.  With two prepared batches a host keeps the GPU busy: while batch A's payload travels to the host (copy
 * engine), batch B is being synthesised -- a serving loop's steady state is then max(synthesis, transfer) per batch
 * instead of their sum.  The host buffers must stay valid (and, to overlap, be pinned) until the wait returns. */
int gtts_batch_submit_host_pcm16(gtts_batch* batch, const float* h_frames, int16_t* h_pcm, float* h_scale);
int gtts_batch_wait(gtts_batch* batch);
/* Order-independent 64-bit checksum per utterance of float32 audio in device memory (batch layout): what the
 * multi-GPU determinism checks of BASELINE config 4 compare instead of the audio. d_sums[n_utt] is device memory. */
int gtts_batch_checksum_device(gtts_batch* batch, const float* d_audio, uint64_t* d_sums, void* cuda_stream);
/* Number of kernel launches the last run issued (for bench.py's gpu_launches claim). */
int gtts_batch_last_launches(const gtts_batch* batch, int32_t* n_out);
/* Determinism: a run is bit-reproducible, and an utterance's audio does not depend on what else is in the batch as long as
 * the same kernel synthesises it.  Two pipelined kernels exist (same operations, different multiply-adds contracted: their
 * outputs differ by <= 1e-7 of full scale, the reference's own FMA on / off noise floor): a batch whose utterances all have
 * one voice and one length takes tube_kernel_v1, any other tube_kernel_v2; the environment variable GTTS_KERNEL=v2 (or v1)
 * pins one for every batch.  The shards of a gtts_multi batch take the whole batch's choice.
 * Name of the synthesis kernel(s) the last run launched ("tube_kernel_v1": batches of one voice and one length,
 * "tube_kernel_v2": ragged / mixed-voice batches, "tube_kernel_v0": short control periods, models 3 / 4, streams), for logs. */
const char* gtts_batch_last_kernel(const gtts_batch* batch);
void gtts_batch_free(gtts_batch* batch);

/* One-call convenience over prepare + run_host + free (what a C caller of the reference's
 * Controller::synthesize + outputBuffer() would use).  Utterance u is written at out[out_offsets[u]] with
 * gtts_output_length() samples; every utterance starts on a multiple of 64 samples (see gtts_batch_layout), so
 * out_capacity must be at least the sum over the utterances of their length rounded up to a multiple of 64 (the call
 * fails with "output buffer too small" otherwise and still fills out_offsets, whose last entry is the size needed).
 * out_offsets[n_utt + 1] is filled if not NULL. */
int gtts_batch_synthesize(gtts_handle* handle, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const float* frames,
			const int64_t* frame_offsets, int64_t n_utt, float* out, int64_t out_capacity,
			int64_t* out_offsets);

/* ---- one batch over the GPUs of a box (BASELINE config 4) ------------------------------------------------
 * The batch is partitioned by utterance (gtts_shard_plan on internal + output samples); every GPU synthesises its
 * utterances from / into the caller's ONE packed frame array and ONE output buffer (utterance u at out_offsets[u],
 * as in gtts_batch_layout), driven by one host thread per GPU.  No collective and no inter-GPU traffic: utterances are
 * independent, and the result is bit for bit what a single GPU produces.  Host buffers
 * should be pinned (they are read / written in place by all GPUs); pageable buffers are pinned for the call. */
typedef struct gtts_multi gtts_multi;
typedef struct gtts_multi_batch gtts_multi_batch;
int gtts_multi_create(const int32_t* devices, int32_t n_devices, gtts_multi** multi_out);
void gtts_multi_destroy(gtts_multi* multi);
int32_t gtts_multi_device_count(const gtts_multi* multi);
int gtts_multi_batch_prepare(gtts_multi* multi, const gtts_voice_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts_multi_batch** batch_out);
/* out_offsets[n_utt + 1], n_out[n_utt], shard_of[n_utt] (which GPU of the list got the utterance); any may be NULL */
int gtts_multi_batch_layout(const gtts_multi_batch* batch, int64_t* out_offsets, int64_t* n_out, int32_t* shard_of);
int gtts_multi_batch_run_host(gtts_multi_batch* batch, const float* h_frames, float* h_out);
int gtts_multi_batch_run_host_pcm16(gtts_multi_batch* batch, const float* h_frames, int16_t* h_pcm, float* h_scale);
void gtts_multi_batch_free(gtts_multi_batch* batch);

/* ---- streaming (one utterance, control frame by control frame; BASELINE config 5) ------------------ */

int gtts_stream_open(gtts_handle* handle, const gtts_voice_config* voice, double control_rate,
			int32_t steps_override, gtts_stream** stream_out);
/* Appends n_frames frames (host float32 [n_frames][16]).  Control periods whose end frame is known
 * are synthesised now; audio that is final is written to out (host, capacity in samples). */
int gtts_stream_push_frames(gtts_stream* stream, const float* frames, int64_t n_frames,
			float* out, int64_t out_capacity, int64_t* n_written);
/* Last control period (duplicated final frame, Controller.cpp:283) + SRC flush (finishSynthesis()). */
int gtts_stream_finish(gtts_stream* stream, float* out, int64_t out_capacity, int64_t* n_written);
/* Back to the state right after open (VocalTractModel::reset(), VocalTractModel0.h:309-326). */
int gtts_stream_reset(gtts_stream* stream);
void gtts_stream_close(gtts_stream* stream);

/* ---- model 5 (BASELINE next row 1) -------------------------------------------------------------------------
 * The same boundary for the reference's model 5, GS::VTM::VocalTractModel5<double, 1> -- the voice directories
 * data/voice/english/5_xxx, which the reference's documentation uses by default:
 *     setAllParameters / execSynthesisStep / finishSynthesis      gama_tts/src/vtm/VocalTractModel5.h:776-792, 527-582, 794-797
 *     (Rosenberg-B source RosenbergBGlottalSource.h:112-150, Butterworth filters, pole-zero radiation impedance
 *     PoleZeroRadiationImpedance.h:143-189, 30 + 21 flow-equation sections :646-730, down-sampling converter
 *     SampleRateConverter.h:362-415, float difference filter * output rate :506-512)
 * driven by the same Controller::synthesize interpolation (Controller.cpp:277-313).  The control frames are the same
 * 16 parameters.  One warp per utterance (gama_tts_b200/csrc/tube5_kernel.cuh).  Internal rates from 50 kHz (below it the
 * reference's radiation impedance refuses to construct) up to about 170 kHz (tracts down to 6.2 cm at 35 deg C: a
 * converter wing of at most 48 taps) are accepted; the output rate must be below the internal rate. */
typedef struct gtts_voice5_config {
	double output_rate;                 /* Hz */
	int32_t waveform;                   /* 0 = Rosenberg-B pulse, 1 = sine */
	int32_t noise_modulation;
	int32_t bypass;                     /* 1 = glottal waveform only */
	int32_t constant_radius_mouth_impedance;
	double glottal_pulse_tp;            /* % */
	double glottal_pulse_tn_min;        /* % */
	double glottal_pulse_tn_max;        /* % */
	double breathiness;                 /* % */
	double vocal_tract_length_offset;   /* cm */
	double vocal_tract_length;          /* cm */
	double temperature;                 /* deg C */
	double loss_factor;                 /* % */
	double mix_offset;                  /* dB */
	double global_radius_coef;
	double global_nasal_radius_coef;
	double nasal_radius[6];             /* nasal_radius_2 .. nasal_radius_7, cm */
	double radius_coef[8];              /* radius_1_coef .. radius_8_coef */
	double glottal_noise_cutoff;        /* Hz */
	double frication_noise_cutoff;      /* Hz */
	double frication_factor;
	double min_glottal_loss;            /* % */
	double max_glottal_loss;            /* % */
	double glottal_lowpass_cutoff;      /* Hz */
	double mouth_impedance_radius;      /* cm, with constant_radius_mouth_impedance */
} gtts_voice5_config;

typedef struct gtts5_batch gtts5_batch;

/* VocalTractModel5::internalSampleRate() (VocalTractModel5.h:464-465): a double for this model. */
int gtts5_voice_internal_rate(const gtts_voice5_config* voice, double* fs_out);
/* controlSteps (steps <= 0: rint(fs / control_rate), Controller.cpp:286), internal and output samples of a track. */
int gtts5_output_length(const gtts_voice5_config* voice, double control_rate, int32_t steps, int64_t n_frames,
			int32_t* steps_out, int64_t* n_internal_out, int64_t* n_output_out);
/* Arguments as gtts_batch_prepare. */
int gtts5_batch_prepare(gtts_handle* handle, const gtts_voice5_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts5_batch** batch_out);
/* out_offsets[n_utt + 1] (every utterance starts on a multiple of 64 samples), n_out[n_utt], n_internal[n_utt]; any may be NULL */
int gtts5_batch_layout(const gtts5_batch* batch, int64_t* out_offsets, int64_t* n_out, int64_t* n_internal);
/* Device buffers, asynchronous on cuda_stream. */
int gtts5_batch_run_device(gtts5_batch* batch, const float* d_frames, float* d_out, void* cuda_stream);
/* Host buffers (staged through device memory), synchronised on return. */
int gtts5_batch_run_host(gtts5_batch* batch, const float* h_frames, float* h_out);
/* The reference's output stage (see gtts_batch_run_device_pcm16): per-utterance scale and 16-bit PCM payload. */
int gtts5_batch_run_device_pcm16(gtts5_batch* batch, const float* d_frames, float* d_audio, int16_t* d_pcm,
			float* d_scale, void* cuda_stream);
int gtts5_batch_run_host_pcm16(gtts5_batch* batch, const float* h_frames, int16_t* h_pcm, float* h_scale);
void gtts5_batch_free(gtts5_batch* batch);

/* Model 5 over the GPUs of a box (gtts_multi_create above): the batch partitioned by utterance as in gtts_multi_batch_*, one
 * host thread per GPU, every utterance's audio (or 16-bit payload, with its scale) at its offset of the caller's ONE buffer
 * (layout as gtts5_batch_layout of the whole batch).  Bit for bit the result of one GPU. */
typedef struct gtts5_multi_batch gtts5_multi_batch;
int gtts5_multi_batch_prepare(gtts_multi* multi, const gtts_voice5_config* voices, int32_t n_voices,
			const int32_t* voice_index, double control_rate, const int32_t* steps_override,
			const int64_t* frame_offsets, int64_t n_utt, gtts5_multi_batch** batch_out);
int gtts5_multi_batch_layout(const gtts5_multi_batch* batch, int64_t* out_offsets, int64_t* n_out, int32_t* shard_of);
int gtts5_multi_batch_run_host(gtts5_multi_batch* batch, const float* h_frames, float* h_out);
int gtts5_multi_batch_run_host_pcm16(gtts5_multi_batch* batch, const float* h_frames, int16_t* h_pcm, float* h_scale);
void gtts5_multi_batch_free(gtts5_multi_batch* batch);

/* ---- control-frame generation on the device (BASELINE next row 3) -------------------------------------------------
 * From the event list of a chunk of an utterance to its control frames -- the float32 [frame][16] array every batch
 * call above takes -- in device memory, so that the frames never cross PCIe:
 *     EventList::generateOutput      gama_tts/src/vtm_control_model/EventList.cpp:929-1091
 *     DriftGenerator::drift          gama_tts/src/vtm_control_model/DriftGenerator.cpp:72-84 (its low-pass:
 *                                    gama_tts/src/vtm/Butterworth2LowpassFilter.h:104-113)
 * What stays on the host is what builds the event list (text parser, rule engine, EventList::generateEventList and
 * applyIntonation): the binding copies EventList::list_ into gtts_event records.  Output: bit-identical
 * to the reference's frames (float32), IEEE double arithmetic in its order of evaluation without FMA contraction. */
typedef struct gtts_event {             /* Event, gama_tts/src/vtm_control_model/EventList.h:113-161 */
	int32_t time;                       /* ms from the start of the chunk */
	int32_t has_interp;                 /* interpData present (macro-intonation polynomial of the segment that starts here) */
	double param[16];                   /* parameters; +infinity = Event::EMPTY_PARAMETER */
	double special[16];                 /* specialParameters, same convention */
	double a, b, c, d;                  /* interpData->a .. d */
} gtts_event;

typedef struct gtts_event_config {      /* what generateOutput reads besides the events */
	int32_t control_period;             /* ms (vtm_control_model.txt: control_period) */
	int32_t macro_intonation, micro_intonation, intonation_drift, smooth_intonation;   /* EventList.h:182-192 */
	int32_t reserved;
	double initial_pitch, mean_pitch;   /* EventList::setInitialPitch / setMeanPitch */
	double drift_deviation2, drift_offset;   /* DriftGenerator::setUp: 2 deviation, deviation */
	double drift_seed;                  /* the generator's state: the chaotic seed ... */
	double drift_b0, drift_b1, drift_a1, drift_a2;   /* ... its Butterworth-2 low-pass (Butterworth2LowpassFilter::update) ... */
	double drift_x1, drift_x2, drift_y1, drift_y2;   /* ... and the filter's history */
} gtts_event_config;

typedef struct gtts_events_batch gtts_events_batch;

/* Fills the drift_* members of *config as a fresh generator has them: DriftGenerator::DriftGenerator + setUp(deviation,
 * sample_rate, lowpass_cutoff) (DriftGenerator.cpp:23-34, 55-63; Butterworth2LowPassFilter::update,
 * Butterworth2LowpassFilter.h:80-100) -- called by EventList::setUpDriftGenerator with drift_deviation, the control rate and
 * drift_lowpass_cutoff of vtm_control_model.txt (Controller.cpp:73).  The cutoff must lie in [1, 0.48 sample_rate]. */
int gtts_events_drift_setup(double deviation, double sample_rate, double lowpass_cutoff, gtts_event_config* config);
/* Frames generateOutput makes of an event list (host arithmetic on the event times only; 0 for fewer than 2 events). */
int gtts_events_frame_count(const gtts_event_config* config, const gtts_event* events, int64_t n_events, int64_t* n_frames_out);
/* Plans a batch of n_chunks chunks.
 *   configs[n_chunks]            one per chunk
 *   continues_previous[n_chunks] NULL, or nonzero where chunk c is the next chunk of the utterance of chunk c - 1: its drift
 *                                generator goes on from the state chunk c - 1 left (the drift_* state in configs[c] is
 *                                ignored), as Controller::getParametersFromPhoneticString does (Controller.cpp:119-156)
 *   events, event_offsets        chunk c owns events [event_offsets[c], event_offsets[c + 1]) of the packed host array;
 *                                only the times are read here (the frame layout follows from them)
 * All arrays are host memory and are copied. */
int gtts_events_prepare(gtts_handle* handle, const gtts_event_config* configs, const int32_t* continues_previous,
			const gtts_event* events, const int64_t* event_offsets, int64_t n_chunks, gtts_events_batch** batch_out);
/* frame_offsets[n_chunks + 1]: chunk c produces frames [frame_offsets[c], frame_offsets[c + 1]) of the packed frame array --
 * with the entries of the utterances' first chunks, the frame_offsets argument of gtts_batch_prepare. */
int gtts_events_layout(const gtts_events_batch* batch, int64_t* frame_offsets);
/* Device buffers, asynchronous on cuda_stream: d_events gtts_event [event_offsets[n_chunks]] (8-byte aligned),
 * d_frames float32 [frame_offsets[n_chunks]][16]; d_configs_out NULL or gtts_event_config [n_chunks] receiving the configs with
 * the drift generator's state as each chunk left it.  A following gtts_batch_run_device on the same stream reads d_frames. */
int gtts_events_run_device(gtts_events_batch* batch, const gtts_event* d_events, float* d_frames,
			gtts_event_config* d_configs_out, void* cuda_stream);
/* Host buffers (staged), synchronised on return; configs_out may be NULL. */
int gtts_events_run_host(gtts_events_batch* batch, const gtts_event* h_events, float* h_frames, gtts_event_config* configs_out);
void gtts_events_free(gtts_events_batch* batch);

#ifdef __cplusplus
}
#endif
#endif /* GTTS_B200_H_ */
