"""gama_tts_b200 -- B200-native batched GamaTTS tube model (reference model 0) behind a C ABI.

The product is ``csrc/libgtts_b200.so`` (hand-written sm_100a kernels + C++ host runtime, ABI in
``include/gtts_b200.h``) and the plugin shim ``csrc/libgtts_plugin.so`` the unmodified reference
loads through its own ``VocalTractModelPlugin`` seam.  This package is the thin Python view used by
the tests and the benchmark: it mirrors the reference's ``VocalTractModel`` call pattern
(``gama_tts/src/vtm/VocalTractModel.h:43-71``) for batches:

    synth = TubeSynthesizer(device=0)
    audio = synth.synthesize(voice, [track0, track1, ...])      # Controller::synthesize + outputBuffer()

There is no CPU implementation here; importing works without a GPU, constructing a synthesizer
does not.
"""
import ctypes as C

import numpy as np

from . import capi, voices  # noqa: F401
from .capi import GttsError, check, load, voice_array, voice_config

NUM_PARAMS = 16


def _as_tracks(tracks):
    out = []
    for t in tracks:
        a = np.ascontiguousarray(t, dtype=np.float32).reshape(-1, NUM_PARAMS)
        out.append(a)
    return out


def pack_tracks(tracks):
    """list of [F_u, 16] float32 -> (frames [sum F_u, 16], frame_offsets int64 [U+1])."""
    tracks = _as_tracks(tracks)
    fo = np.zeros(len(tracks) + 1, np.int64)
    if tracks:
        fo[1:] = np.cumsum([len(t) for t in tracks])
    frames = np.concatenate(tracks) if tracks else np.zeros((0, NUM_PARAMS), np.float32)
    return np.ascontiguousarray(frames, np.float32), fo


def internal_rate(voice):
    v = voice_config(voice)
    fs = C.c_int32()
    check(load().gtts_voice_internal_rate(C.byref(v), C.byref(fs)))
    return fs.value


def control_steps(voice, control_rate=voices.DEFAULT_CONTROL_RATE):
    v = voice_config(voice)
    st = C.c_int32()
    check(load().gtts_voice_control_steps(C.byref(v), float(control_rate), C.byref(st)))
    return st.value


def output_length(voice, n_frames, control_rate=voices.DEFAULT_CONTROL_RATE, steps=None):
    """(n_internal, n_output) of a track of n_frames control frames."""
    v = voice_config(voice)
    if steps is None:
        steps = control_steps(voice, control_rate)
    ni, no = C.c_int64(), C.c_int64()
    check(load().gtts_output_length(C.byref(v), int(steps), int(n_frames), C.byref(ni), C.byref(no)))
    return ni.value, no.value


def shard_plan(cost, n_shards):
    """Greedy longest-first assignment of utterances to GPUs; returns int32 [U] shard ids."""
    cost = np.ascontiguousarray(cost, np.int64)
    out = np.zeros(len(cost), np.int32)
    check(load().gtts_shard_plan(cost.ctypes.data, len(cost), int(n_shards), out.ctypes.data))
    return out


class Batch:
    """A prepared batch (gtts_batch): plan + device metadata; run it as often as you like."""

    def __init__(self, synth, voice_list, n_utt, frame_offsets, voice_index, control_rate, steps_override):
        self._lib = load()
        self._synth = synth
        self._h = C.c_void_p()
        va = voice_array(voice_list)
        fo = np.ascontiguousarray(frame_offsets, np.int64)
        vi = None if voice_index is None else np.ascontiguousarray(voice_index, np.int32)
        so = None if steps_override is None else np.ascontiguousarray(steps_override, np.int32)
        check(self._lib.gtts_batch_prepare(synth._h, va, len(voice_list), None if vi is None else vi.ctypes.data,
                                           float(control_rate), None if so is None else so.ctypes.data,
                                           fo.ctypes.data, int(n_utt), C.byref(self._h)))
        self.n_utt = int(n_utt)
        self.frame_offsets = fo
        self.out_offsets = np.zeros(self.n_utt + 1, np.int64)
        self.n_internal = np.zeros(max(self.n_utt, 1), np.int64)
        check(self._lib.gtts_batch_layout(self._h, self.out_offsets.ctypes.data, self.n_internal.ctypes.data))
        self.n_internal = self.n_internal[:self.n_utt]
        self.n_out = np.zeros(max(self.n_utt, 1), np.int64)
        check(self._lib.gtts_batch_lengths(self._h, self.n_out.ctypes.data))
        self.n_out = self.n_out[:self.n_utt]
        self.n_out_total = int(self.out_offsets[-1])     # size of the output buffer (utterances start on 64-sample row pairs)
        self.n_samples_total = int(self.n_out.sum())     # audio samples produced
        self.n_frames_total = int(fo[-1]) if len(fo) else 0

    def run_device(self, d_frames_ptr, d_out_ptr, stream_ptr=0):
        check(self._lib.gtts_batch_run_device(self._h, C.c_void_p(d_frames_ptr), C.c_void_p(d_out_ptr),
                                              C.c_void_p(stream_ptr)))

    def run_device_pcm16(self, d_frames_ptr, d_audio_ptr, d_pcm_ptr, d_scale_ptr=0, stream_ptr=0):
        check(self._lib.gtts_batch_run_device_pcm16(self._h, C.c_void_p(d_frames_ptr), C.c_void_p(d_audio_ptr), C.c_void_p(d_pcm_ptr),
                                                    C.c_void_p(d_scale_ptr), C.c_void_p(stream_ptr)))

    def run_host(self, frames, out=None):
        frames = np.ascontiguousarray(frames, np.float32)
        if out is None:
            out = np.empty(self.n_out_total, np.float32)
        check(self._lib.gtts_batch_run_host(self._h, frames.ctypes.data, out.ctypes.data))
        return out

    def run_host_pcm16(self, frames, want_scale=True):
        """The reference's output stage on the device: (int16 payload in the batch layout, scale per utterance)."""
        frames = np.ascontiguousarray(frames, np.float32)
        pcm = np.zeros(self.n_out_total, np.int16)
        scale = np.zeros(max(self.n_utt, 1), np.float32) if want_scale else None
        check(self._lib.gtts_batch_run_host_pcm16(self._h, frames.ctypes.data, pcm.ctypes.data,
                                                  None if scale is None else scale.ctypes.data))
        return pcm, (None if scale is None else scale[:self.n_utt])

    def run_host_pcm16_ptr(self, frames_ptr, pcm_ptr, scale_ptr=0):
        check(self._lib.gtts_batch_run_host_pcm16(self._h, C.c_void_p(frames_ptr), C.c_void_p(pcm_ptr), C.c_void_p(scale_ptr)))

    def submit_host_pcm16_ptr(self, frames_ptr, pcm_ptr, scale_ptr=0):
        """Queues the whole pipeline on the batch's stream and returns; wait() blocks until it is done."""
        check(self._lib.gtts_batch_submit_host_pcm16(self._h, C.c_void_p(frames_ptr), C.c_void_p(pcm_ptr), C.c_void_p(scale_ptr)))

    def wait(self):
        check(self._lib.gtts_batch_wait(self._h))

    def run_host_ptr(self, frames_ptr, out_ptr):
        check(self._lib.gtts_batch_run_host(self._h, C.c_void_p(frames_ptr), C.c_void_p(out_ptr)))

    def last_launches(self):
        n = C.c_int32()
        check(self._lib.gtts_batch_last_launches(self._h, C.byref(n)))
        return n.value

    def last_kernel(self):
        return self._lib.gtts_batch_last_kernel(self._h).decode()

    def split(self, packed):
        return [packed[self.out_offsets[u]:self.out_offsets[u] + self.n_out[u]] for u in range(self.n_utt)]

    def close(self):
        if self._h:
            self._lib.gtts_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stream:
    """One utterance fed control frame by control frame (gtts_stream)."""

    def __init__(self, synth, voice, control_rate, steps_override=0):
        self._lib = load()
        self._synth = synth
        self._h = C.c_void_p()
        v = voice_config(voice)
        check(self._lib.gtts_stream_open(synth._h, C.byref(v), float(control_rate), int(steps_override), C.byref(self._h)))
        self._fs = internal_rate(voice)
        self._ratio = voice["output_rate"] / self._fs
        self._steps = steps_override if steps_override > 0 else control_steps(voice, control_rate)

    def _cap(self, n_frames):
        return int((n_frames * self._steps + 64) * self._ratio) + 256

    def push(self, frames):
        frames = np.ascontiguousarray(frames, np.float32).reshape(-1, NUM_PARAMS)
        out = np.empty(self._cap(len(frames) + 1), np.float32)
        n = C.c_int64()
        check(self._lib.gtts_stream_push_frames(self._h, frames.ctypes.data, len(frames), out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value].copy()

    def finish(self):
        out = np.empty(self._cap(2), np.float32)
        n = C.c_int64()
        check(self._lib.gtts_stream_finish(self._h, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value].copy()

    def reset(self):
        check(self._lib.gtts_stream_reset(self._h))

    def close(self):
        if self._h:
            self._lib.gtts_stream_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiSynthesizer:
    """One batch over several GPUs of a box (gtts_multi): partitioned by utterance, no collective, one caller buffer."""

    def __init__(self, devices):
        self._lib = load()
        self._h = C.c_void_p()
        dev = np.ascontiguousarray(list(devices), np.int32)
        check(self._lib.gtts_multi_create(dev.ctypes.data, len(dev), C.byref(self._h)))
        self.devices = dev.tolist()

    def synthesize(self, voice_or_voices, tracks, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                   steps_override=None, pcm16=False, return_shards=False):
        vl = [voice_or_voices] if isinstance(voice_or_voices, dict) else list(voice_or_voices)
        frames, fo = pack_tracks(tracks)
        va = voice_array(vl)
        vi = None if voice_index is None else np.ascontiguousarray(voice_index, np.int32)
        so = None if steps_override is None else np.ascontiguousarray(steps_override, np.int32)
        b = C.c_void_p()
        n = len(tracks)
        check(self._lib.gtts_multi_batch_prepare(self._h, va, len(vl), None if vi is None else vi.ctypes.data, float(control_rate),
                                                 None if so is None else so.ctypes.data, fo.ctypes.data, n, C.byref(b)))
        try:
            oo = np.zeros(n + 1, np.int64)
            no = np.zeros(max(n, 1), np.int64)
            sh = np.zeros(max(n, 1), np.int32)
            check(self._lib.gtts_multi_batch_layout(b, oo.ctypes.data, no.ctypes.data, sh.ctypes.data))
            if pcm16:
                out = np.zeros(int(oo[-1]), np.int16)
                scale = np.zeros(max(n, 1), np.float32)
                check(self._lib.gtts_multi_batch_run_host_pcm16(b, frames.ctypes.data, out.ctypes.data, scale.ctypes.data))
                res = ([out[oo[u]:oo[u] + no[u]].copy() for u in range(n)], scale[:n])
            else:
                out = np.zeros(int(oo[-1]), np.float32)
                check(self._lib.gtts_multi_batch_run_host(b, frames.ctypes.data, out.ctypes.data))
                res = [out[oo[u]:oo[u] + no[u]].copy() for u in range(n)]
            return (res, sh[:n].copy()) if return_shards else res
        finally:
            self._lib.gtts_multi_batch_free(b)

    def synthesize5(self, voice_or_voices, tracks, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                    steps_override=None, pcm16=False, return_shards=False):
        """Model-5 voices over the GPUs (gtts5_multi_batch_*): same results as TubeSynthesizer.synthesize5, bit for bit."""
        from .capi import voice5_array
        vl = [voice_or_voices] if isinstance(voice_or_voices, dict) else list(voice_or_voices)
        frames, fo = pack_tracks(tracks)
        va = voice5_array(vl)
        vi = None if voice_index is None else np.ascontiguousarray(voice_index, np.int32)
        so = None if steps_override is None else np.ascontiguousarray(steps_override, np.int32)
        b = C.c_void_p()
        n = len(tracks)
        check(self._lib.gtts5_multi_batch_prepare(self._h, va, len(vl), None if vi is None else vi.ctypes.data, float(control_rate),
                                                  None if so is None else so.ctypes.data, fo.ctypes.data, n, C.byref(b)))
        try:
            oo = np.zeros(n + 1, np.int64)
            no = np.zeros(max(n, 1), np.int64)
            sh = np.zeros(max(n, 1), np.int32)
            check(self._lib.gtts5_multi_batch_layout(b, oo.ctypes.data, no.ctypes.data, sh.ctypes.data))
            if pcm16:
                out = np.zeros(int(oo[-1]) + 1, np.int16)
                scale = np.zeros(max(n, 1), np.float32)
                check(self._lib.gtts5_multi_batch_run_host_pcm16(b, frames.ctypes.data, out.ctypes.data, scale.ctypes.data))
                res = ([out[oo[u]:oo[u] + no[u]].copy() for u in range(n)], scale[:n])
            else:
                out = np.zeros(int(oo[-1]) + 1, np.float32)
                check(self._lib.gtts5_multi_batch_run_host(b, frames.ctypes.data, out.ctypes.data))
                res = [out[oo[u]:oo[u] + no[u]].copy() for u in range(n)]
            return (res, sh[:n].copy()) if return_shards else res
        finally:
            self._lib.gtts5_multi_batch_free(b)

    def close(self):
        if self._h:
            self._lib.gtts_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TubeSynthesizer:
    """One GPU's worth of the tube path (gtts_handle)."""

    def __init__(self, device=0):
        self._lib = load()
        self._h = C.c_void_p()
        check(self._lib.gtts_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def describe(self):
        return self._lib.gtts_describe(self._h).decode()

    def fp64_peak_tflops(self):
        t = C.c_double()
        check(self._lib.gtts_probe_fp64_peak(self._h, C.byref(t)))
        return t.value

    def probe_exp(self, x):
        """The kernels' 2^x and 10^x evaluated on the device (test hook)."""
        x = np.ascontiguousarray(x, np.float64)
        e2, e10 = np.empty_like(x), np.empty_like(x)
        check(self._lib.gtts_probe_exp(self._h, x.ctypes.data, len(x), e2.ctypes.data, e10.ctypes.data))
        return e2, e10

    def prepare(self, voice_or_voices, frame_offsets, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                steps_override=None):
        vl = [voice_or_voices] if isinstance(voice_or_voices, dict) else list(voice_or_voices)
        return Batch(self, vl, len(frame_offsets) - 1, frame_offsets, voice_index, control_rate, steps_override)

    def synthesize(self, voice_or_voices, tracks, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                   steps_override=None):
        """Host-buffer path: tracks (list of [F,16] float32) -> list of float32 audio arrays, the raw
        outputBuffer() of each utterance (before the reference's peak normalisation)."""
        frames, fo = pack_tracks(tracks)
        b = self.prepare(voice_or_voices, fo, voice_index, control_rate, steps_override)
        try:
            return [a.copy() for a in b.split(b.run_host(frames))]
        finally:
            b.close()

    def synthesize_pcm16(self, voice_or_voices, tracks, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE):
        """tracks -> (list of int16 arrays: the WAVE payload the reference would write, float32 scales)."""
        frames, fo = pack_tracks(tracks)
        b = self.prepare(voice_or_voices, fo, voice_index, control_rate)
        try:
            pcm, scale = b.run_host_pcm16(frames)
            return [a.copy() for a in b.split(pcm)], scale
        finally:
            b.close()

    def stream(self, voice, control_rate=voices.DEFAULT_CONTROL_RATE, steps_override=0):
        return Stream(self, voice, control_rate, steps_override)

    # ---- model 5 (gtts5_*: VocalTractModel5<double, 1>) ----
    def prepare5(self, voice_or_voices, frame_offsets, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                 steps_override=None):
        vl = [voice_or_voices] if isinstance(voice_or_voices, dict) else list(voice_or_voices)
        return Batch5(self, vl, frame_offsets, voice_index, control_rate, steps_override)

    def synthesize5(self, voice_or_voices, tracks, voice_index=None, control_rate=voices.DEFAULT_CONTROL_RATE,
                    steps_override=None):
        """Model-5 voices (voices.default_voice5): tracks -> list of float32 audio arrays (raw outputBuffer())."""
        frames, fo = pack_tracks(tracks)
        b = self.prepare5(voice_or_voices, fo, voice_index, control_rate, steps_override)
        try:
            return [a.copy() for a in b.split(b.run_host(frames))]
        finally:
            b.close()

    # ---- control-frame generation (gtts_events_*: EventList::generateOutput) ----
    def prepare_events(self, configs, events, event_offsets, continues_previous=None):
        return EventsBatch(self, configs, events, event_offsets, continues_previous)

    def control_frames(self, configs, event_lists, continues_previous=None):
        """One config record (capi.EVENT_CONFIG_DTYPE) and one event array (capi.EVENT_DTYPE) per chunk ->
        list of float32 [frames, 16] control tracks, one per chunk (EventList::generateOutput on the device)."""
        events, eo = pack_events(event_lists)
        b = self.prepare_events(configs, events, eo, continues_previous)
        try:
            frames, _ = b.run_host(events)
            return [frames[b.frame_offsets[c]:b.frame_offsets[c + 1]].copy() for c in range(b.n_chunks)]
        finally:
            b.close()

    def close(self):
        if self._h:
            self._lib.gtts_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_events(event_lists):
    """list of capi.EVENT_DTYPE arrays -> (packed events, event_offsets[n + 1])."""
    lists = [np.ascontiguousarray(e, capi.EVENT_DTYPE).reshape(-1) for e in event_lists]
    eo = np.zeros(len(lists) + 1, np.int64)
    if lists:
        eo[1:] = np.cumsum([len(e) for e in lists])
    events = np.concatenate(lists) if lists else np.zeros(0, capi.EVENT_DTYPE)
    return np.ascontiguousarray(events), eo


def events_frame_count(config, events):
    """Frames EventList::generateOutput makes of an event list (host arithmetic on the times)."""
    cfg = np.array(config, capi.EVENT_CONFIG_DTYPE).reshape(1)
    ev = np.ascontiguousarray(events, capi.EVENT_DTYPE).reshape(-1)
    n = C.c_int64()
    check(load().gtts_events_frame_count(cfg.ctypes.data, ev.ctypes.data, len(ev), C.byref(n)))
    return int(n.value)


class EventsBatch:
    """gtts_events_batch: a prepared batch of chunks (event lists) whose control frames are generated on the device."""

    def __init__(self, synth, configs, events, event_offsets, continues_previous=None):
        self._lib = synth._lib
        self._synth = synth
        cfgs = np.ascontiguousarray(configs, capi.EVENT_CONFIG_DTYPE).reshape(-1)
        ev = np.ascontiguousarray(events, capi.EVENT_DTYPE).reshape(-1)
        eo = np.ascontiguousarray(event_offsets, np.int64)
        self.n_chunks = len(eo) - 1
        assert len(cfgs) == self.n_chunks
        cp = None if continues_previous is None else np.ascontiguousarray(continues_previous, np.int32)
        self._h = C.c_void_p()
        check(self._lib.gtts_events_prepare(synth._h, cfgs.ctypes.data, None if cp is None else cp.ctypes.data, ev.ctypes.data,
                                            eo.ctypes.data, self.n_chunks, C.byref(self._h)))
        self.event_offsets = eo
        self.frame_offsets = np.zeros(self.n_chunks + 1, np.int64)
        check(self._lib.gtts_events_layout(self._h, self.frame_offsets.ctypes.data))
        self.n_frames_total = int(self.frame_offsets[-1])
        self.n_events_total = int(eo[-1]) if len(eo) else 0
        first = np.ones(self.n_chunks, bool) if cp is None else (cp == 0)
        if self.n_chunks:
            first[0] = True
        # frame_offsets of the utterances (chains of chunks): the frame_offsets argument of gtts_batch_prepare
        self.utterance_frame_offsets = np.concatenate([self.frame_offsets[:-1][first], self.frame_offsets[-1:]])

    def run_host(self, events):
        """-> (frames [n_frames_total, 16] float32, configs with the drift state each chunk left)."""
        ev = np.ascontiguousarray(events, capi.EVENT_DTYPE).reshape(-1)
        assert len(ev) == self.n_events_total
        frames = np.zeros((max(self.n_frames_total, 1), NUM_PARAMS), np.float32)
        cfg_out = np.zeros(max(self.n_chunks, 1), capi.EVENT_CONFIG_DTYPE)
        check(self._lib.gtts_events_run_host(self._h, ev.ctypes.data, frames.ctypes.data, cfg_out.ctypes.data))
        return frames[:self.n_frames_total], cfg_out[:self.n_chunks]

    def run_device(self, d_events_ptr, d_frames_ptr, d_configs_out_ptr=0, stream_ptr=0):
        check(self._lib.gtts_events_run_device(self._h, C.c_void_p(d_events_ptr), C.c_void_p(d_frames_ptr),
                                               C.c_void_p(d_configs_out_ptr), C.c_void_p(stream_ptr)))

    def close(self):
        if self._h:
            self._lib.gtts_events_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch5:
    """gtts5_batch: a prepared batch of model-5 utterances."""

    def __init__(self, synth, voice_list, frame_offsets, voice_index, control_rate, steps_override):
        from .capi import voice5_array
        self._lib = synth._lib
        self._synth = synth
        fo = np.ascontiguousarray(frame_offsets, np.int64)
        self.n_utt = len(fo) - 1
        va = voice5_array(voice_list)
        vi = None if voice_index is None else np.ascontiguousarray(voice_index, np.int32)
        so = None if steps_override is None else np.ascontiguousarray(steps_override, np.int32)
        self._h = C.c_void_p()
        check(self._lib.gtts5_batch_prepare(synth._h, va, len(voice_list), None if vi is None else vi.ctypes.data,
                                            float(control_rate), None if so is None else so.ctypes.data, fo.ctypes.data,
                                            self.n_utt, C.byref(self._h)))
        self.out_offsets = np.zeros(self.n_utt + 1, np.int64)
        self.n_out = np.zeros(max(self.n_utt, 1), np.int64)
        self.n_internal = np.zeros(max(self.n_utt, 1), np.int64)
        check(self._lib.gtts5_batch_layout(self._h, self.out_offsets.ctypes.data, self.n_out.ctypes.data, self.n_internal.ctypes.data))
        self.n_out, self.n_internal = self.n_out[:self.n_utt], self.n_internal[:self.n_utt]
        self.n_out_total = int(self.out_offsets[-1])
        self.n_frames_total = int(fo[-1]) if len(fo) else 0

    def run_host(self, frames):
        frames = np.ascontiguousarray(frames, np.float32).reshape(-1, NUM_PARAMS)
        assert frames.shape[0] == self.n_frames_total
        out = np.zeros(max(self.n_out_total, 1), np.float32)
        check(self._lib.gtts5_batch_run_host(self._h, frames.ctypes.data, out.ctypes.data))
        return out

    def run_device(self, d_frames_ptr, d_out_ptr, stream_ptr=0):
        check(self._lib.gtts5_batch_run_device(self._h, d_frames_ptr, d_out_ptr, stream_ptr))

    def run_host_pcm16(self, frames):
        """-> (int16 payload in the batch layout, float32 scale per utterance): the reference's WAVE data."""
        frames = np.ascontiguousarray(frames, np.float32).reshape(-1, NUM_PARAMS)
        assert frames.shape[0] == self.n_frames_total
        pcm = np.zeros(max(self.n_out_total, 1), np.int16)
        scale = np.zeros(max(self.n_utt, 1), np.float32)
        check(self._lib.gtts5_batch_run_host_pcm16(self._h, frames.ctypes.data, pcm.ctypes.data, scale.ctypes.data))
        return pcm, scale[:self.n_utt]

    def split(self, out):
        return [out[self.out_offsets[u]:self.out_offsets[u] + self.n_out[u]] for u in range(self.n_utt)]

    def close(self):
        if self._h:
            self._lib.gtts5_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
