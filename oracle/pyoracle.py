"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the two CPU checkers.

* ``Oracle``    -> oracle/liboracle.so       (oracle/tube_oracle.c, our plain-C restatement, kind "port")
* ``Reference`` -> oracle/_ref/libgtts_ref.so (the unmodified reference compiled in place, kind "reference")

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Voices are plain dicts keyed like the reference's vtm.txt / variant files (VocalTractModel0.h:266-305).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

VOICE_KEYS_SCALAR = [
    ("output_rate", float), ("waveform", int), ("glottal_pulse_tp", float), ("glottal_pulse_tn_min", float),
    ("glottal_pulse_tn_max", float), ("breathiness", float), ("vocal_tract_length_offset", float),
    ("vocal_tract_length", float), ("temperature", float), ("loss_factor", float), ("mouth_coefficient", float),
    ("nose_coefficient", float), ("throat_cutoff", float), ("throat_volume", float), ("noise_modulation", int),
    ("mix_offset", float), ("global_radius_coef", float), ("global_nasal_radius_coef", float),
    ("aperture_radius", float),
]


class OracleVoice(C.Structure):
    _fields_ = [(k, C.c_double if t is float else C.c_int) for k, t in VOICE_KEYS_SCALAR] + [
        ("nasal_radius", C.c_double * 5), ("radius_coef", C.c_double * 8), ("tube_model", C.c_int)]


def voice_struct(voice):
    s = OracleVoice()
    for k, t in VOICE_KEYS_SCALAR:
        setattr(s, k, t(voice[k]))
    for i in range(5):
        s.nasal_radius[i] = float(voice["nasal_radius_%d" % (i + 1)])
    for i in range(8):
        s.radius_coef[i] = float(voice["radius_%d_coef" % (i + 1)])
    s.tube_model = int(voice.get("tube_model", 0))      # 0: models 0 / 2, 3: model 3, 4: model 4 (same keys)
    return s


def config_text(voice, model=0, extra=None):
    """key = value text in the reference's ConfigurationData format (ConfigurationData.cpp:67-118)."""
    lines = ["model = %d" % model, "log_parameters = false"]   # log_parameters: read by models 2/3/5 only
    for k, v in (extra or {}).items():
        lines.append("%s = %s" % (k, v))
    for k, t in VOICE_KEYS_SCALAR:
        lines.append("%s = %s" % (k, repr(float(voice[k])) if t is float else int(voice[k])))
    for i in range(5):
        lines.append("nasal_radius_%d = %r" % (i + 1, float(voice["nasal_radius_%d" % (i + 1)])))
    for i in range(8):
        lines.append("radius_%d_coef = %r" % (i + 1, float(voice["radius_%d_coef" % (i + 1)])))
    return ("\n".join(lines) + "\n").encode()


def build(ref=True):
    """Compiles liboracle.so (always) and oracle/_ref (only when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []), check=True)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.oracle_synthesize.restype = C.c_long
        L.oracle_synthesize.argtypes = [C.POINTER(OracleVoice), C.c_double, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleVoice)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_finish.argtypes = [C.c_void_p]
        L.oracle_run_track.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_long]
        L.oracle_output_size.restype = C.c_long
        L.oracle_output_size.argtypes = [C.c_void_p]
        L.oracle_output.restype = C.POINTER(C.c_float)
        L.oracle_output.argtypes = [C.c_void_p]
        L.oracle_internal_size.restype = C.c_long
        L.oracle_internal_size.argtypes = [C.c_void_p]
        L.oracle_internal.restype = C.POINTER(C.c_double)
        L.oracle_internal.argtypes = [C.c_void_p]
        L.oracle_internal_rate.restype = C.c_double
        L.oracle_internal_rate.argtypes = [C.c_void_p]
        L.oracle_noise.argtypes = [C.c_void_p, C.c_long]
        L.oracle_fir_taps.argtypes = [C.c_void_p, C.c_int]
        L.oracle_src_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_src_params.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_src_run.restype = C.c_long
        L.oracle_src_run.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.oracle_wavetable.argtypes = [C.POINTER(OracleVoice), C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_constants.argtypes = [C.POINTER(OracleVoice), C.c_void_p]
        L.oracle_pcm16.restype = C.c_float
        L.oracle_pcm16.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_output_scale.restype = C.c_float
        L.oracle_output_scale.argtypes = [C.c_void_p, C.c_long]

    def synthesize(self, voice, frames, control_rate=250.0, return_internal=False):
        frames = _f32(frames).reshape(-1, 16)
        vs = voice_struct(voice)
        m = self.lib.oracle_create(C.byref(vs))
        try:
            self.lib.oracle_run_track(m, control_rate, frames.ctypes.data, frames.shape[0])
            n = self.lib.oracle_output_size(m)
            out = np.ctypeslib.as_array(self.lib.oracle_output(m), shape=(max(n, 1),))[:n].copy()
            if return_internal:
                ni = self.lib.oracle_internal_size(m)
                x = np.ctypeslib.as_array(self.lib.oracle_internal(m), shape=(max(ni, 1),))[:ni].copy()
                return out, x
            return out
        finally:
            self.lib.oracle_destroy(m)

    def pcm16(self, audio):
        """The reference's output stage on a raw output buffer: (int16 payload, normalisation scale)."""
        x = _f32(audio)
        out = np.empty(len(x), np.int16)
        scale = self.lib.oracle_pcm16(x.ctypes.data, len(x), out.ctypes.data)
        return out, float(scale)

    def synthesize_samples(self, voice, params):
        """Per-sample entry (setAllParameters + execSynthesisStep for every row of params)."""
        params = _f32(params).reshape(-1, 16)
        vs = voice_struct(voice)
        m = self.lib.oracle_create(C.byref(vs))
        try:
            for i in range(params.shape[0]):
                self.lib.oracle_step(m, params[i].ctypes.data)
            self.lib.oracle_finish(m)
            n = self.lib.oracle_output_size(m)
            return np.ctypeslib.as_array(self.lib.oracle_output(m), shape=(max(n, 1),))[:n].copy()
        finally:
            self.lib.oracle_destroy(m)

    def internal_rate(self, voice):
        vs = voice_struct(voice)
        m = self.lib.oracle_create(C.byref(vs))
        r = self.lib.oracle_internal_rate(m)
        self.lib.oracle_destroy(m)
        return r

    def noise(self, n):
        out = np.empty(n, np.float64)
        self.lib.oracle_noise(out.ctypes.data, n)
        return out

    def fir_taps(self):
        out = np.zeros(512, np.float64)
        n = self.lib.oracle_fir_taps(out.ctypes.data, 512)
        return out[:n].copy()

    def src_tables(self):
        h = np.empty(3328, np.float64)
        dh = np.empty(3328, np.float64)
        self.lib.oracle_src_tables(h.ctypes.data, dh.ctypes.data)
        return h, dh

    def src_params(self, in_rate, out_rate):
        inc, pinc, pad = C.c_uint(), C.c_uint(), C.c_int()
        self.lib.oracle_src_params(in_rate, out_rate, C.byref(inc), C.byref(pinc), C.byref(pad))
        return inc.value, pinc.value, pad.value

    def src_run(self, in_rate, out_rate, x):
        x = np.ascontiguousarray(x, np.float64)
        cap = int(len(x) * out_rate / in_rate * 1.5) + 4096
        out = np.empty(cap, np.float32)
        n = self.lib.oracle_src_run(in_rate, out_rate, x.ctypes.data, len(x), out.ctypes.data, cap)
        assert n <= cap
        return out[:n].copy()

    def wavetable(self, voice, amplitude=-1.0):
        vs = voice_struct(voice)
        t = np.empty(512, np.float64)
        s = np.empty(5, np.float64)
        self.lib.oracle_wavetable(C.byref(vs), amplitude, t.ctypes.data, s.ctypes.data)
        return t, s

    def constants(self, voice):
        vs = voice_struct(voice)
        out = np.zeros(32, np.float64)
        n = self.lib.oracle_constants(C.byref(vs), out.ctypes.data)
        return out[:n].copy()


class Reference:
    """The unmodified reference (any of its models: 0, 1, 2, ...) through oracle/ref_harness.cpp."""

    def __init__(self, variant=""):
        path = os.path.join(HERE, "_ref", "libgtts_ref%s.so" % variant)
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/gama_tts/src"):
                build(ref=True)
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_synthesize.argtypes = [C.c_char_p, C.c_double, C.c_void_p, C.c_long,
                                     C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long), C.POINTER(C.c_double)]
        L.ref_synthesize_samples.argtypes = [C.c_char_p, C.c_void_p, C.c_long,
                                             C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long)]
        L.ref_interactive.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_long, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_noise.argtypes = [C.c_void_p, C.c_long]
        L.ref_fir_taps.argtypes = [C.c_void_p, C.c_int]
        L.ref_src_tables.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_src_run.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_long,
                                  C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long)]
        L.ref_wavetable.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p]
        L.ref_model0_constants.argtypes = [C.c_char_p, C.c_void_p]
        L.ref_pcm16.argtypes = [C.c_void_p, C.c_long, C.c_float, C.c_void_p, C.POINTER(C.c_float)]

    def _take(self, p, n):
        out = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
        self.lib.ref_free(p)
        return out

    def synthesize(self, voice, frames, control_rate=250.0, model=0, extra=None):
        """extra: additional config keys, e.g. {"dll_path": ...} with model=2000 (the plugin seam)."""
        frames = _f32(frames).reshape(-1, 16)
        p, n, rate = C.POINTER(C.c_float)(), C.c_long(), C.c_double()
        rc = self.lib.ref_synthesize(config_text(voice, model, extra), control_rate, frames.ctypes.data, frames.shape[0],
                                     C.byref(p), C.byref(n), C.byref(rate))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return self._take(p, n)

    def synthesize5(self, voice5, frames, control_rate=250.0, extra=None, model=5):
        """Model 5 (or the plugin, model=2000 + extra dll_path) on a model-5 voice dict (gama_tts_b200.voices.default_voice5)."""
        frames = _f32(frames).reshape(-1, 16)
        p, n, rate = C.POINTER(C.c_float)(), C.c_long(), C.c_double()
        rc = self.lib.ref_synthesize(config_text5(voice5, extra, model), control_rate, frames.ctypes.data, frames.shape[0],
                                     C.byref(p), C.byref(n), C.byref(rate))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return self._take(p, n)

    def interactive(self, voice, params, callback_frames=1024, model=0, extra=None):
        """The editor's interactive call pattern (InteractiveAudio.cpp:131-186) on per-step parameter rows."""
        params = _f32(params).reshape(-1, 16)
        p, n = C.POINTER(C.c_float)(), C.c_long()
        rc = self.lib.ref_interactive(config_text(voice, model, extra), params.ctypes.data, params.shape[0], callback_frames,
                                      C.byref(p), C.byref(n))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return self._take(p, n)

    def pcm16(self, audio, rate=48000.0):
        """The reference's own writer (Controller::writeOutputToFile + WAVEFileWriter) on a raw output buffer."""
        x = _f32(audio)
        out = np.empty(len(x), np.int16)
        scale = C.c_float()
        if self.lib.ref_pcm16(x.ctypes.data, len(x), rate, out.ctypes.data, C.byref(scale)):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out, float(scale.value)

    def synthesize_samples(self, voice, params, model=0):
        params = _f32(params).reshape(-1, 16)
        p, n = C.POINTER(C.c_float)(), C.c_long()
        rc = self.lib.ref_synthesize_samples(config_text(voice, model), params.ctypes.data, params.shape[0],
                                             C.byref(p), C.byref(n))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return self._take(p, n)

    def batch(self, voices, frames, frame_offsets, n_threads=1, control_rate=250.0, out_offsets=None, model=0):
        """voices: one dict (shared) or a list of U dicts. Returns (seconds, n_out_each, out or None)."""
        frames = _f32(frames).reshape(-1, 16)
        fo = np.ascontiguousarray(frame_offsets, np.int64)
        U = len(fo) - 1
        vl = [voices] if isinstance(voices, dict) else list(voices)
        # a voice dict with model 5's keys (gama_tts_b200.voices.default_voice5) is written with them, model = 5
        texts = (C.c_char_p * len(vl))(*[config_text5(v) if "glottal_noise_cutoff" in v else config_text(v, model) for v in vl])
        n_each = np.zeros(U, np.int64)
        out = None
        oo_ptr = None
        if out_offsets is not None:
            oo = np.ascontiguousarray(out_offsets, np.int64)
            out = np.zeros(int(oo[-1]), np.float32)
            oo_ptr = oo.ctypes.data
        sec = C.c_double()
        rc = self.lib.ref_batch(texts, len(vl), control_rate, frames.ctypes.data, fo.ctypes.data, U, n_threads,
                                out.ctypes.data if out is not None else None, oo_ptr, n_each.ctypes.data, C.byref(sec))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return sec.value, n_each, out

    def noise(self, n):
        out = np.empty(n, np.float64)
        self.lib.ref_noise(out.ctypes.data, n)
        return out

    def fir_taps(self):
        out = np.zeros(512, np.float64)
        n = self.lib.ref_fir_taps(out.ctypes.data, 512)
        return out[:n].copy()

    def src_tables(self, in_rate=20034.0, out_rate=48000.0):
        h = np.empty(3328, np.float64)
        dh = np.empty(3328, np.float64)
        incs = (C.c_uint * 3)()
        pad = C.c_int()
        self.lib.ref_src_tables(in_rate, out_rate, h.ctypes.data, dh.ctypes.data, incs, C.byref(pad))
        return h, dh, tuple(incs), pad.value

    def src_run(self, in_rate, out_rate, x):
        x = np.ascontiguousarray(x, np.float64)
        p, n = C.POINTER(C.c_float)(), C.c_long()
        rc = self.lib.ref_src_run(in_rate, out_rate, x.ctypes.data, len(x), C.byref(p), C.byref(n))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return self._take(p, n)

    def wavetable(self, fs, tp, tn_min, tn_max, amplitude=-1.0, sine=0):
        t = np.empty(512, np.float64)
        s = np.empty(5, np.float64)
        self.lib.ref_wavetable(sine, float(fs), tp, tn_min, tn_max, amplitude, t.ctypes.data, s.ctypes.data)
        return t, s

    def constants(self, voice):
        out = np.zeros(32, np.float64)
        n = self.lib.ref_model0_constants(config_text(voice, 0), out.ctypes.data)
        if n < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out[:n].copy()


# ---- model 5 (VocalTractModel5<double, 1>) ------------------------------------------------------------------------
VOICE5_KEYS_SCALAR = [
    ("output_rate", float), ("waveform", int), ("glottal_pulse_tp", float), ("glottal_pulse_tn_min", float),
    ("glottal_pulse_tn_max", float), ("breathiness", float), ("vocal_tract_length_offset", float),
    ("vocal_tract_length", float), ("temperature", float), ("loss_factor", float), ("noise_modulation", int),
    ("mix_offset", float), ("global_radius_coef", float), ("global_nasal_radius_coef", float),
]
VOICE5_KEYS_TAIL = [
    ("glottal_noise_cutoff", float), ("frication_noise_cutoff", float), ("frication_factor", float),
    ("min_glottal_loss", float), ("max_glottal_loss", float), ("glottal_lowpass_cutoff", float), ("bypass", int),
    ("constant_radius_mouth_impedance", int), ("mouth_impedance_radius", float),
]


class Oracle5Voice(C.Structure):
    _fields_ = ([(k, C.c_double if t is float else C.c_int) for k, t in VOICE5_KEYS_SCALAR]
                + [("nasal_radius", C.c_double * 6), ("radius_coef", C.c_double * 8)]
                + [(k, C.c_double if t is float else C.c_int) for k, t in VOICE5_KEYS_TAIL])


def voice5_struct(voice):
    s = Oracle5Voice()
    for k, t in VOICE5_KEYS_SCALAR + VOICE5_KEYS_TAIL:
        setattr(s, k, t(voice[k]))
    for i in range(6):
        s.nasal_radius[i] = float(voice["nasal_radius_%d" % (i + 2)])
    for i in range(8):
        s.radius_coef[i] = float(voice["radius_%d_coef" % (i + 1)])
    return s


def config_text5(voice, extra=None, model=5):
    """A model-5 voice as the reference's key = value text (data/voice/english/5_xxx/vtm.txt + variant/xxx.txt)."""
    lines = ["model = %d" % model, "log_parameters = false"]
    for k, v in (extra or {}).items():
        lines.append("%s = %s" % (k, v))
    for k, t in VOICE5_KEYS_SCALAR + VOICE5_KEYS_TAIL:
        if k == "constant_radius_mouth_impedance":
            lines.append("%s = %s" % (k, "true" if voice[k] else "false"))
        else:
            lines.append("%s = %s" % (k, repr(float(voice[k])) if t is float else int(voice[k])))
    for i in range(6):
        lines.append("nasal_radius_%d = %r" % (i + 2, float(voice["nasal_radius_%d" % (i + 2)])))
    for i in range(8):
        lines.append("radius_%d_coef = %r" % (i + 1, float(voice["radius_%d_coef" % (i + 1)])))
    return ("\n".join(lines) + "\n").encode()


class Oracle5:
    """oracle/tube5_oracle.c through ctypes."""

    def __init__(self):
        build(ref=False)
        L = self.lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.oracle5_synthesize.restype = C.c_long
        L.oracle5_synthesize.argtypes = [C.POINTER(Oracle5Voice), C.c_double, C.c_int, C.c_void_p, C.c_long, C.c_void_p,
                                         C.c_long, C.POINTER(C.c_double)]

    def synthesize(self, voice, frames, control_rate=250.0, steps=0):
        frames = _f32(frames).reshape(-1, 16)
        vs = voice5_struct(voice)
        rate = C.c_double()
        probe = np.zeros(1, np.float32)
        n = self.lib.oracle5_synthesize(C.byref(vs), control_rate, steps, frames.ctypes.data, frames.shape[0], probe.ctypes.data, 0,
                                        C.byref(rate))
        out = np.zeros(max(n, 1), np.float32)
        n2 = self.lib.oracle5_synthesize(C.byref(vs), control_rate, steps, frames.ctypes.data, frames.shape[0], out.ctypes.data, n,
                                         C.byref(rate))
        assert n2 == n
        return out[:n]

    def synthesize_samples(self, voice, params):
        return self.synthesize(voice, params, steps=1)[...]

    def internal_rate(self, voice):
        vs = voice5_struct(voice)
        rate = C.c_double()
        probe = np.zeros(1, np.float32)
        self.lib.oracle5_synthesize(C.byref(vs), 250.0, 0, None, 0, probe.ctypes.data, 0, C.byref(rate))
        return float(rate.value)


# ---- control-frame generation (oracle/events_oracle.c) -------------------------------------------------------------
EVENT_DTYPE = np.dtype([("time", "<i4"), ("has_interp", "<i4"), ("param", "<f8", 16), ("special", "<f8", 16),
                        ("a", "<f8"), ("b", "<f8"), ("c", "<f8"), ("d", "<f8")])
EVENT_CONFIG_DTYPE = np.dtype([("control_period", "<i4"), ("macro_intonation", "<i4"), ("micro_intonation", "<i4"),
                               ("intonation_drift", "<i4"), ("smooth_intonation", "<i4"), ("pad_", "<i4"),
                               ("initial_pitch", "<f8"), ("mean_pitch", "<f8"), ("drift_deviation2", "<f8"),
                               ("drift_offset", "<f8"), ("drift_seed", "<f8"), ("drift_b0", "<f8"), ("drift_b1", "<f8"),
                               ("drift_a1", "<f8"), ("drift_a2", "<f8"), ("drift_x1", "<f8"), ("drift_x2", "<f8"),
                               ("drift_y1", "<f8"), ("drift_y2", "<f8")])


class OracleEvents:
    """oracle/events_oracle.c through ctypes: event list -> control frames (EventList::generateOutput)."""

    def __init__(self):
        build(ref=False)
        L = self.lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.oracle_events_generate.restype = C.c_long
        L.oracle_events_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long]

    def generate(self, cfg, events):
        """cfg: EVENT_CONFIG_DTYPE scalar (not modified), events: EVENT_DTYPE array -> (frames [F, 16] float32,
        cfg after the call, i.e. with the drift generator's state advanced)."""
        cfg = np.array(cfg, dtype=EVENT_CONFIG_DTYPE).reshape(1).copy()
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        before = cfg.copy()
        n = self.lib.oracle_events_generate(cfg.ctypes.data, ev.ctypes.data, len(ev), None, 0)
        frames = np.zeros((max(n, 1), 16), np.float32)
        cfg = before
        n2 = self.lib.oracle_events_generate(cfg.ctypes.data, ev.ctypes.data, len(ev), frames.ctypes.data, n)
        assert n2 == n
        return frames[:n], cfg[0]
