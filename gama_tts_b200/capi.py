"""ctypes view of include/gtts_b200.h: structures, prototypes and the library loader.

Python is test / benchmark orchestration only; the product is the C-ABI library
``gama_tts_b200/csrc/libgtts_b200.so`` (CUDA kernels + C++ host runtime).  Loading fails loudly when
the library is missing -- there is no Python or CPU implementation to fall back to.
"""
import ctypes as C

import numpy as np
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTTS_LIB_PATH") or os.path.join(HERE, "csrc", "libgtts_b200.so")   # GTTS_LIB_PATH: another build (tools/ab_build.sh)

GTTS_OK, GTTS_ERR_INVALID, GTTS_ERR_CUDA, GTTS_ERR_NO_DEVICE, GTTS_ERR_NOMEM, GTTS_ERR_UNSUPPORTED = range(6)


class VoiceConfig(C.Structure):
    """struct gtts_voice_config"""
    _fields_ = [
        ("output_rate", C.c_double), ("waveform", C.c_int32), ("noise_modulation", C.c_int32),
        ("glottal_pulse_tp", C.c_double), ("glottal_pulse_tn_min", C.c_double), ("glottal_pulse_tn_max", C.c_double),
        ("breathiness", C.c_double), ("vocal_tract_length_offset", C.c_double), ("vocal_tract_length", C.c_double),
        ("temperature", C.c_double), ("loss_factor", C.c_double), ("mouth_coefficient", C.c_double),
        ("nose_coefficient", C.c_double), ("throat_cutoff", C.c_double), ("throat_volume", C.c_double),
        ("mix_offset", C.c_double), ("global_radius_coef", C.c_double), ("global_nasal_radius_coef", C.c_double),
        ("aperture_radius", C.c_double), ("nasal_radius", C.c_double * 5), ("radius_coef", C.c_double * 8),
        ("tube_model", C.c_int32), ("reserved_", C.c_int32),
    ]


def voice_config(voice):
    """dict keyed like the reference's vtm.txt / variant files -> struct gtts_voice_config."""
    s = VoiceConfig()
    for name, ctype in VoiceConfig._fields_:
        if name == "nasal_radius":
            for i in range(5):
                s.nasal_radius[i] = float(voice["nasal_radius_%d" % (i + 1)])
        elif name == "radius_coef":
            for i in range(8):
                s.radius_coef[i] = float(voice["radius_%d_coef" % (i + 1)])
        elif name in ("tube_model", "reserved_"):
            setattr(s, name, int(voice.get(name, 0)))        # 0: models 0 / 2, 3: model 3, 4: model 4
        elif ctype is C.c_int32:
            setattr(s, name, int(voice[name]))
        else:
            setattr(s, name, float(voice[name]))
    return s


class Voice5Config(C.Structure):
    """struct gtts_voice5_config (model 5)"""
    _fields_ = [
        ("output_rate", C.c_double), ("waveform", C.c_int32), ("noise_modulation", C.c_int32), ("bypass", C.c_int32),
        ("constant_radius_mouth_impedance", C.c_int32),
        ("glottal_pulse_tp", C.c_double), ("glottal_pulse_tn_min", C.c_double), ("glottal_pulse_tn_max", C.c_double),
        ("breathiness", C.c_double), ("vocal_tract_length_offset", C.c_double), ("vocal_tract_length", C.c_double),
        ("temperature", C.c_double), ("loss_factor", C.c_double), ("mix_offset", C.c_double),
        ("global_radius_coef", C.c_double), ("global_nasal_radius_coef", C.c_double),
        ("nasal_radius", C.c_double * 6), ("radius_coef", C.c_double * 8),
        ("glottal_noise_cutoff", C.c_double), ("frication_noise_cutoff", C.c_double), ("frication_factor", C.c_double),
        ("min_glottal_loss", C.c_double), ("max_glottal_loss", C.c_double), ("glottal_lowpass_cutoff", C.c_double),
        ("mouth_impedance_radius", C.c_double),
    ]


def voice5_config(voice):
    """dict keyed like data/voice/english/5_xxx/vtm.txt + variant -> struct gtts_voice5_config."""
    s = Voice5Config()
    for name, ctype in Voice5Config._fields_:
        if name == "nasal_radius":
            for i in range(6):
                s.nasal_radius[i] = float(voice["nasal_radius_%d" % (i + 2)])
        elif name == "radius_coef":
            for i in range(8):
                s.radius_coef[i] = float(voice["radius_%d_coef" % (i + 1)])
        elif ctype is C.c_int32:
            setattr(s, name, int(voice[name]))
        else:
            setattr(s, name, float(voice[name]))
    return s


def voice5_array(voices):
    arr = (Voice5Config * len(voices))()
    for i, v in enumerate(voices):
        arr[i] = voice5_config(v)
    return arr


def voice_array(voices):
    arr = (VoiceConfig * len(voices))()
    for i, v in enumerate(voices):
        arr[i] = voice_config(v)
    return arr


# struct gtts_event / gtts_event_config (include/gtts_b200.h) as numpy record types
EVENT_DTYPE = np.dtype([("time", "<i4"), ("has_interp", "<i4"), ("param", "<f8", 16), ("special", "<f8", 16),
                        ("a", "<f8"), ("b", "<f8"), ("c", "<f8"), ("d", "<f8")])
EVENT_CONFIG_DTYPE = np.dtype([("control_period", "<i4"), ("macro_intonation", "<i4"), ("micro_intonation", "<i4"),
                               ("intonation_drift", "<i4"), ("smooth_intonation", "<i4"), ("reserved", "<i4"),
                               ("initial_pitch", "<f8"), ("mean_pitch", "<f8"), ("drift_deviation2", "<f8"),
                               ("drift_offset", "<f8"), ("drift_seed", "<f8"), ("drift_b0", "<f8"), ("drift_b1", "<f8"),
                               ("drift_a1", "<f8"), ("drift_a2", "<f8"), ("drift_x1", "<f8"), ("drift_x2", "<f8"),
                               ("drift_y1", "<f8"), ("drift_y2", "<f8")])
assert EVENT_DTYPE.itemsize == 296 and EVENT_CONFIG_DTYPE.itemsize == 128

EXPORTS = [
    "gtts_last_error", "gtts_abi_version", "gtts_voice_internal_rate", "gtts_voice_control_steps",
    "gtts_output_length", "gtts_shard_plan", "gtts_probe_fir_taps", "gtts_probe_src_tables",
    "gtts_probe_voice_constants", "gtts_probe_fp64_peak", "gtts_probe_exp", "gtts_create", "gtts_destroy", "gtts_describe", "gtts_batch_prepare",
    "gtts_batch_layout", "gtts_batch_lengths", "gtts_batch_run_device", "gtts_batch_run_host", "gtts_batch_run_device_pcm16",
    "gtts_batch_run_host_pcm16", "gtts_batch_submit_host_pcm16", "gtts_batch_wait", "gtts_batch_checksum_device", "gtts_batch_last_launches", "gtts_batch_last_kernel",
    "gtts_multi_create", "gtts_multi_destroy", "gtts_multi_device_count", "gtts_multi_batch_prepare", "gtts_multi_batch_layout",
    "gtts_multi_batch_run_host", "gtts_multi_batch_run_host_pcm16", "gtts_multi_batch_free",
    "gtts_batch_free", "gtts_batch_synthesize", "gtts_stream_open", "gtts_stream_push_frames",
    "gtts_stream_finish", "gtts_stream_reset", "gtts_stream_close",
    "gtts5_voice_internal_rate", "gtts5_output_length", "gtts5_batch_prepare", "gtts5_batch_layout", "gtts5_batch_run_device",
    "gtts5_batch_run_host", "gtts5_batch_run_device_pcm16", "gtts5_batch_run_host_pcm16", "gtts5_batch_free",
    "gtts5_multi_batch_prepare", "gtts5_multi_batch_layout", "gtts5_multi_batch_run_host", "gtts5_multi_batch_run_host_pcm16",
    "gtts5_multi_batch_free",
    "gtts_events_drift_setup", "gtts_events_frame_count", "gtts_events_prepare", "gtts_events_layout", "gtts_events_run_device", "gtts_events_run_host",
    "gtts_events_free",
]

_lib = None


def load():
    """Loads libgtts_b200.so and declares the prototypes of include/gtts_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    PV = C.POINTER(VoiceConfig)
    L.gtts_last_error.restype = C.c_char_p
    L.gtts_describe.restype = C.c_char_p
    L.gtts_describe.argtypes = [vp]
    L.gtts_voice_internal_rate.argtypes = [PV, C.POINTER(i32)]
    L.gtts_voice_control_steps.argtypes = [PV, dbl, C.POINTER(i32)]
    L.gtts_output_length.argtypes = [PV, i32, i64, C.POINTER(i64), C.POINTER(i64)]
    L.gtts_shard_plan.argtypes = [vp, i64, i32, vp]
    L.gtts_probe_fir_taps.argtypes = [vp, i32, C.POINTER(i32)]
    L.gtts_probe_src_tables.argtypes = [vp, vp]
    L.gtts_probe_voice_constants.argtypes = [PV, vp, i32, C.POINTER(i32)]
    L.gtts_probe_fp64_peak.argtypes = [vp, C.POINTER(dbl)]
    L.gtts_probe_exp.argtypes = [vp, vp, i32, vp, vp]
    L.gtts_create.argtypes = [i32, C.POINTER(vp)]
    L.gtts_destroy.argtypes = [vp]
    L.gtts_destroy.restype = None
    L.gtts_batch_prepare.argtypes = [vp, PV, i32, vp, dbl, vp, vp, i64, C.POINTER(vp)]
    L.gtts_batch_layout.argtypes = [vp, vp, vp]
    L.gtts_batch_lengths.argtypes = [vp, vp]
    L.gtts_batch_run_device.argtypes = [vp, vp, vp, vp]
    L.gtts_batch_run_host.argtypes = [vp, vp, vp]
    L.gtts_batch_run_device_pcm16.argtypes = [vp, vp, vp, vp, vp, vp]
    L.gtts_batch_run_host_pcm16.argtypes = [vp, vp, vp, vp]
    L.gtts_batch_submit_host_pcm16.argtypes = [vp, vp, vp, vp]
    L.gtts_batch_wait.argtypes = [vp]
    L.gtts_batch_checksum_device.argtypes = [vp, vp, vp, vp]
    L.gtts_multi_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.gtts_multi_destroy.argtypes = [vp]
    L.gtts_multi_destroy.restype = None
    L.gtts_multi_device_count.argtypes = [vp]
    L.gtts_multi_batch_prepare.argtypes = [vp, PV, i32, vp, dbl, vp, vp, i64, C.POINTER(vp)]
    L.gtts_multi_batch_layout.argtypes = [vp, vp, vp, vp]
    L.gtts_multi_batch_run_host.argtypes = [vp, vp, vp]
    L.gtts_multi_batch_run_host_pcm16.argtypes = [vp, vp, vp, vp]
    L.gtts_multi_batch_free.argtypes = [vp]
    L.gtts_multi_batch_free.restype = None
    L.gtts_batch_last_launches.argtypes = [vp, C.POINTER(i32)]
    L.gtts_batch_last_kernel.argtypes = [vp]
    L.gtts_batch_last_kernel.restype = C.c_char_p
    L.gtts_batch_free.argtypes = [vp]
    L.gtts_batch_free.restype = None
    L.gtts_batch_synthesize.argtypes = [vp, PV, i32, vp, dbl, vp, vp, i64, vp, i64, vp]
    L.gtts_stream_open.argtypes = [vp, PV, dbl, i32, C.POINTER(vp)]
    L.gtts_stream_push_frames.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.gtts_stream_finish.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.gtts_stream_reset.argtypes = [vp]
    L.gtts_stream_close.argtypes = [vp]
    L.gtts_stream_close.restype = None
    PV5 = C.POINTER(Voice5Config)
    L.gtts5_voice_internal_rate.argtypes = [PV5, C.POINTER(dbl)]
    L.gtts5_output_length.argtypes = [PV5, dbl, i32, i64, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    L.gtts5_batch_prepare.argtypes = [vp, PV5, i32, vp, dbl, vp, vp, i64, C.POINTER(vp)]
    L.gtts5_batch_layout.argtypes = [vp, vp, vp, vp]
    L.gtts5_batch_run_device.argtypes = [vp, vp, vp, vp]
    L.gtts5_batch_run_host.argtypes = [vp, vp, vp]
    L.gtts5_batch_run_device_pcm16.argtypes = [vp, vp, vp, vp, vp, vp]
    L.gtts5_batch_run_host_pcm16.argtypes = [vp, vp, vp, vp]
    L.gtts5_batch_free.argtypes = [vp]
    L.gtts5_batch_free.restype = None
    L.gtts5_multi_batch_prepare.argtypes = [vp, PV5, i32, vp, dbl, vp, vp, i64, C.POINTER(vp)]
    L.gtts5_multi_batch_layout.argtypes = [vp, vp, vp, vp]
    L.gtts5_multi_batch_run_host.argtypes = [vp, vp, vp]
    L.gtts5_multi_batch_run_host_pcm16.argtypes = [vp, vp, vp, vp]
    L.gtts5_multi_batch_free.argtypes = [vp]
    L.gtts5_multi_batch_free.restype = None
    L.gtts_events_drift_setup.argtypes = [dbl, dbl, dbl, vp]
    L.gtts_events_frame_count.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.gtts_events_prepare.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(vp)]
    L.gtts_events_layout.argtypes = [vp, vp]
    L.gtts_events_run_device.argtypes = [vp, vp, vp, vp, vp]
    L.gtts_events_run_host.argtypes = [vp, vp, vp, vp]
    L.gtts_events_free.argtypes = [vp]
    L.gtts_events_free.restype = None
    _lib = L
    return L


class GttsError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("gtts error %d: %s" % (code, text))
        self.code = code


def check(rc):
    if rc != GTTS_OK:
        raise GttsError(rc, load().gtts_last_error().decode(errors="replace"))
