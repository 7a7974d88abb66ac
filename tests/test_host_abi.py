"""CPU-side checks of the product library: it loads, exports every symbol include/gtts_b200.h declares,
its host-only entry points agree with the oracle, and device entry points fail loudly without a GPU
(no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gama_tts_b200 as g
from gama_tts_b200 import capi
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import default_voice, random_voice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(product_lib):
    hdr = open(os.path.join(ROOT, "include", "gtts_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gtts5?_[a-z0-9_]+)\s*\(", hdr)))
    assert declared and set(declared) == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(product_lib, name), name
    assert product_lib.gtts_abi_version() == 2


def test_voice_struct_layout_matches_header():
    # 2 int32 + 17 doubles + 5 + 8 doubles + 2 int32 (tube_model, reserved)
    assert C.sizeof(capi.VoiceConfig) == 8 * (1 + 1 + 16 + 5 + 8 + 1)


def test_internal_rate_and_steps(product_lib):
    for var, fs, steps in (("male", 20034, 80), ("female", 23373, 93), ("large_child", 28047, 112),
                           ("small_child", 35059, 140), ("baby", 46746, 187)):
        v = default_voice(var)
        assert g.internal_rate(v) == fs
        assert g.control_steps(v, 250.0) == steps


def test_output_length_matches_oracle(product_lib, oracle):
    rng = np.random.Generator(np.random.PCG64(3))
    v = default_voice("male")
    assert g.output_length(v, 2500) == (200000, 479250)
    assert g.output_length(v, 250) == (20000, 47981)
    assert g.output_length(v, 0) == (0, 63)
    for i in range(6):
        voice = random_voice(rng)
        n_frames = int(rng.integers(0, 40))
        ni, no = g.output_length(voice, n_frames)
        out = oracle.synthesize(voice, T.synthetic_track(i, n_frames))
        assert no == len(out) and ni == n_frames * g.control_steps(voice)


def test_host_tables_match_oracle(product_lib, oracle, golden):
    taps = np.zeros(64)
    n = C.c_int32()
    capi.check(product_lib.gtts_probe_fir_taps(taps.ctypes.data, 64, C.byref(n)))
    assert n.value == 49 and np.array_equal(taps[:49], oracle.fir_taps())
    assert np.array_equal(taps[:49], golden.kat("fir_taps"))
    h, dh = np.zeros(3328), np.zeros(3328)
    capi.check(product_lib.gtts_probe_src_tables(h.ctypes.data, dh.ctypes.data))
    # the FIR taps are shipped constants (bit-identical to the reference's design above); the SRC table is computed
    # from its closed form with the library's Bessel function: equal to the reference's table to a few ulp
    oh, odh = oracle.src_tables()
    assert np.abs(h - oh).max() <= 4e-15 * np.abs(oh).max() and (np.abs(h - oh) <= 4e-15 * np.abs(oh)).all()
    assert np.abs(dh - odh).max() <= 3e-15


def test_voice_constants_match_reference_kat(product_lib, golden):
    for var in ("male", "female", "large_child", "small_child", "baby"):
        v = capi.voice_config(default_voice(var))
        out = np.zeros(32)
        n = C.c_int32()
        capi.check(product_lib.gtts_probe_voice_constants(C.byref(v), out.ctypes.data, 32, C.byref(n)))
        ref = golden.kat("constants_" + var)       # 24 values from the reference's private members
        assert np.allclose(out[:24], ref, rtol=1e-15, atol=0)
        assert np.array_equal(out[:18], ref[:18])


def test_shard_plan_balanced(product_lib):
    rng = np.random.Generator(np.random.PCG64(7))
    cost = np.exp(rng.uniform(np.log(250), np.log(5000), 4096)).astype(np.int64)
    for shards in (1, 2, 4, 8):
        s = g.shard_plan(cost, shards)
        assert s.min() == 0 and s.max() == shards - 1
        loads = np.bincount(s, weights=cost, minlength=shards)
        assert loads.max() / loads.mean() < 1.01
    assert len(g.shard_plan(np.zeros(0, np.int64), 4)) == 0


def test_invalid_arguments_are_reported(product_lib):
    v = default_voice("male")
    with pytest.raises(capi.GttsError) as e:
        g.output_length(v, 10, steps=0)
    assert e.value.code == capi.GTTS_ERR_INVALID
    bad = dict(v, output_rate=0.0)
    with pytest.raises(capi.GttsError) as e:
        g.output_length(bad, 10)
    assert e.value.code == capi.GTTS_ERR_INVALID


def test_output_length_of_down_sampling_voices(product_lib, oracle):
    # fs_int > 48 kHz (tract shorter than 7.3 cm): the closed form covers the down-sampling branch (pad > 13)
    for length in (7.2, 6.0, 4.5, 3.0):
        short = dict(default_voice("baby"), vocal_tract_length=length)
        tr = np.tile(np.array([-12, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.84, 1.15, 1.31, 1.59, 1.59, 2.61, 0.1], np.float32), (3, 1))
        assert g.output_length(short, 3)[1] == len(oracle.synthesize(short, tr))


def test_no_cpu_fallback_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GttsError) as e:
        g.TubeSynthesizer(0)
    assert e.value.code in (capi.GTTS_ERR_NO_DEVICE, capi.GTTS_ERR_CUDA)


def test_product_does_not_reference_the_oracle():
    # The product tree must not import, link or execute anything under oracle/ (or the test emulator).
    bad = []
    for base in ("gama_tts_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"oracle/|pyoracle|liboracle|libgtts_ref|simt_emu\.h", text):
                        bad.append(os.path.join(dp, f))
    # tube_kernel.cuh only *mentions* the emulator in a comment; it must not include it
    bad = [b for b in bad if not b.endswith("tube_kernel.cuh")]
    assert not bad, bad


def test_model5_host_entry_points(product_lib, oracle5):
    # gtts5_* host-only helpers: internal rate (a double for this model), control steps and output length against
    # the model-5 oracle; what the reference refuses is refused, what this path does not implement says so
    from gama_tts_b200.voices import default_voice5, random_voice5
    assert C.sizeof(capi.Voice5Config) == 8 + 4 * 4 + 8 * (11 + 6 + 8 + 7)
    rng = np.random.Generator(np.random.PCG64(2))
    for v in [default_voice5(n) for n in ("male", "female", "large_child", "small_child", "baby")] + [random_voice5(rng)]:
        vc = capi.voice5_config(v)
        fs = C.c_double()
        capi.check(product_lib.gtts5_voice_internal_rate(C.byref(vc), C.byref(fs)))
        assert fs.value == oracle5.internal_rate(v)
        tr = T.synthetic_track(4, 7)
        steps, n_int, n_out = C.c_int32(), C.c_int64(), C.c_int64()
        capi.check(product_lib.gtts5_output_length(C.byref(vc), 250.0, 0, len(tr), C.byref(steps), C.byref(n_int), C.byref(n_out)))
        assert steps.value == int(np.rint(fs.value / 250.0)) and n_int.value == steps.value * len(tr)
        assert n_out.value == len(oracle5.synthesize(v, tr))
    assert abs(oracle5.internal_rate(default_voice5("male")) - 60411.428571428) < 1e-6
    n_out = C.c_int64()
    bad = capi.voice5_config(dict(default_voice5("male"), vocal_tract_length=22.0))
    assert product_lib.gtts5_output_length(C.byref(bad), 250.0, 0, 5, None, None, C.byref(n_out)) == capi.GTTS_ERR_INVALID
    short = capi.voice5_config(dict(default_voice5("male"), vocal_tract_length=5.0))
    assert product_lib.gtts5_output_length(C.byref(short), 250.0, 0, 5, None, None, C.byref(n_out)) == capi.GTTS_ERR_UNSUPPORTED


def test_models_3_and_4_host_entry_points(product_lib, oracle):
    # tube_model 3 / 4 (VocalTractModel2<double, 3>, VocalTractModel4<double, 1>): internal rate (30 delays along the
    # tract), control steps and output length against the oracle; anything else is refused; a tract too short for the
    # kernels' converter ring says so
    for tm in (3, 4):
        for var, fs in (("male", 60102), ("female", 70119), ("baby", 140239)):
            v = dict(default_voice(var), tube_model=tm)
            assert g.internal_rate(v) == fs == int(oracle.internal_rate(v))
            tr = T.synthetic_track(4, 7)
            assert g.output_length(v, len(tr))[1] == len(oracle.synthesize(v, tr))
    n_int, n_out = C.c_int64(), C.c_int64()
    bad = capi.voice_config(dict(default_voice("male"), tube_model=1))
    assert product_lib.gtts_output_length(C.byref(bad), 80, 5, C.byref(n_int), C.byref(n_out)) == capi.GTTS_ERR_INVALID
    short = capi.voice_config(dict(default_voice("male"), tube_model=4, vocal_tract_length=5.0))
    assert product_lib.gtts_output_length(C.byref(short), 80, 5, C.byref(n_int), C.byref(n_out)) != capi.GTTS_OK


# ---- control-frame generation: host entry points (gtts_events_*) -------------------------------------------------------

def _golden_events():
    z = np.load(os.path.join(ROOT, "tests", "golden", "events_v1.npz"))
    return z, [str(n) for n in z["names"]]


def test_event_structs_match_header():
    hdr = open(os.path.join(ROOT, "include", "gtts_b200.h")).read()
    assert "typedef struct gtts_event {" in hdr and "typedef struct gtts_event_config {" in hdr
    from oracle.pyoracle import EVENT_DTYPE, EVENT_CONFIG_DTYPE
    assert capi.EVENT_DTYPE.itemsize == EVENT_DTYPE.itemsize == 296
    assert capi.EVENT_CONFIG_DTYPE.itemsize == EVENT_CONFIG_DTYPE.itemsize == 128
    assert [capi.EVENT_CONFIG_DTYPE.fields[n][1] for n in capi.EVENT_CONFIG_DTYPE.names] == \
           [EVENT_CONFIG_DTYPE.fields[n][1] for n in EVENT_CONFIG_DTYPE.names]


def test_events_frame_count_and_drift_setup_match_reference_fixtures(product_lib):
    # the frame counts and the drift generator's constants the unmodified reference had (tests/golden/events_v1.npz)
    from gama_tts_b200.events import event_config, synthetic_events
    from oracle.pyoracle import OracleEvents
    z, names = _golden_events()
    for n in names:
        cfg = z["cfg_" + n].astype(capi.EVENT_CONFIG_DTYPE)
        assert g.events_frame_count(cfg, z["ev_" + n]) == len(z["frames_" + n]), n
    fresh = event_config()
    ref = z["cfg_hello_0"]
    for k in ("drift_deviation2", "drift_offset", "drift_seed", "drift_b0", "drift_b1", "drift_a1", "drift_a2", "drift_x1", "drift_y2"):
        assert fresh[k] == ref[k], k
    cfg = np.zeros(1, capi.EVENT_CONFIG_DTYPE)
    assert product_lib.gtts_events_drift_setup(4.0, 250.0, 0.5, cfg.ctypes.data) == capi.GTTS_ERR_INVALID     # below 1 Hz
    assert product_lib.gtts_events_drift_setup(4.0, 250.0, 121.0, cfg.ctypes.data) == capi.GTTS_ERR_INVALID   # above 0.48 fs
    # synthetic lists, also with events closer than a control period: the host's count is what the oracle produces
    o = OracleEvents()
    for seed in range(8):
        ev = synthetic_events(seed, 3 + seed, tight=seed % 2 == 1)
        for period in (1, 4, 10):
            c = event_config(control_period=period)
            assert g.events_frame_count(c, ev) == len(o.generate(c, ev)[0]), (seed, period)
    assert g.events_frame_count(fresh, ev[:1]) == 0 and g.events_frame_count(fresh, ev[:0]) == 0


def test_events_prepare_fails_loudly_without_a_gpu_or_on_bad_input(product_lib):
    from gama_tts_b200.events import event_config, synthetic_events
    ev = synthetic_events(1, 3)
    eo = np.array([0, len(ev)], np.int64)
    cfg = np.array([event_config()], capi.EVENT_CONFIG_DTYPE)
    h = C.c_void_p()
    assert product_lib.gtts_events_prepare(None, cfg.ctypes.data, None, ev.ctypes.data, eo.ctypes.data, 1, C.byref(h)) == capi.GTTS_ERR_INVALID
    assert product_lib.gtts_events_run_host(None, ev.ctypes.data, None, None) == capi.GTTS_ERR_INVALID
    assert product_lib.gtts_events_run_device(None, None, None, None, None) == capi.GTTS_ERR_INVALID
    n = C.c_int64()
    bad = cfg.copy()
    bad["control_period"] = 0
    assert product_lib.gtts_events_frame_count(bad.ctypes.data, ev.ctypes.data, len(ev), C.byref(n)) == capi.GTTS_ERR_INVALID
