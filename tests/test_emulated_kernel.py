"""The device source (gama_tts_b200/csrc/tube_kernel.cuh) compiled for the host under the SIMT emulator
of tests/simt_emu, against the oracle.  This checks the kernel's blocking, ring indexing, lane roles
and shuffles in the GPU-less container; the real parity gate is tests/test_gpu_parity.py on the B200.
With -ffp-contract=off and pow() the emulated kernel follows the oracle's arithmetic exactly except for
the analytic wavetable fall segment (voices with tn_min != tn_max)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import full_scale_error
from gama_tts_b200 import pack_tracks
from gama_tts_b200 import tracks as T
from gama_tts_b200.capi import voice_array
from gama_tts_b200.voices import default_voice, random_voice

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "simt_emu")], check=True)
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    L.emu_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]

    def run(voices, vidx, tracks, warps=4, ctas=1, rate=250.0, steps=None):
        va = voice_array(voices)
        frames, fo = pack_tracks(tracks)
        vi = np.ascontiguousarray(vidx, np.int32)
        so = None if steps is None else np.ascontiguousarray(steps, np.int32)
        oo = np.zeros(len(tracks) + 1, np.int64)
        ol = np.zeros(len(tracks) + 1, np.int64)
        args = [va, len(voices), vi.ctypes.data, rate, None if so is None else so.ctypes.data, frames.ctypes.data,
                fo.ctypes.data, len(tracks)]
        assert L.emu_batch(*args, None, oo.ctypes.data, ol.ctypes.data, warps, ctas) == 0, L.emu_last_error()
        out = np.zeros(int(oo[-1]), np.float32)
        assert L.emu_batch(*args, out.ctypes.data, oo.ctypes.data, ol.ctypes.data, warps, ctas) == 0
        return [out[oo[i]:oo[i] + ol[i]] for i in range(len(tracks))]
    return run


def test_emulated_kernel_matches_oracle(emu, oracle, real_tracks):
    rng = np.random.Generator(np.random.PCG64(5))
    hello, shells = real_tracks[0], real_tracks[2]
    voices = [default_voice("male"), default_voice("female"), default_voice("baby"), random_voice(rng), random_voice(rng)]
    tracks = [hello[:90], shells[230:290], shells[640:690], T.synthetic_track(3, 50), shells[900:950], hello[:1], hello[:0]]
    vidx = [0, 1, 2, 3, 4, 3, 0]
    res = emu(voices, vidx, tracks, warps=4)
    for vi, tr, out in zip(vidx, tracks, res):
        ref = oracle.synthesize(voices[vi], tr)
        assert len(out) == len(ref)
        assert full_scale_error(out, ref) <= 1e-9
        if voices[vi]["glottal_pulse_tn_min"] == voices[vi]["glottal_pulse_tn_max"]:
            assert np.array_equal(out, ref)


def test_emulated_kernel_steps_override(emu, oracle, real_tracks):
    # steps = 1: every frame is one internal sample (what the plugin shim sends)
    v = default_voice("male")
    params = np.repeat(real_tracks[0][100:112], 9, axis=0)
    out = emu([v], [0], [params], warps=1, steps=[1])[0]
    assert np.array_equal(out, oracle.synthesize_samples(v, params))


def _pipelined_runner(symbol):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "simt_emu")], check=True)
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    fn = getattr(L, symbol)
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]

    def run(voices, vidx, tracks, ctas=1, rate=250.0, steps=None):
        va = voice_array(voices)
        frames, fo = pack_tracks(tracks)
        vi = np.ascontiguousarray(vidx, np.int32)
        so = None if steps is None else np.ascontiguousarray(steps, np.int32)
        oo = np.zeros(len(tracks) + 1, np.int64)
        ol = np.zeros(len(tracks) + 1, np.int64)
        args = [va, len(voices), vi.ctypes.data, rate, None if so is None else so.ctypes.data, frames.ctypes.data,
                fo.ctypes.data, len(tracks)]
        assert fn(*args, None, oo.ctypes.data, ol.ctypes.data, ctas) == 0, L.emu_last_error()
        out = np.full(int(oo[-1]), np.nan, np.float32)
        assert fn(*args, out.ctypes.data, oo.ctypes.data, ol.ctypes.data, ctas) == 0, L.emu_last_error()
        return [out[oo[i]:oo[i] + ol[i]] for i in range(len(tracks))]
    return run


@pytest.fixture(scope="module", params=["v1", "v2"])
def emu_v1(request):
    """The pipelined kernels under the emulator: tube_kernel_v1.cuh (CTA barrier per iteration) and
    tube_kernel_v2.cuh (decoupled roles, walker warps, two-output SRC) run the same tests."""
    return _pipelined_runner("emu_batch_" + request.param)


@pytest.fixture(scope="module")
def emu_v2():
    return _pipelined_runner("emu_batch_v2")


def test_emulated_pipelined_kernel_matches_oracle(emu_v1, oracle, real_tracks):
    # the warp-specialised pipeline (tube_kernel_v1.cuh): 10 utterances over 7 slots, so slots are re-used;
    # ragged lengths, partial last blocks, frame boundaries inside and at the end of blocks, 5 voices
    rng = np.random.Generator(np.random.PCG64(5))
    hello, shells = real_tracks[0], real_tracks[2]
    voices = [default_voice("male"), default_voice("female"), default_voice("baby"), random_voice(rng), random_voice(rng)]
    tracks = [hello[:40], shells[230:262], shells[640:670], T.synthetic_track(3, 25), shells[900:930], hello[:1],
              hello[:0], hello[100:133], hello[200:209], shells[300:320]]
    vidx = [0, 1, 2, 3, 4, 3, 0, 0, 1, 2]
    res = emu_v1(voices, vidx, tracks)
    for vi, tr, out in zip(vidx, tracks, res):
        ref = oracle.synthesize(voices[vi], tr)
        assert len(out) == len(ref) and not np.isnan(out).any()
        assert full_scale_error(out, ref) <= 1e-9
        if voices[vi]["glottal_pulse_tn_min"] == voices[vi]["glottal_pulse_tn_max"]:
            assert np.array_equal(out, ref)


def test_emulated_pipelined_kernel_control_rates(emu_v1, oracle):
    v = default_voice("small_child")          # fs_int 35059: 35 / 70 / 105 / 140 steps per control period
    tr = T.synthetic_track(9, 30)
    for period in (1, 2, 3, 4):
        out = emu_v1([v], [0], [tr], rate=1000.0 / period)[0]
        assert np.array_equal(out, oracle.synthesize(v, tr, control_rate=1000.0 / period))


def test_emulated_pipelined_kernel_aligned_batch(emu_v1, oracle):
    # equally long utterances of one voice step in lockstep: the slot-shared SRC task is used
    v = default_voice("male")
    tracks = [T.synthetic_track(50 + i, 12) for i in range(9)] + [T.synthetic_track(70, 5)]
    for tr, out in zip(tracks, emu_v1([v], [0] * 10, tracks)):
        assert np.array_equal(out, oracle.synthesize(v, tr))


def test_emulated_pipelined_kernel_unaligned_output(emu_v1, oracle, monkeypatch):
    # utterances that do not start on a 32-sample row (the planner never lays them out so, a caller of the
    # kernel could): partial first rows, per-slot row phases; 4 equal utterances keep the shared-SRC path busy
    v = default_voice("male")
    tracks = [T.synthetic_track(40 + i, 14) for i in range(3)] + [T.synthetic_track(50, 9)]
    refs = [oracle.synthesize(v, tr) for tr in tracks]
    for shift in ("0", "13"):
        monkeypatch.setenv("EMU_OUT_SHIFT", shift)
        for out, ref in zip(emu_v1([v], [0] * len(tracks), tracks), refs):
            assert np.array_equal(out, ref)


def test_emulated_pipelined_kernel_per_sample_mode(emu_v1, oracle, real_tracks):
    # steps = 1 (what the plugin shim records: the reference's parameters of every internal sample) runs on the
    # pipelined kernel too: the "walk" is a copy of the frames.  Lengths that are not multiples of the block.
    v = default_voice("male")
    hello = real_tracks[0]
    params = [np.repeat(hello[40:44], 80, axis=0)[:300], np.repeat(hello[100:102], 80, axis=0)[:95],
              T.synthetic_track(5, 33), hello[:1]]
    outs = emu_v1([v], [0] * len(params), params, steps=[1] * len(params))
    for p, out in zip(params, outs):
        assert np.array_equal(out, oracle.synthesize_samples(v, p))


def _loud_track(n=40, seed=11):
    # glottal volume ramps through 60 dB up to 78 dB and back: with tn_min != tn_max the closure point falls below
    # the end of the rise segment and the reference zeroes part of it for good (WavetableGlottalSource.h:162-184)
    tr = T.synthetic_track(seed, n).copy()
    tr[:, 1] = np.concatenate([np.linspace(50, 78, n // 2), np.linspace(78, 40, n - n // 2)]).astype(np.float32)
    return tr


def test_emulated_kernels_reproduce_rise_segment_corruption(emu, emu_v2, oracle):
    v = dict(default_voice("male"))
    v["glottal_pulse_tn_min"], v["glottal_pulse_tn_max"] = 16.0, 32.0
    tr = _loud_track()
    ref = oracle.synthesize(v, tr)
    quiet = tr.copy()
    quiet[:, 1] = np.minimum(quiet[:, 1], 60.0)
    assert full_scale_error(oracle.synthesize(v, quiet)[-2000:], ref[-2000:]) > 1e-3     # the corruption is audible and lasting
    for run in (lambda: emu([v], [0], [tr], warps=1)[0], lambda: emu_v2([v], [0], [tr])[0]):
        out = run()
        assert len(out) == len(ref)
        assert full_scale_error(out, ref) <= 1e-9


def test_emulated_stream_on_pipelined_kernel(oracle, real_tracks):
    # gtts_stream_* on the pipelined kernel: the utterance arrives in pushes of 1..7 frames, every chunk (whole
    # 32-sample blocks) is a launch whose roles resume from the state the previous chunk saved; the result is the
    # batch result, i.e. the oracle's, bit for bit
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "simt_emu")], check=True)
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    L.emu_stream_v2.restype = C.c_longlong
    L.emu_stream_v2.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    rng = np.random.Generator(np.random.PCG64(12))
    for voice, track in ((default_voice("male"), real_tracks[0][:24]), (default_voice("baby"), real_tracks[2][640:652]),
                         (random_voice(rng), T.synthetic_track(3, 9))):
        ref = oracle.synthesize(voice, track)
        for sizes in ([1] * len(track), rng.integers(1, 8, len(track)).tolist(), [len(track)]):
            va = voice_array([voice])
            fr = np.ascontiguousarray(track, np.float32)
            ps = np.ascontiguousarray(sizes, np.int32)
            out = np.full(len(ref) + 64, np.nan, np.float32)
            n = L.emu_stream_v2(va, 250.0, fr.ctypes.data, len(fr), ps.ctypes.data, len(ps), out.ctypes.data, len(out))
            assert n == len(ref), (n, len(ref), L.emu_last_error())
            assert full_scale_error(out[:n], ref) <= 1e-9
            if voice["glottal_pulse_tn_min"] == voice["glottal_pulse_tn_max"]:
                assert np.array_equal(out[:n], ref)


def test_emulated_kernels_down_sampling_branch(emu, emu_v2, oracle):
    # vocal tracts shorter than 7.3 cm: the internal rate exceeds the output rate and the SRC runs its down-sampling
    # loop (SampleRateConverter.h:362-415: variable tap count, rint per output, 2 * pad > 26 flush zeros)
    for length in (7.0, 5.5, 3.2):
        v = dict(default_voice("baby"))
        v["vocal_tract_length"] = length
        tr = T.synthetic_track(21, 6)
        ref = oracle.synthesize(v, tr)
        for run in (lambda: emu([v], [0], [tr], warps=1)[0], lambda: emu_v2([v, default_voice("male")], [0, 1, 0], [tr, tr[:3], tr[:2]])[0]):
            out = run()
            assert len(out) == len(ref)
            assert np.array_equal(out, ref)


def test_emulated_wide_batch_kernel_matches_oracle(oracle, real_tracks):
    # tube_kernel_v3.cuh (one thread per utterance, 32 utterances of a warp in lockstep): 70 ragged utterances, each
    # its own voice (dynamic wavetable), a sine voice, one without noise modulation, an empty and a one-frame track,
    # control periods of 1 sample and of a non-default rate; three groups, so a warp takes more than one
    # (the FIR and SRC sums run on four / two accumulators: a float32 output sample may round the other way)
    WIDE_TOL = 5e-8
    emu = _pipelined_runner("emu_batch_v3")
    rng = np.random.Generator(np.random.PCG64(11))
    voices = [random_voice(rng) for _ in range(66)] + [default_voice("male"), default_voice("female"), default_voice("baby"),
                                                        default_voice("small_child")]
    voices[5]["waveform"] = 1
    voices[7]["noise_modulation"] = 0
    hello = real_tracks[0]
    tracks = [T.synthetic_track(100 + i, int(rng.integers(20, 90))) for i in range(66)] + [hello[:60], hello[:1], hello[:0], hello[100:140]]
    res = emu(voices, list(range(70)), tracks)
    for v, tr, out in zip(voices, tracks, res):
        ref = oracle.synthesize(v, tr)
        assert len(out) == len(ref) and not np.isnan(out).any()
        assert full_scale_error(out, ref) <= WIDE_TOL
    v = default_voice("small_child")
    tr = T.synthetic_track(9, 30)
    assert full_scale_error(emu([v], [0], [tr], rate=500.0)[0], oracle.synthesize(v, tr, control_rate=500.0)) <= WIDE_TOL
    params = np.repeat(hello[100:112], 9, axis=0)
    assert full_scale_error(emu([v], [0], [params], steps=[1])[0], oracle.synthesize_samples(v, params)) <= WIDE_TOL


def test_emulated_model5_kernel_matches_oracle(oracle5, real_tracks):
    # tube5_kernel.cuh under the emulator against oracle/tube5_oracle.c (itself bit-identical to the reference's model 5):
    # with host libm and no FMA contraction the kernel's blocking, lane roles, shuffles and converter reproduce it
    # bit for bit -- all source / output variants, a randomised voice, the shortest tract (39-tap wings), empty and
    # one-frame tracks, two warps sharing the queue
    from gama_tts_b200.capi import voice5_array
    from gama_tts_b200.voices import default_voice5, random_voice5
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "simt_emu")], check=True)
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    fn = L.emu_batch_m5
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    hello, shells = real_tracks[0], real_tracks[2]
    rng = np.random.Generator(np.random.PCG64(3))
    voices = [default_voice5("male"), random_voice5(rng), default_voice5("baby"),
              dict(default_voice5("male"), constant_radius_mouth_impedance=1), dict(default_voice5("male"), waveform=1),
              dict(default_voice5("male"), bypass=1), dict(default_voice5("female"), noise_modulation=0)]
    tracks = [hello[:24], shells[236:252], hello[104:116], hello[:9], hello[:9], hello[:9], hello[44:56], hello[:0], hello[:1]]
    vidx = [0, 1, 2, 3, 4, 5, 6, 0, 0]
    va = voice5_array(voices)
    frames, fo = pack_tracks(tracks)
    vi = np.ascontiguousarray(vidx, np.int32)
    oo = np.zeros(len(tracks) + 1, np.int64)
    ol = np.zeros(len(tracks) + 1, np.int64)
    args = [va, len(voices), vi.ctypes.data, 250.0, None, frames.ctypes.data, fo.ctypes.data, len(tracks)]
    assert fn(*args, None, oo.ctypes.data, ol.ctypes.data, 2) == 0, L.emu_last_error()
    out = np.full(int(oo[-1]), np.nan, np.float32)
    assert fn(*args, out.ctypes.data, oo.ctypes.data, ol.ctypes.data, 2) == 0, L.emu_last_error()
    for u, (v_i, tr) in enumerate(zip(vidx, tracks)):
        ref = oracle5.synthesize(voices[v_i], tr)
        got = out[oo[u]:oo[u] + ol[u]]
        assert len(got) == len(ref)
        assert np.array_equal(got, ref), u


def test_emulated_pipelined_kernel_per_sample_mode(emu_v2, oracle, real_tracks):
    # steps = 1 (what the plugin shim sends: one row per internal sample) on the pipelined kernel, together with a
    # normally stepped utterance in the same CTA: the lanes of a per-sample walk leave walk_block early, the others
    # go on to its warp votes
    v = default_voice("male")
    params = np.repeat(real_tracks[0][100:104], 40, axis=0)
    tr = real_tracks[0][100:120]
    outs = emu_v2([v], [0, 0], [params, tr], steps=[1, 0])
    assert np.array_equal(outs[0], oracle.synthesize_samples(v, params))
    assert np.array_equal(outs[1], oracle.synthesize(v, tr))


def test_emulated_general_kernel_models_3_and_4(emu, oracle, real_tracks):
    # tube_model 3 (three-sample section delay: three interleaved wave states) and 4 (30 + 18 sections, lane = section)
    # on the general kernel, at three times the internal rate (down-sampling converter, up to 39 taps a wing for the
    # shortest tract), mixed with model-0 utterances in one CTA: bit-identical to the oracle
    rng = np.random.Generator(np.random.PCG64(4))
    hello, shells = real_tracks[0], real_tracks[2]
    base = [default_voice("male"), default_voice("baby"), random_voice(rng)]
    voices = [dict(v, tube_model=m) for m in (3, 4, 0) for v in base]
    tracks = [hello[:20], hello[104:114], T.synthetic_track(6, 14)] * 3
    res = emu(voices, list(range(9)), tracks, warps=2)
    for v, tr, out in zip(voices, tracks, res):
        ref = oracle.synthesize(v, tr)
        assert len(out) == len(ref)
        assert full_scale_error(out, ref) <= 1e-9
        if v["glottal_pulse_tn_min"] == v["glottal_pulse_tn_max"]:
            assert np.array_equal(out, ref)


# ---- control-frame generation (events_kernel.cuh) ----------------------------------------------------------------------

@pytest.fixture(scope="module")
def emu_events(emu):
    from gama_tts_b200 import capi
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    L.emu_events.argtypes = [C.c_void_p] * 8 + [C.c_int]

    def run(cfgs, event_lists, continues=None, warps=3):
        cfgs = np.ascontiguousarray(cfgs, capi.EVENT_CONFIG_DTYPE).reshape(-1)
        ev, eo = g_pack_events(event_lists)
        cp = None if continues is None else np.ascontiguousarray(continues, np.int32)
        fo = np.zeros(len(event_lists) + 1, np.int64)
        args = [cfgs.ctypes.data, None if cp is None else cp.ctypes.data, ev.ctypes.data, eo.ctypes.data, len(event_lists)]
        assert L.emu_events(*args, None, fo.ctypes.data, None, warps) == 0, L.emu_last_error()
        frames = np.zeros((max(int(fo[-1]), 1), 16), np.float32)
        out = np.zeros(max(len(event_lists), 1), capi.EVENT_CONFIG_DTYPE)
        assert L.emu_events(*args, frames.ctypes.data, fo.ctypes.data, out.ctypes.data, warps) == 0, L.emu_last_error()
        return [frames[fo[c]:fo[c + 1]] for c in range(len(event_lists))], out[:len(event_lists)]
    return run


def g_pack_events(event_lists):
    from gama_tts_b200 import pack_events
    return pack_events(event_lists)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_emulated_events_kernel_reference_fixtures(emu_events):
    # the frames the unmodified reference front end produced (tests/golden/events_v1.npz), the two-chunk utterances as
    # chains (the second chunk's own drift state is wiped: it must come from the first chunk): bit for bit
    from gama_tts_b200 import capi
    z = np.load(os.path.join(HERE, "golden", "events_v1.npz"))
    names = [str(n) for n in z["names"]]
    cfgs = np.array([z["cfg_" + n] for n in names]).astype(capi.EVENT_CONFIG_DTYPE)
    cont = np.array([n.endswith("_1") for n in names], np.int32)
    for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
        cfgs[k][cont == 1] = 0.0
    frames, out = emu_events(cfgs, [z["ev_" + n] for n in names], cont)
    for i, n in enumerate(names):
        assert frames[i].shape == z["frames_" + n].shape, n
        assert np.array_equal(_bits(frames[i]), _bits(z["frames_" + n])), n
    # the state a first chunk leaves is the state the reference's second chunk started from
    i = names.index("shells_0")
    for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
        assert out[k][i] == z["cfg_shells_1"][k], k


def test_emulated_events_kernel_synthetic_lists_vs_oracle(emu_events):
    # synthetic lists: every flag combination, control periods 1 / 4 / 10 ms, events closer than a period or at equal
    # times, special parameters, lists of 0 / 1 / 2 events, a three-chunk utterance
    from gama_tts_b200.events import event_config, synthetic_events
    from oracle.pyoracle import OracleEvents
    o = OracleEvents()
    cfgs, lists, cont = [], [], []
    for seed in range(16):
        cfgs.append(event_config(control_period=(4, 1, 10)[seed % 3], macro=seed & 1, micro=(seed >> 1) & 1,
                                 drift=(seed >> 2) & 1, smooth=(seed >> 3) & 1))
        lists.append(synthetic_events(100 + seed, 2 + seed % 7, special_rate=0.05, tight=seed % 2 == 0))
        cont.append(0)
    for n in (0, 1, 2):
        cfgs.append(event_config())
        lists.append(synthetic_events(7, 3)[:n])
        cont.append(0)
    for k in range(3):
        cfgs.append(event_config())
        lists.append(synthetic_events(200 + k, 5))
        cont.append(int(k > 0))
    # lists longer than the kernel's occupancy-mask window (512 rows): the window moves, and scans run past its end
    for k, n_post in enumerate((150, 400)):
        cfgs.append(event_config(smooth=k))
        lists.append(synthetic_events(300 + k, n_post, special_rate=0.004, tight=k == 1))
        cont.append(0)
    # signed zeros: the kernel adds a zero delta where the reference skips it; values and targets of -0.0 / +0.0 and
    # ramps that land exactly on zero must still give the reference's bits (mean pitch 0 keeps a zero pitch a zero)
    rng = np.random.Generator(np.random.PCG64(5))
    for k in range(6):
        ev = synthetic_events(400 + k, 6)
        vals = np.array([-0.0, 0.0, 1.0, -1.0, 0.25, -0.25])
        keep = ~np.isinf(ev["param"])
        ev["param"] = np.where(keep, vals[rng.integers(0, 6, ev["param"].shape)], np.inf)
        keep = rng.random(ev["special"].shape) < 0.2
        ev["special"] = np.where(keep, vals[rng.integers(0, 6, ev["special"].shape)], np.inf)
        ev["time"] = 8 * np.arange(len(ev))
        cfgs.append(event_config(macro=0, micro=1, drift=0, mean_pitch=0.0 if k % 2 else -0.0))
        lists.append(ev)
        cont.append(0)
    frames, out = emu_events(np.array(cfgs), lists, cont)
    carried = None
    for i, (c, ev) in enumerate(zip(cfgs, lists)):
        if cont[i]:
            c = c.copy()
            for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
                c[k] = carried[k]
        want, carried = o.generate(c, ev)
        assert frames[i].shape == want.shape, i
        assert np.array_equal(_bits(frames[i]), _bits(want)), i
        for k in ("drift_seed", "drift_y1"):
            assert out[k][i] == carried[k], (i, k)


def test_events_plan_refusals():
    # the host planner behind gtts_events_prepare (events_host.cpp, here through the emulator's entry point, which needs
    # no device): what it refuses and why
    from gama_tts_b200 import capi, pack_events
    from gama_tts_b200.events import event_config, synthetic_events
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "simt_emu")], check=True)
    L = C.CDLL(os.path.join(HERE, "simt_emu", "libemu_tube.so"))
    L.emu_last_error.restype = C.c_char_p
    L.emu_events.argtypes = [C.c_void_p] * 8 + [C.c_int]

    def plan(cfgs, lists, offsets=None):
        cfgs = np.ascontiguousarray(cfgs, capi.EVENT_CONFIG_DTYPE).reshape(-1)
        ev, eo = pack_events(lists)
        if offsets is not None:
            eo = np.ascontiguousarray(offsets, np.int64)
        fo = np.zeros(len(eo), np.int64)
        rc = L.emu_events(cfgs.ctypes.data, None, ev.ctypes.data, eo.ctypes.data, len(eo) - 1, None, fo.ctypes.data, None, 1)
        return rc, L.emu_last_error().decode(), fo

    ev = synthetic_events(3, 4)
    rc, _, fo = plan([event_config()], [ev])
    assert rc == 0 and fo[1] > 0
    bad = event_config()
    bad["control_period"] = 0
    rc, msg, _ = plan([bad], [ev])
    assert rc == capi.GTTS_ERR_INVALID and "control_period" in msg
    neg = ev.copy()
    neg["time"][2] = -4
    rc, msg, _ = plan([event_config()], [neg])
    assert rc == capi.GTTS_ERR_INVALID and "time" in msg
    rc, msg, _ = plan([event_config(), event_config()], [ev, ev], offsets=[0, len(ev), len(ev) - 1])
    assert rc == capi.GTTS_ERR_INVALID and "decrease" in msg
    rc, msg, _ = plan([event_config()], [ev], offsets=[1, len(ev)])
    assert rc == capi.GTTS_ERR_INVALID and "event_offsets[0]" in msg
    far = ev.copy()
    far["time"][-1] = 1 << 30
    huge = event_config()
    huge["control_period"] = 1
    rc, msg, _ = plan([huge], [far])
    assert rc == 0                                              # 2^30 ms at a period of 1 ms: the largest chunk accepted
    many = np.repeat(ev[:2], 600000)                            # 1.2 M events at the same two times: a frame each, period 1000 ms
    many["time"] = 0
    slow = event_config()
    slow["control_period"] = 1000
    rc, msg, _ = plan([slow], [many])
    assert rc == capi.GTTS_ERR_INVALID and "too long" in msg    # its clock would pass 2^30 ms
    # an empty batch and a batch of empty lists are fine: no frames
    rc, _, fo = plan([event_config(), event_config()], [ev[:0], ev[:1]])
    assert rc == 0 and fo[-1] == 0
