"""The drop-in boundary: the reference's own plugin loader (VocalTractModelPlugin.cpp:57-91, compiled
unmodified into oracle/_ref with ENABLE_VTM_PLUGINS) loads gama_tts_b200/csrc/libgtts_plugin.so through
`model = 2000` / `dll_path`, and the reference-side Controller loop drives it."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, full_scale_error
from gama_tts_b200.voices import default_voice

PLUGIN = os.path.join(ROOT, "gama_tts_b200", "csrc", "libgtts_plugin.so")


def _need_plugin():
    if not os.path.exists(PLUGIN):
        pytest.skip("libgtts_plugin.so not built (needs the reference headers at build time)")


def test_plugin_exports_the_two_symbols(product_lib):
    _need_plugin()
    L = C.CDLL(PLUGIN)
    assert hasattr(L, "GAMA_TTS_construct_vocal_tract_model")
    assert hasattr(L, "GAMA_TTS_destruct_vocal_tract_model")
    L.GAMA_TTS_construct_vocal_tract_model.restype = C.c_void_p
    L.GAMA_TTS_construct_vocal_tract_model.argtypes = [C.c_void_p, C.c_int]
    assert L.GAMA_TTS_construct_vocal_tract_model(None, 0) is None


def test_plugin_fails_loudly_without_gpu(reference, real_tracks, product_lib):
    import torch
    _need_plugin()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    # no device -> construct returns NULL -> the reference throws "Could not construct the vocal tract model"
    with pytest.raises(RuntimeError, match="Could not construct"):
        reference.synthesize(default_voice("male"), real_tracks[0][:5], model=2000, extra={"dll_path": PLUGIN})


@pytest.mark.gpu
def test_reference_drives_plugin_through_its_own_seam(reference, oracle, real_tracks, product_lib):
    _need_plugin()
    for var, track in (("male", real_tracks[0]), ("female", real_tracks[1][300:420])):
        v = default_voice(var)
        out = reference.synthesize(v, track, model=2000, extra={"dll_path": PLUGIN})
        builtin = reference.synthesize(v, track, model=0)
        assert len(out) == len(builtin)
        assert full_scale_error(out, builtin) <= 2e-7
        assert full_scale_error(out, oracle.synthesize(v, track)) <= 2e-7


@pytest.mark.gpu
def test_reference_drives_plugin_model5_voice(reference, oracle5, real_tracks, product_lib):
    # a 5_xxx voice (model 5's configuration keys) with model = 2000: the shim stands in for VocalTractModel5
    from gama_tts_b200.voices import default_voice5
    _need_plugin()
    for var, track in (("male", real_tracks[0][:200]), ("female", real_tracks[1][300:380])):
        v = default_voice5(var)
        out = reference.synthesize5(v, track, model=2000, extra={"dll_path": PLUGIN})
        builtin = reference.synthesize5(v, track)
        assert len(out) == len(builtin)
        assert full_scale_error(out, builtin) <= 2e-7
        assert np.array_equal(builtin, oracle5.synthesize(v, track))


@pytest.mark.gpu
def test_interactive_call_pattern_through_the_seam(reference, real_tracks, product_lib):
    # The editor's pattern (InteractiveAudio.cpp:131-186): the model is constructed interactive, stepped until
    # outputBuffer() holds a callback's worth of samples, drained, stepped on.  The plugin serves it in chunks of 64
    # internal samples through gtts_stream_*; what the host hears must be the built-in model's samples.
    _need_plugin()
    v = default_voice("male")
    track = real_tracks[0][:150]
    steps = 80
    cur = np.repeat(track, steps, axis=0).astype(np.float32)                 # a held-parameter stand-in for the editor's filters
    ramp = np.linspace(0.0, 1.0, steps, endpoint=False, dtype=np.float32)[None, :, None]
    nxt = np.concatenate([track[1:], track[-1:]])
    params = (track[:, None, :] + (nxt - track)[:, None, :] * ramp).reshape(-1, 16).astype(np.float32)
    assert params.shape == cur.shape
    for frames in (1024, 256):
        builtin = reference.interactive(v, params, callback_frames=frames, model=0)
        plugin = reference.interactive(v, params, callback_frames=frames, model=2000, extra={"dll_path": PLUGIN})
        n = min(len(builtin), len(plugin))
        assert n > 0.9 * len(params) * 48000 / 20034 and abs(len(builtin) - len(plugin)) <= 64 * 3 + frames
        assert full_scale_error(plugin[:n], builtin[:n]) <= 2e-7


@pytest.mark.gpu
def test_reference_drives_plugin_as_models_3_and_4(reference, real_tracks, product_lib):
    # plugin_tube_model = 3 | 4 next to model = 2000: the shim stands in for VocalTractModel2<double, 3> /
    # VocalTractModel4<double, 1> (same configuration keys as model 0)
    _need_plugin()
    v = default_voice("male")
    track = real_tracks[0][:120]
    for tm in (3, 4):
        out = reference.synthesize(v, track, model=2000, extra={"dll_path": PLUGIN, "plugin_tube_model": tm})
        builtin = reference.synthesize(v, track, model=tm)
        assert len(out) == len(builtin)
        assert full_scale_error(out, builtin) <= 2e-7


STOCK = os.path.join(ROOT, "oracle", "_ref", "gama_tts")
STOCK_VOICE = os.path.join(ROOT, "oracle", "_ref", "voice_0_male")


def _stock_run(tmp_path, name, params_txt, model, extra=""):
    """Runs the UNMODIFIED reference program `gama_tts vtm <voice dir> <parameter file> <out.wav>` (main.cpp:286-337)
    on a copy of the shipped voice whose vtm.txt selects `model`; returns (exit code, stderr, int16 payload)."""
    import shutil
    import subprocess
    voice = tmp_path / ("voice_" + name)
    shutil.copytree(STOCK_VOICE, voice)
    vtm = (voice / "vtm.txt").read_text().replace("model = 0", "model = %d" % model) + extra
    (voice / "vtm.txt").write_text(vtm)
    wav = tmp_path / (name + ".wav")
    r = subprocess.run([STOCK, "vtm", str(voice), str(params_txt), str(wav)], capture_output=True, text=True)
    data = np.fromfile(wav, np.int16)[22:] if wav.exists() else np.zeros(0, np.int16)     # 44-byte header
    return r.returncode, r.stderr, data


@pytest.mark.gpu
def test_stock_gama_tts_binary_drives_the_plugin(tmp_path, real_tracks, product_lib):
    # the stock command-line program (every source of the reference, compiled unmodified
    # with its plugin option, oracle/Makefile) loads libgtts_plugin.so through `model = 2000` / `dll_path` and writes a
    # WAVE file; the same program with its built-in model 0 is the comparison.  Both read the same parameter file, so
    # the 16-bit payloads agree to one LSB (the float32 audio agrees to 2e-7 of full scale).
    _need_plugin()
    if not (os.path.exists(STOCK) and os.path.isdir(STOCK_VOICE)):
        pytest.skip("oracle/_ref/gama_tts not built (needs the reference sources at build time)")
    params = tmp_path / "params.txt"
    np.savetxt(params, real_tracks[0], fmt="%.9g")
    rc0, err0, builtin = _stock_run(tmp_path, "builtin", params, 0)
    assert rc0 == 0, err0
    rc1, err1, plugin = _stock_run(tmp_path, "plugin", params, 2000, "\ndll_path = %s\n" % PLUGIN)
    assert rc1 == 0, err1
    assert len(plugin) == len(builtin) > 1000
    assert np.abs(plugin.astype(np.int32) - builtin.astype(np.int32)).max() <= 1
    assert (plugin == builtin).mean() > 0.99


def test_stock_gama_tts_binary_reports_a_missing_gpu(tmp_path, real_tracks, product_lib):
    import torch
    _need_plugin()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not (os.path.exists(STOCK) and os.path.isdir(STOCK_VOICE)):
        pytest.skip("oracle/_ref/gama_tts not built")
    params = tmp_path / "params.txt"
    np.savetxt(params, real_tracks[0][:20], fmt="%.9g")
    rc, err, data = _stock_run(tmp_path, "nogpu", params, 2000, "\ndll_path = %s\n" % PLUGIN)
    assert rc != 0 and "Could not construct" in err and len(data) == 0


# ---- control-frame generation through the reference's own front end ----------------------

REF_EVENTS = os.path.join(ROOT, "oracle", "_ref", "ref_events")
VOICE_DIR = os.path.join(ROOT, "oracle", "_ref", "voice_0_male")
PRODUCT = os.path.join(ROOT, "gama_tts_b200", "csrc", "libgtts_b200.so")


def _run_ref_events(tmp_path, text, flags=None, lib=True):
    import subprocess
    if not (os.path.exists(REF_EVENTS) and os.path.isdir(VOICE_DIR)):
        pytest.skip("oracle/_ref/ref_events not built (needs the reference sources at build time)")
    env = dict(os.environ)
    if lib:
        env["GTTS_LIB"] = PRODUCT
    if flags:
        env["REF_EVENTS_FLAGS"] = flags
    return subprocess.run([REF_EVENTS, VOICE_DIR, str(tmp_path / "events.bin"), text], env=env, capture_output=True, text=True,
                          timeout=300)


@pytest.mark.gpu
def test_reference_front_end_drives_the_control_frame_kernel(tmp_path, product_lib):
    # The unmodified reference front end (text parser, rule engine, generateEventList, applyIntonation) builds the event
    # lists of a text; inside its process EventList::list_ is copied into gtts_event records and gtts_events_prepare /
    # gtts_events_run_host take the place of generateOutput (the binding of INTEGRATION.md, compiled into
    # oracle/ref_events.cpp); the frames must be the reference's own, bit for bit.  The chunks of the utterance are one
    # batch, later chunks chained to the first (continues_previous), so the drift generator's state is carried on the device.
    cases = [("Hello world.", None),
             ("She sells sea shells by the sea shore. The shells she sells are surely sea shells. Are they?", None),
             ("Why did the old clock stop? Nobody wound it.", "1,1,1,0"),
             ("Monotone machines speak like this, all day long.", "0,0,1,1"),
             ("In 1984, 3 of 12 ships sailed at 6:45.", "1,0,0,1")]
    for text, flags in cases:
        r = _run_ref_events(tmp_path, text, flags)
        assert r.returncode == 0, (text, r.stdout[-300:], r.stderr[-300:])
        line = [l for l in r.stdout.splitlines() if l.startswith("gpu_check")]
        assert line and line[0].split()[-2:] == ["mismatches", "0"], (text, r.stdout[-300:])
        assert int(line[0].split()[4]) > 100                         # frames


def test_reference_front_end_reports_a_missing_gpu(tmp_path, product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run_ref_events(tmp_path, "Hello world.")
    assert r.returncode == 4 and "gtts_create" in r.stderr               # no CPU path behind the binding
