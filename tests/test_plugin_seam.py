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
