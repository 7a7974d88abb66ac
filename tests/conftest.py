import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
HAVE_REFERENCE_TREE = os.path.isdir("/root/reference/gama_tts/src")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # gpu-marked tests are skipped (not failed) where no CUDA device exists, whatever -m says
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _build_product():
    lib = os.path.join(ROOT, "gama_tts_b200", "csrc", "libgtts_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "gama_tts_b200", "csrc")], check=True)
    return lib


@pytest.fixture(scope="session")
def product_lib():
    _build_product()
    from gama_tts_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled in place (oracle/_ref); skipped when it is not available."""
    from oracle.pyoracle import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError, subprocess.CalledProcessError) as e:
        pytest.skip("oracle/_ref not available: %s" % e)


@pytest.fixture(scope="session")
def reference_nofma():
    from oracle.pyoracle import Reference
    try:
        return Reference("_nofma")
    except (FileNotFoundError, OSError, subprocess.CalledProcessError) as e:
        pytest.skip("oracle/_ref not available: %s" % e)


class Golden:
    def __init__(self):
        self.z = np.load(os.path.join(GOLDEN, "golden_v1.npz"))
        self.names = [str(n) for n in self.z["names"]]

    def case(self, name):
        return (json.loads(str(self.z[name + "/voice"])), self.z[name + "/track"],
                self.z[name + "/ref"], self.z[name + "/ref_nofma"])

    def kat(self, key):
        return self.z["kat/" + key]


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def real_tracks():
    z = np.load(os.path.join(GOLDEN, "real_tracks.npz"))
    return [z["track%d" % i] for i in range(4)]


@pytest.fixture(scope="session")
def synth():
    """The CUDA path on cuda:0 (gpu tests only)."""
    _build_product()
    import gama_tts_b200 as g
    return g.TubeSynthesizer(0)


def full_scale_error(test, ref):
    """max |test - ref| after the reference's own normalisation: peak of the reference -> 0.95
    (VTMUtil.cpp:48-67), i.e. the error as a fraction of full scale."""
    test = np.asarray(test, np.float64)
    ref = np.asarray(ref, np.float64)
    peak = np.abs(ref).max() if len(ref) else 0.0
    if peak < 1e-30:
        return float(np.abs(test - ref).max()) if len(ref) else 0.0
    return float(np.abs(test - ref).max() * 0.95 / peak)


def snr_db(test, ref):
    test = np.asarray(test, np.float64)
    ref = np.asarray(ref, np.float64)
    noise = ((test - ref) ** 2).sum()
    sig = (ref ** 2).sum()
    if noise == 0:
        return float("inf")
    return float(10 * np.log10(sig / noise))


class Golden5:
    """tests/golden/golden5_v1.npz: the reference's model 5 on committed inputs (tools/make_golden5.py)."""

    def __init__(self):
        import json
        self.z = np.load(os.path.join(GOLDEN, "golden5_v1.npz"))
        self.names = [str(n) for n in self.z["names"]]
        self._json = json

    def case(self, name):
        return (self._json.loads(str(self.z["voice_" + name])), self.z["track_" + name], self.z["ref_" + name],
                self.z["nofma_" + name])


@pytest.fixture(scope="session")
def golden5():
    return Golden5()


@pytest.fixture(scope="session")
def oracle5():
    from oracle.pyoracle import Oracle5
    return Oracle5()
