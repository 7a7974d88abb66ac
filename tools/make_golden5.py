#!/usr/bin/env python
"""Regenerates tests/golden/golden5_v1.npz: outputs of the UNMODIFIED reference's model 5 (oracle/_ref, needs
/root/reference) on committed inputs -- the pins of oracle/tube5_oracle.c and of the model-5 kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gama_tts_b200 import tracks as T                      # noqa: E402
from gama_tts_b200.voices import default_voice5, random_voice5   # noqa: E402
from oracle.pyoracle import Reference                       # noqa: E402


def main():
    ref, ref_nofma = Reference(), Reference("_nofma")
    rt = np.load(os.path.join(ROOT, "tests", "golden", "real_tracks.npz"))
    hello, shells = rt["track0"], rt["track2"]
    rng = np.random.Generator(np.random.PCG64(55))
    cases = {
        "male_hello": (default_voice5("male"), hello[:100]),
        "female_shells": (default_voice5("female"), shells[230:300]),
        "large_child": (default_voice5("large_child"), hello[100:150]),
        "small_child": (default_voice5("small_child"), shells[640:690]),
        "baby": (default_voice5("baby"), hello[180:220]),
        "random_a": (random_voice5(rng), T.synthetic_track(41, 70)),
        "random_b": (random_voice5(rng), shells[900:960]),
        "constant_mouth_radius": (dict(default_voice5("male"), constant_radius_mouth_impedance=1), hello[:40]),
        "sine": (dict(default_voice5("male"), waveform=1), hello[:40]),
        "bypass": (dict(default_voice5("male"), bypass=1), hello[:40]),
        "no_modulation": (dict(default_voice5("male"), noise_modulation=0), hello[40:80]),
        "empty": (default_voice5("male"), hello[:0]),
        "one_frame": (default_voice5("male"), hello[:1]),
    }
    out = {"names": np.array(list(cases))}
    for name, (v, tr) in cases.items():
        a, b = ref.synthesize5(v, tr), ref_nofma.synthesize5(v, tr)
        out["voice_" + name] = np.array(json.dumps(v))
        out["track_" + name] = np.ascontiguousarray(tr, np.float32)
        out["ref_" + name] = a
        out["nofma_" + name] = b
        print(name, len(tr), len(a), "peak %.4g" % (np.abs(a).max() if len(a) else 0), "fma==nofma", np.array_equal(a, b))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden5_v1.npz"), **out)


if __name__ == "__main__":
    main()
