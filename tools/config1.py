#!/usr/bin/env python
"""BASELINE config 1: the reference's own CPU-runnable case, one sentence ("Hello world.", voice 0_male/male).
The control track was captured once from `gama_tts tts -p` (tests/golden/real_tracks.npz); this prints the
reference's synthesis-loop wall time on this host, ours for the same single utterance, and the parity."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def best(fn, n=5):
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return min(t), r


def main():
    import json
    import gama_tts_b200 as g
    from gama_tts_b200.voices import default_voice
    from pyoracle import Reference
    from conftest import full_scale_error
    z = np.load(os.path.join(ROOT, "tests", "golden", "real_tracks.npz"))
    track = z["track0"]
    v = default_voice("male")
    ref = Reference()
    t_ref, y_ref = best(lambda: ref.synthesize(v, track))
    synth = g.TubeSynthesizer(0)
    synth.synthesize(v, [track])
    t_gpu, y = best(lambda: synth.synthesize(v, [track])[0])
    audio = len(y_ref) / 48000.0
    print(json.dumps({"config": "1: 'Hello world.' (%d frames, %.2f s of audio), voice 0_male/male" % (len(track), audio),
                      "reference_synth_loop_ms": t_ref * 1e3, "reference_x_realtime": audio / t_ref,
                      "b200_single_utterance_ms": t_gpu * 1e3, "b200_x_realtime": audio / t_gpu,
                      "full_scale_error": full_scale_error(y, y_ref), "samples": int(len(y))}))


if __name__ == "__main__":
    main()
