#!/usr/bin/env python
"""The drop-in seam, timed: the reference's own Controller loop (oracle/_ref, unmodified) synthesising one real
sentence with its built-in model 0 and with this library loaded through `model = 2000` / `dll_path`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from gama_tts_b200.voices import default_voice
    from pyoracle import Reference
    from conftest import full_scale_error
    plugin = os.path.join(ROOT, "gama_tts_b200", "csrc", "libgtts_plugin.so")
    z = np.load(os.path.join(ROOT, "tests", "golden", "real_tracks.npz"))
    ref = Reference()
    v = default_voice("male")
    out = {}
    for name in ("track0", "track1"):
        track = z[name]

        def best(fn, n=5):
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                r = fn()
                ts.append(time.perf_counter() - t0)
            return min(ts) * 1e3, r
        ref.synthesize(v, track, model=2000, extra={"dll_path": plugin})
        t_plug, y = best(lambda: ref.synthesize(v, track, model=2000, extra={"dll_path": plugin}))
        t_cpu, y0 = best(lambda: ref.synthesize(v, track, model=0))
        out[name] = {"frames": int(len(track)), "audio_s": len(y0) / 48000.0, "builtin_model0_ms": t_cpu,
                     "through_plugin_on_b200_ms": t_plug, "full_scale_error": full_scale_error(y, y0)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
