#!/usr/bin/env python
"""Wide-batch kernel (tube_kernel_v3: one thread per utterance) against the oracle and against the pipelined kernel
on the config-3 slice of bench.py.  python tools/v3_probe.py [--utts 16384] [--parity 256]"""
import argparse
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=16384)
    ap.add_argument("--parity", type=int, default=256)
    ap.add_argument("--kernels", default="v2,v3")
    a = ap.parse_args()
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import random_voice, default_voice
    from conftest import full_scale_error
    from oracle.pyoracle import Oracle
    import bench
    synth = g.TubeSynthesizer(0)
    res = {}
    if a.parity > 0:
        os.environ["GTTS_KERNEL"] = "v3"
        rng = np.random.Generator(np.random.PCG64(21))
        voices = [random_voice(np.random.Generator(np.random.PCG64(500 + u))) for u in range(a.parity - 2)]
        voices += [default_voice("male"), default_voice("female")]
        tracks = [T.synthetic_track(900 + u, int(rng.integers(20, 400))) for u in range(a.parity)]
        outs = synth.synthesize(voices, tracks, voice_index=np.arange(a.parity))
        orc = Oracle()
        worst = 0.0
        for v, tr, out in zip(voices, tracks, outs):
            ref = orc.synthesize(v, tr)
            assert len(ref) == len(out)
            worst = max(worst, full_scale_error(out, ref))
        res["parity"] = {"utterances": a.parity, "worst_full_scale_error": worst}
        print(json.dumps(res["parity"]), flush=True)
    import torch
    peak = synth.fp64_peak_tflops()
    args = types.SimpleNamespace(config3_utts=a.utts)
    for kern in a.kernels.split(","):
        os.environ["GTTS_KERNEL"] = kern
        r = bench.config34_leg(synth, 0, 1, None, torch.cuda.synchronize, lambda x: x, peak, args)
        res[kern] = r
        print(kern, json.dumps({k: r[k] for k in ("utterances", "ms", "value", "roofline_frac", "finite")}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "v3_probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
