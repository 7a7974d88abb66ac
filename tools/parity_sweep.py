#!/usr/bin/env python
"""One-off wide parity sweep on a GPU box: python tools/parity_sweep.py [n_utt] [seed]
Ragged config-3-shaped batch (one random voice per utterance), EVERY utterance checked against the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice, random_voice
    from pyoracle import Oracle
    from conftest import full_scale_error
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    rng = np.random.Generator(np.random.PCG64(seed))
    oracle = Oracle()
    lengths = T.config3_lengths(n, seed=seed, lo=20, hi=600)
    names = ("male", "female", "large_child", "small_child", "baby")
    voices = [random_voice(rng) if i % 4 else default_voice(names[(i // 4) % 5]) for i in range(n)]
    seeds = [T.synthetic_track(seed * 1000 + i, 600) for i in range(48)]
    tracks = [seeds[i % 48][int(rng.integers(0, 600 - lengths[i] + 1)):][: int(lengths[i])] for i in range(n)]
    synth = g.TubeSynthesizer(0)
    t0 = time.time()
    outs = synth.synthesize(voices, tracks, voice_index=np.arange(n))
    print("synthesised %d utterances in %.2f s" % (n, time.time() - t0))
    worst, bad = 0.0, 0
    t0 = time.time()
    for i in range(n):
        ref = oracle.synthesize(voices[i], tracks[i])
        if len(ref) != len(outs[i]):
            print("LENGTH", i, len(ref), len(outs[i]))
            bad += 1
            continue
        e = full_scale_error(outs[i], ref)
        worst = max(worst, e)
        if e > 2e-7:
            bad += 1
            if bad <= 10:
                d = np.abs(outs[i] - ref)
                w = np.nonzero(d > 1e-6 * max(np.abs(ref).max(), 1e-30))[0]
                print("utt %d frames %d err %.3g fs %d bad samples %d first %s" % (i, len(tracks[i]), e, g.internal_rate(voices[i]), len(w), w[:3]))
    print("oracle pass %.1f s; worst error %.3g of full scale; %d of %d utterances above 2e-7" % (time.time() - t0, worst, bad, n))


if __name__ == "__main__":
    main()
