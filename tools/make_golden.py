#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built in place from
/root/reference by oracle/Makefile) -- run in the build container, where /root/reference exists.

    python tools/make_golden.py [--cli /tmp/gtts_cli]

* real_tracks.npz : control tracks captured once from the reference front end
  (`echo "<sentence>" | gama_tts tts -p track.txt data/voice/english/0_male out.wav`; the front end
  is non-deterministic, so the capture is frozen here).  If --cli has no s1..s4.txt the existing
  file is kept.
* golden_v1.npz   : for each case a voice (JSON), a float32 track and the reference's raw float32
  outputBuffer() from the default (FMA-contracted) build and from the -ffp-contract=off build,
  plus component known-answer vectors (noise, FIR taps, SRC table samples, constants).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gama_tts_b200 import tracks as T  # noqa: E402
from gama_tts_b200.voices import default_voice, random_voice  # noqa: E402
from oracle.pyoracle import Reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SENTENCES = ["Hello world.", "The quick brown fox jumps over the lazy dog.",
             "She sells sea shells by the sea shore, and the shells she sells are sea shells I am sure.",
             "Nine men and many more names."]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cli", default="/tmp/gtts_cli")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)

    real_path = os.path.join(GOLD, "real_tracks.npz")
    files = [os.path.join(args.cli, "s%d.txt" % (i + 1)) for i in range(len(SENTENCES))]
    if all(os.path.exists(f) for f in files):
        real = {"track%d" % i: np.loadtxt(f, dtype=np.float32) for i, f in enumerate(files)}
        np.savez_compressed(real_path, sentences=np.array(SENTENCES), **real)
    real = np.load(real_path)
    hello, fox, shells = real["track0"], real["track1"], real["track2"]

    ref, ref_nofma = Reference(), Reference("_nofma")
    rng = np.random.Generator(np.random.PCG64(1234))
    rv1, rv2 = random_voice(rng), random_voice(rng)
    sine = default_voice("male")
    sine["waveform"] = 1
    nomod = default_voice("female")
    nomod["noise_modulation"] = 0
    cases = [
        ("vowel_aa_male", default_voice("male"), np.tile(T.VOWEL_AA, (250, 1))),
        ("hello_world_male", default_voice("male"), hello),
        ("shells_female", default_voice("female"), shells[200:330]),
        ("fox_large_child", default_voice("large_child"), fox[300:400]),
        ("fox_small_child", default_voice("small_child"), fox[500:580]),
        ("shells_baby", default_voice("baby"), shells[600:700]),
        ("synthetic_random_voice_1", rv1, T.synthetic_track(11, 120)),
        ("synthetic_random_voice_2", rv2, T.synthetic_track(12, 100)),
        ("synthetic_sine_male", sine, T.synthetic_track(13, 60)),
        ("synthetic_nomod_female", nomod, T.synthetic_track(14, 60)),
        ("single_frame", default_voice("male"), hello[100:101]),
        ("empty", default_voice("male"), hello[:0]),
    ]
    out = {"names": np.array([c[0] for c in cases])}
    for name, voice, track in cases:
        track = np.ascontiguousarray(track, np.float32).reshape(-1, 16)
        out[name + "/voice"] = np.array(json.dumps(voice))
        out[name + "/track"] = track
        out[name + "/ref"] = ref.synthesize(voice, track)
        out[name + "/ref_nofma"] = ref_nofma.synthesize(voice, track)
        print("%-28s frames %4d  out %6d  peak %.3e" % (name, len(track), len(out[name + "/ref"]),
                                                         np.abs(out[name + "/ref"]).max() if len(out[name + "/ref"]) else 0))
    # component KATs (from the no-FMA build: the exact-IEEE values)
    n = ref_nofma.noise(1000000)
    out["kat/noise_first16"] = n[:16]
    out["kat/noise_999999"] = n[999999:1000000]
    out["kat/noise_sum"] = np.array([n.sum()])
    out["kat/fir_taps"] = ref_nofma.fir_taps()
    h, dh, incs, pad = ref_nofma.src_tables(20034.0, 48000.0)
    out["kat/src_h_every64"] = h[::64]
    out["kat/src_dh_every64"] = dh[::64]
    out["kat/src_inc_male"] = np.array([incs[0], pad])
    for var in ("male", "female", "large_child", "small_child", "baby"):
        out["kat/constants_" + var] = ref_nofma.constants(default_voice(var))
    x = np.random.Generator(np.random.PCG64(5)).standard_normal(5000)
    out["kat/src_in"] = x
    out["kat/src_out_20034"] = ref_nofma.src_run(20034.0, 48000.0, x)
    np.savez_compressed(os.path.join(GOLD, "golden_v1.npz"), **out)
    print("wrote", os.path.join(GOLD, "golden_v1.npz"), os.path.getsize(os.path.join(GOLD, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
