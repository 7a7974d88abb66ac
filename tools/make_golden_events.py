#!/usr/bin/env python
"""Regenerates tests/golden/events_v1.npz: event lists and the control frames the UNMODIFIED reference front end made of
them (oracle/_ref/ref_events, built by oracle/Makefile from every reference source; needs /root/reference): the pins of
oracle/events_oracle.c and of the device path (gtts_events_*)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EVENT_DTYPE = np.dtype([("time", "<i4"), ("has_interp", "<i4"), ("param", "<f8", 16), ("special", "<f8", 16),
                        ("a", "<f8"), ("b", "<f8"), ("c", "<f8"), ("d", "<f8")])
CONFIG_DTYPE = np.dtype([("control_period", "<i4"), ("macro_intonation", "<i4"), ("micro_intonation", "<i4"),
                         ("intonation_drift", "<i4"), ("smooth_intonation", "<i4"), ("pad_", "<i4"),
                         ("initial_pitch", "<f8"), ("mean_pitch", "<f8"), ("drift_deviation2", "<f8"), ("drift_offset", "<f8"),
                         ("drift_seed", "<f8"), ("drift_b0", "<f8"), ("drift_b1", "<f8"), ("drift_a1", "<f8"), ("drift_a2", "<f8"),
                         ("drift_x1", "<f8"), ("drift_x2", "<f8"), ("drift_y1", "<f8"), ("drift_y2", "<f8")])


def parse(path):
    """-> list of (config record, events record array, frames [F, 16] float32), one per chunk."""
    raw = open(path, "rb").read()
    pos, chunks = 0, []
    while pos < len(raw):
        head = np.frombuffer(raw, "<i4", 7, pos)
        assert head[0] == 0x45564E54
        pos += 28
        dbl = np.frombuffer(raw, "<f8", 13, pos)
        pos += 104
        cfg = np.zeros((), CONFIG_DTYPE)
        cfg["control_period"], cfg["macro_intonation"], cfg["micro_intonation"] = head[2], head[3], head[4]
        cfg["intonation_drift"], cfg["smooth_intonation"] = head[5], head[6]
        for k, name in enumerate(CONFIG_DTYPE.names[6:]):
            cfg[name] = dbl[k]
        ev = np.frombuffer(raw, EVENT_DTYPE, int(head[1]), pos).copy()
        pos += EVENT_DTYPE.itemsize * int(head[1])
        n_frames = int(np.frombuffer(raw, "<i4", 1, pos)[0])
        pos += 4
        frames = np.frombuffer(raw, "<f4", n_frames * 16, pos).reshape(n_frames, 16).copy()
        pos += n_frames * 64
        chunks.append((cfg, ev, frames))
    return chunks


TEXTS = {   # name: (voice, text, intonation flags "macro,micro,drift,smooth" or None = the voice's own)
    "hello": ("0_male", "Hello world.", None),
    "shells": ("0_male", "She sells sea shells by the sea shore. The shells she sells are surely sea shells.", None),
    "question": ("0_male", "Is the quick brown fox really jumping over the lazy dog, or not?", None),
    "numbers_5male": ("5_male", "In 1984, 3 of 12 ships sailed at 6:45.", None),
    "linear": ("0_male", "Why did the old clock stop? Nobody wound it.", "1,1,1,0"),
    "nodrift": ("0_male", "A short one, with a comma.", "1,1,0,1"),
    "flat": ("0_male", "Monotone machines speak like this.", "0,0,1,1"),
    "micro_only": ("5_male", "Only the micro intonation remains!", "0,1,0,0"),
}


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_events")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/ref_events"], check=True)
    out = {"names": []}
    for name, (voice, text, flags) in TEXTS.items():
        with tempfile.NamedTemporaryFile(suffix=".bin") as tmp:
            subprocess.run([exe, os.path.join("/root/reference/data/voice/english", voice), tmp.name, text], check=True,
                           env=dict(os.environ, **({"REF_EVENTS_FLAGS": flags} if flags else {})), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for k, (cfg, ev, frames) in enumerate(parse(tmp.name)):
                key = "%s_%d" % (name, k)
                out["names"].append(key)
                out["cfg_" + key], out["ev_" + key], out["frames_" + key] = cfg, ev, frames
                print(key, "events", len(ev), "frames", len(frames), "period", int(cfg["control_period"]),
                      "flags", int(cfg["macro_intonation"]), int(cfg["micro_intonation"]), int(cfg["intonation_drift"]), int(cfg["smooth_intonation"]))
    out["names"] = np.array(out["names"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "events_v1.npz"), **out)


if __name__ == "__main__":
    main()
