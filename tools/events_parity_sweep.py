#!/usr/bin/env python
"""Wide one-off parity sweep of the control-frame kernel on a GPU box: N random chunks (every flag combination, control
periods 1 / 2 / 4 / 5 / 10 ms, grid-aligned and arbitrary event times incl. equal ones, sparse and dense special
parameters, chains of 1-4 chunks, lists of 0-700 events), every chunk compared with the oracle bit for bit (NaN as NaN).
`python tools/events_parity_sweep.py [n_chunks] [seed]` prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gama_tts_b200 as g  # noqa: E402
from gama_tts_b200.events import event_config, synthetic_events  # noqa: E402
from oracle.pyoracle import OracleEvents  # noqa: E402

STATE = ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2")


def bits(a):
    a = np.ascontiguousarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7fc00000), a.view(np.uint32))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.Generator(np.random.PCG64(seed))
    synth, o = g.TubeSynthesizer(0), OracleEvents()
    done = frames_total = mismatched = nan_chunks = 0
    t0 = time.time()
    while done < n:
        m = min(20000, n - done)
        cfgs, lists, cont = [], [], []
        left = 0
        for i in range(m):
            f = int(rng.integers(0, 16))
            cfgs.append(event_config(control_period=int(rng.choice([1, 2, 4, 4, 4, 5, 10])), macro=f & 1, micro=(f >> 1) & 1,
                                     drift=(f >> 2) & 1, smooth=(f >> 3) & 1, initial_pitch=float(rng.normal(-20, 3)),
                                     mean_pitch=float(rng.normal(-16, 3))))
            size = int(rng.choice([0, 1, 2, 3, 8, 20, 40, 175])) if rng.random() < 0.5 else int(rng.integers(1, 30))
            ev = synthetic_events(int(rng.integers(0, 1 << 31)), size, special_rate=float(rng.choice([0.0, 0.005, 0.02, 0.2])),
                                  tight=bool(rng.random() < 0.25))
            if size == 0:
                ev = ev[:int(rng.integers(0, 3))]
            lists.append(ev)
            cont.append(1 if left > 0 and i > 0 else 0)
            left = left - 1 if left > 0 else int(rng.choice([0, 0, 0, 1, 2, 3]))
        got = synth.control_frames(np.array(cfgs), lists, cont)
        carried = None
        for i in range(m):
            c = cfgs[i].copy()
            if cont[i]:
                for k in STATE:
                    c[k] = carried[k]
            want, carried = o.generate(c, lists[i])
            frames_total += len(want)
            nan_chunks += bool(np.isnan(want).any())
            if got[i].shape != want.shape or not np.array_equal(bits(got[i]), bits(want)):
                mismatched += 1
        done += m
    print(json.dumps({"chunks": done, "frames": frames_total, "mismatched_chunks": mismatched, "chunks_with_nan": int(nan_chunks),
                      "seed": seed, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
