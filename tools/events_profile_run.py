#!/usr/bin/env python
"""The workload of the events_kernel ncu captures: N chunks of P postures each, a few launches, CUDA-event time."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gama_tts_b200 as g  # noqa: E402
from gama_tts_b200.events import event_config, synthetic_events  # noqa: E402

if __name__ == "__main__":
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=37888)
    ap.add_argument("--postures", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lib", default=None, help="time another build of the library (tools/ab_build.sh)")
    a = ap.parse_args()
    if a.lib:
        from gama_tts_b200 import capi
        capi.LIB_PATH = os.path.abspath(a.lib)
    synth = g.TubeSynthesizer(0)
    base = [synthetic_events(1000 + k, a.postures) for k in range(128)]
    events, eo = g.pack_events([base[u % 128] for u in range(a.chunks)])
    eb = synth.prepare_events(np.array([event_config()] * a.chunks), events, eo)
    d_events = torch.from_numpy(events.view(np.uint8)).cuda()
    d_frames = torch.empty(eb.n_frames_total * 16, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(2):
        eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.reps):
        eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print("chunks %d events %d frames %d: %.4f ms per launch, %.3g frames/s, %.1f GB/s algorithmic" % (
        a.chunks, eb.n_events_total, eb.n_frames_total, ms, eb.n_frames_total / ms * 1e3,
        (events.nbytes + eb.n_frames_total * 64) / ms * 1e-6))
