#!/usr/bin/env python
"""Where the end-to-end time of BASELINE config 2 goes: python tools/e2e_probe.py [--utts 1024] [--frames 2500]
Times, on one GPU: the kernel with device output, the kernel storing straight into pinned host memory,
the H2D copy of the frames, a D2H copy of the whole output, and gtts_batch_run_host (zero-copy and staged)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=2500)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--lib", default=None, help="time another build of the library (tools/ab_build.sh)")
    args = ap.parse_args()
    import torch
    if args.lib:
        from gama_tts_b200 import capi
        capi.LIB_PATH = os.path.abspath(args.lib)
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice
    base = [T.synthetic_track(20240 + u, args.frames) for u in range(min(args.utts, 32))]
    frames = np.concatenate([base[u % len(base)] for u in range(args.utts)])
    fo = np.arange(args.utts + 1, dtype=np.int64) * args.frames
    synth = g.TubeSynthesizer(0)
    b = synth.prepare(default_voice("male"), fo)
    h_frames = torch.from_numpy(frames).pin_memory()
    h_out = torch.empty(b.n_out_total, dtype=torch.float32).pin_memory()
    d_frames = h_frames.cuda()
    d_out = torch.zeros(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, label):
        best = 1e9
        for _ in range(args.reps):
            torch.cuda.synchronize()
            e0.record(s)
            fn()
            e1.record(s)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("%-44s %8.2f ms" % (label, best))
        return best

    def wall(fn, label):
        best = 1e9
        for _ in range(args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        print("%-44s %8.2f ms" % (label, best))
        return best

    gb = b.n_out_total * 4 / 1e9
    print("output %.3f GB, frames %.3f GB" % (gb, frames.nbytes / 1e9))
    timed(lambda: b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream), "kernel, device frames -> device out")
    timed(lambda: b.run_device(d_frames.data_ptr(), h_out.data_ptr(), s.cuda_stream), "kernel, device frames -> pinned host out")
    timed(lambda: b.run_device(h_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream), "kernel, pinned host frames -> device out")
    timed(lambda: b.run_device(h_frames.data_ptr(), h_out.data_ptr(), s.cuda_stream), "kernel, pinned host frames -> pinned host out")
    t = timed(lambda: d_frames.copy_(h_frames, non_blocking=True), "H2D frames")
    t = timed(lambda: h_out.copy_(d_out, non_blocking=True), "D2H output")
    print("   D2H rate %.1f GB/s" % (gb / (t * 1e-3)))
    wall(lambda: b.run_host_ptr(h_frames.data_ptr(), h_out.data_ptr()), "gtts_batch_run_host (default)")
    os.environ["GTTS_HOST_OUTPUT"] = "staged"
    wall(lambda: b.run_host_ptr(h_frames.data_ptr(), h_out.data_ptr()), "gtts_batch_run_host (staged)")


if __name__ == "__main__":
    main()
