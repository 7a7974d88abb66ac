#!/usr/bin/env python
"""Small, fixed workload for ncu: `python tools/profile_run.py [--utts U] [--frames F] [--reps R]`.
Runs the batch R times on device-resident buffers and prints the CUDA-event time per launch."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--lib", default=None, help="time another build of the library (tools/ab_build.sh)")
    ap.add_argument("--model5", action="store_true", help="model-5 utterances (voice 5_male) on tube5_kernel")
    ap.add_argument("--mixed", action="store_true", help="config-3-like: every utterance its own randomised voice, ragged lengths (0.5 .. 1.5 x --frames)")
    args = ap.parse_args()
    import torch
    if args.lib:
        from gama_tts_b200 import capi
        capi.LIB_PATH = os.path.abspath(args.lib)
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice
    base = [T.synthetic_track(20240 + u, args.frames) for u in range(min(args.utts, 64))]
    frames = np.concatenate([base[u % len(base)] for u in range(args.utts)])
    fo = np.arange(args.utts + 1, dtype=np.int64) * args.frames
    synth = g.TubeSynthesizer(0)
    if args.mixed:
        from gama_tts_b200.voices import random_voice
        rng = np.random.Generator(np.random.PCG64(1))
        lens = rng.integers(args.frames // 2, args.frames * 3 // 2 + 1, args.utts)
        long_base = [T.synthetic_track(20240 + u, args.frames * 2) for u in range(64)]
        frames = np.concatenate([long_base[u % 64][:lens[u]] for u in range(args.utts)])
        fo = np.zeros(args.utts + 1, np.int64)
        fo[1:] = np.cumsum(lens)
        voices_l = [random_voice(np.random.Generator(np.random.PCG64(7 + u))) for u in range(args.utts)]
        b = synth.prepare(voices_l, fo, voice_index=np.arange(args.utts, dtype=np.int32))
        steps = float(b.n_internal.sum()) / float(lens.sum())
        args.frames = float(lens.mean())
    elif args.model5:
        from gama_tts_b200.voices import default_voice5
        b = synth.prepare5(default_voice5("male"), fo)
        b.n_samples_total = int(b.n_out.sum())
        steps = int(b.n_internal[0] // args.frames)
    else:
        b = synth.prepare(default_voice("male"), fo)
        steps = 80
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.zeros(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(args.reps):
        e0.record(s)
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("launch %d: %.3f ms  (%.1f ns per internal sample per utterance, %.0f audio-s/s)" %
              (r, ms, ms * 1e6 / (args.frames * steps), b.n_samples_total / 48000.0 / (ms * 1e-3)))
    print("checksum", float(d_out[::997].double().abs().sum()))


if __name__ == "__main__":
    main()
