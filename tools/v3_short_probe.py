#!/usr/bin/env python
"""Many short utterances (the regime the planner gives to tube_kernel_v3): pipelined kernel vs thread-per-utterance.
python tools/v3_short_probe.py [--utts 37888] [--frames 250]"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--utts", type=int, default=37888); ap.add_argument("--frames", type=int, default=250)
a = ap.parse_args()
import torch
import gama_tts_b200 as g
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import random_voice
U = a.utts
rng = np.random.Generator(np.random.PCG64(3))
lens = rng.integers(a.frames // 2, a.frames * 3 // 2 + 1, U)
voices = [random_voice(np.random.Generator(np.random.PCG64(7 + u))) for u in range(U)]
uniq = [T.synthetic_track(7 + i, a.frames * 2) for i in range(64)]
fo = np.zeros(U + 1, np.int64); fo[1:] = np.cumsum(lens)
frames = np.empty((int(fo[-1]), 16), np.float32)
for u in range(U): frames[fo[u]:fo[u + 1]] = uniq[u % 64][:lens[u]]
synth = g.TubeSynthesizer(0)
d_frames = torch.from_numpy(frames).cuda()
res = {}
for kern in ("v2", "v3", "auto"):
    if kern == "auto": os.environ.pop("GTTS_KERNEL", None)
    else: os.environ["GTTS_KERNEL"] = kern
    b = synth.prepare(voices, fo, voice_index=np.arange(U, dtype=np.int32))
    d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(3):
        e0.record(s); b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream); e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    res[kern] = {"ms": ms, "audio_s_per_s": b.n_samples_total / 48000.0 / (ms * 1e-3), "checksum": float(d_out[::4099].double().abs().sum())}
    print(kern, res[kern], flush=True)
    b.close(); del d_out
json.dump({"utterances": U, "frames": a.frames, **res}, open(os.path.join(ROOT, "gpurun_out", "v3_short_probe.json"), "w"), indent=1)
