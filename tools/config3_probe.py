#!/usr/bin/env python
"""Times a slice of BASELINE config 3 (ragged, one randomised voice per utterance), device-resident (development aid).
python tools/config3_probe.py [--utts 8192]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--utts", type=int, default=8192); a = ap.parse_args()
import torch
import gama_tts_b200 as g
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import random_voice
U = a.utts
lengths = T.config3_lengths()[:U]
voices = [random_voice(np.random.Generator(np.random.PCG64(7 + u))) for u in range(U)]
uniq = [T.synthetic_track(7 + i, 5000) for i in range(64)]
fo = np.zeros(U + 1, np.int64); fo[1:] = np.cumsum(lengths)
frames = np.empty((int(fo[-1]), 16), np.float32)
for u in range(U): frames[fo[u]:fo[u + 1]] = uniq[u % 64][:lengths[u]]
synth = g.TubeSynthesizer(0)
b = synth.prepare(voices, fo, voice_index=np.arange(U, dtype=np.int32))
d_frames = torch.from_numpy(frames).cuda(); d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(3):
    e0.record(s); b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream); e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ni, ns = int(b.n_internal.sum()), b.n_samples_total
    print("launch %d: %.1f ms, %.0f audio-s/s, %.2f TFLOP/s" % (r, ms, ns / 48000.0 / (ms * 1e-3), (384.0 * ni + 106.0 * ns) / (ms * 1e-3) * 1e-12))
