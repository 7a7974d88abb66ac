#!/usr/bin/env python
"""Does a device->host copy on another stream slow the synthesis kernel down?  (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gama_tts_b200 as g
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import default_voice

U, F = 1024, 2500
base = [T.synthetic_track(20240 + u, F) for u in range(64)]
frames = np.concatenate([base[u % 64] for u in range(U)])
fo = np.arange(U + 1, dtype=np.int64) * F
synth = g.TubeSynthesizer(0)
b = synth.prepare(default_voice("male"), fo)
d_frames = torch.from_numpy(frames).cuda()
d_out = torch.zeros(b.n_out_total, dtype=torch.float32, device="cuda")
d_pcm = torch.zeros(b.n_out_total, dtype=torch.int16, device="cuda")
h_pcm = torch.empty(b.n_out_total, dtype=torch.int16).pin_memory()
h_in = torch.empty(b.n_out_total // 2, dtype=torch.int16).pin_memory()
d_in = torch.empty(b.n_out_total // 2, dtype=torch.int16, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def kernel_ms(copy=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if copy == "d2h":
        with torch.cuda.stream(s2):
            for _ in range(2): h_pcm.copy_(d_pcm, non_blocking=True)
    if copy == "h2d":
        with torch.cuda.stream(s2):
            for _ in range(4): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s1):
        e0.record(s1)
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s1.cuda_stream)
        e1.record(s1)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(2): kernel_ms()
print("kernel alone           %.2f ms" % kernel_ms())
print("kernel + D2H copies    %.2f ms" % kernel_ms("d2h"))
print("kernel + H2D copies    %.2f ms" % kernel_ms("h2d"))
print("kernel alone           %.2f ms" % kernel_ms())

def copy_ms(with_kernel):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if with_kernel:
        with torch.cuda.stream(s1):
            b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s1.cuda_stream)
            b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s1.cuda_stream)
    with torch.cuda.stream(s2):
        e0.record(s2)
        h_pcm.copy_(d_pcm, non_blocking=True)
        e1.record(s2)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print("D2H copy alone         %.2f ms (%.1f GB/s)" % (copy_ms(False), h_pcm.numel() * 2 / copy_ms(False) * 1e-6))
print("D2H copy + kernel      %.2f ms" % copy_ms(True))
print("D2H copy + kernel      %.2f ms" % copy_ms(True))
