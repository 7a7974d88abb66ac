#!/usr/bin/env python
"""Side measurements for BASELINE configs 3 and 5 (bench.py stays the contract benchmark on config 2).

  config 3-like: U utterances of log-uniform length (250..5000 frames), every utterance its own randomised
                 voice (tract length, glottal pulse, nasal radii, breathiness); tracks cycle over 256 unique
                 synthetic tracks to keep host-side generation short.  Device-resident timing.
  config 5:      one utterance streamed through gtts_stream_* control frame by control frame (and in chunks).

    python tools/bench_configs.py [--utts 8192] [--stream-frames 3000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=8192)
    ap.add_argument("--stream-frames", type=int, default=3000)
    args = ap.parse_args()
    import torch
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice, random_voice

    synth = g.TubeSynthesizer(0)
    out = {}

    # ---- config 3-like ---------------------------------------------------------------------------------
    U = args.utts
    lengths = T.config3_lengths(U, seed=7)
    rng = np.random.Generator(np.random.PCG64(7))
    voices = [random_voice(rng) for _ in range(U)]
    uniq = [T.synthetic_track(7 + i, 5000) for i in range(256)]
    fo = np.zeros(U + 1, np.int64)
    fo[1:] = np.cumsum(lengths)
    frames = np.empty((int(fo[-1]), 16), np.float32)
    for u in range(U):
        frames[fo[u]:fo[u + 1]] = uniq[u % 256][:lengths[u]]
    t0 = time.perf_counter()
    b = synth.prepare(voices, fo, voice_index=np.arange(U, dtype=np.int32))
    prep = time.perf_counter() - t0
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    e0.record(s)
    b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    audio = b.n_samples_total / 48000.0
    out["config3_like"] = {"utterances": U, "frames_total": int(fo[-1]), "audio_seconds": audio, "ms": ms,
                           "audio_s_per_s": audio / (ms * 1e-3), "prepare_s": prep,
                           "internal_samples": int(b.n_internal.sum()),
                           "finite": bool(torch.isfinite(d_out[::1009]).all().item())}
    b.close()
    del d_frames, d_out

    # ---- config 5: streaming, one utterance --------------------------------------------------------------
    v = default_voice("male")
    track = T.synthetic_track(99, args.stream_frames)
    for chunk in (1, 25, 250):
        st = synth.stream(v)
        n = 0
        t0 = time.perf_counter()
        for i in range(0, len(track), chunk):
            n += len(st.push(track[i:i + chunk]))
        n += len(st.finish())
        dt = time.perf_counter() - t0
        st.close()
        out["config5_stream_chunk_%d" % chunk] = {"frames": len(track), "audio_seconds": n / 48000.0, "wall_s": dt,
                                                  "audio_s_per_s": n / 48000.0 / dt,
                                                  "us_per_push": dt / max(1, len(track) // chunk) * 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
