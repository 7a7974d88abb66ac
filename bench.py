#!/usr/bin/env python
"""Benchmark of the hot path: batched tube-model synthesis (BASELINE.json metric
"audio-seconds synthesized per wall-second").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 2 -- 1,024 synthetic control tracks of 2,500 frames (10 s)
per GPU, voice 0_male/male (fs_int 20,034 Hz, 80 internal samples per 4 ms control frame), i.e.
204.8 M internal samples -> 490.8 M float32 output samples (10,224 audio-seconds) per step per GPU.
One step = one pass of the whole batch through the tube path.  N > 1 (torchrun, one rank per GPU):
every rank synthesises its own 1,024 utterances (weak scaling, no collective on the data path; NCCL
is used only for the barrier and the max-over-ranks of the timings).

value  = audio-seconds / wall-second with the control tracks already resident in HBM (CUDA events on
         the launching stream around K launches, max over ranks).
e2e    = the same through the C ABI with pinned HOST buffers, every step's input transfer (the kernel
         reads the frames in place over PCIe) and output transfer inside the timed region.  The output
         is what the reference's pipeline ends in: the peak-normalised 16-bit PCM payload
         (gtts_batch_submit_host_pcm16 / gtts_batch_wait on two alternating batches, so that one
         step's payload travels while the next step is synthesised).  e2e.variants also gives the
         float32 output (gtts_batch_run_host, the round-1 definition) and the unpipelined PCM call;
         e2e.ceiling is the plain device->host copy rate measured on the same box (all ranks at once)
         and what it allows for this payload.
roofline: FP64 FMA pipe. achieved = algorithmic flops per launch (384 per internal sample + 106 per
         output sample) / launch duration; peak = the DFMA peak measured on this
         GPU by gtts_probe_fp64_peak just before the timed region.
cpu_baseline: the unmodified reference compiled in place (oracle/_ref, kind "reference"; the oracle
         port if that build is absent) on the host cores, bounded sample of the same tracks.
--impl reference: the reference arm -- the reference's own CPU implementation, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_UTT = 1024
N_FRAMES = 2500
SEED0 = 20240
FLOP_PER_INTERNAL = 384.0     # SURVEY.md section 8d / BASELINE.md section 4
FLOP_PER_OUTPUT = 106.0
FP64_PEAK_FALLBACK_TFLOPS = 34.1   # measured on this pool's B200 (profiles/fp64_pipe_r01.json)
METRIC = "audio_seconds_per_wall_second"
UNIT = "audio-s/s"


def make_tracks(rank, n_utt, n_frames):
    from gama_tts_b200 import tracks as T
    cache = os.path.join(tempfile.gettempdir(), "gtts_bench_tracks_r%d_%d_%d.npy" % (rank, n_utt, n_frames))
    if os.path.exists(cache):
        try:
            a = np.load(cache)
            if a.shape == (n_utt * n_frames, 16):
                return a
        except Exception:
            pass
    a = np.concatenate([T.synthetic_track(SEED0 + rank * n_utt + u, n_frames) for u in range(n_utt)])
    try:
        np.save(cache, a)
    except Exception:
        pass
    return a


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel on this workload, from the committed ncu capture (or None)."""
    path = os.path.join(ROOT, "profiles", "ncu_r02_traffic_bench.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "ncu_r01_traffic_bench.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = os.path.join(tempfile.gettempdir(), "gtts_clocks_%d_%d.csv" % (os.getpid(), gpu_index))
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, smax = [], set(), None
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            busy = [x for x in sm if x > 0.5 * max(sm)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(frames, n_frames, n_utt_sample, threads):
    """Times the reference's CPU path on the first n_utt_sample tracks. Returns (audio_s_per_s, kind, seconds)."""
    from gama_tts_b200.voices import default_voice
    from oracle import pyoracle
    voice = default_voice("male")
    fo = np.arange(n_utt_sample + 1, dtype=np.int64) * n_frames
    sub = frames[:n_utt_sample * n_frames]
    try:
        ref = pyoracle.Reference()
        sec, n_each, _ = ref.batch(voice, sub, fo, n_threads=threads)
        kind = "reference"
        audio = float(n_each.sum()) / voice["output_rate"]
    except (FileNotFoundError, OSError, subprocess.CalledProcessError):
        # the in-place build of the reference is absent: time the oracle port, single thread
        orc = pyoracle.Oracle()
        t0 = time.perf_counter()
        n = 0
        for u in range(n_utt_sample):
            n += len(orc.synthesize(voice, sub[u * n_frames:(u + 1) * n_frames]))
        sec = time.perf_counter() - t0
        kind = "port"
        threads = 1
        audio = n / voice["output_rate"]
    return audio / sec, kind, sec, threads


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    n_sample = min(N_UTT, threads * 16)
    frames = make_tracks(0, N_UTT, N_FRAMES)
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, sec, used = cpu_reference_run(frames, N_FRAMES, n_sample, threads)
        if i >= args.warmup:
            vals.append((v, sec))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([s for _, s in vals])) * 1e3
    sample = "%d of the %d tracks (10 s each) per step, %d threads" % (n_sample, N_UTT, used)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "BASELINE config 2: 1024 synthetic control tracks x 2500 frames (10 s) per GPU, voice 0_male/male, "
                        "fs_int 20034 Hz, output 48 kHz float32",
            "utterances_per_gpu": N_UTT, "frames_per_utterance": N_FRAMES, "control_rate_hz": 250,
            "cache": "inputs (164 MB) + outputs (1.96 GB) per step exceed the 126 MB L2; no explicit flush"}


def config34_leg(synth, rank, world, dist, barrier, max_over_ranks, peak, args):
    """BASELINE config 3 (N = 1) / config 4 (N > 1): a FIXED batch -- the first --config3-utts utterances of the
    65,536-utterance draw (log-uniform 250..5,000 frames, every utterance its own randomised voice) -- sharded by
    utterance over the ranks with gtts_shard_plan (strong scaling: the batch does not grow with N).  Device-resident
    timing, 3 warm-ups.  With N > 1, rank 0 also runs the whole batch on its one GPU and the per-utterance checksums
    of the shards are compared with it (config 4: bitwise equal, no collective on the data path)."""
    import torch
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.sharding import shard_utterances, utterance_cost
    from gama_tts_b200.voices import random_voice
    U = args.config3_utts
    lengths = T.config3_lengths()[:U]
    voices = [random_voice(np.random.Generator(np.random.PCG64(7 + u))) for u in range(U)]
    uniq = [T.synthetic_track(7 + i, 5000) for i in range(128)]      # tracks cycle over 128 unique ones (host time)

    def build(ids):
        fo = np.zeros(len(ids) + 1, np.int64)
        fo[1:] = np.cumsum(lengths[ids])
        frames = np.empty((int(fo[-1]), 16), np.float32)
        for k, u in enumerate(ids):
            frames[fo[k]:fo[k + 1]] = uniq[u % 128][:lengths[u]]
        b = synth.prepare([voices[u] for u in ids], fo, voice_index=np.arange(len(ids), dtype=np.int32))
        return b, torch.from_numpy(frames).cuda()

    def checksums(b, d_out):
        sums = torch.zeros(max(b.n_utt, 1), dtype=torch.int64, device="cuda")
        g.check(g.load().gtts_batch_checksum_device(b._h, d_out.data_ptr(), sums.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return sums[:b.n_utt]

    cost = utterance_cost(voices, np.arange(U), lengths)
    shards = shard_utterances(cost, world)
    mine = shards[rank]
    b, d_frames = build(mine)
    d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(s)
    for _ in range(reps):
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    e1.record(s)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
    finite = bool(torch.isfinite(d_out[::1009]).all().item())
    mine_sums = checksums(b, d_out)
    n_internal = int(b.n_internal.sum())
    n_samples = int(b.n_samples_total)
    tot = torch.tensor([n_internal, n_samples], dtype=torch.int64, device="cuda")
    bitwise = None
    if dist is not None:
        dist.all_reduce(tot)
        allsums = torch.zeros(U, dtype=torch.int64, device="cuda")
        allsums[torch.from_numpy(mine).cuda()] = mine_sums
        dist.all_reduce(allsums)                   # every entry is contributed by exactly one rank
        del d_out, d_frames
        b.close()
        if rank == 0:
            bw, dfw = build(np.arange(U))
            d_all = torch.empty(bw.n_out_total, dtype=torch.float32, device="cuda")
            bw.run_device(dfw.data_ptr(), d_all.data_ptr(), s.cuda_stream)
            bitwise = bool(torch.equal(checksums(bw, d_all), allsums))
            bw.close()
    n_internal, n_samples = int(tot[0]), int(tot[1])
    audio = n_samples / 48000.0
    flops = FLOP_PER_INTERNAL * n_internal + FLOP_PER_OUTPUT * n_samples
    ach = flops / (ms * 1e-3) * 1e-12
    return {"workload": "first %d utterances of the BASELINE config 3 draw (log-uniform 250..5000 frames, one randomised voice "
                        "each), sharded by utterance over %d GPU(s) with gtts_shard_plan" % (U, world),
            "utterances": U, "n_gpus": world, "scaling": "strong", "audio_seconds": audio, "ms": ms,
            "value": audio / (ms * 1e-3), "unit": UNIT, "warmup": 3, "reps": reps, "finite": finite,
            "roofline_frac": ach / (peak * world), "achieved_tflops": ach,
            "config4_bitwise_equal": bitwise}


def model5_leg(synth, rank, peak):
    """BASELINE next row 1, measured the same way on rank 0: a batch of model-5 utterances (voice 5_male, 60,411 Hz
    internal rate, 48 kHz output), 12 per SM x 2 s each, device-resident, 3 warm-ups.  Work per unit: 397 flop
    per internal sample + 168 per output sample (33-tap down-sampling converter), specials not counted."""
    import torch
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice5
    if rank != 0:
        return None
    n_utt, n_frames = 1776, 500
    tracks = [T.synthetic_track(SEED0 + 100000 + (u % 96), n_frames) for u in range(n_utt)]
    frames, fo = g.pack_tracks(tracks)
    b = synth.prepare5(default_voice5("male"), fo)
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(s)
    for _ in range(reps):
        b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_internal, n_samples = int(b.n_internal.sum()), int(b.n_out.sum())
    audio = n_samples / 48000.0
    ach = (397.0 * n_internal + 168.0 * n_samples) / (ms * 1e-3) * 1e-12
    finite = bool(torch.isfinite(d_out[::1009]).all().item())
    b.close()
    # the reference's own model 5 on the host cores, a bounded sample of the same tracks
    cpu = None
    try:
        from oracle import pyoracle
        threads = host_threads()
        n_sample = min(n_utt, threads * 16)
        fo_s = np.arange(n_sample + 1, dtype=np.int64) * n_frames
        sec, n_each, _ = pyoracle.Reference().batch(default_voice5("male"), frames[:n_sample * n_frames], fo_s, n_threads=threads)
        cpu = {"value": float(n_each.sum()) / 48000.0 / sec, "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": "first %d of the %d tracks, %.2f s wall" % (n_sample, n_utt, sec)}
    except Exception as e:        # the in-place build of the reference is absent
        cpu = {"unavailable": str(e)[:120]}
    return {"cpu_baseline": cpu, "workload": "%d model-5 utterances (voice 5_male, fs_int 60411.43 Hz) x %d frames (2 s), one GPU" % (n_utt, n_frames),
            "kernel": "tube5_kernel", "utterances": n_utt, "audio_seconds": audio, "ms": ms, "value": audio / (ms * 1e-3),
            "unit": UNIT, "warmup": 3, "reps": reps, "finite": finite, "roofline_frac": ach / peak, "achieved_tflops": ach}


def models34_leg(synth, rank):
    """Models 3 and 4 (gtts_voice_config::tube_model, general kernel: one warp per utterance) on rank 0: 1,184 utterances
    x 1 s of the male voice at 60,102 Hz internal rate, device-resident, 2 warm-ups."""
    import torch
    import gama_tts_b200 as g
    from gama_tts_b200 import tracks as T
    from gama_tts_b200.voices import default_voice
    if rank != 0:
        return None
    out = {}
    n_utt, n_frames = 1184, 250
    tracks = [T.synthetic_track(SEED0 + 200000 + (u % 64), n_frames) for u in range(n_utt)]
    frames, fo = g.pack_tracks(tracks)
    d_frames = torch.from_numpy(frames).cuda()
    s = torch.cuda.current_stream()
    for tm in (3, 4):
        b = synth.prepare(dict(default_voice("male"), tube_model=tm), fo)
        d_out = torch.empty(b.n_out_total, dtype=torch.float32, device="cuda")
        for _ in range(2):
            b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(2):
            b.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        audio = b.n_samples_total / 48000.0
        out["model%d" % tm] = {"workload": "%d utterances x %d frames (1 s), voice 0_male/male as model %d" % (n_utt, n_frames, tm),
                               "kernel": "tube_kernel_v0", "ms": ms, "value": audio / (ms * 1e-3), "unit": UNIT,
                               "finite": bool(torch.isfinite(d_out[::1009]).all().item())}
        b.close()
        del d_out
    return out


def events_leg(synth, rank):
    """BASELINE next row 3 on rank 0: control-frame generation on the device (events_kernel: EventList::generateOutput).
    (a) the event lists of the headline workload's shape (1,024 utterances of about 10 s) alone and chained in front of the
    synthesis on one stream, (b) a batch that fills the GPU (37,888 chunks of about 2 s), whose algorithmic bytes (296 per
    event read, 64 per frame written) over its time are set against the measured HBM copy rate.  CPU: the oracle port of
    generateOutput, one thread, on a sample."""
    import torch
    import gama_tts_b200 as g
    from gama_tts_b200.events import event_config, synthetic_events
    from gama_tts_b200.voices import default_voice
    if rank != 0:
        return None
    s = torch.cuda.current_stream()

    def timed(fn, warm=3, reps=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def batch(n_chunks, n_postures, distinct):
        base = [synthetic_events(SEED0 + 300000 + k, n_postures) for k in range(distinct)]
        lists = [base[u % distinct] for u in range(n_chunks)]
        events, eo = g.pack_events(lists)
        cfgs = np.array([event_config()] * n_chunks)
        eb = synth.prepare_events(cfgs, events, eo)
        return eb, events, cfgs, lists

    out = {}
    # (a) the headline shape, chained in front of the synthesis
    eb, events, cfgs, lists = batch(N_UTT, 74, 96)
    d_events = torch.from_numpy(events.view(np.uint8)).cuda()
    d_frames = torch.empty(eb.n_frames_total * 16, dtype=torch.float32, device="cuda")
    tb = synth.prepare(default_voice("male"), eb.frame_offsets)
    d_out = torch.empty(tb.n_out_total, dtype=torch.float32, device="cuda")
    ms_ev = timed(lambda: eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream))
    ms_tube = timed(lambda: tb.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream), warm=2, reps=3)

    def chained():
        eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream)
        tb.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    ms_both = timed(chained, warm=2, reps=3)
    audio = tb.n_samples_total / 48000.0
    out["headline_shape"] = {
        "workload": "%d utterances, %d events -> %d frames (%.1f s each on average), voice 0_male/male" % (
            N_UTT, eb.n_events_total, eb.n_frames_total, eb.n_frames_total * 0.004 / N_UTT),
        "events_kernel_ms": ms_ev, "synthesis_ms": ms_tube, "events_then_synthesis_ms": ms_both,
        "value": audio / (ms_both * 1e-3), "unit": UNIT, "finite": bool(torch.isfinite(d_out[::1009]).all().item()),
        "h2d_bytes_events": int(events.nbytes), "h2d_bytes_frames_avoided": int(eb.n_frames_total * 64)}
    # CPU: the oracle port on a sample of the same lists
    try:
        from oracle.pyoracle import OracleEvents
        o = OracleEvents()
        t0 = time.perf_counter()
        n = 0
        for k in range(96):
            n += len(o.generate(cfgs[k], lists[k])[0])
        sec = (time.perf_counter() - t0) / 2      # generate() runs the list twice (count, then fill)
        out["cpu_baseline"] = {"value": n / sec, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "96 lists, %.3f s" % sec}
    except Exception as e:
        out["cpu_baseline"] = {"unavailable": str(e)[:120]}
    tb.close()
    eb.close()
    del d_out, d_frames, d_events
    # (c) BASELINE config 2 itself, from event lists: 1,024 utterances of exactly 2,500 frames, so the synthesis is the
    # headline's (one voice, one length).  End to end from HOST event lists to the HOST 16-bit payload, synchronous: pinned
    # events -> device, drift + frame passes, synthesis, output stage, payload -> pinned host; against the same utterances
    # entering as host FRAMES (gtts_batch_run_host_pcm16: the frames cross PCIe instead of the events).
    base = [synthetic_events(SEED0 + 310000 + k, 74, duration_ms=10000) for k in range(96)]
    lists = [base[u % 96] for u in range(N_UTT)]
    events, eo = g.pack_events(lists)
    eb = synth.prepare_events(np.array([event_config()] * N_UTT), events, eo)
    assert eb.n_frames_total == N_UTT * N_FRAMES
    tb = synth.prepare(default_voice("male"), eb.frame_offsets)
    h_events = torch.from_numpy(events.view(np.uint8)).pin_memory()
    d_events = torch.empty_like(h_events, device="cuda")
    d_frames = torch.empty(eb.n_frames_total * 16, dtype=torch.float32, device="cuda")
    d_audio = torch.empty(tb.n_out_total, dtype=torch.float32, device="cuda")
    d_pcm = torch.empty(tb.n_out_total, dtype=torch.int16, device="cuda")
    h_pcm = torch.empty(tb.n_out_total, dtype=torch.int16).pin_memory()
    h_frames = torch.empty(eb.n_frames_total * 16, dtype=torch.float32).pin_memory()

    def from_events():
        d_events.copy_(h_events, non_blocking=True)
        eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream)
        tb.run_device_pcm16(d_frames.data_ptr(), d_audio.data_ptr(), d_pcm.data_ptr(), 0, s.cuda_stream)
        h_pcm.copy_(d_pcm, non_blocking=True)
        torch.cuda.synchronize()

    def wall(fn, warm=2, reps=5):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3
    def payload_sum():      # over the utterances' own samples (the up to 63 samples between two utterances are never written)
        return sum(int(h_pcm[int(tb.out_offsets[u]):int(tb.out_offsets[u] + tb.n_out[u])].to(torch.int64).sum().item())
                   for u in range(0, N_UTT, 37))
    ms_events_e2e = wall(from_events)
    sum_events = payload_sum()
    h_frames.copy_(d_frames)
    torch.cuda.synchronize()
    ms_frames_e2e = wall(lambda: tb.run_host_pcm16_ptr(h_frames.data_ptr(), h_pcm.data_ptr(), 0))
    sum_frames = payload_sum()
    ms_dev = timed(lambda: (eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream),
                            tb.run_device(d_frames.data_ptr(), d_audio.data_ptr(), s.cuda_stream)), warm=2, reps=5)
    ms_synth = timed(lambda: tb.run_device(d_frames.data_ptr(), d_audio.data_ptr(), s.cuda_stream), warm=1, reps=5)
    audio = tb.n_samples_total / 48000.0
    out["config2_from_events"] = {
        "workload": "BASELINE config 2 from event lists: %d utterances x %d frames (10 s), %d events, voice 0_male/male" % (
            N_UTT, N_FRAMES, eb.n_events_total),
        "kernel_used": tb.last_kernel(), "device_resident_ms": ms_dev, "synthesis_alone_ms": ms_synth,
        "value": audio / (ms_dev * 1e-3), "unit": UNIT,
        "e2e_from_host_events": {"ms": ms_events_e2e, "value": audio / (ms_events_e2e * 1e-3), "h2d_bytes": int(events.nbytes),
                                 "d2h_bytes": int(tb.n_samples_total * 2)},
        "e2e_from_host_frames": {"ms": ms_frames_e2e, "value": audio / (ms_frames_e2e * 1e-3), "h2d_bytes": int(eb.n_frames_total * 64),
                                 "d2h_bytes": int(tb.n_samples_total * 2)},
        "same_payload": sum_events == sum_frames}
    tb.close()
    eb.close()
    del d_audio, d_pcm, d_frames, d_events, h_pcm, h_frames, h_events
    # (b) a batch that fills the GPU
    eb, events, cfgs, lists = batch(37888, 16, 128)
    d_events = torch.from_numpy(events.view(np.uint8)).cuda()
    d_frames = torch.empty(eb.n_frames_total * 16, dtype=torch.float32, device="cuda")
    ms = timed(lambda: eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream))
    nbytes = events.nbytes + eb.n_frames_total * 64
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm, src = 6550.0, "fallback of B200_PROFILING.md"
    gbs = nbytes / (ms * 1e-3) * 1e-9
    out["full_gpu"] = {"workload": "%d chunks, %d events -> %d frames" % (37888, eb.n_events_total, eb.n_frames_total),
                       "ms": ms, "value": eb.n_frames_total / (ms * 1e-3), "unit": "frames/s",
                       "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                    "peak_source": src, "bytes_per_launch_algorithmic": int(nbytes)}}
    eb.close()
    return out


def small_configs_leg(synth, rank):
    """BASELINE configs 1 and 5 on rank 0, so that the driver's run carries them: config 1 = ONE real sentence
    ("Hello world.", 332 control frames captured from the reference's front end, tests/golden/real_tracks.npz) through
    synthesize() -- planning, copies, kernel --, best of 5; config 5 = one utterance streamed control frame by control
    frame through gtts_stream_* (5,000 single-frame pushes, host wall clock), and in pushes of 250 frames."""
    if rank != 0:
        return None
    from gama_tts_b200.voices import default_voice
    out = {}
    v = default_voice("male")
    try:
        hello = np.load(os.path.join(ROOT, "tests", "golden", "real_tracks.npz"))["track0"]
        best = None
        for _ in range(6):
            t0 = time.perf_counter()
            audio = synth.synthesize(v, [hello])[0]
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        out["config1"] = {"workload": "one sentence, %d control frames, voice 0_male/male, one utterance through synthesize()" % len(hello),
                          "audio_seconds": len(audio) / 48000.0, "ms": best * 1e3, "value": len(audio) / 48000.0 / best, "unit": UNIT}
    except Exception as e:
        out["config1"] = {"unavailable": str(e)[:120]}
    from gama_tts_b200 import tracks as T
    track = T.tile_track(T.synthetic_track(99, 3000), 10000)
    res = {}
    for chunk, n in ((1, 5000), (250, 10000)):
        st = synth.stream(v)
        st.push(track[:chunk])                                  # first launch (graph instantiation) outside the clock
        t0 = time.perf_counter()
        produced = 0
        for i in range(chunk, n, chunk):
            produced += len(st.push(track[i:i + chunk]))
        dt = time.perf_counter() - t0
        st.finish()
        st.close()
        pushes = (n - chunk) // chunk
        res["frames_per_push_%d" % chunk] = {"pushes": pushes, "us_per_push": dt / pushes * 1e6,
                                             "value": produced / 48000.0 / dt, "unit": UNIT}
    out["config5"] = dict(res, workload="one utterance streamed through gtts_stream_* (control rate 250 Hz: a frame is 4 ms of audio)")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the config 3 / 4 leg")
    ap.add_argument("--config3-utts", type=int, default=16384)
    ap.add_argument("--no-model5", action="store_true", help="skip the model-5 leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import gama_tts_b200 as g
    from gama_tts_b200.voices import default_voice

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    voice = default_voice("male")
    frames_np = make_tracks(rank, N_UTT, N_FRAMES)
    fo = np.arange(N_UTT + 1, dtype=np.int64) * N_FRAMES
    synth = g.TubeSynthesizer(local_rank)
    batch = synth.prepare(voice, fo)
    n_out = batch.n_out_total                 # buffer size: every utterance starts on a 64-sample row pair
    n_samples = batch.n_samples_total         # audio samples produced
    n_internal = int(batch.n_internal.sum())
    audio_seconds = n_samples / voice["output_rate"]
    flops_per_launch = FLOP_PER_INTERNAL * n_internal + FLOP_PER_OUTPUT * n_samples

    h_frames = torch.from_numpy(frames_np).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    d_frames = h_frames.cuda(non_blocking=True)
    d_out = torch.empty(n_out, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak = synth.fp64_peak_tflops()
    peak_src = "measured by gtts_probe_fp64_peak on this GPU (register-resident DFMA, burst)"
    if not peak or peak <= 0:
        peak, peak_src = FP64_PEAK_FALLBACK_TFLOPS, "fallback: profiles/fp64_pipe_r01.json"

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        batch.run_device(d_frames.data_ptr(), d_out.data_ptr(), sptr)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(args.steps):
        batch.run_device(d_frames.data_ptr(), d_out.data_ptr(), sptr)
        launches += batch.last_launches()
    e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # ---- end to end: pinned host buffers through the C ABI ----------------------------------------------
    def timed_host_loop(step_fn, drain_fn=None):
        for i in range(min(args.warmup, 2)):
            step_fn(i)
        if drain_fn:
            drain_fn()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_fn(i)
        if drain_fn:
            drain_fn()
        torch.cuda.synchronize()
        ms_local = (time.perf_counter() - t0) * 1e3 / args.steps
        barrier()
        return max_over_ranks(ms_local)

    # (a) float32 audio, synchronous call (round-1 definition)
    ms_f32 = timed_host_loop(lambda i: batch.run_host_ptr(h_frames.data_ptr(), h_out.data_ptr()))
    checksum = float(h_out[::4097].double().abs().sum())
    del h_out
    # (b) 16-bit PCM payload, synchronous call
    h_pcm = [torch.empty(n_out, dtype=torch.int16).pin_memory() for _ in range(2)]
    ms_pcm_sync = timed_host_loop(lambda i: batch.run_host_pcm16_ptr(h_frames.data_ptr(), h_pcm[0].data_ptr()))
    # (c) 16-bit PCM payload, two batches in flight
    batch2 = synth.prepare(voice, fo)
    pair = [batch, batch2]

    def pipelined_step(i):
        b = pair[i & 1]
        b.wait()                                   # its previous step (and the host buffer it wrote) is done
        b.submit_host_pcm16_ptr(h_frames.data_ptr(), h_pcm[i & 1].data_ptr())

    ms_pcm_pipe = timed_host_loop(pipelined_step, lambda: (pair[0].wait(), pair[1].wait()))
    pcm_checksum = int(h_pcm[0][::4097].to(torch.int64).abs().sum())
    ms_e2e = ms_pcm_pipe
    # (d) the ceiling: plain device->host copies of the same payload size, all ranks at once
    d_pcm_probe = torch.empty(n_out, dtype=torch.int16, device="cuda")
    for _ in range(2):
        h_pcm[1].copy_(d_pcm_probe, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_pcm[1].copy_(d_pcm_probe, non_blocking=True)
    torch.cuda.synchronize()
    ms_copy_local = (time.perf_counter() - t0) * 1e3 / 3
    barrier()
    ms_copy = max_over_ranks(ms_copy_local)
    d2h_gbs = n_out * 2 / (ms_copy * 1e-3) * 1e-9
    del d_pcm_probe
    clocks = sampler.stop() if sampler is not None else None

    # ---- BASELINE configs 3 / 4: a fixed slice of the 65,536-utterance draw, sharded over the ranks ---------
    cfg34 = None if args.no_config3 else config34_leg(synth, rank, world, dist, barrier, max_over_ranks, peak, args)
    m5 = None if args.no_model5 else model5_leg(synth, rank, peak)
    small = None if args.no_model5 else small_configs_leg(synth, rank)
    m34 = None if args.no_model5 else models34_leg(synth, rank)
    evl = None if args.no_model5 else events_leg(synth, rank)

    value = audio_seconds * world / (ms_dev * 1e-3)
    e2e_value = audio_seconds * world / (ms_e2e * 1e-3)
    achieved_tflops = flops_per_launch / (ms_dev * 1e-3) * 1e-12

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(frames_np.nbytes), "d2h_bytes_per_step": int(n_samples * 2),
                    "output": "peak-normalised 16-bit PCM payload (the reference's WAVE data), two batches in flight",
                    "variants": {
                        "float32_sync": {"value": audio_seconds * world / (ms_f32 * 1e-3), "ms_per_step": ms_f32,
                                         "d2h_bytes_per_step": int(n_samples * 4)},
                        "pcm16_sync": {"value": audio_seconds * world / (ms_pcm_sync * 1e-3), "ms_per_step": ms_pcm_sync,
                                       "d2h_bytes_per_step": int(n_samples * 2)},
                        "pcm16_pipelined": {"value": e2e_value, "ms_per_step": ms_e2e,
                                            "d2h_bytes_per_step": int(n_samples * 2)}},
                    "ceiling": {"d2h_gbs_per_gpu": d2h_gbs, "d2h_gbs_aggregate": d2h_gbs * world,
                                "how": "torch copy_ of the payload size, device -> pinned host, all ranks at once, max over ranks",
                                "value_at_ceiling": audio_seconds * world / (ms_copy * 1e-3),
                                "frac_of_ceiling": ms_copy / ms_e2e},
                    "pcm_checksum": pcm_checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64_fma", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak, "traffic": measured_traffic(), "peak_source": peak_src,
                         "flops_per_launch": flops_per_launch,
                         "hbm_bytes_per_launch_algorithmic": int(frames_np.nbytes + n_samples * 4)},
            "clocks": clocks,
            "kernel": json.loads(synth.describe()),
            "kernel_used": batch.last_kernel(),
            "checksum": checksum,
        }
        if cfg34 is not None:
            line["config3"] = cfg34
        if m5 is not None:
            line["model5"] = m5
        if small is not None:
            line.update(small)
        if m34 is not None:
            line.update(m34)
        if evl is not None:
            line["control_frames"] = evl
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            n_sample = min(N_UTT, threads * 16)
            v, kind, sec, used = cpu_reference_run(frames_np, N_FRAMES, n_sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": kind,
                                    "sample": "first %d of the %d tracks (10 s each), %.2f s wall" % (n_sample, N_UTT, sec)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
