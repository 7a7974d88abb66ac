"""The parity gate: the CUDA path (through the C ABI) against the oracle and the committed golden
vectors of the unmodified reference.  Tolerance (north_star): max abs sample error <= 1e-6 of full
scale, full scale being the reference's own 0.95-peak normalisation (VTMUtil.cpp:48-67)."""
import numpy as np
import pytest

from conftest import full_scale_error, snr_db
import gama_tts_b200 as g
from gama_tts_b200 import tracks as T
from gama_tts_b200.voices import default_voice, random_voice

pytestmark = pytest.mark.gpu

TOL = 1e-6          # of full scale (north_star)
TIGHT = 2e-7        # what we actually expect: the reference's own FMA on/off noise floor is 7e-8


def test_golden_vectors(synth, golden):
    for name in golden.names:
        voice, track, ref, ref_nofma = golden.case(name)
        out = synth.synthesize(voice, [track])[0]
        assert len(out) == len(ref), name
        err = full_scale_error(out, ref)
        assert err <= TOL, (name, err)
        assert err <= TIGHT, (name, err)
        if len(ref) > 100 and np.abs(ref).max() > 0:
            assert snr_db(out, ref) >= 100.0, name


def test_vs_oracle_fresh_tracks_and_voices(synth, oracle):
    rng = np.random.Generator(np.random.PCG64(2024))
    voices = [default_voice(v) for v in ("male", "female", "large_child", "small_child", "baby")]
    voices += [random_voice(rng) for _ in range(11)]
    tracks = [T.synthetic_track(100 + i, int(rng.integers(20, 160))) for i in range(len(voices))]
    outs = synth.synthesize(voices, tracks, voice_index=np.arange(len(voices)))
    for v, tr, out in zip(voices, tracks, outs):
        ref = oracle.synthesize(v, tr)
        assert len(out) == len(ref)
        assert full_scale_error(out, ref) <= TIGHT


def test_ragged_batch_equals_singles(synth, real_tracks):
    # batch == N x single, bitwise, whatever the batch composition and ordering
    v = default_voice("male")
    hello, fox = real_tracks[0], real_tracks[1]
    tracks = [hello, fox[:1], fox[100:163], hello[:0], fox[400:640], hello[50:51], fox[:333]] * 3
    batch = synth.synthesize(v, tracks)
    for tr, out in zip(tracks, batch):
        single = synth.synthesize(v, [tr])[0]
        assert np.array_equal(out, single)
    n_out = [g.output_length(v, len(t))[1] for t in tracks]
    assert [len(o) for o in batch] == n_out


def test_empty_and_tiny_inputs(synth, oracle):
    v = default_voice("male")
    assert synth.synthesize(v, []) == []
    out = synth.synthesize(v, [np.zeros((0, 16), np.float32)])[0]
    assert len(out) == 63 and not out.any()
    one = T.synthetic_track(1, 1)
    assert full_scale_error(synth.synthesize(v, [one])[0], oracle.synthesize(v, one)) <= TIGHT


def test_edge_parameter_values(synth, oracle):
    # silent glottis, zero radii (clamped to 0.01), velum 0 and 1.5, frication at both ends of the tract,
    # tiny negative volumes as the real front end produces them
    v = default_voice("male")
    tr = np.tile(T.VOWEL_AA, (60, 1)).astype(np.float32)
    tr[10:20, 1] = -6.7e-32
    tr[20:30, 7:15] = 0.0
    tr[30:40, 15] = 0.0
    tr[40:50, 15] = 1.5
    tr[5:25, 3] = 10.0
    tr[5:15, 4] = 0.0
    tr[15:25, 4] = 7.0
    tr[25:35, 2] = 10.0
    tr[45:55, 5] = 20000.0
    tr[45:55, 6] = 20000.0
    out = synth.synthesize(v, [tr])[0]
    ref = oracle.synthesize(v, tr)
    assert np.isfinite(out).all()
    assert full_scale_error(out, ref) <= TIGHT


def test_steps_override_per_sample_mode(synth, oracle, real_tracks):
    # steps = 1: the plugin shim's mode, parameters used verbatim for one internal sample each
    v = default_voice("male")
    params = np.repeat(real_tracks[0][100:130], 11, axis=0)
    out = synth.synthesize(v, [params], steps_override=[1])[0]
    ref = oracle.synthesize_samples(v, params)
    assert len(out) == len(ref) and full_scale_error(out, ref) <= TIGHT


def test_control_rates(synth, oracle):
    # control_period 1..4 ms (VTMControlModelConfiguration.cpp:37-41)
    v = default_voice("female")
    tr = T.synthetic_track(9, 50)
    for period in (1, 2, 3, 4):
        out = synth.synthesize(v, [tr], control_rate=1000.0 / period)[0]
        ref = oracle.synthesize(v, tr, control_rate=1000.0 / period)
        assert len(out) == len(ref) and full_scale_error(out, ref) <= TIGHT


def test_streaming_equals_batch(synth, real_tracks):
    v = default_voice("male")
    track = real_tracks[3][:200]
    whole = synth.synthesize(v, [track])[0]
    st = synth.stream(v)
    parts = []
    rng = np.random.Generator(np.random.PCG64(4))
    i = 0
    while i < len(track):
        n = int(rng.integers(1, 9))
        parts.append(st.push(track[i:i + n]))
        i += n
    parts.append(st.finish())
    streamed = np.concatenate(parts)
    assert len(streamed) == len(whole)
    assert np.array_equal(streamed, whole)
    # frame-by-frame (BASELINE config 5 pattern) after a reset
    st.reset()
    parts = [st.push(track[i:i + 1]) for i in range(60)] + [st.finish()]
    assert np.array_equal(np.concatenate(parts), synth.synthesize(v, [track[:60]])[0])
    st.close()


def test_rerun_is_deterministic(synth):
    v = default_voice("male")
    tracks = [T.synthetic_track(40 + i, 64) for i in range(20)]
    a = synth.synthesize(v, tracks)
    b = synth.synthesize(v, tracks)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_device_buffer_entry_point(synth):
    import torch
    v = default_voice("male")
    tracks = [T.synthetic_track(70 + i, 40) for i in range(9)]
    frames, fo = g.pack_tracks(tracks)
    b = synth.prepare(v, fo)
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.zeros(b.n_out_total, dtype=torch.float32, device="cuda")
    b.run_device(d_frames.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = b.run_host(frames)
    # utterances start on 64-sample row pairs of the buffer; the gaps between them are never written
    assert np.all(b.out_offsets % 64 == 0)
    for x, y in zip(b.split(d_out.cpu().numpy()), b.split(host)):
        assert np.array_equal(x, y)
    assert b.last_launches() == 1
    b.close()


def _oracle_many(oracle, jobs, workers=None):
    """oracle.synthesize over (voice, track) jobs on the host's cores (ctypes releases the GIL; every call owns
    its model instance)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or max(1, min(32, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(lambda j: oracle.synthesize(j[0], j[1]), jobs))


@pytest.mark.slow
def test_config2_full_batch_parity(synth, oracle):
    # BASELINE config 2 as benchmarked: 1,024 tracks x 2,500 frames (10 s), voice male, ONE batch.  48 of the 1,024
    # outputs are compared sample by sample with the oracle at the tight bound (every 32nd utterance plus the first
    # and last 8: all slot positions and both ends of the queue), the rest through size-independent properties.
    v = default_voice("male")
    tracks = T.config2_tracks()
    outs = synth.synthesize(v, tracks)
    assert len(outs) == 1024 and all(len(o) == 479250 for o in outs)
    assert all(np.isfinite(o).all() and np.abs(o).max() < 0.1 for o in outs)
    check = sorted(set(range(0, 1024, 32)) | set(range(8)) | set(range(1016, 1024)))
    refs = _oracle_many(oracle, [(v, tracks[i]) for i in check])
    worst = 0.0
    for i, ref in zip(check, refs):
        assert len(ref) == len(outs[i])
        worst = max(worst, full_scale_error(outs[i], ref))
        assert snr_db(outs[i], ref) >= 100.0
    assert worst <= TIGHT, worst
    # the uniform batch runs on tube_kernel_v1, one utterance alone on tube_kernel_v2 (test_pipelined_kernel_choice): the same
    # operations with different multiply-adds contracted, so equal to within the FMA noise floor; bitwise under one kernel
    assert b_last_kernel(synth, v, tracks[:2]) == "tube_kernel_v1"
    assert full_scale_error(outs[7], synth.synthesize(v, [tracks[7]])[0]) <= 1e-7


def b_last_kernel(synth, voice, tracks):
    frames, fo = g.pack_tracks(tracks)
    b = synth.prepare(voice, fo)
    b.run_host(frames)
    name = b.last_kernel()
    b.close()
    return name


def test_pcm16_output_stage_is_bit_exact(synth, oracle, golden, real_tracks):
    # gtts_batch_run_host_pcm16: per-utterance peak normalisation + 16-bit PCM on the device.  For the float32 audio the
    # GPU produced, the payload and the scale are the reference's own (oracle.pcm16 is pinned bit for bit against the
    # reference's WAVEFileWriter in tests/test_oracle.py); against the reference's audio the payload differs by at
    # most one LSB (the float32 audio itself agrees to 2e-7 of full scale).
    v = default_voice("male")
    tracks = [real_tracks[0], real_tracks[1][:1], real_tracks[1][100:163], real_tracks[0][:0], T.synthetic_track(5, 77),
              np.zeros((3, 16), np.float32)]
    audio = synth.synthesize(v, tracks)
    pcm, scale = synth.synthesize_pcm16(v, tracks)
    for a, p, s, tr in zip(audio, pcm, scale, tracks):
        want, want_scale = oracle.pcm16(a)
        assert len(p) == len(a)
        assert np.array_equal(p, want)
        assert s == np.float32(want_scale)
        ref_pcm, _ = oracle.pcm16(oracle.synthesize(v, tr))
        assert np.abs(p.astype(np.int32) - ref_pcm.astype(np.int32)).max(initial=0) <= 1
    assert scale[3] == 0.0 and not pcm[3].any()                 # empty track: 63 zeros, scale 0
    name = golden.names[1]
    voice, track, ref, _ = golden.case(name)
    p, s = synth.synthesize_pcm16(voice, [track])
    ref_pcm, _ = oracle.pcm16(ref)
    assert np.abs(p[0].astype(np.int32) - ref_pcm.astype(np.int32)).max() <= 1


def test_down_sampling_voices(synth, oracle, monkeypatch):
    # tract lengths below 7.31 cm: internal rate above 48 kHz, the SRC's down-sampling branch
    # (SampleRateConverter.h:362-415); mixed with up-sampling voices in one batch, both kernels
    voices = []
    for length in (7.2, 6.0, 4.5, 3.0):
        v = dict(default_voice("baby"))
        v["vocal_tract_length"] = length
        voices.append(v)
    voices.append(default_voice("male"))
    tracks = [T.synthetic_track(600 + i, 30 + 7 * i) for i in range(len(voices))]
    refs = [oracle.synthesize(v, t) for v, t in zip(voices, tracks)]
    for kernel in ("v2", "v0"):
        monkeypatch.setenv("GTTS_KERNEL", kernel)
        outs = synth.synthesize(voices, tracks, voice_index=np.arange(len(voices)))
        for out, ref in zip(outs, refs):
            assert len(out) == len(ref)
            assert full_scale_error(out, ref) <= TIGHT


def test_general_kernel_golden_vectors(synth, golden, monkeypatch):
    # GTTS_KERNEL=v0 forces the general warp-per-utterance kernel (the one streaming uses) on batches
    monkeypatch.setenv("GTTS_KERNEL", "v0")
    for name in golden.names:
        voice, track, ref, _ = golden.case(name)
        out = synth.synthesize(voice, [track])[0]
        assert len(out) == len(ref), name
        assert full_scale_error(out, ref) <= TIGHT, name


def test_both_kernels_agree(synth, monkeypatch):
    # the two kernels evaluate the same operations but contract different multiply-adds into FMAs, so
    # they agree to within the FMA noise floor (<= 1e-7 of full scale), not bitwise
    rng = np.random.Generator(np.random.PCG64(31))
    voices = [default_voice("male"), random_voice(rng), default_voice("baby")]
    tracks = [T.synthetic_track(300 + i, int(rng.integers(1, 120))) for i in range(23)]
    vidx = [i % 3 for i in range(23)]
    a = synth.synthesize(voices, tracks, voice_index=vidx)
    monkeypatch.setenv("GTTS_KERNEL", "v0")
    b = synth.synthesize(voices, tracks, voice_index=vidx)
    for x, y in zip(a, b):
        assert len(x) == len(y)
        assert full_scale_error(x, y) <= 1e-7


def test_wide_batch_kernel_parity(synth, oracle, monkeypatch):
    # GTTS_KERNEL=v3 selects the one-thread-per-utterance kernel (a measured alternative, never the default: DESIGN.md
    # section 4.3): 200 ragged utterances, every one its own randomised voice, against the oracle; the
    # down-sampling voice in the batch stays on the pipelined kernel
    monkeypatch.setenv("GTTS_KERNEL", "v3")
    rng = np.random.Generator(np.random.PCG64(77))
    voices = [random_voice(np.random.Generator(np.random.PCG64(700 + u))) for u in range(197)]
    voices += [default_voice("male"), default_voice("female"), dict(default_voice("baby"), vocal_tract_length=6.0)]
    voices[3]["waveform"] = 1
    tracks = [T.synthetic_track(800 + u, int(rng.integers(1, 300))) for u in range(200)]
    outs = synth.synthesize(voices, tracks, voice_index=np.arange(200))
    refs = _oracle_many(oracle, list(zip(voices, tracks)))
    for out, ref in zip(outs, refs):
        assert len(out) == len(ref)
        assert full_scale_error(out, ref) <= TIGHT


def test_pinned_host_output_is_written_directly(synth):
    # pinned host buffers are device-accessible: the kernel stores the audio straight into them
    # (device->host transfer overlapped with the synthesis); result must equal the staged path
    import torch
    v = default_voice("male")
    tracks = [T.synthetic_track(900 + i, 37 + i) for i in range(12)]
    frames, fo = g.pack_tracks(tracks)
    b = synth.prepare(v, fo)
    staged = b.run_host(frames)                              # pageable numpy buffers -> staged copy
    h_frames = torch.from_numpy(frames).pin_memory()
    h_out = torch.zeros(b.n_out_total, dtype=torch.float32).pin_memory()
    b.run_host_ptr(h_frames.data_ptr(), h_out.data_ptr())
    for x, y in zip(b.split(h_out.numpy()), b.split(staged)):
        assert len(x) > 0 and np.array_equal(x, y)
    # the padding between utterances is left untouched
    gaps = np.ones(b.n_out_total, bool)
    for u in range(b.n_utt):
        gaps[b.out_offsets[u]:b.out_offsets[u] + b.n_out[u]] = False
    assert not h_out.numpy()[gaps].any()
    b.close()


def test_device_exp2_exp10_match_libm(synth):
    # the kernels' branch-free 2^x / 10^x (tube_kernel.cuh) against libm over the ranges the path uses:
    # (pitch + 3) / 12 with pitch in [-30, 30] semitones, (dB - 60) / 20 with dB in (0, 100]
    rng = np.random.Generator(np.random.PCG64(11))
    x = np.concatenate([rng.uniform(-16.0, 16.0, 200000), np.linspace(-3.0, 2.0, 100001),
                        np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-300, -3.0, 15.999, -15.999])])
    e2, e10 = synth.probe_exp(x)
    r2, r10 = np.exp2(x), np.power(10.0, x)
    ulp2 = np.abs(e2 - r2) / np.spacing(r2)
    ulp10 = np.abs(e10 - r10) / np.spacing(r10)
    assert ulp2.max() <= 2.0, ulp2.max()
    assert ulp10.max() <= 2.0, ulp10.max()
    # integer arguments are exact cases of the reference's pow(): 10^-1 * tnDelta can be a tie of rint()
    # (the amplitudes use x = (dB - 60) / 20 in (-3, 0]: 10^-1 and 10^-2 must be libm's values bit for bit;
    # further out libm's own pow() is no longer correctly rounded -- 10^-5 -- and 1 ulp is all that is asked)
    ints = np.arange(-15.0, 16.0)
    e2, e10 = synth.probe_exp(ints)
    assert np.array_equal(e2, np.exp2(ints))
    near = np.abs(ints) <= 4
    assert np.array_equal(e10[near], np.power(10.0, ints[near]))
    assert (np.abs(e10 - np.power(10.0, ints)) <= np.spacing(np.power(10.0, ints))).all()


@pytest.mark.slow
def test_config3_real_shape_parity(synth, oracle):
    # BASELINE config 3 at its real shape: the 65,536-utterance draw (log-uniform 250..5,000 frames, every utterance
    # its own random voice and track, gama_tts_b200.tracks.config3_utterance).  A slice of 4,096 utterances that
    # contains the 16 longest of the whole draw is synthesised as one batch (several waves through the 7 x 148
    # slots, mixed voices in every CTA); 256 random utterances of the slice plus the 16 longest are compared with
    # the oracle, all of them checked for exact length and finiteness.
    lengths = T.config3_lengths()
    longest = np.argsort(lengths, kind="stable")[-16:]
    ids = np.array(sorted(set(longest.tolist()) | set(range(4096 - 16))))[:4096]
    assert len(ids) >= 4080 and set(longest.tolist()) <= set(ids.tolist())
    voices, tracks = zip(*[T.config3_utterance(u, lengths) for u in ids])
    outs = synth.synthesize(list(voices), list(tracks), voice_index=np.arange(len(ids)))
    for v, tr, out in zip(voices, tracks, outs):
        assert len(out) == g.output_length(v, len(tr))[1]
        assert np.isfinite(out).all()
    rng = np.random.Generator(np.random.PCG64(3))
    pos = {int(u): i for i, u in enumerate(ids)}
    check = sorted(set(rng.choice(len(ids), 256, replace=False).tolist()) | {pos[int(u)] for u in longest})
    refs = _oracle_many(oracle, [(voices[i], tracks[i]) for i in check])
    worst = max(full_scale_error(outs[i], ref) for i, ref in zip(check, refs))
    assert all(len(outs[i]) == len(ref) for i, ref in zip(check, refs))
    assert worst <= TIGHT, worst


@pytest.mark.slow
def test_config5_ten_minute_stream_parity(synth, oracle):
    # BASELINE config 5 at full length: ONE utterance of 150,000 control frames (10 min) pushed frame by frame
    # through gtts_stream_*; the whole 28.75 M-sample output against the oracle (phase drift of the oscillator over
    # 12 M internal samples is what a streaming implementation would get wrong).
    v = default_voice("male")
    track = T.tile_track(T.synthetic_track(99, 3000), 150000)
    st = synth.stream(v)
    parts = [st.push(track[i:i + 1]) for i in range(len(track))]
    parts.append(st.finish())
    st.close()
    out = np.concatenate(parts)
    ref = oracle.synthesize(v, track)
    assert len(out) == len(ref) == 28751278
    assert full_scale_error(out, ref) <= TIGHT
    # the tail on its own (an error that grows with time shows here first)
    assert full_scale_error(out[-480000:], ref[-480000:]) <= TIGHT
    assert snr_db(out, ref) >= 100.0


def test_slot_protocol_stress(synth, monkeypatch):
    # The pipelined kernel double-buffers its slot control blocks by iteration parity and refills slots from an
    # atomic queue: stress it with 3,000 short utterances of 1..3 blocks up to a few dozen (random lengths incl.
    # single-frame and empty ones, three voices with different control steps), repeated, and demand bitwise
    # equality with the general kernel's arithmetic-free properties: batch == singles for a sample of utterances,
    # run-to-run determinism, and agreement with the general kernel within the FMA noise floor.
    rng = np.random.Generator(np.random.PCG64(2026))
    voices = [default_voice("male"), default_voice("baby"), default_voice("female")]
    n = 3000
    vidx = rng.integers(0, 3, n)
    base = [T.synthetic_track(7000 + i, 40) for i in range(32)]
    lens = rng.choice([0, 1, 1, 1, 2, 2, 3, 4, 7, 12, 25, 40], n)
    tracks = [base[i % 32][: lens[i]] for i in range(n)]
    a = synth.synthesize(voices, tracks, voice_index=vidx)
    b = synth.synthesize(voices, tracks, voice_index=vidx)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for i in rng.choice(n, 40, replace=False):
        single = synth.synthesize(voices[vidx[i]], [tracks[i]])[0]
        assert np.array_equal(a[i], single), i
    monkeypatch.setenv("GTTS_KERNEL", "v0")
    c = synth.synthesize(voices, tracks, voice_index=vidx)
    for x, y in zip(a, c):
        assert len(x) == len(y)
        assert full_scale_error(x, y) <= 1e-7 or np.abs(y).max() < 1e-12


def test_mixed_control_periods_split_between_kernels(synth, oracle):
    # control periods shorter than one block (steps 20 at 1 kHz control rate for the male voice) go to the general
    # kernel, the rest of the batch stays on the pipelined one: two launches, every utterance right
    v = default_voice("male")
    tracks = [T.synthetic_track(800 + i, 30) for i in range(6)]
    steps = [0, 20, 0, 1, 20, 0]           # 0: derived from the control rate (80)
    frames, fo = g.pack_tracks(tracks)
    b = synth.prepare(v, fo, steps_override=steps)
    outs = b.split(b.run_host(frames))
    assert b.last_launches() == 2
    b.close()
    for tr, st, out in zip(tracks, steps, outs):
        if st == 0:
            ref = oracle.synthesize(v, tr)
        elif st == 1:
            ref = oracle.synthesize_samples(v, tr)
        else:
            ref = oracle.synthesize(v, tr, control_rate=v_rate(v) / st)
        assert len(out) == len(ref) and full_scale_error(out, ref) <= TIGHT


def v_rate(v):
    return float(g.internal_rate(v))


def test_loud_glottis_rise_segment_corruption(synth, oracle, monkeypatch):
    # glottal volume above 60 dB with tn_min != tn_max: the reference's wavetable rewrite zeroes part of the rise
    # segment for good (WavetableGlottalSource.h:162-184); both kernels reproduce it
    v = dict(default_voice("male"))
    v["glottal_pulse_tn_min"], v["glottal_pulse_tn_max"] = 16.0, 32.0
    tr = T.synthetic_track(11, 120).copy()
    tr[:, 1] = np.concatenate([np.linspace(50, 78, 60), np.linspace(78, 40, 60)]).astype(np.float32)
    ref = oracle.synthesize(v, tr)
    assert full_scale_error(synth.synthesize(v, [tr])[0], ref) <= TIGHT
    monkeypatch.setenv("GTTS_KERNEL", "v0")
    assert full_scale_error(synth.synthesize(v, [tr])[0], ref) <= TIGHT


def test_multi_gpu_dispatch_through_the_c_abi(synth):
    # gtts_multi_*: ONE call partitions the batch by utterance over the GPUs of the box (one host thread per GPU, no
    # collective), every utterance's audio lands at its place in ONE caller buffer, bit for bit what one GPU gives
    # (BASELINE config 4).  With one GPU the dispatcher is exercised with that GPU listed twice (two shards).
    import torch
    n_gpu = torch.cuda.device_count()
    devices = list(range(n_gpu)) if n_gpu >= 2 else [0, 0]
    rng = np.random.Generator(np.random.PCG64(45))
    n = 300
    voices = [default_voice("male"), default_voice("female"), random_voice(rng)]
    vidx = rng.integers(0, 3, n)
    tracks = [T.synthetic_track(4500 + i, int(rng.integers(0, 150))) for i in range(n)]
    whole = synth.synthesize(voices, tracks, voice_index=vidx)
    multi = g.MultiSynthesizer(devices)
    outs, shard_of = multi.synthesize(voices, tracks, voice_index=vidx, return_shards=True)
    assert set(shard_of.tolist()) == set(range(len(devices)))
    for a, b in zip(whole, outs):
        assert np.array_equal(a, b)
    (pcm, scale), _ = multi.synthesize(voices, tracks, voice_index=vidx, pcm16=True, return_shards=True)
    want_pcm, want_scale = synth.synthesize_pcm16(voices, tracks, voice_index=vidx)
    for a, b in zip(want_pcm, pcm):
        assert np.array_equal(a, b)
    assert np.array_equal(scale, want_scale)
    multi.close()


def test_sharded_over_gpus_equals_one_gpu(synth):
    # BASELINE config 4: the batch partitioned by utterance over the GPUs of the box (no collective on the data
    # path) gives, utterance by utterance, the bits of the one-GPU run.  Needs >= 2 GPUs.
    import torch
    from gama_tts_b200.sharding import shard_utterances, utterance_cost
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("one GPU on this box")
    rng = np.random.Generator(np.random.PCG64(44))
    n = 600
    voices = [default_voice("male"), default_voice("female"), random_voice(rng)]
    vidx = rng.integers(0, 3, n)
    tracks = [T.synthetic_track(4400 + i, int(rng.integers(10, 200))) for i in range(n)]
    whole = synth.synthesize(voices, tracks, voice_index=vidx)
    cost = utterance_cost(voices, vidx, [len(t) for t in tracks])
    shards = shard_utterances(cost, n_gpu)
    assert sorted(np.concatenate(shards).tolist()) == list(range(n))
    for r, idx in enumerate(shards):
        dev = g.TubeSynthesizer(r)
        part = dev.synthesize(voices, [tracks[i] for i in idx], voice_index=vidx[idx])
        for i, out in zip(idx, part):
            assert np.array_equal(out, whole[i]), (r, i)
        dev.close()


# ---- model 5 (gtts5_*, tube5_kernel.cuh) ------------------------------------------------------------------------------

def test_model5_golden_vectors(synth, golden5):
    # the reference's own model-5 outputs (tests/golden/golden5_v1.npz): all five shipped variants, randomised voices,
    # constant-radius mouth impedance, sine source, bypass, modulation off, empty and one-frame tracks
    for name in golden5.names:
        voice, track, ref, _ = golden5.case(name)
        out = synth.synthesize5(voice, [track])[0]
        assert len(out) == len(ref), name
        assert full_scale_error(out, ref) <= TIGHT, name


def test_model5_ragged_batch_vs_oracle(synth, oracle5):
    # 60 ragged utterances, every one its own randomised voice (internal rates 60-141 kHz), one batch: more warps
    # than one CTA holds, the queue is used; control rates 250 and 500 Hz
    from gama_tts_b200.voices import default_voice5, random_voice5
    rng = np.random.Generator(np.random.PCG64(91))
    voices = [random_voice5(np.random.Generator(np.random.PCG64(900 + u)), base=["male", "female", "baby"][u % 3]) for u in range(58)]
    voices += [default_voice5("small_child"), default_voice5("large_child")]
    tracks = [T.synthetic_track(950 + u, int(rng.integers(1, 200))) for u in range(60)]
    for rate in (250.0, 500.0):
        outs = synth.synthesize5(voices, tracks, voice_index=np.arange(60), control_rate=rate)
        for v, tr, out in zip(voices, tracks, outs):
            ref = oracle5.synthesize(v, tr, control_rate=rate)
            assert len(out) == len(ref)
            assert full_scale_error(out, ref) <= TIGHT


def test_models_3_and_4_multi_gpu_dispatch(synth):
    # models 3 / 4 are voices of the model-0 ABI (tube_model), so gtts_multi_* takes them as they are: bitwise equal to one GPU
    import torch
    rng = np.random.Generator(np.random.PCG64(23))
    voices = [dict(default_voice("male"), tube_model=3), dict(default_voice("female"), tube_model=4), default_voice("male")]
    n = 18
    vidx = np.arange(n) % 3
    tracks = [T.synthetic_track(5300 + i, int(rng.integers(1, 50))) for i in range(n)]
    whole = synth.synthesize(voices, tracks, voice_index=vidx)
    multi = g.MultiSynthesizer(list(range(torch.cuda.device_count())))
    outs = multi.synthesize(voices, tracks, voice_index=vidx)
    multi.close()
    for u in range(n):
        assert np.array_equal(outs[u], whole[u]), u


def test_model5_multi_gpu_dispatch_through_the_c_abi(synth, oracle5):
    # gtts5_multi_*: a ragged model-5 batch of several voices over every GPU of the box (one on the driver's test box: the
    # packing / scattering path is the same), float32 and 16-bit payload, bit for bit the one-GPU result
    import torch
    from gama_tts_b200.voices import default_voice5, random_voice5
    rng = np.random.Generator(np.random.PCG64(17))
    voices = [default_voice5("male"), default_voice5("female"), random_voice5(rng, base="baby")]
    n = 40
    vidx = rng.integers(0, 3, n)
    tracks = [T.synthetic_track(5200 + i, int(rng.integers(0, 90))) for i in range(n)]
    whole = synth.synthesize5(voices, tracks, voice_index=vidx)
    multi = g.MultiSynthesizer(list(range(torch.cuda.device_count())))
    outs, shard_of = multi.synthesize5(voices, tracks, voice_index=vidx, return_shards=True)
    assert set(shard_of.tolist()) == set(range(min(torch.cuda.device_count(), n)))
    for u in range(n):
        assert np.array_equal(outs[u], whole[u]), u
    fo = np.zeros(n + 1, np.int64)
    fo[1:] = np.cumsum([len(t) for t in tracks])
    b = synth.prepare5(voices, fo, vidx)
    want_pcm, want_scale = b.run_host_pcm16(np.concatenate(tracks))
    want_pcm = b.split(want_pcm)
    b.close()
    (pcm, scale) = multi.synthesize5(voices, tracks, voice_index=vidx, pcm16=True)
    assert np.array_equal(scale, want_scale)
    for u in range(n):
        assert np.array_equal(pcm[u], want_pcm[u]), u
    multi.close()
    for u in (0, 13, 39):
        ref = oracle5.synthesize(voices[vidx[u]], tracks[u])
        assert len(ref) == len(whole[u]) and full_scale_error(whole[u], ref) <= TIGHT


def test_model5_rejects_what_the_reference_rejects(synth):
    from gama_tts_b200.voices import default_voice5
    tr = [T.synthetic_track(1, 5)]
    with pytest.raises(g.GttsError) as e:
        synth.synthesize5(dict(default_voice5("male"), vocal_tract_length=22.0), tr)     # internal rate below 50 kHz
    assert e.value.code == g.capi.GTTS_ERR_INVALID
    with pytest.raises(g.GttsError) as e:
        synth.synthesize5(dict(default_voice5("male"), vocal_tract_length=5.0), tr)      # converter wing above 48 taps
    assert e.value.code == g.capi.GTTS_ERR_UNSUPPORTED


def test_model5_pcm16_output_stage_is_bit_exact(synth, oracle, golden5):
    # the reference's output stage on model-5 audio: scale and 16-bit payload of the float32 audio the kernel produced,
    # against the oracle's restatement of the reference's writer (pinned by test_oracle.py) on the same audio
    names = [n for n in golden5.names if n not in ("empty",)]
    voices, tracks = zip(*[(golden5.case(n)[0], golden5.case(n)[1]) for n in names])
    frames, fo = g.pack_tracks(list(tracks))
    b = synth.prepare5(list(voices), fo, voice_index=np.arange(len(names)))
    audio = b.split(b.run_host(frames))
    pcm, scale = b.run_host_pcm16(frames)
    for u, (a, p) in enumerate(zip(audio, b.split(pcm))):
        ref_pcm, ref_scale = oracle.pcm16(a)
        assert scale[u] == np.float32(ref_scale)
        assert np.array_equal(p, ref_pcm), names[u]
    b.close()


# ---- models 2, 3 and 4 (gtts_voice_config::tube_model, general kernel) ------------------------------------------------

def test_models_3_and_4_vs_oracle(synth, oracle, real_tracks):
    # VocalTractModel2<double, 3> and VocalTractModel4<double, 1> (the oracle is bit-identical to the reference's models,
    # tests/test_oracle.py): the five shipped variants and randomised voices, ragged, mixed with model-0 utterances in
    # one batch (those go to the pipelined kernel); control rates 250 and 500 Hz
    rng = np.random.Generator(np.random.PCG64(17))
    base = [default_voice(n) for n in ("male", "female", "large_child", "small_child", "baby")] + [random_voice(rng) for _ in range(5)]
    voices = [dict(v, tube_model=m) for m in (3, 4, 0) for v in base]
    tracks = [real_tracks[i % 4][40 * i: 40 * i + int(rng.integers(1, 120))] for i in range(len(voices))]
    for rate in (250.0, 500.0):
        outs = synth.synthesize(voices, tracks, voice_index=np.arange(len(voices)), control_rate=rate)
        for v, tr, out in zip(voices, tracks, outs):
            ref = oracle.synthesize(v, tr, control_rate=rate)
            assert len(out) == len(ref)
            assert full_scale_error(out, ref) <= TIGHT, v["tube_model"]


def test_models_3_and_4_are_batch_only(synth):
    with pytest.raises(g.GttsError) as e:
        synth.stream(dict(default_voice("male"), tube_model=3))
    assert e.value.code == g.capi.GTTS_ERR_UNSUPPORTED


def test_pipelined_kernel_choice(synth, oracle, monkeypatch):
    # a batch of one voice and one length runs on tube_kernel_v1 (aligned slots: its CTA barrier is free and its code is
    # smaller), anything ragged or of several voices on tube_kernel_v2; both against the oracle, and equal to each other
    monkeypatch.delenv("GTTS_KERNEL", raising=False)
    v = default_voice("male")
    uniform = [T.synthetic_track(900 + i, 50) for i in range(12)]
    frames, fo = g.pack_tracks(uniform)
    b = synth.prepare(v, fo)
    a1 = [x.copy() for x in b.split(b.run_host(frames))]
    assert b.last_kernel() == "tube_kernel_v1"
    b.close()
    ragged = uniform[:11] + [T.synthetic_track(950, 37)]
    frames_r, fo_r = g.pack_tracks(ragged)
    b = synth.prepare(v, fo_r)
    a2 = [x.copy() for x in b.split(b.run_host(frames_r))]
    assert b.last_kernel() == "tube_kernel_v2"
    b.close()
    monkeypatch.setenv("GTTS_KERNEL", "v2")
    b = synth.prepare(v, fo)
    a3 = [x.copy() for x in b.split(b.run_host(frames))]
    assert b.last_kernel() == "tube_kernel_v2"
    b.close()
    for u in range(11):
        assert np.array_equal(a2[u], a3[u]), u                       # the same kernel, whatever the batch around the utterance
        assert full_scale_error(a1[u], a2[u]) <= 1e-7, u             # two kernels: different FMA contraction
    for u in (0, 5, 11):
        assert full_scale_error(a1[u], oracle.synthesize(v, uniform[u])) <= TIGHT


def test_uniform_batches_on_v1_every_variant(synth, oracle, real_tracks, monkeypatch):
    # the kernel BASELINE config 2 runs on (tube_kernel_v1: batches of one voice and one length), on every shipped variant,
    # randomised voices (incl. tn_min != tn_max: the analytic fall segment) and control rates, every utterance against the oracle
    monkeypatch.delenv("GTTS_KERNEL", raising=False)
    rng = np.random.Generator(np.random.PCG64(77))
    voices = [default_voice(v) for v in ("male", "female", "large_child", "small_child", "baby")] + [random_voice(rng) for _ in range(3)]
    for k, v in enumerate(voices):
        rate = (250.0, 250.0, 500.0, 200.0)[k % 4]
        n_frames = 40 + 7 * k
        tracks = [T.synthetic_track(6000 + 16 * k + i, n_frames) for i in range(9)] + [real_tracks[k % len(real_tracks)][:n_frames]]
        frames, fo = g.pack_tracks(tracks)
        b = synth.prepare(v, fo, control_rate=rate)
        outs = [x.copy() for x in b.split(b.run_host(frames))]
        assert b.last_kernel() == "tube_kernel_v1", k
        b.close()
        for tr, out in zip(tracks, outs):
            ref = oracle.synthesize(v, tr, control_rate=rate)
            assert len(out) == len(ref), k
            assert full_scale_error(out, ref) <= TIGHT, k


# ---- control-frame generation on the device (gtts_events_*, events_kernel.cuh) -----------------------------------------

def _bits(a):
    # bit patterns; every NaN as one pattern (events at the time of the frame being made divide 0 by 0 in the reference:
    # x86 returns its negative default NaN, the GPU its canonical one -- a NaN either way)
    a = np.ascontiguousarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7fc00000), a.view(np.uint32))


def test_events_reference_fixtures_bit_exact(synth):
    # the frames the unmodified reference front end produced from its own event lists (tests/golden/events_v1.npz),
    # two-chunk utterances as chains whose second chunk takes the drift state the first one left
    import os
    from gama_tts_b200 import capi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "events_v1.npz"))
    names = [str(n) for n in z["names"]]
    cfgs = np.array([z["cfg_" + n] for n in names]).astype(capi.EVENT_CONFIG_DTYPE)
    cont = np.array([n.endswith("_1") for n in names], np.int32)
    for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
        cfgs[k][cont == 1] = 0.0
    frames = synth.control_frames(cfgs, [z["ev_" + n] for n in names], cont)
    for i, n in enumerate(names):
        assert frames[i].shape == z["frames_" + n].shape, n
        assert np.array_equal(_bits(frames[i]), _bits(z["frames_" + n])), n


def test_events_ragged_batch_vs_oracle(synth):
    # 3,000 chunks (more than the resident warps of the GPU: the queue is exercised), every flag combination, control
    # periods 1 / 4 / 10 ms, events closer than a period, lists of 0 / 1 / 2 events, chains of up to three chunks;
    # 300 of them and every chain checked bit for bit against the oracle, the drift state a chunk leaves included
    from gama_tts_b200.events import event_config, synthetic_events
    from oracle.pyoracle import OracleEvents
    o = OracleEvents()
    rng = np.random.Generator(np.random.PCG64(11))
    cfgs, lists, cont = [], [], []
    for i in range(3000):
        seed = int(rng.integers(0, 1 << 30))
        cfgs.append(event_config(control_period=(4, 4, 1, 10)[i % 4], macro=i & 1, micro=(i >> 1) & 1, drift=(i >> 2) & 1,
                                 smooth=(i >> 3) & 1))
        n_post = int(rng.integers(1, 40))
        ev = synthetic_events(seed, n_post, special_rate=0.03, tight=i % 5 == 0)
        if i % 97 == 0:
            ev = ev[:i // 97 % 3]
        lists.append(ev)
        cont.append(int(i % 7 in (5, 6) and i > 0))
    frames_list = synth.control_frames(np.array(cfgs), lists, cont)
    events, eo = g.pack_events(lists)
    b = synth.prepare_events(np.array(cfgs), events, eo, cont)
    frames, out = b.run_host(events)
    b.close()
    checked = 0
    carried = None
    for i in range(3000):
        in_chain = cont[i] or (i + 1 < 3000 and cont[i + 1])
        if not (in_chain or i % 10 == 0):
            carried = None
            continue
        c = cfgs[i].copy()
        if cont[i]:
            assert carried is not None
            for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
                c[k] = carried[k]
        want, carried = o.generate(c, lists[i])
        assert frames_list[i].shape == want.shape, i
        assert np.array_equal(_bits(frames_list[i]), _bits(want)), i
        assert np.array_equal(_bits(frames[b.frame_offsets[i]:b.frame_offsets[i + 1]]), _bits(want)), i
        for k in ("drift_seed", "drift_x1", "drift_y2"):
            assert out[k][i] == carried[k], (i, k)
        checked += 1
    assert checked >= 300


def test_events_to_audio_without_leaving_the_device(synth, oracle):
    # the chain the row exists for: event lists -> control frames (events_kernel) -> audio (tube kernel) on one stream, the
    # frames never on the host.  The audio equals, bit for bit, what the same kernel makes of the oracle's frames fed from
    # the host, and is within the tolerance of the oracle's audio.
    import torch
    from gama_tts_b200 import capi
    from gama_tts_b200.events import event_config, synthetic_events
    from oracle.pyoracle import OracleEvents
    o = OracleEvents()
    v = default_voice("male")
    cfgs, lists, cont = [], [], []
    for u in range(40):
        chunks = 1 + u % 3
        for k in range(chunks):
            cfgs.append(event_config())
            lists.append(synthetic_events(500 + 10 * u + k, 3 + (u + k) % 6))
            cont.append(int(k > 0))
    events, eo = g.pack_events(lists)
    eb = synth.prepare_events(np.array(cfgs), events, eo, cont)
    fo = eb.utterance_frame_offsets
    assert len(fo) == 41 and fo[-1] == eb.n_frames_total
    tb = synth.prepare(v, fo)
    s = torch.cuda.current_stream()
    d_events = torch.from_numpy(events.view(np.uint8)).cuda()
    d_frames = torch.zeros(eb.n_frames_total * 16, dtype=torch.float32, device="cuda")
    d_out = torch.zeros(tb.n_out_total, dtype=torch.float32, device="cuda")
    eb.run_device(d_events.data_ptr(), d_frames.data_ptr(), 0, s.cuda_stream)
    tb.run_device(d_frames.data_ptr(), d_out.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    audio = tb.split(d_out.cpu().numpy())
    # the oracle's frames, chained chunk by chunk
    want_frames, carried = [], None
    for i, (c, ev) in enumerate(zip(cfgs, lists)):
        c = c.copy()
        if cont[i]:
            for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
                c[k] = carried[k]
        f, carried = o.generate(c, ev)
        want_frames.append(f)
    want_frames = np.concatenate(want_frames)
    assert np.array_equal(_bits(d_frames.cpu().numpy().reshape(-1, 16)), _bits(want_frames))
    host = tb.split(tb.run_host(want_frames))
    for u in range(40):
        assert np.array_equal(audio[u], host[u]), u
    for u in (0, 7, 23, 39):
        ref = oracle.synthesize(v, want_frames[fo[u]:fo[u + 1]])
        assert len(ref) == len(audio[u]) and full_scale_error(audio[u], ref) <= TIGHT, u
    eb.close()
    tb.close()


@pytest.mark.slow
def test_events_full_gpu_batch_properties(synth):
    # 37,888 chunks (8 warps x 3 CTAs x 148 SMs = 3,552 resident warps: every warp takes ten or more chunks from the queue),
    # 128 distinct lists tiled: each distinct list is checked against the oracle bit for bit, and every copy of a list must
    # give the bits of its first copy whatever its place in the batch and whichever warp took it (size-independent property)
    from gama_tts_b200.events import event_config, synthetic_events
    from oracle.pyoracle import OracleEvents
    n, distinct = 37888, 128
    base = [synthetic_events(8000 + k, 4 + k % 9, special_rate=0.03) for k in range(distinct)]
    cfgs = [event_config(macro=k & 1, micro=(k >> 1) & 1, drift=(k >> 2) & 1, smooth=(k >> 3) & 1) for k in range(distinct)]
    lists = [base[u % distinct] for u in range(n)]
    events, eo = g.pack_events(lists)
    b = synth.prepare_events(np.array([cfgs[u % distinct] for u in range(n)]), events, eo)
    frames, out = b.run_host(events)
    fo = b.frame_offsets
    b.close()
    o = OracleEvents()
    for k in range(distinct):
        want, state = o.generate(cfgs[k], base[k])
        first = frames[fo[k]:fo[k + 1]]
        assert first.shape == want.shape and np.array_equal(_bits(first), _bits(want)), k
        assert out["drift_seed"][k] == state["drift_seed"] and out["drift_y1"][k] == state["drift_y1"], k
        for u in range(k + distinct, n, distinct):
            assert np.array_equal(frames[fo[u]:fo[u + 1]].view(np.uint32), first.view(np.uint32)), (k, u)
