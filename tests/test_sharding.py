"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo backend.  Each rank takes its shard
of a ragged batch from the same deterministic plan; together the shards cover every utterance exactly
once, the loads are balanced, and the max-over-ranks reduction used for timing works."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gama_tts_b200 as g
    from gama_tts_b200 import sharding
    from gama_tts_b200.tracks import config3_lengths
    from gama_tts_b200.voices import default_voice, random_voice
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.Generator(np.random.PCG64(11))
    voices = [default_voice("male")] + [random_voice(rng) for _ in range(3)]
    n_frames = config3_lengths(512, seed=7)
    vidx = np.arange(512) % 4
    cost = sharding.utterance_cost(voices, vidx, n_frames)
    shards = sharding.shard_utterances(cost, world)
    mine = shards[rank]
    # what this rank would synthesise: output sizes through the C ABI's host-side closed form
    n_out = sum(g.output_length(voices[vidx[u]], int(n_frames[u]))[1] for u in mine[:40])
    t_max = sharding.max_over_ranks(1.0 + rank, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.tolist(), int(cost[mine].sum()), n_out))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put((gathered, t_max, int(cost.sum())))


def test_two_rank_sharding_gloo(product_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, t_max, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_idx = sorted(gathered[0][0] + gathered[1][0])
    assert all_idx == list(range(512))                       # every utterance exactly once
    loads = [gathered[0][1], gathered[1][1]]
    assert sum(loads) == total and max(loads) / (total / 2) < 1.01
    assert t_max == 2.0                                      # max over ranks of (1 + rank)
    assert gathered[0][2] > 0 and gathered[1][2] > 0


# ---- control-frame generation over two ranks: utterances (chains of chunks) stay whole -------------------------------

def _events_batch():
    from gama_tts_b200.events import event_config, synthetic_events
    rng = np.random.Generator(np.random.PCG64(21))
    cfgs, lists, cont = [], [], []
    for u in range(60):
        for k in range(1 + u % 3):
            cfgs.append(event_config())
            lists.append(synthetic_events(7000 + 10 * u + k, int(rng.integers(1, 12))))
            cont.append(int(k > 0))
    return cfgs, lists, cont


def _events_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import zlib
    import torch.distributed as dist
    from gama_tts_b200 import sharding
    from oracle.pyoracle import OracleEvents
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfgs, lists, cont = _events_batch()
    mine = sharding.shard_event_chunks(cfgs, lists, cont, world)[rank]
    # what this rank's GPU would produce, by the checker: chunk by chunk, the drift state carried along a chain
    o = OracleEvents()
    sums, carried = {}, None
    for c in mine:
        cfg = cfgs[c].copy()
        if cont[c]:
            assert carried is not None, "a chain was split between ranks"
            for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
                cfg[k] = carried[k]
        frames, carried = o.generate(cfg, lists[c])
        sums[int(c)] = (len(frames), zlib.crc32(frames.tobytes()))
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(gathered)


def test_two_rank_event_sharding_gloo(product_lib):
    import zlib
    import torch.multiprocessing as mp
    from oracle.pyoracle import OracleEvents
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_events_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfgs, lists, cont = _events_batch()
    assert sorted(list(gathered[0]) + list(gathered[1])) == list(range(len(lists)))      # every chunk exactly once
    loads = [sum(v[0] for v in g.values()) for g in gathered]
    assert max(loads) / (sum(loads) / 2) < 1.05
    # the same frames as one process makes of the whole batch
    o = OracleEvents()
    carried = None
    merged = {**gathered[0], **gathered[1]}
    for c, (cfg, ev) in enumerate(zip(cfgs, lists)):
        cfg = cfg.copy()
        if cont[c]:
            for k in ("drift_seed", "drift_x1", "drift_x2", "drift_y1", "drift_y2"):
                cfg[k] = carried[k]
        frames, carried = o.generate(cfg, ev)
        assert merged[c] == (len(frames), zlib.crc32(frames.tobytes())), c
