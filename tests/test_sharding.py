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
