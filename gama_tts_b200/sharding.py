"""Multi-GPU partitioning of a batch: by utterance, no data-path collective.

Utterances are independent (private state, fixed RNG seed per instance), so a batch is split across the
GPUs of one box with the greedy longest-first plan of ``gtts_shard_plan`` and every rank runs its own
shard through its own ``TubeSynthesizer``; results stay on the rank that produced them (or are copied to
the host by that rank).  ``torch.distributed`` is used only for the start/stop barrier and for the
max-over-ranks of the timings.
"""
import numpy as np

from . import output_length, shard_plan
from .voices import DEFAULT_CONTROL_RATE


def utterance_cost(voices, voice_index, n_frames, control_rate=DEFAULT_CONTROL_RATE):
    """Work estimate per utterance: internal samples + output samples (both scale the kernel time)."""
    cache = {}
    cost = np.zeros(len(n_frames), np.int64)
    for u, f in enumerate(n_frames):
        vi = 0 if voice_index is None else int(voice_index[u])
        if vi not in cache:
            ni, no = output_length(voices[vi], 1000, control_rate)
            cache[vi] = (ni / 1000.0, no / 1000.0)
        a, b = cache[vi]
        cost[u] = int(f * (a + b)) + 1
    return cost


def shard_utterances(cost, world_size):
    """Returns a list of index arrays, one per rank (ascending utterance order inside a shard)."""
    owner = shard_plan(cost, world_size)
    return [np.nonzero(owner == r)[0] for r in range(world_size)]


def shard_event_chunks(configs, event_lists, continues_previous, world_size):
    """Control-frame generation over several GPUs: the chunks of an utterance share one drift generator and stay
    together; utterances are dealt by their frame counts (host arithmetic on the event times, gtts_events_frame_count).
    Returns a list of chunk-index arrays, one per rank, chunks in their original order."""
    from . import events_frame_count
    n = len(event_lists)
    cont = np.zeros(n, np.int32) if continues_previous is None else np.asarray(continues_previous, np.int32)
    first = np.nonzero((cont == 0) | (np.arange(n) == 0))[0]                 # first chunk of every utterance
    chain_of = np.searchsorted(first, np.arange(n), side="right") - 1
    frames = np.array([events_frame_count(configs[c], event_lists[c]) for c in range(n)], np.int64)
    cost = np.bincount(chain_of, weights=frames, minlength=len(first)).astype(np.int64) + 1
    owner = shard_plan(cost, world_size)
    return [np.nonzero(owner[chain_of] == r)[0] for r in range(world_size)]


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (identity without a process group)."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
