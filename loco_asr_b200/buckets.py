"""Length bucketing, batch packing and FLOP-balanced sharding of utterances (host logic, pure numpy).

The encoder itself is padding-free (utterances are packed back to back), so bucketing is not needed for
correctness or to avoid padded FLOPs; sorting by length keeps the per-utterance tile grids of the attention /
positional-conv kernels dense and gives every launch a predictable shape.  Sharding across GPUs is by the
algorithmic FLOP model (flops.py), not by utterance count.  The only precedent in the
reference is the length-binned batching of lms/src/utils.py:18-38 (idea only).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .config import LocoSpeechT5Config
from .flops import encoder_flops


def frames_of(lengths: Sequence[int], cfg: LocoSpeechT5Config | None = None) -> np.ndarray:
    cfg = cfg or LocoSpeechT5Config()
    t = np.asarray(lengths, dtype=np.int64)
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        t = np.where(t >= k, (t - k) // s + 1, 0)
    return t


def make_batches(lengths: Sequence[int], max_frames: int = 131072, max_utts: int = 32768,
                 cfg: LocoSpeechT5Config | None = None) -> List[np.ndarray]:
    """Sort by length and cut into batches of at most `max_frames` encoder frames (+2 slot rows per utterance).
    Returns index arrays into `lengths`; every utterance appears exactly once."""
    lengths = np.asarray(lengths, dtype=np.int64)
    frames = frames_of(lengths, cfg) + 2
    order = np.argsort(lengths, kind="stable")
    batches, cur, acc = [], [], 0
    for i in order:
        f = int(frames[i])
        if cur and (acc + f > max_frames or len(cur) >= max_utts):
            batches.append(np.asarray(cur, dtype=np.int64))
            cur, acc = [], 0
        cur.append(int(i))
        acc += f
    if cur:
        batches.append(np.asarray(cur, dtype=np.int64))
    return batches


def batch_flops(lengths: Sequence[int], batches: List[np.ndarray], cfg: LocoSpeechT5Config | None = None) -> np.ndarray:
    lengths = np.asarray(lengths, dtype=np.int64)
    cache = {}
    out = np.zeros(len(batches))
    for b, idx in enumerate(batches):
        tot = 0.0
        for n in lengths[idx]:
            n = int(n)
            if n not in cache:
                cache[n] = encoder_flops(n, cfg)
            tot += cache[n]
        out[b] = tot
    return out


def shard_batches(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy: assign batches (heaviest first) to the least-loaded rank.
    Deterministic, so every rank derives the same assignment from the lengths alone."""
    order = sorted(range(len(costs)), key=lambda b: (-costs[b], b))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for b in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        out[r].append(b)
        load[r] += costs[b]
    for r in range(world_size):
        out[r].sort()
    return out


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Utterance-level sharding for one box of `world_size` GPUs: sort by length (stable) and deal the sorted list
    round-robin, so every rank gets the same length distribution -- hence the same FLOPs to within one utterance per
    length class and the same batch shapes -- and every rank can derive every other rank's share from the lengths alone
    (no ids travel).  Returns, per rank, the POSITIONS into `lengths` it owns, shortest first."""
    order = np.argsort(np.asarray(lengths, dtype=np.int64), kind="stable")
    return [order[r::world_size] for r in range(world_size)]


def interleaved_order(n: int) -> List[int]:
    """Visit 0..n-1 with a stride coprime to n so that any prefix is a representative mix of lengths."""
    if n <= 2:
        return list(range(n))
    stride = max(1, int(round(n * 0.381966)))
    while np.gcd(stride, n) != 1:
        stride += 1
    return [(i * stride) % n for i in range(n)]
