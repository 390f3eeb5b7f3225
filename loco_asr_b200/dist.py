"""Multi-GPU plumbing: one process per GPU, utterance-sharded, ONE all-gather of pooled embeddings.

The path is embarrassingly parallel over utterances (weights replicated, no cross-utterance dependency), so
the only collective is the final merge of per-rank pooled embeddings; on NVLink 5/NVSwitch it
is microseconds next to the compute and no kernel/collective fusion is warranted.  Backend: NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_pooled(local_pooled: torch.Tensor, local_ids: Sequence[int], counts: Sequence[int], n_total: int) -> torch.Tensor:
    """Merge per-rank pooled embeddings into the global [n_total, D] matrix, in original utterance order.

    local_pooled: [n_local, D] rows in the order of `local_ids` (global utterance indices owned by this rank).
    counts[r]: number of utterances rank r owns (known on every rank: sharding is deterministic).
    One `all_gather_into_tensor` of the rank-padded blocks + one of the ids (tiny)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        out = torch.empty(n_total, local_pooled.shape[1], dtype=local_pooled.dtype, device=local_pooled.device)
        out[torch.as_tensor(np.asarray(local_ids, dtype=np.int64), device=local_pooled.device)] = local_pooled
        return out
    dev, D = local_pooled.device, local_pooled.shape[1]
    cap = int(max(counts))
    block = torch.zeros(cap, D, dtype=local_pooled.dtype, device=dev)
    block[:local_pooled.shape[0]] = local_pooled
    ids = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    ids[:len(local_ids)] = torch.as_tensor(np.asarray(local_ids, dtype=np.int64), device=dev)
    all_blocks = torch.empty(world * cap, D, dtype=local_pooled.dtype, device=dev)
    all_ids = torch.empty(world * cap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_blocks, block)
    dist.all_gather_into_tensor(all_ids, ids)
    keep = all_ids >= 0
    out = torch.empty(n_total, D, dtype=local_pooled.dtype, device=dev)
    out[all_ids[keep]] = all_blocks[keep]
    return out
