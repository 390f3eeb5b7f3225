"""CPU, world_size 2, gloo: the N > 1 host path -- deterministic sharding + the single all-gather merge."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loco_asr_b200.buckets import make_batches, batch_flops, shard_batches
from loco_asr_b200.dist import gather_pooled
from loco_asr_b200.synth import slurp_shaped_lengths


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_pooled(ids):
    """Stand-in for the encoder: row i is a deterministic function of the global utterance id."""
    ids = torch.as_tensor(np.asarray(ids, dtype=np.int64))
    return torch.stack([ids.float() * 0.5 + k for k in range(8)], dim=1)


def _worker(rank, world, port, n_utts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = slurp_shaped_lengths(n_utts, 11)
    batches = make_batches(lengths, max_frames=8192)
    shards = shard_batches(batch_flops(lengths, batches), world)
    ids = [np.concatenate([batches[b] for b in shards[r]]) if shards[r] else np.zeros(0, np.int64) for r in range(world)]
    counts = [len(x) for x in ids]
    merged = gather_pooled(_fake_pooled(ids[rank]), ids[rank], counts, n_utts)
    ok = torch.equal(merged, _fake_pooled(np.arange(n_utts)))
    q.put((rank, bool(ok), counts))
    dist.destroy_process_group()


def test_two_rank_gather_restores_global_order():
    world, n = 2, 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2] and sum(res[0][2]) == n


def test_single_process_gather_is_a_scatter():
    ids = np.array([3, 0, 2, 1])
    out = gather_pooled(_fake_pooled(ids), ids, [4], 4)
    assert torch.equal(out, _fake_pooled(np.arange(4)))


def test_shard_utterances_is_a_balanced_partition():
    from loco_asr_b200.buckets import shard_utterances
    from loco_asr_b200.flops import total_flops
    lengths = slurp_shaped_lengths(20000, 5)
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lengths, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(lengths)))          # every utterance exactly once
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        assert all(np.all(np.diff(lengths[s]) >= 0) for s in shards)                         # shortest first: dense batches
        fl = [total_flops(lengths[s]) for s in shards]
        assert max(fl) / (sum(fl) / world) < 1.002                                           # FLOP-balanced to 0.2 %
