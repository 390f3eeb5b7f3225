"""Two ranks, NCCL, one box: the sharded product path (shard_utterances -> per-rank encode -> gather_pooled) gives, bit for
bit, the matrix a single GPU computes for the same utterances.  Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  SURVEY.md 8e / BASELINE.json configs[2]."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utts, out_path):
    import torch.distributed as dist
    from loco_asr_b200 import dist as ldist
    from loco_asr_b200.buckets import make_batches, shard_utterances
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.synth import slurp_shaped_lengths, synth_state_dict, synth_waves_by_id

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    ldist.init_from_env()
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device=dev)
    lengths = slurp_shaped_lengths(n_utts, 77)
    shards = shard_utterances(lengths, world)
    mine = shards[rank]
    pooled = torch.empty(len(mine), 768, device=dev)
    r0 = 0
    for idx in make_batches(lengths[mine], max_frames=16384):
        ids = mine[idx]
        wave = synth_waves_by_id(lengths[ids], ids, 5, dev)
        enc.encode_packed(wave, lengths[ids].astype(np.int32), out=pooled[r0:r0 + len(ids)])
        r0 += len(ids)
    merged = ldist.gather_pooled(pooled, mine, [len(s) for s in shards], n_utts)
    if rank == 0:
        # the same utterances on ONE GPU, in their original order and in different batches
        ref = torch.empty(n_utts, 768, device=dev)
        all_ids = np.arange(n_utts)
        for s in range(0, n_utts, 97):
            ids = all_ids[s:s + 97]
            ref[s:s + len(ids)] = enc.encode_packed(synth_waves_by_id(lengths[ids], ids, 5, dev), lengths[ids].astype(np.int32))
        torch.save({"equal": bool(torch.equal(merged, ref)), "finite": bool(torch.isfinite(merged).all()),
                    "maxdiff": float((merged - ref).abs().max())}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_rank_nccl_sharded_encode_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    world, n = 2, 600
    out = str(tmp_path / "res.pt")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = torch.load(out)
    assert res["finite"] and res["equal"], res
