"""BASELINE.json configs[4] on N GPUs (torchrun): the 512-utterance config-5 subset sharded over the ranks, encoded with the
fused IntentClassifier head, logits merged by one all-gather, and the intent argmax compared with the CPU reference stored in
tests/golden/config5_hf.npz (HF module fp32; and the same module under bf16 autocast as the yardstick).  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from loco_asr_b200 import dist as ldist
from loco_asr_b200.buckets import make_batches, shard_utterances
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.head import IntentHead
from loco_asr_b200.synth import synth_head, synth_state_dict, synth_wave

rank, world, local = ldist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
g = np.load(os.path.join(ROOT, "tests", "golden", "config5_hf.npz"))
ids, n_samples = g["ids"], g["n_samples"]
enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device=dev)
w, b = synth_head(3)
enc.set_head(IntentHead(w, b, None, "average"))
shards = shard_utterances(n_samples, world)
mine = shards[rank]
out = torch.zeros(len(mine), 768 + 101, device=dev)
r0 = 0
for idx in make_batches(n_samples[mine], max_frames=16384):
    sel = mine[idx]
    waves = [synth_wave(int(n_samples[i]), 1234, int(ids[i])) for i in sel]
    p, hp, lg = enc.encode_packed(torch.from_numpy(np.concatenate(waves)).to(dev), [len(x) for x in waves], with_head=True)
    out[r0:r0 + len(sel), :768] = p
    out[r0:r0 + len(sel), 768:] = lg
    r0 += len(sel)
merged = ldist.gather_pooled(out, mine, [len(s) for s in shards], len(ids)).cpu()
if rank == 0:
    pooled, logits = merged[:, :768], merged[:, 768:]
    ref = torch.from_numpy(g["pooled"])
    cos = torch.nn.functional.cosine_similarity(pooled, ref, dim=1)
    rel = (pooled - ref).abs().amax(1) / ref.abs().amax(1)
    agree = logits.argmax(1).numpy() == g["argmax"]
    print(json.dumps({"n_gpus": world, "utterances": int(len(ids)), "min_pooled_cosine": float(cos.min()), "max_rel_err": float(rel.max()),
                      "mean_rel_err": float(rel.mean()), "argmax_equal": int(agree.sum()), "argmax_differs": int((~agree).sum()),
                      "largest_fp32_margin_among_differences": float(g["margin"][~agree].max()) if (~agree).any() else 0.0,
                      "max_logit_diff": float((logits - torch.from_numpy(g["logits"])).abs().max()),
                      "hf_bf16_yardstick": {"argmax_differs": int((g["hf_bf16_argmax"] != g["argmax"]).sum()),
                                            "max_rel_err": float(g["hf_bf16_rel_err"].max()), "mean_rel_err": float(g["hf_bf16_rel_err"].mean()),
                                            "min_cosine": float(g["hf_bf16_cosine"].min()), "max_logit_diff": float(g["hf_bf16_logit_diff"].max())},
                      "bits_checksum": int(pooled.contiguous().view(torch.int32).to(torch.int64).sum())}))
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
