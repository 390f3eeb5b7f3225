"""Attention kernel timing by utterance length (two-pipeline tcgen05, one-item tcgen05, mma.sync cross-check), ~64k frames per
batch, per-layer milliseconds.  Uses the LOCO_DEBUG library (the product library has only the product kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict

enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=True)
KERNELS = [k for k in [("p2", 0, 1), ("tc", 0, 0), ("mma.sync", 1, 0), ("product", 0, -1)] if k[0] in os.environ.get("ATTN_SWEEP_KERNELS", "p2,tc,mma.sync,product").split(",")]
for T in [int(x) for x in (sys.argv[1:] or [50, 100, 149, 200, 256, 320, 499, 768, 1024, 1499, 2999])]:
    n_samples = (T - 1) * 320 + 400
    n = max(1, 64000 // (T + 2))
    wave = torch.randn(n * n_samples, device="cuda") * 0.1
    ns = [n_samples] * n
    out = []
    for name, impl, p2 in KERNELS:
        enc.debug_set("attn_impl", impl)
        enc.debug_set("attn_p2", 1 if p2 else 0)
        enc.debug_set("attn_p2_max_frames", 193 if p2 < 0 else 1 << 30)      # "product": the shipped per-tile choice between the two kernels
        for _ in range(2):
            enc.encode_packed(wave, ns)
        enc.profile_enable(True)
        for _ in range(3):
            enc.encode_packed(wave, ns)
        prof = enc.profile_collect()
        enc.profile_enable(False)
        ms, cnt = prof["attention"]
        cnt = 3 * 12                                   # per LAYER (the product mix launches both kernels in a layer)
        flops = 12 * (4.0 * 768 * T * T + 2 * 12 * 64 * min(2 * T - 1, 320) * T) * n / 12      # per layer
        out.append(f"{name} {ms / cnt:.3f} ms ({flops / (ms / cnt * 1e-3) / 1e12:6.1f} TF)")
    print(f"T={T:5d} n={n:5d}  " + "   ".join(out), flush=True)
