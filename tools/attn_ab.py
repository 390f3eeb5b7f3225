"""A/B timing of attention_p2 build variants: LOCO_ASR_LIB=<variant> python tools/attn_ab.py T..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict
enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=False)
enc.debug_set("attn_p2_max_frames", 1 << 30)
res = []
for T in [int(x) for x in sys.argv[1:]]:
    n_samples = (T - 1) * 320 + 400
    n = max(1, 64000 // (T + 2))
    wave = torch.randn(n * n_samples, device="cuda") * 0.1
    ns = [n_samples] * n
    for _ in range(2):
        enc.encode_packed(wave, ns)
    enc.profile_enable(True)
    for _ in range(4):
        enc.encode_packed(wave, ns)
    ms, cnt = enc.profile_collect()["attention"]
    enc.profile_enable(False)
    res.append(f"T={T}: {ms / cnt:.3f}")
print(os.environ.get("LOCO_ASR_LIB"), "  ".join(res), flush=True)
