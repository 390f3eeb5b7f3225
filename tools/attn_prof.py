"""One packed encode at a fixed utterance length (for ncu captures of the attention kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict

T = int(sys.argv[1]) if len(sys.argv) > 1 else 149
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=True)
enc.debug_set("attn_impl", 0)
enc.debug_set("attn_p2", 1 if impl == 2 else 0)
enc.debug_set("attn_p2_max_frames", 1 << 30)
n_samples = (T - 1) * 320 + 400
n = max(1, 64000 // (T + 2))
wave = torch.randn(n * n_samples, device="cuda") * 0.1
enc.encode_packed(wave, [n_samples] * n)
torch.cuda.synchronize()
