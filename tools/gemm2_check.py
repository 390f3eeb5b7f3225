"""CTA-pair GEMM (gemm_impl 2) against the single-CTA tcgen05 kernel on the encoder's shapes (bit-exact expected: same
fp32 accumulation order per output element)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from loco_asr_b200 import _lib
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict

enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=True)
g = torch.Generator(device="cuda").manual_seed(0)
for M, N, K, epi in [(256, 256, 64, _lib.EPI_BIAS), (129, 512, 128, _lib.EPI_BIAS), (1000, 2304, 768, _lib.EPI_BIAS),
                     (777, 3072, 768, _lib.EPI_BIAS_GELU), (640, 768, 3072, _lib.EPI_BIAS_RESIDUAL), (40000, 768, 768, _lib.EPI_BIAS_RESIDUAL)]:
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi == _lib.EPI_BIAS_RESIDUAL else None
    c1 = enc.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=0)
    c2 = enc.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=2)
    torch.cuda.synchronize()
    d = (c1.float() - c2.float()).abs().max().item()
    print(f"M={M} N={N} K={K} epi={epi}: max |diff| = {d}  finite={bool(torch.isfinite(c2.float()).all())}", flush=True)
