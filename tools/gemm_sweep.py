"""Per-shape timing of the tcgen05 GEMM (every GEMM-shaped stage of the encoder at a 64k-frame batch), back-to-back launches
with rotating operands, CUDA events.  LOCO_ASR_LIB=<other .so> python tools/gemm_sweep.py for A/B comparisons."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from loco_asr_b200 import _lib
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict

enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=True)
IMPL = int(os.environ.get("GEMM_IMPL", "0"))      # 0 = tcgen05 single-CTA, 2 = CTA-pair (cta_group::2)
R = 64400
SHAPES = [  # name, M, N, K, epilogue, conv-like
    ("conv1", R * 32, 512, 1536, _lib.EPI_BIAS_GELU, True), ("conv2", R * 16, 512, 1536, _lib.EPI_BIAS_GELU, True),
    ("conv4", R * 4, 512, 1536, _lib.EPI_BIAS_GELU, True), ("conv6", R, 512, 1024, _lib.EPI_BIAS_GELU, True),
    ("qkv", R, 2304, 768, _lib.EPI_BIAS, False), ("out_proj", R, 768, 768, _lib.EPI_BIAS_RESIDUAL, False),
    ("ffn1", R, 3072, 768, _lib.EPI_BIAS_GELU, False), ("ffn2", R, 768, 3072, _lib.EPI_BIAS_RESIDUAL, False),
]
g = torch.Generator(device="cuda").manual_seed(0)
for name, M, N, K, epi, conv in SHAPES:
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    nbuf = 3
    if conv:
        a = [torch.randn(M * 1024 + K + 4096, device="cuda", generator=g).bfloat16() for _ in range(nbuf)]
        lda = 1024
    else:
        a = [torch.randn(M, K, device="cuda", generator=g).bfloat16() for _ in range(nbuf)]
        lda = K
    res = [torch.randn(M, N, device="cuda", generator=g).bfloat16() for _ in range(nbuf)] if epi == _lib.EPI_BIAS_RESIDUAL else [None] * nbuf
    reps = 12 if M > 500000 else 36
    for _ in range(3):
        enc.debug_gemm(a[0], w, bias=bias, residual=res[0], epilogue=epi, impl=IMPL, lda=lda, m=M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        enc.debug_gemm(a[i % nbuf], w, bias=bias, residual=res[i % nbuf], epilogue=epi, impl=IMPL, lda=lda, m=M)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:9s} M={M:8d} N={N:5d} K={K:5d}  {ms:7.4f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del a, res, w
    torch.cuda.empty_cache()
