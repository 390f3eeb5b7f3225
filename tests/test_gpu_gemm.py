"""GPU: the tcgen05/TMEM/TMA GEMM through the C ABI (loco_debug_gemm) against an fp32 torch matmul of the same
bf16 operands and against the SIMT debug GEMM.  Tolerance: one bf16 rounding of the fp32-accumulated result."""
import pytest
import torch

from loco_asr_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [  # M, N, K, epilogue, conv-like overlapping rows
    (1, 256, 64, _lib.EPI_BIAS, False),
    (127, 256, 128, _lib.EPI_BIAS, False),
    (128, 512, 512, _lib.EPI_BIAS, False),
    (129, 768, 768, _lib.EPI_BIAS_RESIDUAL, False),
    (1000, 2304, 768, _lib.EPI_BIAS, False),
    (777, 3072, 768, _lib.EPI_BIAS_GELU, False),
    (640, 768, 3072, _lib.EPI_BIAS_RESIDUAL, False),
    (999, 512, 1536, _lib.EPI_BIAS_GELU, True),
    (4001, 512, 1024, _lib.EPI_BIAS_GELU, True),
    (40000, 768, 768, _lib.EPI_BIAS_RESIDUAL, False),   # > 148 tiles: exercises the persistent loop + TMEM double buffer
]


def _make(M, N, K, epi, conv_like, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    if conv_like:
        flat = torch.randn(M * 1024 + K + 4096, device="cuda", generator=g).bfloat16()
        a_mat, a_arg, lda = torch.as_strided(flat, (M, K), (1024, 1)), flat, 1024
    else:
        a_mat = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        a_arg, lda = a_mat, K
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi == _lib.EPI_BIAS_RESIDUAL else None
    ref = a_mat.float() @ w.float().t() + bias
    if epi == _lib.EPI_BIAS_GELU:
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res.float()
    return a_arg, lda, w, bias, res, ref


@pytest.mark.parametrize("impl", [2, 0], ids=["cta_pair", "single_cta"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"M{c[0]}_N{c[1]}_K{c[2]}_e{c[3]}{'_conv' if c[4] else ''}")
def test_tcgen05_gemm_matches_fp32_reference(encoder, case, impl):
    """Both tcgen05 kernels: the CTA-pair product path (cta_group::2, 256 x 256 tiles) and the single-CTA kernel."""
    M, N, K, epi, conv_like = case
    a, lda, w, bias, res, ref = _make(M, N, K, epi, conv_like, seed=M + N + K)
    c = encoder.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=impl, lda=lda, m=M)
    torch.cuda.synchronize()
    err = (c.float() - ref).abs()
    tol = ref.abs() * 2 ** -8 + 2e-3            # bf16 output rounding (+ accumulation-order slack)
    assert bool((err <= tol).all()), f"max err {float(err.max())} at ref {float(ref.flatten()[err.argmax()])}"
    c2 = encoder.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=1, lda=lda, m=M)
    assert float((c.float() - c2.float()).abs().max()) <= float(ref.abs().max()) * 2 ** -7 + 2e-3


def test_gemm_without_bias(encoder):
    a, lda, w, _, _, _ = _make(300, 512, 1536, _lib.EPI_BIAS_GELU, False, seed=5)
    c = encoder.debug_gemm(a, w, bias=None, epilogue=_lib.EPI_BIAS_GELU, impl=2)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t())
    assert float((c.float() - ref).abs().max()) < 0.03
