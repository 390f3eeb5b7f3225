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
def test_tcgen05_gemm_matches_fp32_reference(debug_encoder, case, impl):
    """Both tcgen05 kernels: the CTA-pair product path (cta_group::2, 256 x 256 tiles) and the single-CTA kernel."""
    M, N, K, epi, conv_like = case
    a, lda, w, bias, res, ref = _make(M, N, K, epi, conv_like, seed=M + N + K)
    c = debug_encoder.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=impl, lda=lda, m=M)
    torch.cuda.synchronize()
    err = (c.float() - ref).abs()
    tol = ref.abs() * 2 ** -8 + 2e-3            # bf16 output rounding (+ accumulation-order slack)
    assert bool((err <= tol).all()), f"max err {float(err.max())} at ref {float(ref.flatten()[err.argmax()])}"
    c2 = debug_encoder.debug_gemm(a, w, bias=bias, residual=res, epilogue=epi, impl=1, lda=lda, m=M)
    assert float((c.float() - c2.float()).abs().max()) <= float(ref.abs().max()) * 2 ** -7 + 2e-3


def test_gemm_without_bias(encoder):
    a, lda, w, _, _, _ = _make(300, 512, 1536, _lib.EPI_BIAS_GELU, False, seed=5)
    c = encoder.debug_gemm(a, w, bias=None, epilogue=_lib.EPI_BIAS_GELU, impl=2)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t())
    assert float((c.float() - ref).abs().max()) < 0.03


def _row_stats(x):
    """(mean, M2) of each 128-column slice of the 768-wide rows of x: the layout the GEMM epilogues exchange."""
    xs = x.double().reshape(x.shape[0], 6, 128)
    mean = xs.mean(dim=2)
    return torch.stack([mean, ((xs - mean[:, :, None]) ** 2).sum(dim=2)], dim=2).float()


@pytest.mark.parametrize("M", [1, 300, 40000])
def test_deferred_layernorm_epilogues(encoder, M):
    """The four epilogues that replace the LayerNorm kernels of the transformer layers, each against torch:
    producer (C = A W^T + b + R and its row statistics), consumer (rstd (A W'^T - mean c1) + c2, optionally GELU), and the
    producer whose residual is itself un-normalised ((R - mean) rstd gamma, beta inside the bias).  Rows carry a mean of
    up to +-3 sigma-units so the mean terms matter."""
    g = torch.Generator(device="cuda").manual_seed(1000 + M)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)
    K = 768
    # ---- producer: out_proj-like, plain residual, statistics out
    a = rnd(M, K).bfloat16()
    w = (rnd(768, K) * 0.05).bfloat16()
    bias = rnd(768) * 0.1
    res = (rnd(M, 768) + 3.0 * rnd(M, 1)).bfloat16()
    u_ref = a.float() @ w.float().t() + bias + res.float()
    u, stats = encoder.debug_gemm_ln(a, w, _lib.EPI_BIAS_RESIDUAL_STATS, bias=bias, residual=res, want_stats=True)
    tol = lambda ref: ref.abs() * 2 ** -8 + 4e-3
    assert bool(((u.float() - u_ref).abs() <= tol(u_ref)).all())
    want = _row_stats(u_ref)
    assert torch.allclose(stats[:, :, 0], want[:, :, 0], rtol=0, atol=2e-4)
    assert torch.allclose(stats[:, :, 1], want[:, :, 1], rtol=2e-3, atol=2e-2)
    # ---- consumers: FFN1-like, A = the un-normalised u with its statistics
    gamma, beta = 1.0 + 0.2 * rnd(768), 0.2 * rnd(768)
    w1 = rnd(3072, 768) * 0.05
    b1 = rnd(3072) * 0.1
    w1f = (w1 * gamma).bfloat16()
    c1 = w1f.float().sum(dim=1)
    c2 = w1 @ beta + b1
    x_ln = torch.nn.functional.layer_norm(u.float(), (768,), gamma, beta, 1e-5)
    ref = x_ln @ w1.t() + b1
    for epi, fn in ((_lib.EPI_LN_BIAS, lambda t: t), (_lib.EPI_LN_BIAS_GELU, torch.nn.functional.gelu)):
        got = encoder.debug_gemm_ln(u, w1f, epi, bias=c2, stats_in=stats, c1=c1)
        r = fn(ref)
        # bf16 rounding of gamma (.) W against fp32 W of the reference, K = 768 terms, plus the output rounding
        assert float((got.float() - r).abs().max()) < 0.06 and float((got.float() - r).abs().mean()) < 4e-3
    # ---- producer with an un-normalised residual: FFN2-like
    a2 = rnd(M, 256).bfloat16()
    w2 = (rnd(768, 256) * 0.05).bfloat16()
    b2 = rnd(768) * 0.1
    v_ref = a2.float() @ w2.float().t() + b2 + x_ln
    v, stats2 = encoder.debug_gemm_ln(a2, w2, _lib.EPI_BIAS_LNRESIDUAL_STATS, bias=b2 + beta, residual=u, stats_in=stats, gamma=gamma,
                                      want_stats=True)
    assert bool(((v.float() - v_ref).abs() <= tol(v_ref) + 4e-3).all())
    want2 = _row_stats(v_ref)
    # (the kernel normalises R with the statistics of the producer's fp32 values, torch with those of the bf16-rounded u)
    assert torch.allclose(stats2[:, :, 0], want2[:, :, 0], rtol=0, atol=4e-3)
    assert torch.allclose(stats2[:, :, 1], want2[:, :, 1], rtol=1e-2, atol=0.2)
