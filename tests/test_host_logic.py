"""CPU: host-side logic -- config validation, FLOP model, bucketing / sharding, synthetic data, GELU approximation."""
import math

import numpy as np
import pytest
import torch

from loco_asr_b200.buckets import make_batches, shard_batches, batch_flops, interleaved_order, frames_of
from loco_asr_b200.config import LocoSpeechT5Config
from loco_asr_b200.flops import encoder_flops, encoder_flops_breakdown
from loco_asr_b200.synth import slurp_shaped_lengths, synth_wave, synth_state_dict, config1_lengths


def test_config_defaults_match_hf():
    from transformers import SpeechT5Config
    hf = SpeechT5Config()
    cfg = LocoSpeechT5Config.from_hf(hf)
    assert cfg == LocoSpeechT5Config()
    assert cfg.num_frames(48000) == 149 and cfg.min_samples == 400


def test_config_rejects_other_shapes():
    for bad in ({"hidden_size": 1024}, {"feat_extract_norm": "layer"}, {"conv_bias": True}, {"hidden_act": "relu"},
                {"conv_kernel": (10, 3, 3, 3, 3, 3, 2)}, {"encoder_max_relative_position": 64}):
        with pytest.raises(ValueError):
            LocoSpeechT5Config.from_hf({**LocoSpeechT5Config().to_dict(), **bad})


def test_flop_model_matches_baseline_table():
    # BASELINE.md section 3
    for sec, gflop in ((1, 13.90), (3, 43.19), (5, 73.13), (10, 151.07), (30, 508.89), (60, 1183.85)):
        assert abs(encoder_flops(16000 * sec) / 1e9 - gflop) < 0.01, sec
    d = encoder_flops_breakdown(48000)
    assert abs(d["conv1_6"] / d["total"] - 0.339) < 0.002
    assert abs((d["qkvo"] + d["ffn"]) / d["total"] - 0.586) < 0.002


def test_slurp_shaped_lengths():
    l = slurp_shaped_lengths(70000, 1234)
    assert l.min() >= 16000 and l.max() <= 160000
    assert 2.6 < np.median(l) / 16000 < 3.0
    assert np.array_equal(l, slurp_shaped_lengths(70000, 1234))
    assert config1_lengths()[0] == 40000 and config1_lengths()[-1] == 56000 and sum(config1_lengths()) == 48 * 16000


def test_batches_cover_every_utterance_once():
    l = slurp_shaped_lengths(5000, 7)
    batches = make_batches(l, max_frames=16384)
    allidx = np.concatenate(batches)
    assert sorted(allidx.tolist()) == list(range(5000))
    fr = frames_of(l) + 2
    for b in batches:
        assert fr[b].sum() <= 16384 or len(b) == 1
    # length-sorted: batch maxima are non-decreasing
    mx = [l[b].max() for b in batches]
    assert mx == sorted(mx)


def test_ragged_and_tiny_batches():
    assert make_batches([], 1024) == []
    one = make_batches([160000], max_frames=100)   # a single utterance larger than the budget still forms a batch
    assert len(one) == 1 and one[0].tolist() == [0]


def test_flop_balanced_sharding():
    l = slurp_shaped_lengths(20000, 3)
    batches = make_batches(l, max_frames=16384)
    costs = batch_flops(l, batches)
    for world in (2, 4, 8):
        shards = shard_batches(costs, world)
        assert sorted(sum(shards, [])) == list(range(len(batches)))
        load = np.array([costs[s].sum() for s in shards])
        assert load.max() / load.mean() < 1.08, (world, load)
        assert shards == shard_batches(costs, world)     # deterministic on every rank


def test_interleaved_order_is_a_permutation():
    for n in (1, 2, 3, 10, 179, 180, 256):
        assert sorted(interleaved_order(n)) == list(range(n))


def test_synth_wave_is_deterministic():
    a, b = synth_wave(16000, 5, 9), synth_wave(16000, 5, 9)
    assert np.array_equal(a, b) and a.dtype == np.float32 and np.abs(a).max() < 1.0
    assert not np.array_equal(a, synth_wave(16000, 5, 10))


def test_state_dict_has_hf_keys():
    from oracle.hf_reference import build_hf_encoder
    sd = synth_state_dict(seed=0)
    model = build_hf_encoder(sd)     # raises on any missing / unexpected key
    assert set(sd) - {"prenet.masked_spec_embed"} <= set(model.state_dict())


def test_gelu_sigmoid_quintic_error_bound():
    """numpy emulation (fp32) of csrc/common.cuh gelu_erf -- v * sigmoid(2 u(v)), u an odd quintic, v^2 clamped
    at 64 -- against the exact erf GELU over [-20, 20]."""
    from scipy.special import erf
    v = np.linspace(-20, 20, 800001).astype(np.float32)
    v2 = np.minimum(v * v, np.float32(64))
    t = np.float32(-2.3011213394570755) + v2 * (np.float32(-0.10677572400266595) + v2 * np.float32(0.0010142630552579922))
    g = (v / (np.float32(1) + np.exp2((v * t).astype(np.float32)))).astype(np.float32)
    ref = 0.5 * v.astype(np.float64) * (1 + erf(v.astype(np.float64) / math.sqrt(2)))
    assert np.abs(g - ref).max() < 3e-5
    big = np.abs(ref) > 0.05
    assert (np.abs(g - ref)[big] / np.abs(ref[big])).max() < 5e-4        # 8x below one bf16 ulp (2^-8)
    # gelu_erf2 (the packed epilogue form): the same quintic with its constants scaled by -1 / (2 log2 e), evaluated as
    # 0.5 v (1 + tanh(u)).  With an exact tanh it is the same function; the hardware tanh.approx.f32 adds at most
    # 2^-10.987 absolute on tanh, i.e. 2.5e-4 |v| on the result -- still under half a bf16 ulp of |v|.
    c = np.float32(-0.34657359027997264)
    t2 = np.float32(0.79750788) + v2 * (np.float32(0.037005646) + v2 * np.float32(-0.00035151679))
    assert abs(float(np.float32(-2.3011213394570755) * c) - 0.79750788) < 1e-7
    g2 = (np.float32(0.5) * v * (np.float32(1) + np.tanh((v * t2).astype(np.float64)))).astype(np.float32)
    assert np.abs(g2 - ref).max() < 3e-5
    worst = np.abs(g2 - ref) + 2.0 ** -10.987 * 0.5 * np.abs(v)
    keep = np.abs(v) <= 8
    assert (worst[keep] / np.maximum(np.abs(v[keep]), 1e-3)).max() < 2.0 ** -9          # half a bf16 ulp of |v|


def test_exp2_polynomial_error_bound():
    """numpy emulation (fp32, bit-level) of csrc/common.cuh ex2_poly2 -- round-to-nearest split through the 1.5 * 2^23 magic
    number, cubic on [-1/2, 1/2], integer add into the exponent field -- against 2^x over the range the softmax feeds it:
    (-inf, 8].  Relative error below 1e-4, i.e. 40x under the bf16 rounding (2^-9) of the probability it produces."""
    x = np.concatenate([np.linspace(-125, 8, 1_000_001), [-126.0, -1e4, -np.inf, 0.0, -0.5, 0.5, -0.49999, 7.99999]]).astype(np.float32)
    xc = np.maximum(x, np.float32(-125))
    magic = np.float32(12582912.0)
    t = (xc + magic).astype(np.float32)
    n = (t - magic).astype(np.float32)
    f = (xc - n).astype(np.float32)
    assert np.abs(f).max() <= 0.5
    p = np.float32(0.055171459913253784) * f + np.float32(0.2426108568906784)
    p = (p.astype(np.float32) * f + np.float32(0.6932609677314758)).astype(np.float32)
    p = (p * f + np.float32(0.9999281167984009)).astype(np.float32)
    bits = p.view(np.int32).astype(np.int64) + ((t.view(np.int32).astype(np.int64) << 23) & 0xFFFFFFFF)
    got = (bits & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    ref = np.exp2(xc.astype(np.float64))
    assert np.isfinite(got).all() and (got > 0).all()
    assert (np.abs(got.astype(np.float64) / ref - 1)).max() < 1e-4
    assert got[np.isneginf(x)][0] < 1e-37                     # masked keys: 2^-125, nothing against a row sum >= 1


def test_polyphase_positional_conv_algebra():
    """The index algebra of posconv_pp.cu restated in numpy for one group: utterances on a timeline with 64 zero frames between
    them, the window de-interleaved by frame phase, weights as [3 zero taps | 128 taps | zero taps], step j multiplying phase
    array j % 4 (rows shifted by j // 4) with the four taps j-3 .. j side by side, column block q holding output phase 3 - q --
    against the reference's Conv1d(k = 128, padding = 64) with the last frame dropped (HF modeling_speecht5.py:355-397, 445-453)."""
    rng = np.random.default_rng(5)
    C, K, P, ROWS, TILE, HALO = 6, 128, 4, 128, 512, 64            # 6 channels instead of 48: the algebra does not depend on it
    lens = [1, 37, 150, 700, 64]
    xs = [rng.standard_normal((t, C)) for t in lens]
    w = rng.standard_normal((K, C, C)) * 0.1                        # [tap][in][out]
    n_steps = K + P - 1
    wp = np.zeros((n_steps + 8, C, C))
    wp[3:3 + K] = w
    total = sum(lens) + HALO * len(lens)
    n_vt = (total + TILE - 1) // TILE
    vmap = -np.ones(n_vt * TILE + 2 * HALO, dtype=np.int64)         # vmap[HALO + v] = packed row of timeline frame v
    packed = np.concatenate(xs)
    pos, row = HALO, 0
    for t in lens:
        vmap[pos:pos + t] = np.arange(row, row + t)
        pos += t + HALO
        row += t
    out = np.zeros_like(packed)
    for vt in range(n_vt):
        win = np.zeros((ROWS * P + K, C))                           # timeline frames vt*TILE - 64 .. + 575
        idx = vmap[vt * TILE: vt * TILE + ROWS * P + K]
        win[idx >= 0] = packed[idx[idx >= 0]]
        phase = [win[b::P] for b in range(P)]                       # x_b[r] = X[4 r + b], 160 rows each
        acc = np.zeros((ROWS, P * C))
        for j in range(n_steps):
            a = phase[j % P][j // P: j // P + ROWS]                 # row s = frame 4 s + j - 64
            b = np.concatenate([wp[j + q] for q in range(P)], axis=1)        # [in][4 * out]: taps j-3+q
            acc += a @ b
        for s in range(ROWS):
            for q in range(P):
                r = vmap[HALO + vt * TILE + P * s + (P - 1 - q)]
                if r >= 0:
                    out[r] = acc[s, q * C:(q + 1) * C]
    ref, row = np.zeros_like(packed), 0
    for x in xs:
        xt = torch.from_numpy(x.T.copy())[None]                     # [1, C, T]
        wt = torch.from_numpy(np.transpose(w, (2, 1, 0)).copy())    # [out, in, tap]
        y = torch.nn.functional.conv1d(xt, wt, padding=K // 2)[0, :, :x.shape[0]]     # SamePad: drop the last frame
        ref[row:row + x.shape[0]] = y.T.numpy()
        row += x.shape[0]
    assert np.abs(out - ref).max() < 1e-10


def test_conv0_split_gemm_algebra():
    """conv0_tc.cu's K = 48 operand layout restated in numpy: A = [x_hi 1 0.. | x_lo 0.. | x_hi 1 0..], B = [w'_hi sh_hi | w'_hi 0 |
    w'_lo sh_lo] with bf16 operands and fp32 accumulation reproduces scale * conv(x, w) + shift to ~2^-16 -- the 3-term split and
    the GroupNorm shift carried in the padding taps."""
    def bf16(a):
        return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy()
    rng = np.random.default_rng(7)
    n_frames, n_ch = 300, 64
    x = rng.standard_normal(5 * n_frames + 5).astype(np.float32) * 0.3
    w = (rng.standard_normal((n_ch, 10)) * 0.2).astype(np.float32)
    scale = (0.5 + rng.random(n_ch) * 3).astype(np.float32)
    shift = rng.standard_normal(n_ch).astype(np.float32)
    taps = np.stack([x[5 * f: 5 * f + 10] for f in range(n_frames)])          # [frames, 10]
    x_hi = bf16(taps)
    x_lo = bf16(taps - x_hi)
    wf = w * scale[:, None]
    w_hi = bf16(wf)
    w_lo = bf16(wf - w_hi)
    sh_hi = bf16(shift)
    sh_lo = bf16(shift - sh_hi)
    A = np.zeros((n_frames, 48), dtype=np.float32)
    B = np.zeros((n_ch, 48), dtype=np.float32)
    A[:, 0:10], A[:, 10] = x_hi, 1.0
    A[:, 16:26] = x_lo
    A[:, 32:42], A[:, 42] = x_hi, 1.0
    B[:, 0:10], B[:, 10] = w_hi, sh_hi
    B[:, 16:26] = w_hi
    B[:, 32:42], B[:, 42] = w_lo, sh_lo
    got = A.astype(np.float64) @ B.T.astype(np.float64)
    ref = taps.astype(np.float64) @ wf.T.astype(np.float64) + shift[None, :].astype(np.float64)
    bound = 2.0 ** -15 * (np.abs(taps).astype(np.float64) @ np.abs(wf.T).astype(np.float64) + np.abs(shift)[None, :])
    assert np.all(np.abs(got - ref) <= bound)
    # one bf16 pass alone is two orders worse: the split is what buys fp32-like accuracy
    one = x_hi.astype(np.float64) @ w_hi.T.astype(np.float64) + shift[None, :]
    assert np.abs(one - ref).max() > 30 * np.abs(got - ref).max()


def test_deferred_layernorm_fold_algebra():
    """api.cu finalize_impl `fold` + the EPI_LN_BIAS epilogue restated in numpy: with W' = bf16(gamma (.) W), c1[n] = sum_k W'[n, k]
    (of the ROUNDED W') and c2 = W beta + b, `rstd_r (u W'^T - mean_r c1) + c2` is LayerNorm(u) W^T + b up to the bf16 rounding of W'
    alone -- the mean term cancels exactly because c1 sums what the MMA multiplies.  Also the residual form the following GEMM
    reads: (R - mean) rstd gamma with beta folded into its bias."""
    def bf16(a):
        return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy().astype(np.float64)
    rng = np.random.default_rng(11)
    M, K, N = 64, 768, 96
    u = bf16(rng.standard_normal((M, K)) * 1.7 + 40.0)               # un-normalised residual sums as the GEMM reads them (bf16), large common offset
    W = rng.standard_normal((N, K)) * 0.05
    b = rng.standard_normal(N)
    gamma = 0.5 + rng.random(K) * 2
    beta = rng.standard_normal(K)
    mean = u.mean(1, keepdims=True)
    rstd = 1.0 / np.sqrt(u.var(1, keepdims=True) + 1e-5)
    Wf = bf16(W * gamma[None, :])
    c1 = Wf.sum(1)
    c2 = W @ beta + b
    got = rstd * (u @ Wf.T - mean * c1[None, :]) + c2[None, :]
    ln = (u - mean) * rstd
    assert np.abs(got - (ln @ Wf.T + c2[None, :])).max() < 1e-9       # exact identity (fp64): the 40.0 offset cancels
    ref = (ln * gamma + beta) @ W.T + b                              # LayerNorm then Linear, as HF computes it
    bound = 2.0 ** -8 * (np.abs(ln) @ np.abs(W * gamma[None, :]).T) + 1e-9
    assert np.all(np.abs(got - ref) <= bound)                        # what is left is the rounding of W' to bf16
    # c1 from the UNROUNDED product would leave mean * (sum W' - sum gamma W) * rstd behind: visible with the 40.0 offset
    c1_bad = (W * gamma[None, :]).sum(1)
    bad = rstd * (u @ Wf.T - mean * c1_bad[None, :]) + c2[None, :]
    assert np.abs(bad - ref).max() > 5 * np.abs(got - ref).max()
    # residual read of the next GEMM (EPI_BIAS_LNRESIDUAL_STATS): out = acc + bias' + (R - mean) rstd gamma, bias' = bias + beta
    acc = rng.standard_normal((M, K))
    bias = rng.standard_normal(K)
    out = acc + (bias + beta)[None, :] + (u - mean) * rstd * gamma[None, :]
    assert np.abs(out - (acc + bias[None, :] + (ln * gamma + beta))).max() < 1e-9


def test_sliced_row_statistics_merge_to_the_row_statistics():
    """The producers write (mean, M2) per 128-column slice of a 768-wide row ([R6, 6, 2] fp32, no atomics); readers merge the six
    with Chan's formula.  The merge equals the two-pass mean / variance of the whole row, also with a large common offset."""
    rng = np.random.default_rng(12)
    x = rng.standard_normal((50, 768)) * 3.0 + 1000.0
    sl = x.reshape(50, 6, 128)
    m_s = sl.mean(2)
    M2_s = ((sl - m_s[..., None]) ** 2).sum(2)
    n, mean, M2 = 0.0, np.zeros(50), np.zeros(50)
    for s in range(6):
        nb = 128.0
        d = m_s[:, s] - mean
        tot = n + nb
        mean = mean + d * nb / tot
        M2 = M2 + M2_s[:, s] + d * d * n * nb / tot
        n = tot
    assert np.abs(mean - x.mean(1)).max() < 1e-9
    assert np.abs(M2 / 768.0 - x.var(1)).max() < 1e-8


def test_lazy_rescale_online_softmax_algebra():
    """The attention kernels' softmax restated in numpy (attention_p2.cu / attention_tc.cu): scores in log2 units, 64-key blocks,
    the running maximum moves only when a block's maximum exceeds it by more than 8 (LOCO_LAZY_RESCALE; P may reach 2^8), P is
    rounded to bf16 BEFORE it is both multiplied into O and summed into l (l = P . 1 on the tensor core), out = O / l.  Whatever
    the rescale schedule -- every new maximum, the lazy rule, or the lazy rule voted per 32-row warp -- the result is the exact
    softmax(S) V up to the bf16 rounding of P, and the weights that were applied sum to exactly one."""
    def bf16(a):
        return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy().astype(np.float64)
    rng = np.random.default_rng(21)
    R, T, D, FK = 64, 333, 64, 64
    S = rng.standard_normal((R, T)) * 6.0 + np.linspace(0, 30, T)[None, :]     # maxima keep growing along the keys: many rescales
    V = bf16(rng.standard_normal((T, D)))
    ref = np.exp2(S - S.max(1, keepdims=True))
    ref = (ref / ref.sum(1, keepdims=True)) @ V

    def run(threshold, warp_vote):
        m = np.full(R, -np.inf)
        O = np.zeros((R, D))
        l = np.zeros(R)
        n_rescales = 0
        for j in range(0, T, FK):
            s = S[:, j:j + FK]
            bm = s.max(1)
            trig = bm > m + threshold
            if warp_vote:                                   # __any_sync: one row's trigger moves every row of its 32-row warp
                trig = np.repeat(trig.reshape(-1, 32).any(1), 32)
            mx = np.where(trig, np.maximum(m, bm), m)
            with np.errstate(invalid="ignore"):
                corr = np.where(trig, np.exp2(m - mx), 1.0)
            corr = np.where(np.isnan(corr), 0.0, corr)      # first block: exp2(-inf - x) = 0
            n_rescales += int(trig.sum())
            O *= corr[:, None]
            l *= corr
            m = mx
            P = bf16(np.exp2(s - m[:, None]))
            assert P.max() <= 2.0 ** threshold * 1.01 + 1.0
            O += P @ V[j:j + FK]
            l += P.sum(1)
        return O / l[:, None], n_rescales

    eager, n_eager = run(0.0, False)
    lazy, n_lazy = run(8.0, False)
    voted, n_voted = run(8.0, True)
    assert n_lazy < n_eager and n_lazy <= n_voted
    scale = np.abs(ref).max()
    for out in (eager, lazy, voted):
        assert np.abs(out - ref).max() <= 2.0 ** -8 * scale * 2
    # the three schedules differ only by roundings of P (2^-9 relative each), never by the schedule itself
    assert np.abs(lazy - eager).max() <= 2.0 ** -8 * scale * 2


def test_groupnorm_statistics_from_waveform_moments():
    """frontend.cu: conv0 is linear with 10 taps and stride 5, so GroupNorm's per-channel mean / biased variance over the T0 output
    frames (HF modeling_speecht5.py:277-281) follow from 10 first and 55 second moments of the WAVEFORM -- no pass over the
    512 x T0 conv output.  Restated in numpy (fp64 reduction, as the kernel's) against torch's conv1d + GroupNorm statistics."""
    rng = np.random.default_rng(5)
    n = 16000 + 7                                                     # ragged tail: samples past the last frame's taps are unused
    x = (rng.standard_normal(n) * 0.1 + 0.03).astype(np.float32)     # a DC offset makes the mean term matter
    w = (rng.standard_normal((512, 10)) * 0.3).astype(np.float32)
    t0 = (n - 10) // 5 + 1
    taps = np.stack([x[k: k + 5 * (t0 - 1) + 1: 5] for k in range(10)]).astype(np.float64)     # [10, T0]: x[5 t + k]
    m1 = taps.mean(1)                                                 # 10 first moments
    iu = np.triu_indices(10)
    m2 = (taps[iu[0]] * taps[iu[1]]).mean(1)                          # 55 second moments R_kk', k <= k'
    assert m2.shape == (55,)
    R = np.zeros((10, 10))
    R[iu] = m2
    R = R + R.T - np.diag(np.diag(R))
    wd = w.astype(np.float64)
    mean = wd @ m1
    var = np.einsum("ck,kl,cl->c", wd, R, wd) - mean ** 2
    y = torch.nn.functional.conv1d(torch.from_numpy(x).double()[None, None], torch.from_numpy(w).double()[:, None, :], stride=5)[0]
    assert y.shape == (512, t0)
    assert np.abs(mean - y.mean(1).numpy()).max() < 1e-12
    assert np.abs(var - y.var(1, unbiased=False).numpy()).max() < 1e-12
    # the folded form conv0_tc.cu consumes: y_norm = scale_c * conv(x, w_c) + shift_c
    gamma, beta = rng.standard_normal(512), rng.standard_normal(512)
    scale = gamma / np.sqrt(var + 1e-5)
    shift = beta - mean * scale
    gn = torch.nn.functional.group_norm(y[None], 512, torch.from_numpy(gamma), torch.from_numpy(beta), eps=1e-5)[0].numpy()
    assert np.abs(scale[:, None] * y.numpy() + shift[:, None] - gn).max() < 1e-9


def test_strided_conv_as_overlapping_row_gemm():
    """conv1-6 (HF modeling_speecht5.py:210-228: Conv1d(512, 512, k, stride 2, no bias) on a channel-major tensor) as the GEMM
    api.cu launches on the TIME-MAJOR activation [T_in, 512]: row t of the A operand is the contiguous strip of k input frames
    starting at frame 2t (row stride lda = 2 * 512, K = k * 512 -- consecutive rows overlap), and the weight is re-laid
    [out][tap * 512 + in] (api.cu finalize_impl).  Same numbers as torch's conv1d for k = 3 and k = 2."""
    rng = np.random.default_rng(9)
    C = 512
    for k, t_in in ((3, 41), (2, 30), (3, 40)):
        x = rng.standard_normal((t_in, C))                                  # time-major, as the kernels store it
        w = rng.standard_normal((C, C, k)) * 0.05                           # HF layout [out][in][tap]
        t_out = (t_in - k) // 2 + 1
        flat = x.reshape(-1)
        lda = 2 * C
        A = np.lib.stride_tricks.as_strided(flat, shape=(t_out, k * C), strides=(lda * flat.itemsize, flat.itemsize))
        Wg = w.transpose(0, 2, 1).reshape(C, k * C)                          # [out][tap * 512 + in]
        got = A @ Wg.T
        ref = torch.nn.functional.conv1d(torch.from_numpy(x.T.copy())[None], torch.from_numpy(w), stride=2)[0].numpy().T
        assert ref.shape == (t_out, C)
        assert np.abs(got - ref).max() < 1e-9
        # rows the GEMM may touch past the last frame: (t_out - 1) * lda + k * C <= t_in * C, so the 8 zeroed pad frames api.cu
        # appends to every conv buffer are only ever read by the slot-padding rows, never by a valid output frame
        assert (t_out - 1) * lda + k * C <= t_in * C
