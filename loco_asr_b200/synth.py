"""Seeded synthetic inputs: waveforms, SLURP-shaped utterance lengths and random-init weights.

There is no network in the build/bench environment, so neither SLURP audio nor the
``microsoft/speecht5_asr`` checkpoint exist; BASELINE.json asks for synthetic 16 kHz utterances and
random-init weights of the SpeechT5-base architecture.

The weight generator deliberately gives every bias / LayerNorm affine / GroupNorm affine a non-trivial
value (HF's default init leaves them at 0 / 1, which would hide bugs in those code paths) and makes
``pe_k`` and the q/k projections large enough that the relative-position bias and the softmax are
far from uniform.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

from .config import LocoSpeechT5Config

SAMPLE_RATE = 16000


# ----------------------------------------------------------------------------------------------
# waveforms
# ----------------------------------------------------------------------------------------------
def synth_wave(n_samples: int, seed: int, idx: int = 0, kind: str = "mix") -> np.ndarray:
    """One utterance: ``0.1*env(t)*sum_m a_m sin(2 pi f_m t + phi_m) + 0.005*N(0,1)`` (kind="mix")
    or ``0.1*N(0,1)`` (kind="noise"); float32, deterministic in (seed, idx)."""
    rng = np.random.default_rng([int(seed), int(idx)])
    if kind == "noise":
        return (0.1 * rng.standard_normal(n_samples)).astype(np.float32)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    f = rng.uniform(80.0, 3080.0, size=5)
    a = rng.uniform(0.0, 1.0, size=5)
    ph = rng.uniform(0.0, 2 * math.pi, size=5)
    fe = rng.uniform(0.5, 2.5)
    env = np.sin(2 * math.pi * fe * t) ** 2
    x = np.zeros(n_samples, dtype=np.float64)
    for m in range(5):
        x += a[m] * np.sin(2 * math.pi * f[m] * t + ph[m])
    x = 0.1 * env * x + 0.005 * rng.standard_normal(n_samples)
    return x.astype(np.float32)


def config1_lengths() -> List[int]:
    """BASELINE.json configs[0]: 16 utterances, L_i = 16000*(2.5 + i/15) samples (2.5..3.5 s)."""
    return [int(round(SAMPLE_RATE * (2.5 + i / 15.0))) for i in range(16)]


def slurp_shaped_lengths(n_utts: int = 70000, seed: int = 1234) -> np.ndarray:
    """BASELINE.json configs[1]: log-normal durations clipped to [1, 10] s, median ~2.8 s, in samples."""
    rng = np.random.default_rng(seed)
    dur = np.exp(rng.normal(math.log(2.8), 0.45, size=n_utts))
    dur = np.clip(dur, 1.0, 10.0)
    return np.round(dur * SAMPLE_RATE).astype(np.int64)


def synth_waves_device(lengths: Sequence[int], seed: int, device, chunk: int = 2048) -> "torch.Tensor":
    """Bulk generator for the bench: a packed float32 waveform tensor [sum(lengths)] built on `device`
    with the same recipe as :func:`synth_wave` (different random stream)."""
    lengths = torch.as_tensor(np.asarray(lengths, dtype=np.int64))
    total = int(lengths.sum())
    out = torch.empty(total, dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    cu = torch.zeros(len(lengths) + 1, dtype=torch.int64)
    cu[1:] = torch.cumsum(lengths, 0)
    for s in range(0, len(lengths), chunk):
        e = min(s + chunk, len(lengths))
        ln = lengths[s:e].to(device)
        n = e - s
        lmax = int(ln.max())
        t = torch.arange(lmax, device=device, dtype=torch.float32)[None, :] / SAMPLE_RATE
        f = torch.rand(n, 5, device=device, generator=g) * 3000.0 + 80.0
        a = torch.rand(n, 5, device=device, generator=g)
        ph = torch.rand(n, 5, device=device, generator=g) * (2 * math.pi)
        fe = torch.rand(n, 1, device=device, generator=g) * 2.0 + 0.5
        x = torch.zeros(n, lmax, device=device)
        for m in range(5):
            x += a[:, m:m + 1] * torch.sin(2 * math.pi * f[:, m:m + 1] * t + ph[:, m:m + 1])
        x = 0.1 * torch.sin(2 * math.pi * fe * t) ** 2 * x
        x += 0.005 * torch.randn(n, lmax, device=device, generator=g)
        mask = torch.arange(lmax, device=device)[None, :] < ln[:, None]
        out[int(cu[s]):int(cu[e])] = x[mask]
        del x, mask
    return out


def _mix64(x: "torch.Tensor") -> "torch.Tensor":
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic; logical shifts emulated with masks)."""
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * -4658895280553007687          # 0xBF58476D1CE4E5B9
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * -7723592293110705685         # 0x94D049BB133111EB
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def _uniform01(h: "torch.Tensor", shift: int = 0) -> "torch.Tensor":
    """24 bits of a hash -> float32 in (0, 1)."""
    return (((h >> shift) & 0xFFFFFF).to(torch.float32) + 0.5) * (1.0 / 16777216.0)


def synth_waves_by_id(lengths: Sequence[int], ids: Sequence[int], seed: int, device, max_elems: int = 1 << 26) -> "torch.Tensor":
    """Packed float32 waveforms [sum(lengths)] on `device`, utterance u a pure function of (seed, ids[u], lengths[u]):
    the same utterance gets the same samples whatever batch, rank or order it is generated in -- what lets a sharded
    multi-GPU pass be compared bit for bit with the single-GPU pass.  Same recipe as :func:`synth_wave` (amplitude-modulated
    5-tone mixture + 0.005 N(0,1)), counter-based random numbers (splitmix64 of (seed, id, sample index))."""
    lengths = np.asarray(lengths, dtype=np.int64)
    ids_t = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=device)
    total = int(lengths.sum())
    out = torch.empty(total, dtype=torch.float32, device=device)
    cu = np.concatenate([[0], np.cumsum(lengths)])
    base = _mix64(ids_t * -7046029254386353131 + int(seed))              # 0x9E3779B97F4A7C15: one stream per (seed, id)
    s = 0
    n_all = len(lengths)
    while s < n_all:
        e, lmax = s + 1, int(lengths[s])
        while e < n_all and max(lmax, int(lengths[e])) * (e - s + 1) <= max_elems:
            lmax = max(lmax, int(lengths[e]))
            e += 1
        n = e - s
        b = base[s:e, None]
        par = _mix64(b + torch.arange(1, 17, device=device, dtype=torch.int64)[None, :] * 6364136223846793005)
        u = _uniform01(par)                                               # [n, 16] utterance parameters
        f, a, ph, fe = u[:, 0:5] * 3000.0 + 80.0, u[:, 5:10], u[:, 10:15] * (2 * math.pi), u[:, 15:16] * 2.0 + 0.5
        t = torch.arange(lmax, device=device, dtype=torch.float32)[None, :] / SAMPLE_RATE
        x = torch.zeros(n, lmax, device=device)
        for m in range(5):
            x += a[:, m:m + 1] * torch.sin(2 * math.pi * f[:, m:m + 1] * t + ph[:, m:m + 1])
        x = 0.1 * torch.sin(2 * math.pi * fe * t) ** 2 * x
        h = _mix64(b ^ (torch.arange(lmax, device=device, dtype=torch.int64)[None, :] * -3335678366873096957 + 0x5851F42D))
        x += 0.005 * torch.sqrt(-2.0 * torch.log(_uniform01(h))) * torch.cos((2 * math.pi) * _uniform01(h, 24))
        del h
        mask = torch.arange(lmax, device=device)[None, :] < torch.as_tensor(lengths[s:e], device=device)[:, None]
        out[int(cu[s]):int(cu[e])] = x[mask]
        del x, mask
        s = e
    return out


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
def synth_state_dict(cfg: LocoSpeechT5Config | None = None, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init encoder weights under the HF ``SpeechT5EncoderWithSpeechPrenet.state_dict()`` key
    names (transformers 5.x spelling: weight-norm as ``parametrizations.weight.original0/1``)."""
    cfg = cfg or LocoSpeechT5Config()
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))

    def normal(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    H, F = cfg.hidden_size, cfg.encoder_ffn_dim
    C = cfg.conv_dim[0]
    # conv feature encoder (kaiming-normal like HF init, modeling_speecht5.py `_init_weights`)
    cin = 1
    for i, (cout, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        sd[f"prenet.feature_encoder.conv_layers.{i}.conv.weight"] = normal(cout, cin, k, std=math.sqrt(2.0 / (cin * k)))
        cin = cout
    sd["prenet.feature_encoder.conv_layers.0.layer_norm.weight"] = normal(C, std=0.1, mean=1.0)
    sd["prenet.feature_encoder.conv_layers.0.layer_norm.bias"] = normal(C, std=0.1)
    sd["prenet.feature_projection.layer_norm.weight"] = normal(C, std=0.1, mean=1.0)
    sd["prenet.feature_projection.layer_norm.bias"] = normal(C, std=0.1)
    sd["prenet.feature_projection.projection.weight"] = normal(H, C, std=1.0 / math.sqrt(C))
    sd["prenet.feature_projection.projection.bias"] = normal(H, std=0.1)
    kp, gp = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    sd["prenet.pos_conv_embed.conv.bias"] = normal(H, std=0.1)
    sd["prenet.pos_conv_embed.conv.parametrizations.weight.original0"] = normal(1, 1, kp, std=0.3, mean=1.0).abs()
    sd["prenet.pos_conv_embed.conv.parametrizations.weight.original1"] = normal(H, H // gp, kp, std=0.02)
    sd["prenet.masked_spec_embed"] = torch.rand(H, generator=g)
    sd["wrapped_encoder.layer_norm.weight"] = normal(H, std=0.1, mean=1.0)
    sd["wrapped_encoder.layer_norm.bias"] = normal(H, std=0.1)
    sd["wrapped_encoder.embed_positions.pe_k.weight"] = normal(2 * cfg.encoder_max_relative_position, H // cfg.encoder_attention_heads, std=0.25)
    for l in range(cfg.encoder_layers):
        p = f"wrapped_encoder.layers.{l}."
        for name, std in (("q_proj", 0.06), ("k_proj", 0.06), ("v_proj", 0.03), ("out_proj", 0.03)):
            sd[p + f"attention.{name}.weight"] = normal(H, H, std=std)
            sd[p + f"attention.{name}.bias"] = normal(H, std=0.1)
        sd[p + "layer_norm.weight"] = normal(H, std=0.1, mean=1.0)
        sd[p + "layer_norm.bias"] = normal(H, std=0.1)
        sd[p + "feed_forward.intermediate_dense.weight"] = normal(F, H, std=0.03)
        sd[p + "feed_forward.intermediate_dense.bias"] = normal(F, std=0.1)
        sd[p + "feed_forward.output_dense.weight"] = normal(H, F, std=0.02)
        sd[p + "feed_forward.output_dense.bias"] = normal(H, std=0.1)
        sd[p + "final_layer_norm.weight"] = normal(H, std=0.1, mean=1.0)
        sd[p + "final_layer_norm.bias"] = normal(H, std=0.1)
    return sd


def synth_head(seed: int = 3, n_classes: int = 101, hidden: int = 768):
    """Random ``Linear(768,101)`` of the downstream IntentClassifier (intent_classifier.py:20-22)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    bound = 1.0 / math.sqrt(hidden)
    w = (torch.rand(n_classes, hidden, generator=g) * 2 - 1) * bound
    b = (torch.rand(n_classes, generator=g) * 2 - 1) * bound
    return w, b


def synth_text_prenet_state_dict(vocab_size: int = 81, hidden_size: int = 768, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init text prenet under the HF ``SpeechT5TextEncoderPrenet`` key names with the ``prenet.`` prefix
    (``embed_tokens.weight``, ``encode_positions.alpha``; the padding row of the embedding is zero like nn.Embedding's)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) + 7919)
    emb = torch.randn(vocab_size, hidden_size, generator=g, dtype=torch.float32)
    emb[1] = 0.0                                   # padding_idx = pad_token_id = 1
    return {"prenet.embed_tokens.weight": emb,
            "prenet.encode_positions.alpha": torch.tensor(1.0 + 0.25 * float(torch.randn((), generator=g)))}
