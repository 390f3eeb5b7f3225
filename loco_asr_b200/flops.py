"""Algorithmic FLOP model of the SpeechT5 speech-encoder forward (SURVEY.md section 8d / BASELINE.md 3).

multiply-add = 2 FLOPs, valid frames only, padding never counted; the relative-position bias is
counted in the table form ``q . pe_k^T`` (R = min(2T-1, 320) columns), not the reference's
materialised ``[T, T, 64]`` contraction.
"""
from __future__ import annotations

from typing import Dict, Iterable

from .config import LocoSpeechT5Config


def encoder_flops_breakdown(n_samples: int, cfg: LocoSpeechT5Config | None = None) -> Dict[str, float]:
    cfg = cfg or LocoSpeechT5Config()
    T = cfg.frame_lengths(n_samples)
    H, F, C = cfg.hidden_size, cfg.encoder_ffn_dim, cfg.conv_dim[0]
    nh = cfg.encoder_attention_heads
    d = H // nh
    t = T[-1]
    R = min(2 * t - 1, 2 * cfg.encoder_max_relative_position) if t > 0 else 0
    out = {
        "conv0": 2.0 * C * cfg.conv_kernel[0] * T[0],
        "conv1_6": sum(2.0 * C * C * cfg.conv_kernel[i] * T[i] for i in range(1, len(T))),
        "proj": 2.0 * C * H * t,
        "pos_conv": 2.0 * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings * t,
        "qkvo": cfg.encoder_layers * 8.0 * H * H * t,
        "ffn": cfg.encoder_layers * 4.0 * H * F * t,
        "attn": cfg.encoder_layers * 4.0 * H * t * t,
        "relpos": cfg.encoder_layers * 2.0 * nh * d * R * t,
    }
    out["total"] = sum(out.values())
    return out


def encoder_flops(n_samples: int, cfg: LocoSpeechT5Config | None = None) -> float:
    return encoder_flops_breakdown(n_samples, cfg)["total"]


def total_flops(lengths: Iterable[int], cfg: LocoSpeechT5Config | None = None) -> float:
    cfg = cfg or LocoSpeechT5Config()
    cache: Dict[int, float] = {}
    tot = 0.0
    for n in lengths:
        n = int(n)
        v = cache.get(n)
        if v is None:
            v = cache[n] = encoder_flops(n, cfg)
        tot += v
    return tot
