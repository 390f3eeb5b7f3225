"""Encoder configuration: the encoder-relevant subset of HF ``SpeechT5Config``.

Field names and defaults follow ``transformers/models/speecht5/configuration_speecht5.py:142-194``
(the config object the reference loads with ``SpeechT5ForSpeechToText.from_pretrained`` at
``speech_text/extract_speecht5_base_embeddings_slurp.py:98``).  The CUDA kernels are built for exactly
one shape family (SpeechT5-base); anything else is rejected loudly -- there is no generic fallback.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import Any, Mapping, Tuple


@dataclass(frozen=True)
class LocoSpeechT5Config:
    hidden_size: int = 768
    encoder_layers: int = 12
    encoder_attention_heads: int = 12
    encoder_ffn_dim: int = 3072
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-5
    feat_extract_norm: str = "group"
    feat_extract_activation: str = "gelu"
    conv_dim: Tuple[int, ...] = (512, 512, 512, 512, 512, 512, 512)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_bias: bool = False
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    max_speech_positions: int = 4000
    encoder_max_relative_position: int = 160
    pad_token_id: int = 1

    # ------------------------------------------------------------------
    @classmethod
    def from_hf(cls, cfg: Any) -> "LocoSpeechT5Config":
        """Accept an HF ``SpeechT5Config`` object, a dict, or a path to ``config.json``."""
        if isinstance(cfg, cls):
            return cfg
        if isinstance(cfg, str):
            with open(cfg, "r") as fh:
                cfg = json.load(fh)
        if not isinstance(cfg, Mapping):
            cfg = cfg.to_dict() if hasattr(cfg, "to_dict") else vars(cfg)
        kw = {}
        for f in cls.__dataclass_fields__:
            if f in cfg and cfg[f] is not None:
                v = cfg[f]
                if isinstance(v, list):
                    v = tuple(v)
                kw[f] = v
        out = cls(**kw)
        out.validate()
        return out

    def to_dict(self):
        d = asdict(self)
        for k, v in d.items():
            if isinstance(v, tuple):
                d[k] = list(v)
        return d

    # ------------------------------------------------------------------
    def validate(self) -> None:
        """Hard-fail on anything but the SpeechT5-base shape family the kernels implement."""
        want = LocoSpeechT5Config()
        problems = []
        for f in ("hidden_size", "encoder_attention_heads", "encoder_ffn_dim", "hidden_act",
                  "feat_extract_norm", "feat_extract_activation", "conv_dim", "conv_stride",
                  "conv_kernel", "conv_bias", "num_conv_pos_embeddings",
                  "num_conv_pos_embedding_groups", "encoder_max_relative_position", "pad_token_id"):
            if getattr(self, f) != getattr(want, f):
                problems.append(f"{f}={getattr(self, f)!r} (kernels are built for {getattr(want, f)!r})")
        if not (1 <= self.encoder_layers <= 48):
            problems.append(f"encoder_layers={self.encoder_layers}")
        if abs(self.layer_norm_eps - 1e-5) > 1e-12:
            problems.append(f"layer_norm_eps={self.layer_norm_eps}")
        if problems:
            raise ValueError("unsupported SpeechT5 encoder config: " + "; ".join(problems))

    # ------------------------------------------------------------------
    def frame_lengths(self, n_samples: int):
        """Frames after each conv layer, ``T_i = floor((T_{i-1} - k_i) / s_i) + 1``
        (HF ``_get_feat_extract_output_lengths``, modeling_speecht5.py:585-598)."""
        out = []
        t = int(n_samples)
        for k, s in zip(self.conv_kernel, self.conv_stride):
            t = (t - k) // s + 1 if t >= k else 0
            out.append(max(t, 0))
        return out

    def num_frames(self, n_samples: int) -> int:
        return self.frame_lengths(n_samples)[-1]

    @property
    def min_samples(self) -> int:
        """Smallest waveform that yields one output frame."""
        n = 1
        for k, s in reversed(list(zip(self.conv_kernel, self.conv_stride))):
            n = (n - 1) * s + k
        return n
