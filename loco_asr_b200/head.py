"""Downstream consumer of the embeddings: the pooling + Linear(768, 101) head of
``speech_text/intent_classifier.py:20-50`` (IntentClassifier), restated for MASKED variable-length input.

The reference pools over ``pad_sequence``-zero-padded [B, T_max, 768] batches without a mask
(``train_classifier.py:47-51``, ``intent_classifier.py:24-36``); on an unpadded sequence the two agree, which is
the case this module computes (each utterance over its own frames).  The masked mean is what the encoder's
``pooled`` output already is (fused into the last LayerNorm, csrc/rowops.cu), so ``average`` is just the Linear.
This class is the plain-torch form (used on stored embeddings and as the tests' restatement of the reference forward);
``LocoSpeechT5Encoder.set_head(head)`` + ``encode_packed(..., with_head=True)`` runs the same pooling + Linear inside the
encoder's last kernel (csrc/rowops.cu ``final_ln_pool_kernel<true>``), so the [T, 768] sequence never leaves the GPU.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class IntentHead:
    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, q: Optional[torch.Tensor] = None, method: str = "average"):
        """weight [101, 768], bias [101] = ``classifier.0.*``; q [1, 768] = ``q`` (attention pooling only)."""
        if method not in ("average", "max", "attention"):
            raise ValueError(f"unknown pooling method {method!r} (reference: average / max / self-attention)")
        self.weight, self.bias, self.q, self.method = weight, bias, q, method

    @classmethod
    def from_state_dict(cls, sd, method: str = "average"):
        return cls(sd["classifier.0.weight"], sd["classifier.0.bias"], sd.get("q"), method)

    def to(self, device):
        self.weight, self.bias = self.weight.to(device), self.bias.to(device)
        self.q = self.q.to(device) if self.q is not None else None
        return self

    def pool(self, hidden: torch.Tensor, frames: Sequence[int]) -> torch.Tensor:
        """hidden: compact [sum T, 768] (``encode_packed(..., return_hidden=True)``); returns [B, 768]."""
        out = []
        off = 0
        for t in frames:
            x = hidden[off:off + int(t)]
            if self.method == "average":
                out.append(x.mean(dim=0))
            elif self.method == "max":
                out.append(x.max(dim=0).values)
            else:
                alpha = torch.softmax(x @ self.q.t().to(x.dtype), dim=0)     # [T, 1], softmax over time
                out.append((alpha * x).sum(dim=0))
            off += int(t)
        return torch.stack(out)

    def logits(self, pooled: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.linear(pooled, self.weight.to(pooled.dtype), self.bias.to(pooled.dtype))

    def predict(self, pooled: torch.Tensor) -> torch.Tensor:
        return self.logits(pooled).argmax(dim=-1)
