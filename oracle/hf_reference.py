"""ORACLE -- TEST INFRASTRUCTURE ONLY: the reference's *own* implementation of the hot path.

LoCo-ASR's encoder arithmetic is the third-party HuggingFace ``transformers`` SpeechT5 module
(reference pin transformers==4.30.2, speech_text/requirements.txt:151; 5.5.0 in this image), reached at
``speech_text/extract_speecht5_base_embeddings_slurp.py:98-108`` as
``SpeechT5ForSpeechToText(...).speecht5.encoder`` == ``SpeechT5EncoderWithSpeechPrenet``
(HF modeling_speecht5.py:1341-1374).  This file imports that module unmodified; it is used to
(1) pin ``oracle/speecht5_oracle.py``, (2) generate ``tests/golden/``, (3) time the reference's CPU
path for ``bench.py``'s ``cpu_baseline`` / ``--impl reference``.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Sequence

import numpy as np
import torch


def build_hf_encoder(state_dict: Dict[str, torch.Tensor] | None = None):
    from transformers import SpeechT5Config
    from transformers.models.speecht5.modeling_speecht5 import SpeechT5EncoderWithSpeechPrenet

    cfg = SpeechT5Config()
    model = SpeechT5EncoderWithSpeechPrenet(cfg).eval()
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        # masked_spec_embed is unused in eval; the sinusoid table is a non-persistent buffer in 5.x
        bad = [k for k in missing if "masked_spec_embed" not in k and "pos_sinusoidal_embed" not in k]
        if bad or unexpected:
            raise RuntimeError(f"state dict mismatch: missing={bad} unexpected={unexpected}")
    return model


def build_hf_text_encoder(state_dict: Dict[str, torch.Tensor]):
    """The reference's text branch: SpeechT5ForTextToSpeech(...).speecht5.encoder == SpeechT5EncoderWithTextPrenet
    (extract_speecht5_base_embeddings_slurp.py:79-88; HF modeling_speecht5.py:1377-1415)."""
    from transformers import SpeechT5Config
    from transformers.models.speecht5.modeling_speecht5 import SpeechT5EncoderWithTextPrenet

    model = SpeechT5EncoderWithTextPrenet(SpeechT5Config()).eval()
    sd = {k: v for k, v in state_dict.items() if k.startswith("wrapped_encoder.") or
          k in ("prenet.embed_tokens.weight", "prenet.encode_positions.alpha")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    bad = [k for k in missing if "encode_positions.pe" not in k]
    if bad or unexpected:
        raise RuntimeError(f"state dict mismatch: missing={bad} unexpected={unexpected}")
    return model


@torch.no_grad()
def hf_encode_text_unpadded(model, token_lists) -> List[torch.Tensor]:
    """One text at a time, no padding (the reference passes no attention mask, so only unpadded calls are well defined)."""
    return [model(torch.as_tensor(t, dtype=torch.long)[None]).last_hidden_state[0] for t in token_lists]


@torch.no_grad()
def hf_encode_unpadded(model, waves: Sequence[np.ndarray]) -> List[torch.Tensor]:
    """One utterance at a time, no padding: the per-utterance ground truth."""
    return [model(input_values=torch.as_tensor(w, dtype=torch.float32)[None]).last_hidden_state[0] for w in waves]


@torch.no_grad()
def hf_stage_taps(model, wave: np.ndarray) -> Dict[str, torch.Tensor]:
    """Stage-by-stage intermediates of the UNMODIFIED HF module for one unpadded utterance, taken with forward hooks (nothing of
    the restatement is involved): the tap names and time-major [T_i, C] layouts of oracle/speecht5_oracle.py.  ``l0_qkv`` is
    [q * 64^-1/2 | k | v] (HF scales q right after q_proj, modeling_speecht5.py:905)."""
    taps: Dict[str, torch.Tensor] = {}
    hooks = []

    def out_hook(name, fn=lambda o: o):
        def h(_m, _i, o):
            taps[name] = fn(o[0] if isinstance(o, tuple) else o).detach().clone()
        return h

    def in_hook(name):
        def h(_m, i):
            taps[name] = i[0][0].detach().clone()
        return h

    pre, enc = model.prenet, model.wrapped_encoder
    for i in (0, 1, 6):
        hooks.append(pre.feature_encoder.conv_layers[i].register_forward_hook(out_hook(f"conv{i}", lambda o: o[0].t())))
    hooks.append(pre.feature_projection.layer_norm.register_forward_hook(out_hook("proj_ln", lambda o: o[0])))
    hooks.append(pre.feature_projection.projection.register_forward_hook(out_hook("proj", lambda o: o[0])))
    hooks.append(pre.pos_conv_embed.register_forward_hook(out_hook("pos_conv", lambda o: o[0])))
    hooks.append(pre.register_forward_hook(out_hook("prenet_out", lambda o: o[0])))
    hooks.append(enc.layer_norm.register_forward_hook(out_hook("enc_in", lambda o: o[0])))
    l0 = enc.layers[0]
    for nm in ("q_proj", "k_proj", "v_proj"):
        hooks.append(getattr(l0.attention, nm).register_forward_hook(out_hook("_" + nm, lambda o: o[0])))
    hooks.append(l0.attention.out_proj.register_forward_pre_hook(in_hook("l0_ctx")))
    hooks.append(l0.layer_norm.register_forward_hook(out_hook("l0_ln1", lambda o: o[0])))
    hooks.append(l0.feed_forward.output_dense.register_forward_pre_hook(in_hook("l0_mid")))
    for l in (0, 5, 11):
        hooks.append(enc.layers[l].register_forward_hook(out_hook(f"layer{l}", lambda o: o[0])))
    try:
        last = model(input_values=torch.as_tensor(wave, dtype=torch.float32)[None]).last_hidden_state[0]
    finally:
        for h in hooks:
            h.remove()
    head_dim = model.config.hidden_size // model.config.encoder_attention_heads
    taps["l0_qkv"] = torch.cat([taps.pop("_q_proj") * head_dim ** -0.5, taps.pop("_k_proj"), taps.pop("_v_proj")], dim=1)
    taps["last_hidden"] = last
    return taps


@torch.no_grad()
def hf_encode_padded_batches(model, waves: Sequence[np.ndarray], batch_size: int = 2):
    """The reference's literal loop: ``batch_size = 2``, ``padding="longest"``, zero padding value and an
    int attention mask (extract_speecht5_base_embeddings_slurp.py:60,67,108;
    HF feature_extraction_speecht5.py:72-90,275-360 with do_normalize=False)."""
    outs = []
    for s in range(0, len(waves), batch_size):
        chunk = waves[s:s + batch_size]
        lmax = max(len(w) for w in chunk)
        iv = torch.zeros(len(chunk), lmax, dtype=torch.float32)
        am = torch.zeros(len(chunk), lmax, dtype=torch.int32)
        for i, w in enumerate(chunk):
            iv[i, :len(w)] = torch.as_tensor(w)
            am[i, :len(w)] = 1
        outs.append(model(input_values=iv, attention_mask=am).last_hidden_state)
    return outs


def time_cpu_reference(model, waves: Sequence[np.ndarray], mode: str = "padded_bs2", repeats: int = 1):
    """Wall-clock audio-seconds/second of the reference CPU path on all host cores."""
    torch.set_num_threads(os.cpu_count() or 1)
    fn = hf_encode_padded_batches if mode == "padded_bs2" else hf_encode_unpadded
    fn(model, waves)  # warm-up over the same shapes (oneDNN builds its primitives per input shape)
    t0 = time.perf_counter()
    for _ in range(repeats):
        fn(model, waves)
    dt = (time.perf_counter() - t0) / repeats
    audio_s = sum(len(w) for w in waves) / 16000.0
    return {"audio_s_per_s": audio_s / dt, "seconds": dt, "audio_s": audio_s,
            "cores": torch.get_num_threads(), "mode": mode}
