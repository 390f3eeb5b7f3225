"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under ``loco_asr_b200/`` imports this module.

A CPU (torch fp32, functional) restatement of the algorithm on LoCo-ASR's hot path: the HuggingFace
SpeechT5 speech-encoder forward that ``speech_text/extract_speecht5_base_embeddings_slurp.py:108`` and
``speech_text/extract_speecht5_finetuned_embeddings_slurp.py:104`` call as
``model.speecht5.encoder(**audios)``.

The arithmetic lives in a third-party dependency that is NOT vendored under /root/reference:
``transformers`` (reference pin ``transformers==4.30.2``, speech_text/requirements.txt:151; this image
has 5.5.0).  ``HF:`` citations below are lines of
``transformers/models/speecht5/modeling_speecht5.py`` (5.5.0).

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
section 4 / 8c), so this oracle is pinned against outputs of the reference's own implementation run
here: ``oracle/hf_reference.py`` imports the HF module, ``oracle/make_golden.py`` commits its outputs
under ``tests/golden/`` and ``tests/test_oracle.py`` checks this restatement against both (live HF
module and committed vectors).

Semantics: ONE utterance at a time, no padding (``encoder(input_values=x[None])``).  In the reference's
padded batches an utterance's result depends on its batch-mates (layer-0 GroupNorm statistics run over
the zero padding, HF:277-281); equal-length batches reproduce the single-utterance result bit-exactly,
so the unpadded run is the only well-defined per-utterance answer.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
HIDDEN = 768
HEADS = 12
HEAD_DIM = 64
MAX_REL = 160
POS_K = 128
POS_GROUPS = 16
EPS = 1e-5
PAD_IDX = 1


def _strip(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Normalise key spellings to ``prenet.* / wrapped_encoder.*`` and fold weight-norm spellings."""
    out = {}
    for k, v in sd.items():
        for pre in ("speecht5.encoder.", "encoder."):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v.detach().to(torch.float32)
    return out


def frame_lengths(n_samples: int) -> List[int]:
    """HF:585-598 ``_get_feat_extract_output_lengths``: T_i = floor((T_{i-1} - k_i)/s_i) + 1."""
    out, t = [], int(n_samples)
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        t = (t - k) // s + 1
        out.append(t)
    return out


def pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """HF:355-383: ``weight_norm(conv, name="weight", dim=2)`` => W = g * v / ||v|| with the norm taken
    over dims (0, 1) separately for each of the 128 taps.  Accepts both spellings (4.30.2
    ``weight_g/weight_v``; 5.x ``parametrizations.weight.original0/original1``)."""
    p = "prenet.pos_conv_embed.conv."
    if p + "weight_g" in sd:
        g, v = sd[p + "weight_g"], sd[p + "weight_v"]
    else:
        g, v = sd[p + "parametrizations.weight.original0"], sd[p + "parametrizations.weight.original1"]
    norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return g * v / norm


def sinusoid_rows(positions: torch.Tensor, dim: int = HIDDEN) -> torch.Tensor:
    """HF:305-320 ``get_embedding``: row p = [sin(p*f_0..f_{h-1}), cos(p*f_0..f_{h-1})] (halves
    concatenated), f_i = exp(-i*ln(1e4)/(h-1)), h = dim/2; row ``padding_idx`` is zero."""
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.int64).float() * -(math.log(10000) / (half - 1)))
    ang = positions.to(torch.int64).float().unsqueeze(1) * f.unsqueeze(0)
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    emb[positions == PAD_IDX] = 0
    return emb


def gelu(x):
    return F.gelu(x)  # exact erf form (ACT2FN["gelu"])


def encode_utterance(sd: Dict[str, torch.Tensor], wave: torch.Tensor, n_layers: Optional[int] = None,
                     taps: Optional[dict] = None) -> torch.Tensor:
    """Full encoder forward for one unpadded utterance.  wave: f32[L]  ->  last_hidden_state f32[T, 768].

    If `taps` is a dict it receives every intermediate (time-major ``[T_i, C]`` layouts) for bisecting
    mismatches stage by stage."""
    sd = _strip(sd)
    tap = (lambda k, v: taps.__setitem__(k, v.detach().clone())) if taps is not None else (lambda k, v: None)
    x = wave.detach().to(torch.float32)[None, None, :]                       # [1, 1, L]

    # --- a2: SpeechT5GroupNormConvLayer, HF:260-281 ------------------------------------------------
    p = "prenet.feature_encoder.conv_layers."
    y = F.conv1d(x, sd[p + "0.conv.weight"], stride=CONV_STRIDE[0])           # [1, 512, T0], no bias
    tap("conv0_raw", y[0].t())
    mean = y.mean(dim=2, keepdim=True)                                        # GroupNorm(512 groups): per channel
    var = y.var(dim=2, unbiased=False, keepdim=True)                          # over time, biased variance
    y = (y - mean) / torch.sqrt(var + EPS)
    y = y * sd[p + "0.layer_norm.weight"][None, :, None] + sd[p + "0.layer_norm.bias"][None, :, None]
    y = gelu(y)
    tap("conv0", y[0].t())
    # --- a3: SpeechT5NoLayerNormConvLayer x6, HF:210-228 ------------------------------------------
    for i in range(1, 7):
        y = gelu(F.conv1d(y, sd[p + f"{i}.conv.weight"], stride=CONV_STRIDE[i]))
        tap(f"conv{i}", y[0].t())
    feats = y[0].t()                                                          # [T, 512] time-major (HF:544)
    T = feats.shape[0]

    # --- a4: SpeechT5FeatureProjection, HF:498-510 ------------------------------------------------
    q = "prenet.feature_projection."
    normed = F.layer_norm(feats, (feats.shape[1],), sd[q + "layer_norm.weight"], sd[q + "layer_norm.bias"], EPS)
    tap("proj_ln", normed)
    h = F.linear(normed, sd[q + "projection.weight"], sd[q + "projection.bias"])
    tap("proj", h)

    # --- a7: SpeechT5PositionalConvEmbedding + SamePad, HF:355-397, 445-453 ------------------------
    w = pos_conv_weight(sd)
    pc = F.conv1d(h.t()[None], w, sd["prenet.pos_conv_embed.conv.bias"], padding=POS_K // 2, groups=POS_GROUPS)
    pc = gelu(pc[:, :, :-1])[0].t()                                           # drop last frame (even kernel)
    tap("pos_conv", pc)
    h = h + pc                                                                # HF:555-556

    # --- a8: sinusoidal positions, HF:285-351; valid frame t uses row t + padding_idx + 1 ----------
    h = h + sinusoid_rows(torch.arange(T) + PAD_IDX + 1)
    tap("prenet_out", h)

    return _wrapped_encoder(sd, h, n_layers, tap)


def _wrapped_encoder(sd, h: torch.Tensor, n_layers, tap) -> torch.Tensor:
    """SpeechT5Encoder.forward (HF:1250-1338) on one unpadded sequence h f32[T, 768] -- shared by the speech and text paths."""
    T = h.shape[0]
    # --- a10: encoder input LayerNorm, HF:1292 ----------------------------------------------------
    e = "wrapped_encoder."
    h = F.layer_norm(h, (HIDDEN,), sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"], EPS)
    tap("enc_in", h)

    # --- a11: relative positions, HF:425-441 ------------------------------------------------------
    pe_k = sd[e + "embed_positions.pe_k.weight"]                              # [320, 64]
    pos = torch.arange(T)
    rel = (pos[:, None] - pos[None, :]).clamp(-MAX_REL, MAX_REL - 1) + MAX_REL  # [T, T]

    n_layers = n_layers if n_layers is not None else 1 + max(
        int(k.split(".")[2]) for k in sd if k.startswith(e + "layers."))
    for l in range(n_layers):
        lp = e + f"layers.{l}."
        # --- a12: SpeechT5Attention, HF:872-986 ---------------------------------------------------
        qh = F.linear(h, sd[lp + "attention.q_proj.weight"], sd[lp + "attention.q_proj.bias"]) * HEAD_DIM ** -0.5
        kh = F.linear(h, sd[lp + "attention.k_proj.weight"], sd[lp + "attention.k_proj.bias"])
        vh = F.linear(h, sd[lp + "attention.v_proj.weight"], sd[lp + "attention.v_proj.bias"])
        if l == 0:
            tap("l0_qkv", torch.cat([qh, kh, vh], dim=1))
        qh = qh.view(T, HEADS, HEAD_DIM).transpose(0, 1)                      # [12, T, 64]
        kh = kh.view(T, HEADS, HEAD_DIM).transpose(0, 1)
        vh = vh.view(T, HEADS, HEAD_DIM).transpose(0, 1)
        scores = qh @ kh.transpose(1, 2)                                      # HF:930
        qt = qh @ pe_k.t()                                                    # [12, T, 320]  (table form of HF:939-945)
        scores = scores + torch.gather(qt, 2, rel[None].expand(HEADS, T, T))
        probs = torch.softmax(scores, dim=-1)                                 # HF:955
        ctx = (probs @ vh).transpose(0, 1).reshape(T, HIDDEN)                 # HF:969-980
        if l == 0:
            tap("l0_ctx", ctx)
        attn = F.linear(ctx, sd[lp + "attention.out_proj.weight"], sd[lp + "attention.out_proj.bias"])
        # --- a14: post-LN residual blocks, HF:1047-1060 -------------------------------------------
        h = F.layer_norm(h + attn, (HIDDEN,), sd[lp + "layer_norm.weight"], sd[lp + "layer_norm.bias"], EPS)
        if l == 0:
            tap("l0_ln1", h)
        # --- a13: SpeechT5FeedForward, HF:989-1010 ------------------------------------------------
        mid = gelu(F.linear(h, sd[lp + "feed_forward.intermediate_dense.weight"], sd[lp + "feed_forward.intermediate_dense.bias"]))
        if l == 0:
            tap("l0_mid", mid)
        ff = F.linear(mid, sd[lp + "feed_forward.output_dense.weight"], sd[lp + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h + ff, (HIDDEN,), sd[lp + "final_layer_norm.weight"], sd[lp + "final_layer_norm.bias"], EPS)
        tap(f"layer{l}", h)
    return h




def scaled_positional_rows(n: int, dim: int = HIDDEN) -> torch.Tensor:
    """SpeechT5ScaledPositionalEncoding table rows 0..n-1 (HF:405-411): sin on even, cos on odd columns."""
    position = torch.arange(0, n).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.int64).float() * -(math.log(10000.0) / dim))
    pe = torch.zeros(n, dim)
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe


def encode_text(sd: Dict[str, torch.Tensor], token_ids: torch.Tensor, n_layers: Optional[int] = None,
                taps: Optional[dict] = None) -> torch.Tensor:
    """Text-modality encoder for one unpadded token sequence: i64[T] -> last_hidden_state f32[T, 768].
    SpeechT5EncoderWithTextPrenet (HF:1377-1415): SpeechT5TextEncoderPrenet (embed_tokens, HF:776-779; emb + alpha * pe[:T],
    HF:418-421) then the wrapped encoder; the reference calls it at extract_speecht5_base_embeddings_slurp.py:88."""
    sd = _strip(sd)
    tap = (lambda k, v: taps.__setitem__(k, v.detach().clone())) if taps is not None else (lambda k, v: None)
    emb = sd["prenet.embed_tokens.weight"][token_ids.long()]
    h = emb + sd["prenet.encode_positions.alpha"] * scaled_positional_rows(emb.shape[0])
    tap("prenet_out", h)
    return _wrapped_encoder(sd, h, n_layers, tap)


def position_bias_reference_form(sd, qh: torch.Tensor) -> torch.Tensor:
    """The reference's materialised form of the bias (HF:432-441 + HF:939-945): gather pe_k into
    [T, T, 64], then contract with q.  Used by the tests to prove the table form above is identical."""
    sd = _strip(sd)
    pe_k = sd["wrapped_encoder.embed_positions.pe_k.weight"]
    T = qh.shape[1]
    pos = torch.arange(T)
    rel = (pos[:, None] - pos[None, :]).clamp(-MAX_REL, MAX_REL - 1) + MAX_REL
    position_bias = pe_k[rel]                                                 # [T, T, 64]
    reshape_q = qh.transpose(0, 1)                                            # [T, 12, 64]
    rel_pos_bias = torch.matmul(reshape_q, position_bias.transpose(-2, -1))  # [T, 12, T]
    return rel_pos_bias.transpose(0, 1)


def pooled(sd, wave: torch.Tensor) -> torch.Tensor:
    """Mean over the utterance's own frames (what ``IntentClassifier.average`` computes on an
    unpadded sequence, intent_classifier.py:24-26)."""
    return encode_utterance(sd, wave).mean(dim=0)


def encode_batch(sd, waves) -> List[torch.Tensor]:
    return [encode_utterance(sd, torch.as_tensor(w)) for w in waves]


def intent_head(pooled_emb: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``IntentClassifier(method="average")`` tail: Linear(768, 101) then argmax
    (intent_classifier.py:20-22, 38-50; train_classifier.py:109,119)."""
    return F.linear(pooled_emb, w, b).argmax(dim=-1)
