"""fairseq SpeechT5 checkpoint keys -> the HF / map_speecht5_hf.py key names the encoder loads.

The reference converts Microsoft's fairseq SpeechT5 checkpoint (``ckpt["model"]``) once, offline, with
``speech_text/map_speecht5_hf.py`` (class ``Mapping``): it instantiates both HF models, searches their parameter names for
every fairseq key and pickles three state dicts -- ``encoder_state_dict`` (:34-105), ``speech_prenet_state_dict`` (:107-168) and
``text_prenet_state_dict`` (:170-181) -- which the extraction scripts later feed to ``encoder.wrapped_encoder.load_state_dict``
and ``encoder.prenet.load_state_dict`` (extract_speecht5_base_embeddings_slurp.py:99-100).

This module produces the same three dicts from the same ``ckpt["model"]`` with a rename table instead of a search over
instantiated models (no HF model is needed), so a fairseq checkpoint can be handed to the B200 encoder directly::

    enc_sd, speech_sd, text_sd = fairseq_to_hf(torch.load(path)["model"])
    encoder.wrapped_encoder.load_state_dict(enc_sd); encoder.prenet.load_state_dict(speech_sd)

Rules (fairseq name -> HF name, the pairs ``Mapping.search_mapping`` / ``map_speech_prenet`` end up with):
  encoder.pos_emb.pe_k.weight                       embed_positions.pe_k.weight
  encoder.layer_norm.{weight,bias}                  layer_norm.{weight,bias}
  encoder.layers.N.self_attn.{q,k,v,out}_proj.*     layers.N.attention.{q,k,v,out}_proj.*
  encoder.layers.N.self_attn_layer_norm.*           layers.N.layer_norm.*
  encoder.layers.N.fc1.* / fc2.*                    layers.N.feed_forward.intermediate_dense.* / output_dense.*
  encoder.layers.N.final_layer_norm.*               layers.N.final_layer_norm.*
  speech_encoder_prenet.mask_emb                    masked_spec_embed
  speech_encoder_prenet.layer_norm.*                feature_projection.layer_norm.*
  speech_encoder_prenet.post_extract_proj.*         feature_projection.projection.*
  speech_encoder_prenet.feature_extractor.conv_layers.I.0.weight      feature_encoder.conv_layers.I.conv.weight
  speech_encoder_prenet.feature_extractor.conv_layers.I.2.{w,b}       feature_encoder.conv_layers.I.layer_norm.{w,b}
  speech_encoder_prenet.pos_conv.0.{bias,weight_g,weight_v}           pos_conv_embed.conv.{bias,weight_g,weight_v}
  text_encoder_prenet.encoder_prenet.0.weight       embed_tokens.weight
  text_encoder_prenet.encoder_prenet.1.alpha        encode_positions.alpha   (the reference leaves HF's own alpha in place, :176-177;
                                                    taking the checkpoint's is the faithful conversion and is what is done here)
Keys the encoder path never reads (decoder, post-nets, quantizer, ``encoder.proj``, CTC head, ``num_updates`` ...) are dropped,
as the reference's search drops them.  The sinusoid / scaled-positional tables (``pos_sinusoidal_embed.weights``,
``encode_positions.pe``) are buffers the reference copies from the HF instance (:163-165, :179); the library builds them itself.
"""
from __future__ import annotations

import re
from typing import Dict, Mapping, Tuple

_ENCODER_RULES = [
    (r"^encoder\.pos_emb\.pe_k\.weight$", r"embed_positions.pe_k.weight"),
    (r"^encoder\.layer_norm\.(weight|bias)$", r"layer_norm.\1"),
    (r"^encoder\.layers\.(\d+)\.self_attn\.(q_proj|k_proj|v_proj|out_proj)\.(weight|bias)$", r"layers.\1.attention.\2.\3"),
    (r"^encoder\.layers\.(\d+)\.self_attn_layer_norm\.(weight|bias)$", r"layers.\1.layer_norm.\2"),
    (r"^encoder\.layers\.(\d+)\.fc1\.(weight|bias)$", r"layers.\1.feed_forward.intermediate_dense.\2"),
    (r"^encoder\.layers\.(\d+)\.fc2\.(weight|bias)$", r"layers.\1.feed_forward.output_dense.\2"),
    (r"^encoder\.layers\.(\d+)\.final_layer_norm\.(weight|bias)$", r"layers.\1.final_layer_norm.\2"),
]
_SPEECH_RULES = [
    (r"^speech_encoder_prenet\.mask_emb$", r"masked_spec_embed"),
    (r"^speech_encoder_prenet\.layer_norm\.(weight|bias)$", r"feature_projection.layer_norm.\1"),
    (r"^speech_encoder_prenet\.post_extract_proj\.(weight|bias)$", r"feature_projection.projection.\1"),
    (r"^speech_encoder_prenet\.feature_extractor\.conv_layers\.(\d+)\.0\.weight$", r"feature_encoder.conv_layers.\1.conv.weight"),
    (r"^speech_encoder_prenet\.feature_extractor\.conv_layers\.(\d+)\.2\.(weight|bias)$", r"feature_encoder.conv_layers.\1.layer_norm.\2"),
    (r"^speech_encoder_prenet\.pos_conv\.0\.(bias|weight_g|weight_v)$", r"pos_conv_embed.conv.\1"),
]
_TEXT_RULES = [
    (r"^text_encoder_prenet\.encoder_prenet\.0\.weight$", r"embed_tokens.weight"),
    (r"^text_encoder_prenet\.encoder_prenet\.1\.alpha$", r"encode_positions.alpha"),
]


def _apply(rules, model: Mapping[str, object]) -> Dict[str, object]:
    out: Dict[str, object] = {}
    for name, value in model.items():
        for pat, repl in rules:
            new, n = re.subn(pat, repl, name)
            if n:
                out[new] = value
                break
    return out


def fairseq_to_hf(model: Mapping[str, object]) -> Tuple[Dict[str, object], Dict[str, object], Dict[str, object]]:
    """``ckpt["model"]`` of a fairseq SpeechT5 checkpoint -> (encoder_state_dict, speech_prenet_state_dict,
    text_prenet_state_dict) with the key names of map_speecht5_hf.py's pickles (no ``wrapped_encoder.`` / ``prenet.`` prefix:
    they go to the two ``load_state_dict`` shims).  Raises if the checkpoint holds no encoder at all."""
    enc, speech, text = _apply(_ENCODER_RULES, model), _apply(_SPEECH_RULES, model), _apply(_TEXT_RULES, model)
    if not enc:
        raise KeyError("no 'encoder.*' tensors found: not a fairseq SpeechT5 checkpoint's ckpt['model']")
    return enc, speech, text


def hf_to_fairseq_name(prefixed_hf_key: str) -> str | None:
    """Inverse of the rename table for one key of the full HF encoder state dict (``wrapped_encoder.*`` / ``prenet.*``);
    used by the tests to synthesise a fairseq-style checkpoint.  None for keys fairseq has no counterpart of."""
    inv = [
        (r"^wrapped_encoder\.embed_positions\.pe_k\.weight$", r"encoder.pos_emb.pe_k.weight"),
        (r"^wrapped_encoder\.layer_norm\.(weight|bias)$", r"encoder.layer_norm.\1"),
        (r"^wrapped_encoder\.layers\.(\d+)\.attention\.(q_proj|k_proj|v_proj|out_proj)\.(weight|bias)$", r"encoder.layers.\1.self_attn.\2.\3"),
        (r"^wrapped_encoder\.layers\.(\d+)\.layer_norm\.(weight|bias)$", r"encoder.layers.\1.self_attn_layer_norm.\2"),
        (r"^wrapped_encoder\.layers\.(\d+)\.feed_forward\.intermediate_dense\.(weight|bias)$", r"encoder.layers.\1.fc1.\2"),
        (r"^wrapped_encoder\.layers\.(\d+)\.feed_forward\.output_dense\.(weight|bias)$", r"encoder.layers.\1.fc2.\2"),
        (r"^wrapped_encoder\.layers\.(\d+)\.final_layer_norm\.(weight|bias)$", r"encoder.layers.\1.final_layer_norm.\2"),
        (r"^prenet\.masked_spec_embed$", r"speech_encoder_prenet.mask_emb"),
        (r"^prenet\.feature_projection\.layer_norm\.(weight|bias)$", r"speech_encoder_prenet.layer_norm.\1"),
        (r"^prenet\.feature_projection\.projection\.(weight|bias)$", r"speech_encoder_prenet.post_extract_proj.\1"),
        (r"^prenet\.feature_encoder\.conv_layers\.(\d+)\.conv\.weight$", r"speech_encoder_prenet.feature_extractor.conv_layers.\1.0.weight"),
        (r"^prenet\.feature_encoder\.conv_layers\.(\d+)\.layer_norm\.(weight|bias)$", r"speech_encoder_prenet.feature_extractor.conv_layers.\1.2.\2"),
        (r"^prenet\.pos_conv_embed\.conv\.bias$", r"speech_encoder_prenet.pos_conv.0.bias"),
        (r"^prenet\.pos_conv_embed\.conv\.(?:weight_g|parametrizations\.weight\.original0)$", r"speech_encoder_prenet.pos_conv.0.weight_g"),
        (r"^prenet\.pos_conv_embed\.conv\.(?:weight_v|parametrizations\.weight\.original1)$", r"speech_encoder_prenet.pos_conv.0.weight_v"),
        (r"^prenet\.embed_tokens\.weight$", r"text_encoder_prenet.encoder_prenet.0.weight"),
        (r"^prenet\.encode_positions\.alpha$", r"text_encoder_prenet.encoder_prenet.1.alpha"),
    ]
    for pat, repl in inv:
        new, n = re.subn(pat, repl, prefixed_hf_key)
        if n:
            return new
    return None
