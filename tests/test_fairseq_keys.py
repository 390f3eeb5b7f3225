"""fairseq SpeechT5 checkpoint names -> HF names (loco_asr_b200/fairseq_keys.py, the table form of the reference's
speech_text/map_speecht5_hf.py:34-181).  No fairseq checkpoint exists offline, so a fairseq-style ``ckpt["model"]`` is synthesised
by renaming the seeded HF-style state dict backwards (plus tensors the encoder path must ignore)."""
import pytest
import torch

from loco_asr_b200.fairseq_keys import fairseq_to_hf, hf_to_fairseq_name
from loco_asr_b200.synth import synth_state_dict, synth_text_prenet_state_dict


def fake_fairseq_model(seed=0):
    hf = dict(synth_state_dict(seed=seed))
    hf.update(synth_text_prenet_state_dict(seed=seed))
    model = {}
    for k, v in hf.items():
        fk = hf_to_fairseq_name(k)
        assert fk is not None, k
        model[fk] = v
    # what else such a checkpoint holds and the encoder path never reads (map_speecht5_hf.py's search finds no HF name for them)
    for junk in ("decoder.layers.0.self_attn.q_proj.weight", "encoder.proj.weight", "encoder.version", "speech_decoder_prenet.layers.0.0.weight",
                 "text_decoder_postnet.output_projection.weight", "speech_encoder_prenet.feature_extractor.conv_layers.0.2.running_mean_typo",
                 "encoder.layers.0.self_attn.k_proj.weight_extra", "quantizer.vars"):
        model[junk] = torch.zeros(1)
    return hf, model


def test_rename_table_recovers_the_hf_state_dict():
    hf, model = fake_fairseq_model()
    enc, speech, text = fairseq_to_hf(model)
    got = {"wrapped_encoder." + k: v for k, v in enc.items()}
    got.update({"prenet." + k: v for k, v in speech.items()})
    got.update({"prenet." + k: v for k, v in text.items()})
    spell = {"prenet.pos_conv_embed.conv.weight_g": "prenet.pos_conv_embed.conv.parametrizations.weight.original0",
             "prenet.pos_conv_embed.conv.weight_v": "prenet.pos_conv_embed.conv.parametrizations.weight.original1"}
    got = {spell.get(k, k): v for k, v in got.items()}
    assert set(got) == set(hf)                        # nothing missing, none of the junk tensors let through
    assert all(got[k] is hf[k] for k in hf)           # the same tensors under the HF names
    assert len(enc) == 2 + 1 + 12 * 16 and len(text) == 2


def test_mapped_dicts_load_into_the_reference_modules():
    """The three dicts go where the reference puts them (extract_speecht5_base_embeddings_slurp.py:99-100, :86-88): into the HF
    module's ``wrapped_encoder`` / ``prenet`` sub-modules, old weight-norm spelling included."""
    from oracle.hf_reference import build_hf_encoder, build_hf_text_encoder
    hf, model = fake_fairseq_model()
    enc, speech, text = fairseq_to_hf(model)
    m = build_hf_encoder(None)
    missing, unexpected = m.wrapped_encoder.load_state_dict(enc, strict=False)
    assert not missing and not unexpected
    missing, unexpected = m.prenet.load_state_dict(speech, strict=False)
    assert not unexpected and all("pos_sinusoidal_embed" in k for k in missing)
    ref = build_hf_encoder({k: v for k, v in hf.items() if "embed_tokens" not in k and "encode_positions" not in k})
    x = torch.randn(1, 8000)
    with torch.no_grad():
        assert torch.equal(m(input_values=x).last_hidden_state, ref(input_values=x).last_hidden_state)
    t = build_hf_text_encoder({**{"wrapped_encoder." + k: v for k, v in enc.items()}, **{"prenet." + k: v for k, v in text.items()}})
    assert torch.equal(t.prenet.embed_tokens.weight, hf["prenet.embed_tokens.weight"])


def test_not_a_fairseq_checkpoint():
    with pytest.raises(KeyError):
        fairseq_to_hf({"model.decoder.weight": torch.zeros(1)})
