"""CPU: pin the oracle restatement (oracle/speecht5_oracle.py) against the reference's own implementation --
the committed golden vectors produced by the HF module (tests/golden, oracle/make_golden.py) and the live HF
module imported here."""
import os

import numpy as np
import pytest
import torch

from loco_asr_b200.synth import synth_state_dict, synth_wave, config1_lengths, synth_head
from oracle import speecht5_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sd():
    return synth_state_dict(seed=0)


def test_frame_length_formula():
    # 1 s -> 49, 3 s -> 149, 10 s -> 499, 30 s -> 1499, 60 s -> 2999 frames
    for sec, frames in ((1, 49), (3, 149), (10, 499), (30, 1499), (60, 2999)):
        assert O.frame_lengths(16000 * sec)[-1] == frames
    assert O.frame_lengths(48000) == [9599, 4799, 2399, 1199, 599, 299, 149]
    assert O.frame_lengths(400)[-1] == 1 and O.frame_lengths(399)[-1] == 0


def test_oracle_matches_golden_config1(sd):
    """BASELINE.json configs[0]: pooled embeddings of the 16 utterances as the HF module produced them."""
    g = np.load(os.path.join(GOLD, "config1_hf.npz"))
    lengths = config1_lengths()
    assert list(g["lengths"]) == lengths
    for i in (0, 5, 11, 15):
        h = O.encode_utterance(sd, torch.from_numpy(synth_wave(lengths[i], 0, i)))
        assert h.shape[0] == int(g["n_frames"][i])
        np.testing.assert_allclose(h.mean(0).numpy(), g["pooled"][i], atol=2e-5)
        np.testing.assert_allclose(h[0].numpy(), g["first_frame"][i], atol=5e-5)
        np.testing.assert_allclose(h[-1].numpy(), g["last_frame"][i], atol=5e-5)


def test_oracle_taps_match_golden(sd):
    g = np.load(os.path.join(GOLD, "short_taps.npz"))
    for name in ("a", "b"):
        w = synth_wave(int(g[f"{name}_n_samples"]), 0, int(g[f"{name}_idx"]), kind="noise" if name == "a" else "mix")
        taps = {}
        h = O.encode_utterance(sd, torch.from_numpy(w), taps=taps)
        np.testing.assert_allclose(h.numpy(), g[f"{name}_hf_last_hidden"], atol=5e-5)
        for k in g.files:
            if k.startswith(name + "_") and k[2:] in taps:
                ref = g[k]
                np.testing.assert_allclose(taps[k[2:]].numpy()[:ref.shape[0]], ref, atol=5e-5, err_msg=k)


def test_oracle_matches_live_hf_module(sd):
    from oracle.hf_reference import build_hf_encoder, hf_encode_unpadded
    model = build_hf_encoder(sd)
    for n, idx in ((16000, 7), (23456, 8)):
        w = synth_wave(n, 3, idx)
        ref = hf_encode_unpadded(model, [w])[0]
        mine = O.encode_utterance(sd, torch.from_numpy(w))
        assert float((ref - mine).abs().max()) < 2e-5


def test_relpos_table_form_equals_reference_form(sd):
    """q . pe_k[clip(i-j)+160]^T gathered from the [T,320] table == the reference's [T,T,64] contraction."""
    torch.manual_seed(0)
    T = 200  # > 160 so both clip boundaries are exercised
    qh = torch.randn(12, T, 64)
    pe_k = O._strip(sd)["wrapped_encoder.embed_positions.pe_k.weight"]
    pos = torch.arange(T)
    rel = (pos[:, None] - pos[None, :]).clamp(-160, 159) + 160
    table = torch.gather(qh @ pe_k.t(), 2, rel[None].expand(12, T, T))
    ref = O.position_bias_reference_form(sd, qh)
    assert float((table - ref).abs().max()) < 1e-4


def test_padded_batch_leaks_padding_into_short_utterance(sd):
    """Why the oracle is the unpadded run: the reference's own padded batches change a short utterance's result
    (GroupNorm statistics over the zero padding), while equal-length batches are exact."""
    g = np.load(os.path.join(GOLD, "config1_hf.npz"))
    cos = torch.nn.functional.cosine_similarity(torch.from_numpy(g["pooled"]), torch.from_numpy(g["pooled_padded_bs2"]), dim=1)
    assert float(cos.min()) < 0.99999          # some utterance is perturbed by its longer batch-mate
    assert float(cos[1::2].min()) > 0.99999    # the longer one of each pair is not padded -> unchanged


def test_intent_head_argmax_is_stable_under_bf16_noise(sd):
    g = np.load(os.path.join(GOLD, "config1_hf.npz"))
    w, b = synth_head(3)
    p = torch.from_numpy(g["pooled"])
    a = O.intent_head(p, w, b)
    assert a.shape == (16,)
    assert torch.equal(a, O.intent_head(p.bfloat16().float(), w, b))


def test_restatement_matches_config5_golden_subset():
    """Two of the 512 config-5 utterances (the shortest ones, to stay within seconds): the functional restatement with the
    seed-1 weights against the HF pooled embeddings in tests/golden/config5_hf.npz (fp32)."""
    g = np.load(os.path.join(GOLD, "config5_hf.npz"))
    sd = synth_state_dict(seed=1)
    for i in np.argsort(g["n_samples"], kind="stable")[:2]:
        x = torch.from_numpy(synth_wave(int(g["n_samples"][i]), 1234, int(g["ids"][i])))
        got = O.encode_utterance(sd, x).mean(0)
        ref = torch.from_numpy(g["pooled"][i])
        assert float((got - ref).abs().max() / ref.abs().max()) < 5e-5
        logit = torch.nn.functional.linear(got, *synth_head(3))
        assert float((logit - torch.from_numpy(g["logits"][i])).abs().max()) < 1e-4


def test_restatement_matches_hf_on_a_60s_segment():
    """T = 2999: both clamp regions of the relative-position table are in play.  The restatement's table form
    (q . pe_k^T, gathered) against the HF module's [T, T, 64] contraction, as stored in tests/golden/long60_hf.npz."""
    g = np.load(os.path.join(GOLD, "long60_hf.npz"))
    w = synth_wave(int(g["n_samples"]), int(g["wave_seed"]), int(g["wave_idx"]))
    h = O.encode_utterance(synth_state_dict(seed=0), torch.from_numpy(w))
    assert h.shape[0] == int(g["n_frames"]) == 2999
    ref = torch.from_numpy(g["pooled"])
    assert float((h.mean(0) - ref).abs().max() / ref.abs().max()) < 5e-5
    assert float((h[torch.from_numpy(g["rows"])] - torch.from_numpy(g["hidden_rows"])).abs().max()) < 5e-4
