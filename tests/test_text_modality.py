"""Text modality: the `-m text` branch of the reference's extraction scripts,
``model.speecht5.encoder(texts.input_ids)`` on SpeechT5ForTextToSpeech (extract_speecht5_base_embeddings_slurp.py:79-93)
== HF SpeechT5EncoderWithTextPrenet (modeling_speecht5.py:1377-1415).

CPU: the oracle restatement against the committed outputs of the HF module (tests/golden/text_hf.npz, made by
oracle/make_golden.py) and against the live HF module.  GPU: the CUDA path through the C ABI against the oracle; tolerance
as for the speech path (bf16 operands, fp32 accumulation): pooled cosine >= 0.999, pooled max relative error < 3e-2."""
import os

import numpy as np
import pytest
import torch

from oracle import speecht5_oracle as O
from oracle.make_golden import TEXT_LENGTHS, text_state_dict, text_tokens

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tsd():
    return text_state_dict(0)


def test_text_oracle_matches_golden_hf_vectors(tsd):
    g = np.load(os.path.join(GOLD, "text_hf.npz"))
    toks = text_tokens(0)
    assert list(g["lengths"]) == TEXT_LENGTHS and np.array_equal(np.concatenate(toks), g["tokens"])
    for i, t in enumerate(toks):
        h = O.encode_text(tsd, torch.from_numpy(t))
        assert h.shape == (len(t), 768)
        np.testing.assert_allclose(h.mean(0).numpy(), g["pooled"][i], atol=2e-5)
        np.testing.assert_allclose(h[0].numpy(), g["first_row"][i], atol=5e-5)
        np.testing.assert_allclose(h[-1].numpy(), g["last_row"][i], atol=5e-5)


def test_text_oracle_matches_live_hf_module(tsd):
    from oracle.hf_reference import build_hf_text_encoder, hf_encode_text_unpadded
    model = build_hf_text_encoder(tsd)
    toks = [np.array([5, 9, 33, 2]), np.arange(3, 60)]
    for t, ref in zip(toks, hf_encode_text_unpadded(model, toks)):
        assert float((O.encode_text(tsd, torch.from_numpy(t)) - ref).abs().max()) < 2e-5


def test_scaled_positional_table_is_interleaved_sin_cos():
    pe = O.scaled_positional_rows(5)
    assert torch.allclose(pe[0, 0::2], torch.zeros(384)) and torch.allclose(pe[0, 1::2], torch.ones(384))
    assert abs(float(pe[3, 0]) - np.sin(3.0)) < 1e-6 and abs(float(pe[3, 1]) - np.cos(3.0)) < 1e-6


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def text_encoder(tsd):
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return LocoSpeechT5Encoder.from_state_dict(tsd, device="cuda:0")


@pytest.mark.gpu
def test_text_path_matches_golden_and_oracle(text_encoder, tsd):
    g = np.load(os.path.join(GOLD, "text_hf.npz"))
    toks = text_tokens(0)
    packed = torch.from_numpy(np.concatenate(toks)).to(torch.int32).cuda()
    pooled, hidden, info = text_encoder.encode_text_packed(packed, [len(t) for t in toks], return_hidden=True)
    torch.cuda.synchronize()
    pooled, hidden = pooled.cpu(), hidden.cpu()
    ref = torch.from_numpy(g["pooled"])
    cos = torch.nn.functional.cosine_similarity(pooled, ref, dim=1)
    rel = (pooled - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
    print(f"text: min cosine {float(cos.min()):.6f}, max rel err {float(rel.max()):.5f}")
    assert float(cos.min()) >= 0.999 and float(rel.max()) < 3e-2
    off = 0
    for i, t in enumerate(toks):
        want = O.encode_text(tsd, torch.from_numpy(t))
        got = hidden[off:off + len(t)]
        assert float((got - want).abs().max() / want.abs().max()) < 4e-2, i
        off += len(t)


@pytest.mark.gpu
def test_text_path_with_fused_classifier_head(text_encoder):
    """-m text embeddings feed the same IntentClassifier (train_classifier.py); the fused head serves both modalities."""
    from loco_asr_b200.head import IntentHead
    from loco_asr_b200.synth import synth_head
    w, b = synth_head(11)
    q = torch.randn(1, 768, generator=torch.Generator().manual_seed(11)) * 0.08
    toks = text_tokens(0)
    packed = torch.from_numpy(np.concatenate(toks)).to(torch.int32).cuda()
    for method in ("average", "max", "attention"):
        head = IntentHead(w, b, q, method)
        text_encoder.set_head(head)
        pooled, hidden, info = text_encoder.encode_text_packed(packed, [len(t) for t in toks], return_hidden=True, with_head=True)
        torch.cuda.synchronize()
        want_p = head.pool(hidden.cpu(), info["frames"].tolist())
        want_l = head.logits(want_p)
        got_p, got_l = info["head_pooled"].cpu(), info["logits"].cpu()
        assert torch.allclose(got_p, want_p, rtol=0, atol=2e-5 * float(want_p.abs().max())), method
        assert torch.allclose(got_l, want_l, rtol=0, atol=1e-4 * max(1.0, float(want_l.abs().max()))), method
        assert torch.equal(got_l.argmax(dim=1), want_l.argmax(dim=1))


@pytest.mark.gpu
def test_text_result_does_not_depend_on_workspace_contents(text_encoder):
    """As for the speech path: a NaN-filled workspace changes nothing (slot padding rows are written before they are read)."""
    import ctypes as C
    toks = text_tokens(0)
    nt = np.ascontiguousarray(np.asarray([len(t) for t in toks], dtype=np.int32))
    packed = torch.from_numpy(np.concatenate(toks)).to(torch.int32).cuda()
    want_p, want_h, info = text_encoder.encode_text_packed(packed, nt, return_hidden=True)
    ws = torch.full((info["workspace_bytes"],), 0xFF, dtype=torch.uint8, device="cuda")
    pooled = torch.empty(len(toks), 768, device="cuda")
    hidden = torch.empty(info["total_frames"], 768, device="cuda")
    with torch.cuda.device(text_encoder.device):
        rc = text_encoder._lib.loco_encode_text(text_encoder._h, packed.data_ptr(), nt.ctypes.data, len(toks), pooled.data_ptr(),
                                                hidden.data_ptr(), ws.data_ptr(), ws.numel(),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(pooled, want_p) and torch.equal(hidden, want_h)


@pytest.mark.gpu
def test_text_call_surface_and_errors(text_encoder, tsd):
    """encoder(input_ids) as the reference calls it (no mask: every position is a token), with a mask, and the errors."""
    from loco_asr_b200._lib import LocoError
    ids = torch.tensor([[7, 12, 40, 2, 1, 1], [9, 9, 9, 9, 9, 2]])
    out = text_encoder(ids)
    assert out.last_hidden_state.shape == (2, 6, 768)
    want = O.encode_text(tsd, ids[0])                      # the pad tokens ARE part of the sequence, as in the reference
    assert float((out.last_hidden_state[0].cpu() - want).abs().max() / want.abs().max()) < 4e-2
    masked = text_encoder(ids, attention_mask=torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]]))
    want4 = O.encode_text(tsd, ids[0, :4])
    assert float((masked.last_hidden_state[0, :4].cpu() - want4).abs().max() / want4.abs().max()) < 4e-2
    assert float(masked.last_hidden_state[0, 4:].abs().max()) == 0.0
    with pytest.raises(IndexError):
        text_encoder(torch.tensor([[3, 500]]))
    with pytest.raises(LocoError):                          # a text-only handle has no speech prenet
        text_encoder.encode_packed(torch.zeros(16000, device="cuda"), [16000])


@pytest.mark.gpu
def test_speech_only_handle_rejects_text(encoder):
    from loco_asr_b200._lib import LocoError
    with pytest.raises(LocoError):
        encoder.encode_text_packed(torch.tensor([3, 4, 2], dtype=torch.int32, device="cuda"), [3])


@pytest.mark.gpu
def test_extract_cli_text_modality_writes_reference_format(tmp_path, capsys):
    """`-m text` of the drop-in CLI on synthetic sentences: same folder layout / file names / dict keys the reference's
    text branch writes (:79-93), readable the way slurp_embeddings_and_targets.py:19-28 reads them."""
    import pickle
    from loco_asr_b200 import extract
    extract.main(["-m", "text", "-s", "devel", "--synthetic", "5", "--out-root", str(tmp_path)])
    folder = extract.output_folder(str(tmp_path), "base", "devel", "text")
    files = sorted(os.listdir(folder))
    assert files == [f"synth{i}_embedding_and_target.pickle" for i in range(5)]
    with open(os.path.join(folder, files[0]), "rb") as fh:
        d = pickle.load(fh)
    assert set(d) == {"id", "embedding", "target"} and d["embedding"].dtype == np.float32
    assert d["embedding"].ndim == 2 and d["embedding"].shape[1] == 768 and d["target"].shape == (101,) and d["target"].sum() == 1
    assert extract.main(["-m", "text", "-s", "devel", "--synthetic", "5", "--out-root", str(tmp_path)]) is None   # resume: nothing to do
    assert "already exist" in capsys.readouterr().out
