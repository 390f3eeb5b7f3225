"""CPU: the drop-in CLI's host logic -- output file format (what slurp_embeddings_and_targets.py:19-28 reads),
label encoding order (extract...:32-36), resume, dataset index parsing; and the masked head vs the reference's."""
import json
import os
import pickle

import numpy as np
import torch

from loco_asr_b200 import extract
from loco_asr_b200.head import IntentHead


def test_label_binarizer_matches_reference_encoding():
    classes = ["weather_query", "alarm_set", "play_music", "addcontact"]
    f = extract.make_label_binarizer(classes)
    t = f(["alarm_set", "weather_query"])
    assert t.shape == (2, 4) and t.dtype.kind == "i"
    order = sorted(classes)                      # LabelEncoder sorts its classes
    assert t[0].argmax() == order.index("alarm_set") and t[1].argmax() == order.index("weather_query")
    assert t.sum() == 2


def test_written_file_is_what_the_reference_dataset_reads(tmp_path):
    folder = extract.output_folder(str(tmp_path), "base", "test", "audio")
    assert folder.endswith(os.path.join("speecht5_base", "test", "audio"))
    assert extract.output_folder("x", "finetuned", "train", "audio") == os.path.join("x", "speecht5", "train", "audio")
    os.makedirs(folder)
    target = np.eye(101, dtype=np.int64)[7]
    path = extract.write_item(folder, 4242, np.ones((1, 768), np.float64), target)
    assert os.path.basename(path) == "4242_embedding_and_target.pickle"
    # the reference's SLURPEmbeddingsTargets.__getitem__, verbatim semantics
    with open(path, "rb") as fh:
        d = pickle.load(fh)
    slurp_id, emb, tgt = d["id"], torch.from_numpy(d["embedding"]), torch.from_numpy(d["target"])
    assert slurp_id == 4242 and emb.dtype == torch.float32 and emb.shape == (1, 768) and tgt.shape == (101,)
    # train_classifier.py:47-51 collate + IntentClassifier.average / max on a T = 1 "sequence"
    batch = torch.nn.utils.rnn.pad_sequence([emb, emb * 2], batch_first=True)
    assert torch.equal(batch.mean(dim=1, keepdim=True)[0, 0], emb[0]) and batch.max(dim=1, keepdim=True).values.shape == (2, 1, 768)
    assert not [f for f in os.listdir(folder) if f.endswith(".tmp")]


def test_slurp_index_reader(tmp_path):
    d = tmp_path / "slurp" / "dataset" / "slurp"
    d.mkdir(parents=True)
    rows = [{"slurp_id": 1, "sentence": "a", "intent": "alarm_set", "recordings": [{"file": "r1.flac"}, {"file": "r1-headset.flac"}]},
            {"slurp_id": 2, "sentence": "b", "intent": "play_music", "recordings": [{"file": "r2.flac"}]}]
    (d / "test.jsonl").write_text("\n".join(json.dumps(r) for r in rows) + "\n")
    items = extract.read_slurp_index(str(tmp_path / "slurp"), "test")
    assert [i[0] for i in items] == [1, 2] and items[0][1].endswith(os.path.join("audio", "slurp_real", "r1.flac"))
    assert items[1][2] == "play_music"


def test_masked_head_equals_reference_classifier_on_unpadded_input():
    import sys
    torch.manual_seed(0)
    w, b, q = torch.randn(101, 768) * 0.03, torch.randn(101) * 0.1, torch.randn(1, 768) * 0.05
    frames = [5, 12, 1]
    hidden = torch.randn(sum(frames), 768)
    for method in ("average", "max", "attention"):
        head = IntentHead(w, b, q, method)
        got = head.logits(head.pool(hidden, frames))
        off = 0
        for u, t in enumerate(frames):                       # reference forward on one unpadded [1, T, 768] sequence
            x = hidden[off:off + t][None]
            if method == "average":
                pooled = x.mean(dim=1, keepdim=True)
            elif method == "max":
                pooled = x.max(dim=1, keepdim=True).values
            else:
                alpha = torch.softmax(x @ q.t(), dim=1)
                pooled = alpha.permute(0, 2, 1) @ x
            ref = torch.nn.functional.linear(pooled, w, b)[0, 0]
            assert torch.allclose(got[u], ref, atol=1e-5), method
            off += t


def test_weight_file_loaders_round_trip(tmp_path):
    """--weights accepts safetensors (the hub format of microsoft/speecht5_asr), torch files and the reference's pickles."""
    import pickle
    import torch
    from safetensors.torch import save_file
    from loco_asr_b200 import extract
    sd = {"speecht5.encoder.wrapped_encoder.layer_norm.weight": torch.arange(768, dtype=torch.float32),
          "speecht5.decoder.something": torch.ones(3)}
    save_file(sd, str(tmp_path / "model.safetensors"))
    torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    with open(tmp_path / "encoder_state_dict.pickle", "wb") as fh:
        pickle.dump(sd, fh)
    for name in ("model.safetensors", "pytorch_model.bin", "encoder_state_dict.pickle"):
        got = extract.load_weights_file(str(tmp_path / name))
        assert set(got) == set(sd) and torch.equal(got["speecht5.encoder.wrapped_encoder.layer_norm.weight"], sd["speecht5.encoder.wrapped_encoder.layer_norm.weight"])


def test_do_normalize_matches_hf_feature_extractor():
    from transformers import SpeechT5FeatureExtractor
    from loco_asr_b200 import extract
    x = (np.random.default_rng(0).standard_normal(4000) * 0.3 + 0.1).astype(np.float32)
    fe = SpeechT5FeatureExtractor(do_normalize=True)
    ref = fe(audio=[x], sampling_rate=16000, return_tensors="np")["input_values"][0]
    np.testing.assert_allclose(extract.zero_mean_unit_var(x), ref, atol=1e-5)
