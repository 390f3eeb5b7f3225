"""Host logic of the streaming extractor: bounded look-ahead batching, rank slicing, format guard (no GPU)."""
import os
import pickle
import threading

import numpy as np
import pytest

from loco_asr_b200 import extract
from loco_asr_b200.buckets import frames_of


def test_stream_batches_cover_every_item_once_within_the_frame_budget():
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 160000, size=300)
    todo = [(f"id{i}", i, "x") for i in range(len(lens))]
    alive, peak, lock = [0], [0], threading.Lock()

    def decode(i):
        with lock:
            alive[0] += 1
            peak[0] = max(peak[0], alive[0])
        return np.zeros(int(lens[i]), np.float32)

    seen = []
    for idx, waves in extract.stream_batches(todo, decode, lambda n: int(frames_of([n])[0]), max_frames=8192, decoders=4):
        assert [len(w) for w in waves] == [int(lens[i]) for i in idx]
        assert sum(int(frames_of([len(w)])[0]) + 2 for w in waves) <= 8192 or len(idx) == 1
        seen += idx
        with lock:
            alive[0] -= len(idx)
    assert seen == list(range(len(lens)))          # arrival order is kept; every utterance exactly once
    assert peak[0] < 200                           # decoded audio alive at once stays bounded (not the whole split)


class _FakeEncoder:
    """Stands in for the CUDA encoder: pooled row = [n_samples, 0, ...]; records what it was asked to encode."""
    device = "cpu"

    def __init__(self):
        self.calls = []

    def encode_host_pipelined(self, batches):
        import torch
        for wave, ns in batches:
            self.calls.append(list(ns))
            out = torch.zeros(len(ns), 768)
            out[:, 0] = torch.tensor([float(n) for n in ns])
            yield out


def test_run_pooled_writes_each_rank_its_share_and_refuses_to_mix(tmp_path):
    lens = [16000 + 320 * i for i in range(10)]
    items = [(f"u{i}", i, "a" if i % 2 else "b") for i in range(10)]
    binarize = extract.make_label_binarizer(["a", "b", "c"])
    folder = str(tmp_path / "out")
    for rank in range(2):
        n = extract.run(_FakeEncoder(), items, lambda i: np.zeros(lens[i], np.float32), binarize, folder, pooled="average",
                        max_frames=200, rank=rank, world=2, log=lambda *a: None)
        assert n == 5
    files = sorted(os.listdir(folder))
    assert len(files) == 10
    for i in (0, 3, 9):
        with open(extract.output_path(folder, f"u{i}"), "rb") as fh:
            d = pickle.load(fh)
        assert d["id"] == f"u{i}" and d["pooling"] == "average" and d["embedding"].shape == (1, 768)
        assert d["embedding"][0, 0] == lens[i] and d["target"].tolist() == ([1, 0, 0] if i % 2 else [0, 1, 0])
    assert extract.run(_FakeEncoder(), items, lambda i: None, binarize, folder, pooled="average", log=lambda *a: None) == 0   # resume
    with pytest.raises(SystemExit):
        extract.run(_FakeEncoder(), items, lambda i: None, binarize, folder, pooled=None, log=lambda *a: None)
