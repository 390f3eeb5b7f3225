"""Drop-in for ``speech_text/extract_speecht5_{base,finetuned}_embeddings_slurp.py`` (audio and text modalities).

Same CLI (``-m {audio,text} -s {train,devel,test,train_synthetic}``), same output files
(``extracted/{speecht5|speecht5_base}/<split>/<modality>/<slurp_id>_embedding_and_target.pickle`` holding
``{"id", "embedding": np.float32[T, 768], "target": one-hot[101]}``, reference :70-77,:111-113) that
``slurp_embeddings_and_targets.py:19-28`` and ``train_classifier.py`` read back -- but the loop is restructured
for throughput: waveforms are decoded by worker threads while the GPU works (nothing is materialised up front), cut into
padding-free batches of <= 131072 frames, encoded by the CUDA library, and written by a pool of writer threads;
existing files are skipped (resume).  Under ``torchrun`` every rank takes every world-th utterance and writes its own files.

``embedding`` is the reference's [T, 768] ``last_hidden_state`` (the utterance's own T frames) by default, so every pooling
of ``train_classifier.py`` (average / max / self-attention over frames) sees what it expects.  ``--pooled average|max``
writes that pooling's [1, 768] vector instead (3 KB instead of ~460 KB per utterance, and the encoder's fused pooling never
sends the sequence to HBM); such files carry a ``"pooling"`` key and a folder never mixes formats.  A classifier trained on
pooled files must use the same ``--pooling``: pooling a single already-pooled row again is the identity.

Run from the reference's ``speech_text/`` directory (so ``slurp_data`` / ``intent_classes`` import), or pass
``--classes-file`` and ``--synthetic N`` to exercise the pipeline without the SLURP corpus (not available offline).
``-m text`` (reference :79-93) runs the sentences through the text prenet + the same encoder (``loco_encode_text``); it needs
the SpeechT5 tokenizer (``--tokenizer``: a local ``SpeechT5Processor`` / ``SpeechT5Tokenizer`` directory, default the hub name
the reference uses) and ``text_prenet_state_dict.pickle`` in the mapping folder.  Each sentence is encoded alone and
unpadded; the reference's ``padding="longest"`` batches pass no attention mask, so its padded rows see the pad tokens.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

SAMPLE_RATE = 16000


# ------------------------------------------------------------------------------------------------ labels
def make_label_binarizer(classes: Sequence[str]):
    """The reference's target encoding (extract...:32-36): LabelEncoder over ALL_CLASSES (sorted), then a
    LabelBinarizer over the numeric labels; ``target`` = one-hot int64[len(classes)]."""
    from sklearn.preprocessing import LabelBinarizer, LabelEncoder
    enc = LabelEncoder()
    numeric = enc.fit_transform(list(classes))
    binar = LabelBinarizer()
    binar.fit(numeric)
    return lambda intents: binar.transform(enc.transform(list(intents)))


def load_classes(path: Optional[str]) -> List[str]:
    if path:
        with open(path) as fh:
            txt = fh.read().strip()
        return json.loads(txt) if txt.startswith("[") else [l.strip() for l in txt.splitlines() if l.strip()]
    try:
        from intent_classes import ALL_CLASSES       # reference speech_text/intent_classes.py:1 (101 labels)
        return list(ALL_CLASSES)
    except ImportError as e:
        raise SystemExit("intent label inventory not found: run from the reference's speech_text/ directory "
                         "(intent_classes.py) or pass --classes-file") from e


# ------------------------------------------------------------------------------------------------ data
def read_slurp_index(data_path: str, split: str, modality: str = "audio"):
    """(slurp_id, audio_path | sentence, intent) per utterance; mirrors SLURPDataset.prepare_data (slurp_data.py:19-53),
    including its recording choice (always ``recordings[0]`` -- the "headset" test there inspects dict keys)."""
    text_file = os.path.join(data_path, "dataset", "slurp", split + ".jsonl")
    audio_dir = os.path.join(data_path, "audio", "slurp_synth" if split == "train_synthetic" else "slurp_real")
    items = []
    with open(text_file) as fh:
        for line in fh:
            line = line.strip()
            if line:
                it = json.loads(line)
                key = it["sentence"] if modality == "text" else os.path.join(audio_dir, it["recordings"][0]["file"])
                items.append((it["slurp_id"], key, it["intent"]))
    return items


def load_weights_file(path: str):
    """HF-keyed state dict from a checkpoint file: ``model.safetensors`` (what ``microsoft/speecht5_asr`` ships), a torch
    ``.bin`` / ``.pt`` file, or one of the reference's pickles (map_speecht5_hf.py output).  Decoder / head tensors of a
    full-model checkpoint are skipped later by the encoder's loader."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    if path.endswith(".pickle") or path.endswith(".pkl"):
        with open(path, "rb") as fh:
            return pickle.load(fh)
    import torch
    sd = torch.load(path, map_location="cpu")
    return sd.get("state_dict", sd) if isinstance(sd, dict) else sd


def zero_mean_unit_var(x: np.ndarray) -> np.ndarray:
    """``do_normalize=True`` of SpeechT5FeatureExtractor (HF feature_extraction_speecht5.py:119-137):
    (x - mean) / sqrt(var + 1e-7) over the utterance's own samples.  The processor's default is False; the encoder output
    barely depends on it (conv0 has no bias and its GroupNorm renormalises every channel over time), but it is the
    reference's preprocessing switch, so it is offered."""
    x = np.asarray(x, dtype=np.float32)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def load_audio(path: str) -> np.ndarray:
    """Decode + resample to 16 kHz mono float32 (reference: ``librosa.load(path, sr=16000)``, :55-57)."""
    try:
        import librosa
        return librosa.load(path, sr=SAMPLE_RATE)[0].astype(np.float32)
    except ImportError:
        pass
    try:
        import soundfile as sf
        x, sr = sf.read(path, dtype="float32", always_2d=True)
        x = x.mean(axis=1)
    except ImportError:
        import wave
        with wave.open(path, "rb") as w:
            sr, n, ch, width = w.getframerate(), w.getnframes(), w.getnchannels(), w.getsampwidth()
            if width != 2:
                raise SystemExit("only 16-bit PCM WAV can be decoded without librosa/soundfile: " + path)
            x = np.frombuffer(w.readframes(n), dtype=np.int16).reshape(-1, ch).mean(axis=1).astype(np.float32) / 32768.0
    if sr != SAMPLE_RATE:
        from scipy.signal import resample_poly
        from math import gcd
        g = gcd(int(sr), SAMPLE_RATE)
        x = resample_poly(x, SAMPLE_RATE // g, int(sr) // g).astype(np.float32)
    return x


# ------------------------------------------------------------------------------------------------ writer
def output_folder(root: str, version: str, split: str, modality: str) -> str:
    return os.path.join(root, "speecht5" if version == "finetuned" else "speecht5_base", split, modality)


def output_path(folder: str, slurp_id) -> str:
    return os.path.join(folder, f"{slurp_id}_embedding_and_target.pickle")


def write_item(folder: str, slurp_id, embedding: np.ndarray, target: np.ndarray, pooling: Optional[str] = None) -> str:
    """One file per utterance, the reference's dict (:111-113).  A pooled file also says which pooling it holds
    (``"pooling"``; the reference's reader ignores extra keys) so that a classifier is never trained on a mix."""
    path = output_path(folder, slurp_id)
    tmp = path + ".tmp"
    item = {"id": slurp_id, "embedding": np.ascontiguousarray(embedding, dtype=np.float32), "target": target}
    if pooling:
        item["pooling"] = pooling
    blob = pickle.dumps(item, protocol=pickle.HIGHEST_PROTOCOL)
    fd = os.open(tmp, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.write(fd, blob)
    finally:
        os.close(fd)
    os.replace(tmp, path)      # a killed run never leaves a half-written file for the resume check to trust
    return path


def write_many(folder: str, ids, embeddings, targets, pooling: Optional[str] = None) -> int:
    """A chunk of files from ONE writer task: per-file Python overhead is what limits the writer (a thread per file fights over the
    interpreter lock: 2 k files/s from 8 threads, 20 k files/s from one loop), so tasks are per batch, not per utterance."""
    for k, slurp_id in enumerate(ids):
        write_item(folder, slurp_id, embeddings[k], targets[k], pooling)
    return len(ids)


class WriterProcs:
    """Writer PROCESSES fed through pipes (``python -m loco_asr_b200._writer``; numpy only, no CUDA state).  Threads do not work
    here: every file costs four system calls, each releases the interpreter lock, and with the feeding thread busy each
    re-acquisition waits for a switch interval -- measured 0.6 ms per open(), 1.9 k files/s, against 20 k files/s for the same
    loop in a process of its own.  A full pipe blocks ``submit``: that is the back-pressure."""

    def __init__(self, writers: int):
        import subprocess
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        self.procs = [subprocess.Popen([sys.executable, "-m", "loco_asr_b200._writer"], stdin=subprocess.PIPE, env=env)
                      for _ in range(max(1, writers))]
        self.k = 0

    def submit(self, folder, ids, embeddings, targets, pooling=None):
        import struct
        blob = pickle.dumps((folder, list(ids), embeddings, list(targets), pooling), protocol=pickle.HIGHEST_PROTOCOL)
        p = self.procs[self.k % len(self.procs)]
        self.k += 1
        p.stdin.write(struct.pack("<Q", len(blob)))
        p.stdin.write(blob)

    def close(self):
        rc = 0
        for p in self.procs:
            p.stdin.close()
        for p in self.procs:
            rc |= p.wait()
        if rc:
            raise RuntimeError("an embedding writer process failed; some files were not written")


def check_folder_format(folder: str, pooling: Optional[str]):
    """Refuse to add files of one format to a folder that holds the other: ``train_classifier.py --pooling max|attention``
    over already-averaged [1, 768] rows runs without error but computes something else than pooling over frames."""
    if not os.path.isdir(folder):
        return
    for name in os.listdir(folder):
        if not name.endswith("_embedding_and_target.pickle"):
            continue
        with open(os.path.join(folder, name), "rb") as fh:
            have = pickle.load(fh).get("pooling")
        if have != pooling:
            raise SystemExit(f"{folder} already holds {'pooled (' + have + ')' if have else 'full-sequence'} files; this run would add "
                             f"{'pooled (' + pooling + ')' if pooling else 'full-sequence'} ones.  Use another --out-root or the same format.")
        return


# ------------------------------------------------------------------------------------------------ main
def stream_batches(todo, waves_fn, frames_fn, max_frames: int, decoders: int = 8, lookahead: int = 2):
    """Decode / produce waveforms in worker threads and cut them into padding-free batches as they arrive: yields
    ``(indices into todo, [waveforms])`` with at most `max_frames` encoder frames.  At most `lookahead` batches' worth of
    decoded audio is alive at a time (the reference decodes serially inside its collate function, :55-57; materialising a
    whole split first costs > 10 GB of host memory for SLURP train)."""
    from collections import deque
    pool = ThreadPoolExecutor(max_workers=decoders) if decoders > 1 else None   # cheap decoders run inline: threads only add hand-offs
    pending = deque()
    it = iter(range(len(todo)))
    budget = max(64, decoders * 4) if pool else 1

    def refill():
        while len(pending) < budget:
            i = next(it, None)
            if i is None:
                return
            pending.append((i, pool.submit(waves_fn, todo[i][1]) if pool else None))

    refill()
    idx, waves, frames = [], [], 0
    while pending:
        i, fut = pending.popleft()
        w = fut.result() if pool else waves_fn(todo[i][1])
        refill()
        f = int(frames_fn(len(w))) + 2
        if idx and frames + f > max_frames:
            yield idx, waves
            idx, waves, frames = [], [], 0
        idx.append(i)
        waves.append(w)
        frames += f
    if idx:
        yield idx, waves
    if pool:
        pool.shutdown()


def run(encoder, items, waves_fn, binarize, folder: str, pooled: Optional[str] = None, max_frames: int = 131072,
        writers: int = 2, decoders: int = 8, resume: bool = True, rank: int = 0, world: int = 1, log=print):
    """items: [(slurp_id, key, intent)]; waves_fn(key) -> float32 waveform.  ``pooled``: None writes the reference's
    ``embedding`` [T, 768] (last_hidden_state of the utterance's own frames); "average" / "max" write the pooled [1, 768]
    vector (150x smaller; "average" rides the copy-overlapped bulk path).  Under torchrun every rank takes every `world`-th
    item and writes its own files; no collective is needed because the files are the result."""
    import torch
    from .buckets import frames_of

    os.makedirs(folder, exist_ok=True)
    check_folder_format(folder, pooled)
    mine = items[rank::world]
    todo = [it for it in mine if not (resume and os.path.exists(output_path(folder, it[0])))]
    if len(todo) < len(mine):
        log(f"resume: {len(mine) - len(todo)} of {len(mine)} outputs already exist")
    if not todo:
        return 0
    targets = binarize([it[2] for it in todo])
    pool = WriterProcs(writers)
    cuda = torch.cuda.is_available()
    def frames_fn(n):        # HF _get_feat_extract_output_lengths in plain integers (numpy costs 70 us per call here)
        for k, st in ((10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2)):
            n = (n - k) // st + 1 if n >= k else 0
        return n
    if pooled is None:
        max_frames = min(max_frames, 32768)      # [sum T, 768] fp32 crosses PCIe and sits in host memory: 100 MB per batch

    def sorted_batches():
        for idx, waves in stream_batches(todo, waves_fn, frames_fn, max_frames, decoders):
            order = sorted(range(len(idx)), key=lambda k: len(waves[k]))     # short to long inside the batch: dense tiles
            yield [idx[k] for k in order], [waves[k] for k in order]

    if pooled == "average":
        # H2D of the next batch overlaps the encode of the current one (encode_host_pipelined); decode runs ahead in threads
        metas = []
        ring = [None] * 4           # pinned staging buffers, reused: a waveform is copied once, straight into pinned memory
                                    # (concatenate + pin_memory per batch were two more passes over 170 MB and a cudaHostAlloc)

        def host_batches():
            for b, (idx, waves) in enumerate(sorted_batches()):
                n = sum(len(w) for w in waves)
                k = b % len(ring)   # batch b - 4 has been encoded and read back by now (results are consumed one batch late)
                if ring[k] is None or ring[k].numel() < n:
                    ring[k] = torch.empty(max(n, 1 << 20), dtype=torch.float32)
                    if cuda:
                        ring[k] = ring[k].pin_memory()
                dst = ring[k].numpy()
                off = 0
                for w in waves:
                    dst[off:off + len(w)] = w
                    off += len(w)
                metas.append(idx)
                yield (ring[k][:n], [len(w) for w in waves])

        for b, pooled_host in enumerate(encoder.encode_host_pipelined(host_batches())):
            arr = pooled_host.numpy().copy()[:, None, :]          # [B, 1, 768]: the pinned buffer goes back to the pipeline
            idx = metas[b]
            for c in range(0, len(idx), 512):
                part = idx[c:c + 512]
                pool.submit(folder, [todo[i][0] for i in part], arr[c:c + 512], [targets[i] for i in part], "average")
            metas[b] = None
    else:
        if pooled == "max":
            encoder.set_head(method="max")
        elif pooled is not None:
            raise ValueError("pooled must be None, 'average' or 'max' (self_attention pooling needs the trained classifier's q: "
                             "use full-sequence files or LocoSpeechT5Encoder.set_head)")
        for idx, waves in sorted_batches():
            dev_wave = torch.from_numpy(np.concatenate(waves)).to(encoder.device)
            ns = [len(w) for w in waves]
            if pooled is None:
                _, hidden, info = encoder.encode_packed(dev_wave, ns, return_hidden=True)
                hidden = hidden.cpu().numpy()
                offs = np.concatenate([[0], np.cumsum(info["frames"])])
                for c in range(0, len(idx), 64):
                    part = range(c, min(c + 64, len(idx)))
                    pool.submit(folder, [todo[idx[j]][0] for j in part], [hidden[offs[j]:offs[j + 1]] for j in part],
                                [targets[idx[j]] for j in part], None)
            else:
                _, head_pooled, _ = encoder.encode_packed(dev_wave, ns, with_head=True)
                arr = head_pooled.cpu().numpy()[:, None, :]
                pool.submit(folder, [todo[i][0] for i in idx], arr, [targets[i] for i in idx], "max")
    pool.close()
    return len(todo)


def load_tokenizer(name_or_path: str):
    """sentence -> int64 token ids, as ``processor(text=...)`` gives the reference (:59): SpeechT5Tokenizer, character-level
    sentencepiece, vocabulary 81, ``</s>`` appended."""
    try:
        from transformers import SpeechT5Tokenizer
        tok = SpeechT5Tokenizer.from_pretrained(name_or_path)
    except Exception as e:  # no network / no cache in this environment
        raise SystemExit(f"cannot load the SpeechT5 tokenizer from {name_or_path!r} ({type(e).__name__}: {e}); "
                         "pass --tokenizer <local directory> or use --synthetic") from e
    return lambda sentence: np.asarray(tok(sentence)["input_ids"], dtype=np.int64)


def run_text(encoder, items, tokens_fn, binarize, folder: str, pooled: Optional[str] = None, max_tokens: int = 131072,
             writers: int = 8, resume: bool = True, rank: int = 0, world: int = 1, log=print):
    """Text modality: items [(slurp_id, sentence-key, intent)]; tokens_fn(key) -> int token ids.  Same files as run()."""
    import torch
    from .buckets import make_batches

    os.makedirs(folder, exist_ok=True)
    check_folder_format(folder, pooled)
    if pooled not in (None, "average"):
        raise ValueError("text modality: pooled must be None or 'average'")
    full_sequence = pooled is None
    items = items[rank::world]
    todo = [it for it in items if not (resume and os.path.exists(output_path(folder, it[0])))]
    if len(todo) < len(items):
        log(f"resume: {len(items) - len(todo)} of {len(items)} outputs already exist")
    if not todo:
        return 0
    toks = [np.asarray(tokens_fn(it[1]), dtype=np.int64) for it in todo]
    lengths = [len(t) for t in toks]
    targets = binarize([it[2] for it in todo])
    pool = WriterProcs(writers)
    order = sorted(range(len(todo)), key=lambda i: lengths[i])
    start = 0
    while start < len(order):
        end, total = start, 0
        while end < len(order) and (end == start or total + lengths[order[end]] <= max_tokens):
            total += lengths[order[end]]
            end += 1
        idx = order[start:end]
        start = end
        packed = torch.from_numpy(np.concatenate([toks[i] for i in idx])).to(torch.int32).to(encoder.device)
        nt = [lengths[i] for i in idx]
        if full_sequence:
            pooled, hidden, _ = encoder.encode_text_packed(packed, nt, return_hidden=True)
            hidden = hidden.cpu().numpy()
            off = 0
            for j, i in enumerate(idx):
                pool.submit(folder, [todo[i][0]], [hidden[off:off + nt[j]]], [targets[i]], None)
                off += nt[j]
        else:
            arr = encoder.encode_text_packed(packed, nt).cpu().numpy()
            pool.submit(folder, [todo[i][0] for i in idx], arr[:, None, :], [targets[i] for i in idx], "average")
    pool.close()
    return len(todo)


def main(argv=None):
    p = argparse.ArgumentParser(description="Extract SpeechT5 encoder embeddings from SLURP (B200 CUDA path)")
    p.add_argument("--modality", "-m", choices=["text", "audio"], required=True)
    p.add_argument("--split", "-s", choices=["train", "devel", "test", "train_synthetic"], required=True)
    p.add_argument("--version", choices=["base", "finetuned"], default="base",
                   help="base = extract_speecht5_base_... (folder extracted/speecht5_base), finetuned = extracted/speecht5")
    p.add_argument("--data-path", default="slurp")
    p.add_argument("--out-root", default="extracted")
    p.add_argument("--weights", help="torch / pickle file with an HF-keyed encoder state dict (full model or encoder)")
    p.add_argument("--mapping-dir", default="extracted/speecht5/mapping",
                   help="folder with encoder_state_dict.pickle / speech_prenet_state_dict.pickle (reference :41-49)")
    p.add_argument("--classes-file")
    p.add_argument("--pooled", choices=["average", "max"], default=None,
                   help="write the pooled [1, 768] vector instead of the reference's [T, 768] (recorded in the file as 'pooling')")
    p.add_argument("--full-sequence", action="store_true", help="(default) write [T, 768] like the reference")
    p.add_argument("--decoders", type=int, default=8, help="waveform decode threads")
    p.add_argument("--writers", type=int, default=2, help="pickle writer processes (tasks are whole chunks of files)")
    p.add_argument("--report", help="write a JSON throughput report of this run (utterances, audio-s, seconds, files) to this path")
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic SLURP-shaped utterances and random-init weights")
    p.add_argument("--synthetic-decoder", choices=["synth", "slice"], default="synth",
                   help="synthetic waveforms: 'synth' = the seeded generator the tests compare against (8 ms of numpy each), "
                        "'slice' = memcpy-speed windows of one recording (throughput runs: the host pipeline, not numpy, is measured)")
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--max-frames", type=int, default=131072)
    p.add_argument("--do-normalize", action="store_true", help="zero-mean / unit-variance waveforms (the feature extractor's do_normalize)")
    p.add_argument("--tokenizer", default="microsoft/speecht5_asr", help="text modality: SpeechT5 tokenizer name or local directory")
    a = p.parse_args(argv)
    print(f"Extracting {a.modality} embeddings from SLURP {a.split} set using SpeechT5 (loco_asr_b200)")

    import time
    import torch
    from .encoder import LocoSpeechT5Encoder
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:          # torchrun: one process per GPU, each with its own share of the utterances and its own writers
        a.device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    if a.pooled and a.full_sequence:
        raise SystemExit("--pooled and --full-sequence exclude each other")
    enc = LocoSpeechT5Encoder(device=a.device)
    text = a.modality == "text"
    if a.synthetic and text:
        from .synth import synth_state_dict, synth_text_prenet_state_dict
        enc.load_state_dict({k: v for k, v in synth_state_dict(seed=1).items() if k.startswith("wrapped_encoder.")})
        enc.load_state_dict(synth_text_prenet_state_dict(seed=1))
        classes = load_classes(a.classes_file) if (a.classes_file or "intent_classes" in sys.modules) else [f"intent_{i:03d}" for i in range(101)]
        rng = np.random.default_rng(1234)
        sents = [np.concatenate([rng.integers(4, 81, size=int(rng.integers(8, 90))), [2]]) for _ in range(a.synthetic)]
        items = [(f"synth{i}", i, classes[i % len(classes)]) for i in range(a.synthetic)]
        waves_fn = lambda i: sents[i]
    elif a.synthetic:
        from .synth import slurp_shaped_lengths, synth_state_dict, synth_wave
        enc.load_state_dict(synth_state_dict(seed=1))
        classes = load_classes(a.classes_file) if (a.classes_file or "intent_classes" in sys.modules) else [f"intent_{i:03d}" for i in range(101)]
        lens = slurp_shaped_lengths(a.synthetic, 1234)
        items = [(f"synth{i}", i, classes[i % len(classes)]) for i in range(a.synthetic)]
        if a.synthetic_decoder == "slice":
            # stands in for reading decoded PCM: a window of one 11 s recording, scaled -- memcpy-speed, like a warm page cache
            # (synth_wave itself costs ~8 ms of numpy per utterance, more than the GPU spends on 100 of them)
            base = synth_wave(176000, 1234, 0)
            waves_fn = lambda i: base[(int(i) * 7919) % (176000 - int(lens[i]) + 1):][:int(lens[i])] * np.float32(0.5 + 0.001 * (int(i) % 997))
        else:
            waves_fn = lambda i: synth_wave(int(lens[i]), 1234, int(i))
    else:
        classes = load_classes(a.classes_file)
        if a.weights:
            enc.load_state_dict(load_weights_file(a.weights))
        else:
            with open(os.path.join(a.mapping_dir, "encoder_state_dict.pickle"), "rb") as fh:
                enc.wrapped_encoder.load_state_dict(pickle.load(fh))
            prenet_file = "text_prenet_state_dict.pickle" if text else "speech_prenet_state_dict.pickle"   # reference :44-49
            with open(os.path.join(a.mapping_dir, prenet_file), "rb") as fh:
                enc.prenet.load_state_dict(pickle.load(fh))
        items = read_slurp_index(a.data_path, a.split, a.modality)
        waves_fn = load_tokenizer(a.tokenizer) if text else load_audio
    if a.do_normalize and not text:
        raw_fn = waves_fn
        waves_fn = lambda key: zero_mean_unit_var(raw_fn(key))
    enc.finalize()
    print(f"{a.split} set size: {len(items)}")
    folder = output_folder(a.out_root, a.version, a.split, a.modality)
    t0 = time.perf_counter()
    if text:
        n = run_text(enc, items, waves_fn, make_label_binarizer(classes), folder, pooled=a.pooled, max_tokens=a.max_frames,
                     writers=a.writers, rank=rank, world=world)
    else:
        n = run(enc, items, waves_fn, make_label_binarizer(classes), folder, pooled=a.pooled, max_frames=a.max_frames,
                writers=a.writers, decoders=a.decoders, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"wrote {n} files{'' if world == 1 else f' (rank {rank} of {world})'} in {dt:.1f} s\nDone!")
    if a.report and not text:
        audio_s = float(sum(lens[i] for i in range(rank, a.synthetic, world))) / SAMPLE_RATE if a.synthetic else None
        with open(a.report + (f".rank{rank}" if world > 1 else ""), "w") as fh:
            json.dump({"files_written": n, "seconds": dt, "utterances_per_s": n / dt if dt > 0 else None, "audio_s": audio_s,
                       "e2e_files_audio_s_per_s": audio_s / dt if (audio_s and dt > 0 and n) else None, "rank": rank, "world": world,
                       "format": a.pooled or "full-sequence", "decoders": a.decoders, "writers": a.writers,
                       "timed": "decode/synthesis + H2D + encode + D2H + one pickle per utterance on disk (wall clock)"}, fh)


if __name__ == "__main__":
    main()
