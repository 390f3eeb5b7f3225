"""Drop-in for ``speech_text/extract_speecht5_{base,finetuned}_embeddings_slurp.py`` (audio and text modalities).

Same CLI (``-m {audio,text} -s {train,devel,test,train_synthetic}``), same output files
(``extracted/{speecht5|speecht5_base}/<split>/<modality>/<slurp_id>_embedding_and_target.pickle`` holding
``{"id", "embedding": np.float32[T, 768], "target": one-hot[101]}``, reference :70-77,:111-113) that
``slurp_embeddings_and_targets.py:19-28`` and ``train_classifier.py`` read back -- but the loop is restructured
for throughput: utterances are length-bucketed into padding-free batches of <= 64k frames, encoded by the CUDA
library, and written by a pool of writer threads; existing files are skipped (resume).

``embedding`` is the pooled [1, 768] vector by default (3 KB instead of ~460 KB per utterance; ``pad_sequence`` +
``mean(dim=1)`` / ``max`` / attention pooling in the reference classifier all accept T = 1); ``--full-sequence``
writes the reference's [T, 768].

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


def write_item(folder: str, slurp_id, embedding: np.ndarray, target: np.ndarray) -> str:
    path = output_path(folder, slurp_id)
    tmp = path + ".tmp"
    with open(tmp, "wb") as handle:
        pickle.dump({"id": slurp_id, "embedding": np.ascontiguousarray(embedding, dtype=np.float32), "target": target},
                    handle, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, path)      # a killed run never leaves a half-written file for the resume check to trust
    return path


# ------------------------------------------------------------------------------------------------ main
def run(encoder, items, waves_fn, binarize, folder: str, full_sequence: bool = False, max_frames: int = 131072,
        writers: int = 8, resume: bool = True, log=print):
    """items: [(slurp_id, key, intent)]; waves_fn(key) -> float32 waveform."""
    import torch
    from .buckets import make_batches

    os.makedirs(folder, exist_ok=True)
    todo = [it for it in items if not (resume and os.path.exists(output_path(folder, it[0])))]
    if len(todo) < len(items):
        log(f"resume: {len(items) - len(todo)} of {len(items)} outputs already exist")
    if not todo:
        return 0
    waves = [waves_fn(it[1]) for it in todo]
    lengths = [len(w) for w in waves]
    targets = binarize([it[2] for it in todo])
    pool = ThreadPoolExecutor(max_workers=writers)
    futures = []
    batches = make_batches(lengths, max_frames=max_frames)
    if not full_sequence:
        # pooled output: H2D of the next batch overlaps the encode of the current one (encode_host_pipelined)
        def host_batches():
            for idx in batches:
                host = torch.from_numpy(np.concatenate([waves[i] for i in idx]))
                yield (host.pin_memory() if torch.cuda.is_available() else host, [lengths[i] for i in idx])
        for idx, pooled in zip(batches, encoder.encode_host_pipelined(host_batches())):
            pooled = pooled.numpy()
            for j, i in enumerate(idx):
                futures.append(pool.submit(write_item, folder, todo[i][0], pooled[j:j + 1].copy(), targets[i]))
        batches = []
    for idx in batches:
        host = torch.from_numpy(np.concatenate([waves[i] for i in idx]))
        ns = [lengths[i] for i in idx]
        if full_sequence:
            pooled, hidden, info = encoder.encode_packed(host.to(encoder.device), ns, return_hidden=True)
            hidden = hidden.cpu().numpy()
            off = 0
            for j, i in enumerate(idx):
                t = int(info["frames"][j])
                futures.append(pool.submit(write_item, folder, todo[i][0], hidden[off:off + t].copy(), targets[i]))
                off += t
        else:
            pooled = encoder.encode_host(host.pin_memory() if torch.cuda.is_available() else host, ns).numpy()
            for j, i in enumerate(idx):
                futures.append(pool.submit(write_item, folder, todo[i][0], pooled[j:j + 1].copy(), targets[i]))
    for f in futures:
        f.result()
    pool.shutdown()
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


def run_text(encoder, items, tokens_fn, binarize, folder: str, full_sequence: bool = False, max_tokens: int = 131072,
             writers: int = 8, resume: bool = True, log=print):
    """Text modality: items [(slurp_id, sentence-key, intent)]; tokens_fn(key) -> int token ids.  Same files as run()."""
    import torch
    from .buckets import make_batches

    os.makedirs(folder, exist_ok=True)
    todo = [it for it in items if not (resume and os.path.exists(output_path(folder, it[0])))]
    if len(todo) < len(items):
        log(f"resume: {len(items) - len(todo)} of {len(items)} outputs already exist")
    if not todo:
        return 0
    toks = [np.asarray(tokens_fn(it[1]), dtype=np.int64) for it in todo]
    lengths = [len(t) for t in toks]
    targets = binarize([it[2] for it in todo])
    pool = ThreadPoolExecutor(max_workers=writers)
    futures = []
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
                futures.append(pool.submit(write_item, folder, todo[i][0], hidden[off:off + nt[j]].copy(), targets[i]))
                off += nt[j]
        else:
            pooled = encoder.encode_text_packed(packed, nt).cpu().numpy()
            for j, i in enumerate(idx):
                futures.append(pool.submit(write_item, folder, todo[i][0], pooled[j:j + 1].copy(), targets[i]))
    for f in futures:
        f.result()
    pool.shutdown()
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
    p.add_argument("--full-sequence", action="store_true", help="write [T, 768] like the reference instead of pooled [1, 768]")
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic SLURP-shaped utterances and random-init weights")
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--max-frames", type=int, default=131072)
    p.add_argument("--do-normalize", action="store_true", help="zero-mean / unit-variance waveforms (the feature extractor's do_normalize)")
    p.add_argument("--tokenizer", default="microsoft/speecht5_asr", help="text modality: SpeechT5 tokenizer name or local directory")
    a = p.parse_args(argv)
    print(f"Extracting {a.modality} embeddings from SLURP {a.split} set using SpeechT5 (loco_asr_b200)")

    import torch
    from .encoder import LocoSpeechT5Encoder
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
    if text:
        n = run_text(enc, items, waves_fn, make_label_binarizer(classes), folder, full_sequence=a.full_sequence, max_tokens=a.max_frames)
    else:
        n = run(enc, items, waves_fn, make_label_binarizer(classes), folder, full_sequence=a.full_sequence, max_frames=a.max_frames)
    print(f"wrote {n} files\nDone!")


if __name__ == "__main__":
    main()
