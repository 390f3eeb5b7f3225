"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz`` from the reference's own
implementation (the unmodified HF ``SpeechT5EncoderWithSpeechPrenet``, see oracle/hf_reference.py) on
seeded synthetic inputs.  Run in the build container:  ``python -m oracle.make_golden``.

Fixtures
  config1_hf.npz   BASELINE.json configs[0]: 16 utterances of 2.5-3.5 s, weights seed 0, waveform seed 0.
                   pooled f32[16,768] (mean over own frames), n_frames, first/last frame of each utterance,
                   and the reference's literal batch_size=2 padded run (informational).
  text_hf.npz      text modality: 6 seeded token sequences (1..120 tokens) through the unmodified HF
                   SpeechT5EncoderWithTextPrenet, weights seed 0 + text prenet seed 0: pooled, first / last token rows.
  config5_hf.npz   BASELINE.json configs[4] / SURVEY.md 8(d) "Config 5": a fixed 512-utterance subset (every ~137th) of the
                   70k SLURP-shaped set (lengths seed 1234, weights seed 1, waveform seed 1234): the HF module's pooled embedding
                   (fp32), logits / argmax / top-2 margin of IntentClassifier(average) with the seed-3 random Linear(768,101), and
                   what the same module does under bf16 autocast (argmax, rel err, cosine).   ``python -m oracle.make_golden --config5-only``
  long60_hf.npz    BASELINE.json configs[3]: one 60 s segment (T = 2999; the HF module materialises 2.3 GB of position_bias for it),
                   weights seed 0, waveform seed 21: pooled, and 16 evenly spaced rows of last_hidden_state.
                   ``python -m oracle.make_golden --long60-only``
  short_taps.npz   one 0.4 s noise utterance (T=19, all taps) and one 1.3 s utterance (T=64, three taps): stage-by-stage intermediates of
                   the HF module itself, taken with forward hooks (hf_reference.hf_stage_taps); the restatement's taps are asserted
                   equal to them to 2e-5 while the fixture is written.   ``python -m oracle.make_golden --taps-only``
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from loco_asr_b200.synth import synth_state_dict, synth_text_prenet_state_dict, synth_wave, config1_lengths  # noqa: E402
from oracle import speecht5_oracle as O  # noqa: E402
from oracle.hf_reference import (build_hf_encoder, build_hf_text_encoder, hf_encode_padded_batches,  # noqa: E402
                                 hf_encode_text_unpadded, hf_encode_unpadded)

TAP_KEYS = ["conv0", "conv1", "conv6", "proj_ln", "proj", "pos_conv", "prenet_out", "enc_in", "l0_qkv", "l0_ctx",
            "l0_ln1", "l0_mid", "layer0", "layer5", "layer11"]


TEXT_LENGTHS = [1, 7, 23, 64, 65, 120]


def text_tokens(seed: int = 0, vocab: int = 81):
    """Seeded token sequences: ids 4..vocab-1 with the </s> id 2 at the end (SpeechT5 tokenizer convention)."""
    rng = np.random.default_rng(1000 + seed)
    return [np.concatenate([rng.integers(4, vocab, size=n - 1), [2]]).astype(np.int64) for n in TEXT_LENGTHS]


def text_state_dict(seed: int = 0):
    sd = {k: v for k, v in synth_state_dict(seed=seed).items() if k.startswith("wrapped_encoder.")}
    sd.update(synth_text_prenet_state_dict(seed=seed))
    return sd


def make_text(out_dir):
    sd = text_state_dict(0)
    model = build_hf_text_encoder(sd)
    toks = text_tokens(0)
    hs = hf_encode_text_unpadded(model, toks)
    np.savez(os.path.join(out_dir, "text_hf.npz"),
             lengths=np.asarray(TEXT_LENGTHS, dtype=np.int64), tokens=np.concatenate(toks),
             pooled=torch.stack([h.mean(0) for h in hs]).numpy().astype(np.float32),
             first_row=torch.stack([h[0] for h in hs]).numpy().astype(np.float32),
             last_row=torch.stack([h[-1] for h in hs]).numpy().astype(np.float32), weights_seed=0)


def config5_ids(n_total: int = 70000, n: int = 512):
    return np.linspace(0, n_total - 1, n).astype(np.int64)


def make_config5(out_dir):
    """Also runs the SAME unmodified HF module under torch.autocast(bfloat16) -- the reference's own arithmetic at the CUDA path's
    operand precision (bf16 GEMM / conv operands, fp32 softmax / LayerNorm) -- and stores what that alone does to the pooled
    embedding and to the intent argmax: the yardstick for the CUDA path's bf16 error (tests/test_gpu_parity.py)."""
    from loco_asr_b200.synth import slurp_shaped_lengths, synth_head
    lengths = slurp_shaped_lengths(70000, 1234)
    ids = config5_ids()
    model = build_hf_encoder(synth_state_dict(seed=1))
    w, b = synth_head(3)
    pooled, pooled16 = [], []
    for k in range(0, len(ids), 16):
        waves = [synth_wave(int(lengths[i]), 1234, int(i)) for i in ids[k:k + 16]]
        pooled += [h.mean(0) for h in hf_encode_unpadded(model, waves)]
        with torch.autocast("cpu", dtype=torch.bfloat16):
            pooled16 += [h.float().mean(0) for h in hf_encode_unpadded(model, waves)]
        print(f"config5: {k + 16}/{len(ids)}", flush=True)
    pooled, pooled16 = torch.stack(pooled), torch.stack(pooled16)
    logits = torch.nn.functional.linear(pooled, w, b)
    logits16 = torch.nn.functional.linear(pooled16, w, b)
    top2 = logits.topk(2, dim=1).values
    np.savez_compressed(os.path.join(out_dir, "config5_hf.npz"), ids=ids, n_samples=lengths[ids].astype(np.int64),
                        pooled=pooled.numpy().astype(np.float32), logits=logits.numpy().astype(np.float32),
                        argmax=logits.argmax(dim=1).numpy().astype(np.int64),
                        margin=(top2[:, 0] - top2[:, 1]).numpy().astype(np.float32),
                        hf_bf16_argmax=logits16.argmax(dim=1).numpy().astype(np.int64),
                        hf_bf16_rel_err=((pooled16 - pooled).abs().amax(1) / pooled.abs().amax(1)).numpy().astype(np.float32),
                        hf_bf16_cosine=torch.nn.functional.cosine_similarity(pooled16, pooled, dim=1).numpy().astype(np.float32),
                        hf_bf16_logit_diff=(logits16 - logits).abs().amax(1).numpy().astype(np.float32),
                        weights_seed=1, wave_seed=1234, head_seed=3)


def make_long60(out_dir):
    model = build_hf_encoder(synth_state_dict(seed=0))
    w = synth_wave(960000, 21, 7)
    h = hf_encode_unpadded(model, [w])[0]
    rows = np.linspace(0, h.shape[0] - 1, 16).astype(np.int64)
    np.savez(os.path.join(out_dir, "long60_hf.npz"), n_samples=960000, wave_seed=21, wave_idx=7, n_frames=h.shape[0],
             pooled=h.mean(0).numpy().astype(np.float32), rows=rows, hidden_rows=h[rows].numpy().astype(np.float32))


def make_taps(out_dir, sd, model):
    """Stage taps of the HF module itself (forward hooks, oracle/hf_reference.hf_stage_taps); the restatement is only checked
    against them here, it contributes nothing to the fixture."""
    from oracle.hf_reference import hf_stage_taps
    taps_out = {}
    for name, n, idx in (("a", 6400, 100), ("b", 20800, 101)):
        w = synth_wave(n, 0, idx, kind="noise" if name == "a" else "mix")
        ref = hf_stage_taps(model, w)
        mine = {}
        O.encode_utterance(sd, torch.from_numpy(w), taps=mine)
        taps_out[f"{name}_n_samples"] = n
        taps_out[f"{name}_idx"] = idx
        taps_out[f"{name}_hf_last_hidden"] = ref["last_hidden"].numpy().astype(np.float32)
        for k in (TAP_KEYS if name == "a" else ["pos_conv", "enc_in", "layer0"]):
            err = float((mine[k] - ref[k]).abs().max() / ref[k].abs().max())
            assert err < 2e-5, (k, err)
            v = ref[k].numpy().astype(np.float32)
            if k in ("conv0", "conv1"):
                v = v[:64]  # first 64 frames are enough to pin layout + GroupNorm statistics
            taps_out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(out_dir, "short_taps.npz"), **taps_out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_state_dict(seed=0)
    model = build_hf_encoder(sd)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--long60-only" in sys.argv:
        make_long60(out_dir)
        return
    if "--config5-only" in sys.argv:
        make_config5(out_dir)
        return
    if "--taps-only" in sys.argv:
        make_taps(out_dir, sd, model)
        return
    make_text(out_dir)
    if "--text-only" in sys.argv:
        return

    if "--with-config5" in sys.argv:
        make_config5(out_dir)
    lengths = config1_lengths()
    waves = [synth_wave(n, 0, i) for i, n in enumerate(lengths)]
    hs = hf_encode_unpadded(model, waves)
    pooled = torch.stack([h.mean(0) for h in hs]).numpy()
    first = torch.stack([h[0] for h in hs]).numpy()
    last = torch.stack([h[-1] for h in hs]).numpy()
    padded = hf_encode_padded_batches(model, waves, batch_size=2)
    pooled_padded = []
    for b, out in enumerate(padded):
        for j in range(out.shape[0]):
            t = hs[2 * b + j].shape[0]
            pooled_padded.append(out[j, :t].mean(0))
    np.savez(os.path.join(out_dir, "config1_hf.npz"),
             lengths=np.asarray(lengths, dtype=np.int64),
             n_frames=np.asarray([h.shape[0] for h in hs], dtype=np.int64),
             pooled=pooled.astype(np.float32), first_frame=first.astype(np.float32),
             last_frame=last.astype(np.float32),
             pooled_padded_bs2=torch.stack(pooled_padded).numpy().astype(np.float32),
             weights_seed=0, wave_seed=0)

    make_taps(out_dir, sd, model)
    for f in os.listdir(out_dir):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
