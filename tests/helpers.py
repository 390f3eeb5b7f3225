"""Shared by tests/ and tools/diag.py: run the CUDA encoder and compare stage buffers with oracle taps."""
from __future__ import annotations

import numpy as np
import torch

from loco_asr_b200.synth import synth_wave
from oracle import speecht5_oracle as O

# stage buffer name in the workspace -> (oracle tap, conv layer index or None)
STAGES = [("conv0", "conv0", 0), ("conv1", "conv1", 1), ("conv2", "conv2", 2), ("conv3", "conv3", 3),
          ("conv4", "conv4", 4), ("conv5", "conv5", 5), ("conv6", "conv6", 6), ("proj_ln", "proj_ln", None),
          ("proj", "proj", None), ("pos_conv", "pos_conv", None), ("enc_in", "enc_in", None)]
LAYER0_STAGES = [("qkv", "l0_qkv"), ("ctx", "l0_ctx"), ("ln1", "l0_ln1"), ("mid", "l0_mid")]


def make_waves(lengths, seed=0, base_idx=300):
    return [synth_wave(n, seed, base_idx + i, kind="noise" if i % 3 == 2 else "mix") for i, n in enumerate(lengths)]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b|  (the 'max relative error' BASELINE.md reports)."""
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def cosine(a, b) -> float:
    return float(torch.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0))


def oracle_taps(sd, waves, n_layers=None):
    out = []
    for w in waves:
        taps = {}
        h = O.encode_utterance(sd, torch.from_numpy(w), n_layers=n_layers, taps=taps)
        taps["final"] = h
        out.append(taps)
    return out


def run_encoder(enc, waves, return_hidden=True):
    lengths = [len(w) for w in waves]
    wave = torch.from_numpy(np.concatenate(waves)).to(enc.device)
    pooled, hidden, info = enc.encode_packed(wave, lengths, return_hidden=True)
    torch.cuda.synchronize()
    return pooled.cpu(), hidden.cpu(), info


def stage_rows(enc, name, info, u, conv_level=None):
    """Rows of utterance u in a slot-packed stage buffer (see include/loco_asr.h: loco_debug_buffer)."""
    buf = enc.debug_buffer(name).float().cpu()
    r0 = int(info["rows"][u])
    if conv_level is not None:
        r0 <<= (6 - conv_level)
    return buf, r0


def compare_stages(enc, info, taps_list, stages, report):
    worst = {}
    for name, tap, lvl in stages:
        buf = enc.debug_buffer(name).float().cpu()
        for u, taps in enumerate(taps_list):
            ref = taps[tap]
            if name == "qkv":   # the library folds log2(e) into the (already 1/8-scaled) q projection
                ref = torch.cat([ref[:, :768] * 1.4426950408889634, ref[:, 768:]], dim=1)
            r0 = int(info["rows"][u]) << ((6 - lvl) if lvl is not None else 0)
            got = buf[r0:r0 + ref.shape[0]]
            e, c = rel_err(got, ref), cosine(got, ref)
            worst[name] = max(worst.get(name, 0.0), e)
            report(f"  stage {name:9s} utt {u}: rel_err {e:.5f} cosine {c:.6f} (rows {ref.shape[0]})")
    return worst
