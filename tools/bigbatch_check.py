"""Bit-identity of utterances inside one large packed launch vs encoded alone, by batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from loco_asr_b200.encoder import LocoSpeechT5Encoder
from loco_asr_b200.synth import synth_state_dict

enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=0), device="cuda:0", debug=True)
for key, val in [a.split("=") for a in sys.argv[2:]]:
    enc.debug_set(key, int(val))
for max_frames in [int(x) for x in sys.argv[1].split(",")]:
    gen = torch.Generator(device="cuda").manual_seed(77)
    rng = np.random.default_rng(77)
    lengths, frames = [], 0
    while True:
        n = int(rng.integers(16000, 160001))
        t = (n - 400) // 320 + 1
        if frames + t + 2 > max_frames:
            break
        lengths.append(n)
        frames += t + 2
    wave = torch.randn(int(np.sum(lengths)), device="cuda", generator=gen) * 0.1
    pooled, hidden, info = enc.encode_packed(wave, lengths, return_hidden=True)
    torch.cuda.synchronize()
    offs = np.concatenate([[0], np.cumsum(lengths)])
    foffs = np.concatenate([[0], np.cumsum(info["frames"])])
    bad = []
    us = sorted(set([0, 1, 2, len(lengths) // 4, len(lengths) // 2, 3 * len(lengths) // 4, len(lengths) - 2, len(lengths) - 1]))
    for u in us:
        p1, h1, _ = enc.encode_packed(wave[int(offs[u]):int(offs[u + 1])].contiguous(), [lengths[u]], return_hidden=True)
        d = float((p1[0] - pooled[u]).abs().max())
        hd = (h1 - hidden[int(foffs[u]):int(foffs[u + 1])]).abs().amax(dim=1)
        nz = torch.nonzero(hd > 0).flatten()
        bad.append((u, int(info["frames"][u]), d, int(nz.numel()), int(nz[0]) if nz.numel() else -1, int(nz[-1]) if nz.numel() else -1))
    print(max_frames, len(lengths), "utts; (u, T, pooled maxdiff, frames differing, first, last):", bad, flush=True)
