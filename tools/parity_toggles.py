"""Per-approximation error budget on the config-5 subset (512 utterances, tests/golden/config5_hf.npz): every approximation the
CUDA path stacks on top of bf16 operands is switched off in turn (library variants built by tools/build_variant.sh, or debug
knobs) and the pooled embeddings / fused-head logits are compared with the fp32 HF reference.  Prints a markdown table."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [
    ("product path (libloco_asr.so)", None, {}, False),
    ("LOCO_DEBUG build, same kernels", None, {}, True),
    ("GELU: ex2 + rcp form instead of tanh.approx", "lib_gelu_exact.so", {}, True),
    ("GELU: library erff", "lib_gelu_erf.so", {}, True),
    ("softmax: exp2f instead of ex2.approx", "lib_ex2_precise.so", {}, True),
    ("softmax: rescale at every new maximum (lazy threshold 0)", "lib_rescale0.so", {}, True),
    ("LayerNorm kernels instead of deferred LayerNorm (no gamma-folded bf16 weights, normalised bf16 residual stream)", None, {"ln_impl": 1}, True),
    ("attention: one-item tcgen05 kernel for every utterance", None, {"attn_p2": 0}, True),
    ("attention: mma.sync cross-check kernel for every utterance", None, {"attn_impl": 1}, True),
    ("positional conv: mma.sync cross-check kernel", None, {"posconv_impl": 1}, True),
]

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import numpy as np, torch
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.head import IntentHead
    from loco_asr_b200.synth import synth_head, synth_state_dict, synth_wave
    knobs, debug = json.loads(sys.argv[2]), sys.argv[3] == "1"
    g = np.load(os.path.join(ROOT, "tests", "golden", "config5_hf.npz"))
    ids, n_samples = g["ids"], g["n_samples"]
    enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0", debug=debug)
    for k, v in knobs.items():
        enc.debug_set(k, v)
    w, b = synth_head(3)
    enc.set_head(IntentHead(w, b, None, "average"))
    order = np.argsort(n_samples, kind="stable")
    pooled, logits = torch.empty(len(ids), 768), torch.empty(len(ids), 101)
    for k in range(0, len(ids), 64):
        sel = order[k:k + 64]
        waves = [synth_wave(int(n_samples[i]), 1234, int(ids[i])) for i in sel]
        p, hp, lg = enc.encode_packed(torch.from_numpy(np.concatenate(waves)).cuda(), [len(x) for x in waves], with_head=True)
        pooled[sel], logits[sel] = p.cpu(), lg.cpu()
    ref = torch.from_numpy(g["pooled"])
    cos = torch.nn.functional.cosine_similarity(pooled, ref, dim=1)
    rel = (pooled - ref).abs().amax(1) / ref.abs().amax(1)
    agree = logits.argmax(1).numpy() == g["argmax"]
    print(json.dumps({"min_cos": float(cos.min()), "max_rel": float(rel.max()), "mean_rel": float(rel.mean()), "flips": int((~agree).sum()),
                      "max_logit_diff": float((logits - torch.from_numpy(g["logits"])).abs().max())}))
    sys.exit(0)

import numpy as np
g = np.load(os.path.join(ROOT, "tests", "golden", "config5_hf.npz"))
print("| variant | min pooled cosine | max rel err | mean rel err | intent argmax differs (of 512) | max abs logit diff |")
print("|---|---|---|---|---|---|")
print(f"| *yardstick: the HF module itself under torch.autocast(bfloat16), CPU* | {float(g['hf_bf16_cosine'].min()):.6f} | {float(g['hf_bf16_rel_err'].max()):.5f} | "
      f"{float(g['hf_bf16_rel_err'].mean()):.5f} | {int((g['hf_bf16_argmax'] != g['argmax']).sum())} | {float(g['hf_bf16_logit_diff'].max()):.5f} |")
for name, lib, knobs, debug in VARIANTS:
    env = dict(os.environ)
    if lib:
        env["LOCO_ASR_LIB"] = os.path.join(ROOT, "tools", "_libs", lib)
    r = subprocess.run([sys.executable, __file__, "--one", json.dumps(knobs), "1" if debug else "0"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"| {name} | {d['min_cos']:.6f} | {d['max_rel']:.5f} | {d['mean_rel']:.5f} | {d['flips']} | {d['max_logit_diff']:.5f} |", flush=True)
    except Exception:
        print(f"| {name} | failed: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else r.stdout[-200:]} |", flush=True)
