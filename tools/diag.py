"""GPU bring-up diagnostics: each phase runs in its own process so a trap in one kernel cannot hide the rest.
    python tools/diag.py            # all phases, report -> gpurun_out/diag.log
    python tools/diag.py <phase>    # one phase in this process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PHASES = ["env", "gemm_simt", "pipeline_simt", "gemm_tc", "pipeline_tc", "full_tc"]


def log(*a):
    print(*a, flush=True)


def gemm_cases():
    import torch
    from loco_asr_b200 import _lib
    return [  # (M, N, K, epilogue, conv_like)
        (128, 256, 64, _lib.EPI_BIAS, False),
        (128, 256, 256, _lib.EPI_BIAS, False),
        (300, 768, 768, _lib.EPI_BIAS, False),
        (1000, 2304, 768, _lib.EPI_BIAS, False),
        (777, 3072, 768, _lib.EPI_BIAS_GELU, False),
        (640, 768, 3072, _lib.EPI_BIAS_RESIDUAL, False),
        (999, 512, 1536, _lib.EPI_BIAS_GELU, True),
        (20000, 768, 768, _lib.EPI_BIAS_RESIDUAL, False),
    ]


def phase_gemm(impl):
    import torch
    from loco_asr_b200 import _lib
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    enc = LocoSpeechT5Encoder(device="cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    ok = True
    for (M, N, K, epi, conv_like) in gemm_cases():
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g) * 0.1
        if conv_like:  # overlapping rows: row t = 1536 contiguous elements starting at t*1024
            flat = (torch.randn(M * 1024 + K + 4096, device="cuda", generator=g)).bfloat16()
            a_mat = torch.as_strided(flat, (M, K), (1024, 1))
            a_arg, lda = flat, 1024
        else:
            a_mat = torch.randn(M, K, device="cuda", generator=g).bfloat16()
            a_arg, lda = a_mat, K
        res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if epi == _lib.EPI_BIAS_RESIDUAL else None
        ref = a_mat.float() @ w.float().t() + bias
        if epi == _lib.EPI_BIAS_GELU:
            ref = torch.nn.functional.gelu(ref)
        if res is not None:
            ref = ref + res.float()
        t0 = time.time()
        c = enc.debug_gemm(a_arg, w, bias=bias, residual=res, epilogue=epi, impl=impl, lda=lda, m=M)
        torch.cuda.synchronize()
        err = float((c.float() - ref).abs().max())
        scale = float(ref.abs().max())
        bad = err > 0.02 * scale + 0.02
        ok &= not bad
        log(f"  gemm impl={impl} M={M} N={N} K={K} epi={epi} conv={conv_like}: max_abs_err {err:.4f} (ref max {scale:.2f}) "
            f"{'FAIL' if bad else 'ok'}  [{(time.time()-t0)*1e3:.1f} ms]")
    return ok


def phase_pipeline(impl, n_layers_stop):
    import torch
    import helpers as H
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.synth import synth_state_dict
    sd = synth_state_dict(seed=0)
    enc = LocoSpeechT5Encoder.from_state_dict(sd, device="cuda:0", debug=True)
    enc.debug_set("gemm_impl", impl)
    lengths = [6400, 20800, 48000, 9000, 33000]
    waves = H.make_waves(lengths)
    ok = True
    if n_layers_stop is not None:
        enc.debug_set("stop_after_layer", 0)
        taps = H.oracle_taps(sd, waves, n_layers=1)
        pooled, hidden, info = H.run_encoder(enc, waves)
        log(f"  frames {info['frames'].tolist()} rows {info['rows'].tolist()} ws {info['workspace_bytes']/1e6:.1f} MB")
        worst = H.compare_stages(enc, info, taps, H.STAGES, log)
        worst.update(H.compare_stages(enc, info, taps, [(a, b, None) for a, b in H.LAYER0_STAGES], log))
        off = 0
        for u, t in enumerate(taps):
            T = t["final"].shape[0]
            e = H.rel_err(hidden[off:off + T], t["final"])
            log(f"  layer0 output utt {u}: rel_err {e:.5f} cosine {H.cosine(hidden[off:off+T], t['final']):.6f}")
            worst["layer0"] = max(worst.get("layer0", 0), e)
            off += T
        log("  worst per stage: " + ", ".join(f"{k}={v:.4f}" for k, v in worst.items()))
        ok = all(v < 0.05 for v in worst.values())
    else:
        taps = H.oracle_taps(sd, waves)
        pooled, hidden, info = H.run_encoder(enc, waves)
        off = 0
        for u, t in enumerate(taps):
            T = t["final"].shape[0]
            ref_p = t["final"].mean(0)
            cos = H.cosine(pooled[u], ref_p)
            e = H.rel_err(pooled[u], ref_p)
            eh = H.rel_err(hidden[off:off + T], t["final"])
            log(f"  utt {u} T={T}: pooled cosine {cos:.6f} pooled rel_err {e:.5f} hidden rel_err {eh:.5f} "
                f"hidden cosine {H.cosine(hidden[off:off+T], t['final']):.6f}")
            ok &= cos >= 0.999
            off += T
        log(f"  launches {enc.launch_count}")
    return ok


def run_phase(name):
    import torch
    if name == "env":
        log(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs",
            os.cpu_count(), "cpus")
        os.system("nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; free -g | head -2")
        return True
    if name == "gemm_simt":
        return phase_gemm(1)
    if name == "gemm_tc":
        return phase_gemm(0)
    if name == "pipeline_simt":
        return phase_pipeline(1, 0)
    if name == "pipeline_tc":
        return phase_pipeline(0, 0)
    if name == "full_tc":
        return phase_pipeline(0, None)
    if name == "full_simt":
        return phase_pipeline(1, None)
    raise SystemExit("unknown phase " + name)


def main():
    if len(sys.argv) > 1:
        try:
            ok = run_phase(sys.argv[1])
        except Exception:
            traceback.print_exc()
            ok = False
        log(f"PHASE {sys.argv[1]}: {'PASS' if ok else 'FAIL'}")
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w") as fh:
        for ph in PHASES:
            fh.write(f"===== {ph} =====\n")
            fh.flush()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), ph], capture_output=True, text=True, timeout=420)
                fh.write(r.stdout[-12000:] + r.stderr[-6000:])
            except subprocess.TimeoutExpired as e:
                fh.write(f"TIMEOUT after 420 s\n{(e.stdout or b'')[-4000:]}\n{(e.stderr or b'')[-4000:]}\n")
            fh.flush()
    print(open(os.path.join(ROOT, "gpurun_out", "diag.log")).read()[-6000:])


if __name__ == "__main__":
    main()
