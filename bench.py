#!/usr/bin/env python
"""bench.py -- SpeechT5-encoder audio-seconds/second on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (`config.workload`): BASELINE.json configs[1] -- the SLURP-shaped synthetic set: 70,000 utterances,
16 kHz, log-normal durations clipped to [1, 10] s (median ~2.8 s), random-init SpeechT5-base weights,
length-bucketed into padding-free batches of <= 131072 encoder frames (84 batches).  A *step* is one pass of the hot path over
one such batch; steps walk the batches in an interleaved order so any prefix is a representative length mix.
With no flags one full pass over the set is timed (K = number of batches).

  value     audio-seconds encoded per second of device time, waveforms already resident in HBM
  e2e       same metric through the public bulk API with HOST buffers (LocoSpeechT5Encoder.encode_host_pipelined ->
            loco_encode): the pinned-host -> device copy of every step's waveforms and the device -> host copy of its
            pooled embeddings happen inside the timed region (the next step's H2D overlaps the current encode)
  roofline  tcgen05 GEMM kernel: algorithmic GEMM FLOPs / its summed CUDA-event time, vs the measured bf16 peak
  cpu_baseline / --impl reference: the reference's own CPU path (HF SpeechT5 module, batch_size=2 padded
            loop, all host cores) on a bounded sample of the same workload.

N > 1 (torchrun), default `--scaling strong`: the job is the SAME utterances as at N = 1 (those of the K timed batches of the one
70k set), sharded over the ranks by length (sorted, dealt round-robin: equal FLOPs and batch shapes per rank), re-bucketed per
rank, encoded, and merged back into job order by ONE NCCL all-gather inside the timed region; `outputs.pooled_bits_checksum`
is then equal at every N.  `--scaling weak` keeps round 1's one-set-per-rank run.  time = max over ranks of device time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=0, help="timed steps (batches); 0 = one full pass over the set")
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["loco", "reference"], default="loco")
    p.add_argument("--utts", type=int, default=70000)
    p.add_argument("--max-frames", type=int, default=131072)
    p.add_argument("--e2e-steps", type=int, default=16)
    p.add_argument("--cpu-sample", type=int, default=384, help="utterances in the bounded CPU-baseline sample (~15 s of CPU work on 16 cores)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--no-stage-events", action="store_true", help="debug: time the steps without the per-launch CUDA events (no roofline)")
    p.add_argument("--gemm-impl", type=int, default=-1, help="debug: force GEMM kernel (0 single-CTA tcgen05, 2 CTA-pair tcgen05)")
    p.add_argument("--ln-impl", type=int, default=-1, help="debug: 0 deferred LayerNorm in the GEMM epilogues, 1 LayerNorm kernels")
    p.add_argument("--conv0-impl", type=int, default=-1, help="debug: force conv0 kernel (0 tcgen05, 1 mma.sync)")
    p.add_argument("--posconv-impl", type=int, default=-1, help="debug: force positional-conv kernel (0 polyphase tcgen05, 1 mma.sync, 2 one-phase tcgen05)")
    p.add_argument("--attn-impl", type=int, default=-1, help="debug: force attention kernel (0 tcgen05, 1 mma.sync)")
    p.add_argument("--scaling", choices=["strong", "weak"], default="strong",
                   help="N > 1: strong = ONE set sharded over the ranks (the product path: shard -> encode -> all-gather -> un-permute); "
                        "weak = every rank owns a full set (round-1 behaviour)")
    p.add_argument("--workload", choices=["slurp", "long30", "long60"], default="slurp",
                   help="slurp = BASELINE configs[1] (the metric's workload); long30/long60 = configs[3] (256 x 30 s / 128 x 60 s)")
    return p.parse_args()


WORKLOAD = "SpeechT5-base speech encoder, SLURP-shaped synthetic set (BASELINE.json configs[1]): {n} utts 1-10 s log-normal, masked mean-pool"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv is not None and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        reasons = [n for b, n in self.REASONS.items() if self.mask & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"tflops_sustained": p.get("bf16_tflops_sustained"), "tflops_burst": p.get("bf16_tflops"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(lengths, sample_ids, seed, repeats=1):
    """Time the reference's own CPU implementation (HF module, bs=2 padded loop, all cores) on `sample_ids`."""
    from loco_asr_b200.synth import synth_state_dict, synth_wave
    from oracle.hf_reference import build_hf_encoder, time_cpu_reference, hf_encode_unpadded
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_hf_encoder(synth_state_dict(seed=1))
    waves = [synth_wave(int(lengths[i]), seed, int(i)) for i in sample_ids]
    res = time_cpu_reference(model, waves, mode="padded_bs2", repeats=repeats)
    return model, waves, res, hf_encode_unpadded


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from loco_asr_b200.synth import slurp_shaped_lengths
    lengths = slurp_shaped_lengths(args.utts, args.seed)
    order = np.argsort(lengths, kind="stable")
    K = args.steps or 4
    W = max(args.warmup, 1)
    # bounded sample: the same `per_step` utterances (spread over the sorted length distribution) every step, so
    # oneDNN's per-shape primitive setup is paid in the warm-up, not in the timed steps (favours the reference)
    per_step = 16 if K <= 30 else max(2, int(480 / K) // 2 * 2)
    from loco_asr_b200.synth import synth_state_dict, synth_wave
    from oracle.hf_reference import build_hf_encoder, hf_encode_padded_batches
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_hf_encoder(synth_state_dict(seed=1))
    ids = order[np.linspace(0, len(order) - 1, per_step + 2).astype(int)[1:-1]]
    waves = [synth_wave(int(lengths[i]), args.seed, int(i)) for i in ids]
    for s in range(W):
        hf_encode_padded_batches(model, waves, batch_size=2)
    tot_audio, tot_t = 0.0, 0.0
    for s in range(K):
        t0 = time.perf_counter()
        hf_encode_padded_batches(model, waves, batch_size=2)
        tot_t += time.perf_counter() - t0
        tot_audio += sum(len(w) for w in waves) / 16000.0
    v = tot_audio / tot_t
    cores = torch.get_num_threads()
    sample_desc = (f"each step = the same {per_step} utterances spread over the sorted length distribution "
                   f"({tot_audio / K:.0f} audio-s per step), HF SpeechT5 module fp32, batch_size=2 padding=longest loop, warm shapes")
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * tot_t / K, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the workload of the loco arm, named the same way; how the reference runs it (CPU, batch_size=2, padding="longest",
        # bounded sample) is in cpu_baseline.sample, not in config, so that the two arms' configs compare equal on `workload`
        "config": {"workload": WORKLOAD.format(n=args.utts)},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample_desc},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ loco arm
def newest_gemm_traffic():
    """DRAM bytes per GEMM launch from the newest committed ncu capture of this same command (profiles/rNN*_gemm_traffic.json,
    written by tools/ncu_shares.py); never measured in this run -- a number taken under a profiler is not a bench value."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
    for path in reversed(files):
        try:
            with open(path) as fh:
                tj = json.load(fh)
            return float(tj["dram_bytes_per_launch"]), f"{os.path.basename(path)}: {tj['source']}"
        except (OSError, KeyError, ValueError):
            continue
    return None, None


def bits_checksum(t):
    """Order-independent checksum of an fp32 matrix: the int64 sum of its bit patterns (exact, associative), so the
    un-permuted [n, 768] result of a sharded pass can be compared with the single-GPU pass bit for bit."""
    return int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    from loco_asr_b200 import dist as ldist
    from loco_asr_b200.buckets import make_batches, interleaved_order, shard_utterances
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.flops import encoder_flops_breakdown
    from loco_asr_b200.synth import slurp_shaped_lengths, synth_state_dict, synth_waves_by_id

    rank, world, local = ldist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encoder has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    strong = args.scaling == "strong"
    dbg = args.attn_impl >= 0 or args.gemm_impl >= 0 or args.ln_impl >= 0 or args.posconv_impl >= 0 or args.conv0_impl >= 0       # kernel switches exist only in the LOCO_DEBUG build
    enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device=dev, debug=dbg)
    if args.attn_impl >= 0:
        enc.debug_set("attn_impl", args.attn_impl)
    if args.gemm_impl >= 0:
        enc.debug_set("gemm_impl", args.gemm_impl)
    if args.conv0_impl >= 0:
        enc.debug_set("conv0_impl", args.conv0_impl)
    if args.posconv_impl >= 0:
        enc.debug_set("posconv_impl", args.posconv_impl)
    if args.ln_impl >= 0:
        enc.debug_set("ln_impl", args.ln_impl)
    set_seed = args.seed if strong else args.seed + rank       # strong: ONE set for the box; weak: a set per rank
    if args.workload == "slurp":
        lengths = slurp_shaped_lengths(args.utts, set_seed)
    else:
        n_long, sec = (256, 30) if args.workload == "long30" else (128, 60)
        lengths = np.full(n_long, sec * 16000, dtype=np.int64)
        args.utts = n_long
    gbatches = make_batches(lengths, max_frames=args.max_frames)     # the set's batches: one batch = one step at N = 1
    nb = len(gbatches)
    order = interleaved_order(nb)
    W = max(args.warmup, 0)
    K = args.steps if args.steps > 0 else nb
    gsteps = [order[i % nb] for i in range(W + K)]

    # ---- the job: the utterances of the K timed global batches.  N = 1 (and weak scaling) encodes them batch by batch.
    # Strong scaling shards them over the ranks (sorted by length, dealt round-robin), every rank re-buckets its share into
    # batches of <= max_frames, encodes them, and ONE all-gather merges the pooled embeddings back into job order.
    job_utts = np.concatenate([gbatches[b] for b in gsteps[W:]])
    n_job = len(job_utts)
    if strong and world > 1:
        shard_pos = shard_utterances(lengths[job_utts], world)
        my_pos = shard_pos[rank]                                    # positions into job_utts, shortest first
        counts = [len(p) for p in shard_pos]
        my_ids = job_utts[my_pos]
        lb = make_batches(lengths[my_ids], max_frames=args.max_frames)
        timed = [(my_ids[idx], my_pos[idx]) for idx in lb]          # (utterance ids, job positions) per local batch
        warm = [timed[i % len(timed)] for i in range(W)]
    else:
        my_pos = np.arange(n_job)
        counts = [n_job]
        pos0 = np.concatenate([[0], np.cumsum([len(gbatches[b]) for b in gsteps[W:]])])
        timed = [(gbatches[b], np.arange(pos0[i], pos0[i + 1])) for i, b in enumerate(gsteps[W:])]
        warm = [(gbatches[b], None) for b in gsteps[:W]]
    wave_seed = args.seed * 1000 + (0 if strong else rank * 100003)
    cache = {}

    def batch_inputs(ids):
        key = (int(ids[0]), int(ids[-1]), len(ids))
        if key not in cache:
            cache[key] = (synth_waves_by_id(lengths[ids], ids, wave_seed, dev), np.ascontiguousarray(lengths[ids].astype(np.int32)))
        return cache[key]

    for ids, _ in warm + timed:
        batch_inputs(ids)
    my_flops = 0.0
    fcache = {}
    fl = {"gemm": 0.0, "total": 0.0, "attention": 0.0, "pos_conv": 0.0}
    for ids, _ in timed:
        for n in lengths[ids]:
            n = int(n)
            if n not in fcache:
                fcache[n] = encoder_flops_breakdown(n)
            d = fcache[n]
            fl["gemm"] += d["conv1_6"] + d["proj"] + d["qkvo"] + d["ffn"]
            fl["attention"] += d["attn"] + d["relpos"]
            fl["pos_conv"] += d["pos_conv"]
            fl["total"] += d["total"]
    my_flops = fl["total"]
    n_mine = sum(len(ids) for ids, _ in timed)
    cap = n_mine
    if world > 1 and not strong:      # weak scaling: the rank blocks of the all-gather must have one size
        c = torch.tensor([n_mine], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        cap = int(c.item())
    pooled_local = torch.zeros(cap, 768, dtype=torch.float32, device=dev)
    local_pos = np.concatenate([pos for _, pos in timed])             # job position of every local row
    row0 = np.concatenate([[0], np.cumsum([len(ids) for ids, _ in timed])])

    def run_batch(ids, out=None):
        wave, ns = batch_inputs(ids)
        return enc.encode_packed(wave, ns, out=out)

    def merge():
        if world > 1 and strong:
            return ldist.gather_pooled(pooled_local[:n_mine], local_pos, counts, n_job)
        if world > 1:       # weak scaling: every rank's own set, concatenated rank by rank
            g = torch.empty(world * cap, 768, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(g, pooled_local)
            return g
        return pooled_local

    for ids, _ in warm:
        run_batch(ids)
    if world > 1:
        merge()             # warm the communicator (NCCL connects lazily on the first collective)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    enc.profile_enable(not args.no_stage_events)
    launches0 = enc.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i, (ids, _) in enumerate(timed):
        run_batch(ids, out=pooled_local[row0[i]:row0[i + 1]])
    merged = merge()        # the single collective of the path
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.finish()
    prof = enc.profile_collect()
    enc.profile_enable(False)
    launches = enc.launch_count - launches0
    # ---- the timed outputs themselves (untimed check): every pooled embedding finite, three of the timed batches re-encoded
    # give the same bits (the path is deterministic), and the merged matrix's order-independent bit checksum -- equal at every
    # N for the same job, because an utterance's result does not depend on the batch or the rank it travelled in
    out_finite = bool(torch.isfinite(merged).all())
    recheck = sorted({0, len(timed) // 2, len(timed) - 1})
    out_repro = all(torch.equal(run_batch(timed[i][0]), pooled_local[row0[i]:row0[i + 1]]) for i in recheck)
    outputs = {"utterances": int(merged.shape[0]), "all_finite": out_finite, "bit_reproducible": out_repro,
               "rechecked_batches": len(recheck), "pooled_checksum": float(merged.double().sum()),
               "pooled_bits_checksum": bits_checksum(merged),
               "order": "job order (utterances of the timed batches, un-permuted after the all-gather)" if strong or world == 1 else "rank-major"}
    timed_audio = float(lengths[np.concatenate([ids for ids, _ in timed])].sum()) / 16000.0
    t = torch.tensor([ms, timed_audio, my_flops], dtype=torch.float64, device=dev)
    imbalance = None
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        per_rank_ms = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(per_rank_ms, t)
        ms, timed_audio = float(tmax[0]), float(tsum[1])
        imbalance = {"flops_max_over_mean": float(tmax[2] / (tsum[2] / world)),
                     "per_rank_ms": [round(float(x[0]), 3) for x in per_rank_ms],
                     "per_rank_tflop": [round(float(x[2]) / 1e12, 3) for x in per_rank_ms]}
    value = timed_audio / (ms / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): algorithmic GEMM FLOPs / summed event time (this rank's) ---------
    peaks = load_peaks()
    gemm_ms, gemm_n = prof["gemm"]
    gemm_traffic, traffic_src = newest_gemm_traffic()
    achieved = fl["gemm"] / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak = peaks["tflops_sustained"]
    my_ms = float(t[0])
    roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel (CTA-pair tcgen05/TMEM/TMA bf16 GEMM, cta_group::2, all epilogues)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "traffic": gemm_traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)" if gemm_traffic else None,
                "traffic_source": traffic_src, "launches": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                "algorithmic_flops_per_launch": fl["gemm"] / max(gemm_n, 1),
                "share_of_step": gemm_ms / my_ms if my_ms else None}
    stage_ms = {k: round(v[0] / max(len(timed), 1), 4) for k, v in prof.items()}
    whole = {"tflops_per_gpu": fl["total"] / (my_ms / 1e3) / 1e12,
             "frac_of_sustained_peak": fl["total"] / (my_ms / 1e3) / 1e12 / peak if peak else None,
             "frac_of_burst_peak": fl["total"] / (my_ms / 1e3) / 1e12 / peaks["tflops_burst"] if peaks["tflops_burst"] else None}

    # ---- e2e: host buffers through the public bulk API, same sharding, same merge ----------------------------------------
    E = min(args.e2e_steps, len(timed)) if args.e2e_steps > 0 else len(timed)
    e2e_b = list(range(E)) if not (strong and world > 1) else list(range(len(timed)))[:max(1, (args.e2e_steps + world - 1) // world)]
    host_w = {i: batch_inputs(timed[i][0])[0].cpu().pin_memory() for i in e2e_b}
    host_p = {i: torch.empty(len(timed[i][0]), 768, dtype=torch.float32).pin_memory() for i in e2e_b}
    lens_of = {i: batch_inputs(timed[i][0])[1] for i in e2e_b}
    # the bulk-extraction call a user makes (extract.py uses it too): host batches in, pooled host tensors out, with batch
    # i+1's H2D copy overlapping batch i's encode.  Every step's waveform H2D and pooled D2H happen inside the timed region.
    for _ in enc.encode_host_pipelined([(host_w[i], lens_of[i], host_p[i]) for i in e2e_b[:2]]):
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_out = 0
    for out in enc.encode_host_pipelined([(host_w[i], lens_of[i], host_p[i]) for i in e2e_b]):
        n_out += out.shape[0]
    e1.record()
    torch.cuda.synchronize()
    assert n_out == sum(len(timed[i][0]) for i in e2e_b)
    e2e_ms = e0.elapsed_time(e1)
    e2e_audio = sum(float(lens_of[i].sum()) for i in e2e_b) / 16000.0
    t2 = torch.tensor([e2e_ms, e2e_audio], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t2.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t2.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_ms, e2e_audio = float(tmax[0]), float(tsum[1])
    e2e = {"value": e2e_audio / (e2e_ms / 1e3), "unit": "audio-s/s", "steps": len(e2e_b),
           "h2d_bytes_per_step": int(np.mean([host_w[i].numel() * 4 for i in e2e_b])),
           "d2h_bytes_per_step": int(np.mean([host_p[i].numel() * 4 for i in e2e_b])),
           "api": "LocoSpeechT5Encoder.encode_host_pipelined (loco_encode through the C ABI; pinned host buffers, H2D of batch i+1 overlapped with the encode of batch i)"}

    # ---- CPU baseline (rank 0, N = 1 only) + parity spot check ---------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b0 = e2e_b[0]
        ids_in_batch = np.linspace(0, len(lens_of[b0]) - 1, 4).astype(int)
        order_all = np.argsort(lengths, kind="stable")
        sample_ids = order_all[np.linspace(0, len(order_all) - 1, args.cpu_sample).astype(int)]
        model, _, res, hf_unpadded = cpu_reference_run(lengths, sample_ids, args.seed)
        cpu_baseline = {"value": res["audio_s_per_s"], "unit": "audio-s/s", "cores": res["cores"], "kind": "reference",
                        "sample": f"{args.cpu_sample} utterances spread over the sorted length distribution "
                                  f"({res['audio_s']:.0f} audio-s, {res['seconds']:.1f} s of CPU), HF SpeechT5 module fp32, "
                                  "reference-style batch_size=2 padding=longest loop"}
        # parity of the timed GPU output against the reference module run unpadded on the same waveforms
        cu = np.concatenate([[0], np.cumsum(lens_of[b0])])
        hw = host_w[b0].numpy()
        cos = []
        for i in ids_in_batch:
            ref = hf_unpadded(model, [hw[cu[i]:cu[i + 1]]])[0].mean(0)
            cos.append(float(torch.nn.functional.cosine_similarity(ref, host_p[b0][i], dim=0)))
        parity = {"min_pooled_cosine": min(cos), "n": len(cos), "against": "HF SpeechT5 module fp32, unpadded"}

    if rank == 0:
        if args.workload == "slurp":
            workload = WORKLOAD.format(n=args.utts)
            if world > 1:
                workload += (f"; utterance-sharded across {world} B200 with one NCCL all-gather of pooled embeddings "
                             "(BASELINE.json configs[2] shape, configs[4] extraction path)")
        else:
            workload = f"SpeechT5-base speech encoder, long-context segments (BASELINE.json configs[3]): {args.utts} x {int(lengths[0]) // 16000} s"
        print(json.dumps({
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload,
                       "job": (f"the {n_job} utterances of the {K} timed batches of the one {args.utts}-utterance set; "
                               + (f"sorted by length, dealt round-robin to {world} ranks, re-bucketed per rank into {len(timed)} batches, merged un-permuted by one all-gather"
                                  if strong and world > 1 else "encoded batch by batch" + ("" if world == 1 else " (every rank its own set: weak scaling)"))),
                       "utterances_in_set": int(args.utts), "batches_in_set": nb, "local_batches": len(timed),
                       "max_frames_per_batch": args.max_frames, "audio_s_per_step": timed_audio / K,
                       "weights": "random-init SpeechT5-base (seed 1)", "accumulate": "fp32",
                       "l2": "inputs larger than L2 (per-batch working set ~20 GB)",
                       "parallelism": f"dp{world} utterance-sharded, one all-gather of pooled embeddings"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "stage_ms_per_step": stage_ms, "whole_step": whole, "parity": parity, "outputs": outputs,
            "shard_balance": imbalance,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
