"""Turn an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
bench.py into the files kept under profiles/: the per-launch list, one per-kernel share table per captured step, and
the GEMM DRAM-traffic figure bench.py's roofline.traffic reads.

    python tools/ncu_shares.py gpurun_out/r01e_launches_raw.csv profiles/r01e "code state note"
"""
import csv
import json
import re
import sys
from collections import OrderedDict

raw, prefix = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
KERNELS = "row_frames|wave_moments|gn_finalize|conv0_tc|gemm_tc|layernorm_kernel|posconv_pp|prenet_ln|attention|final_ln_pool"
CMD = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
       f"-k 'regex:{KERNELS}' --launch-skip 70 -c 210 --csv python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline")


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("unnamed>::", "").replace("loco::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    return name.replace("(EpilogueKind)", "")


lines = [l for l in open(raw) if l.startswith('"')]
launches = OrderedDict()
for row in csv.DictReader(lines):
    k = int(row["ID"])
    d = launches.setdefault(k, {"kernel": short(row["Kernel Name"]), "grid": row["Grid Size"]})
    d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
rows = list(launches.values())
with open(prefix + "_launches_ncu.csv", "w") as f:
    f.write("idx,kernel,grid,duration_us,dram_read_MB,dram_write_MB\n")
    for i, r in enumerate(rows):
        f.write(f'{i},{r["kernel"]},"{r["grid"]}",{r["gpu__time_duration.sum"] / 1e3:.1f},'
                f'{r["dram__bytes_read.sum"] / 1e6:.2f},{r["dram__bytes_write.sum"] / 1e6:.2f}\n')

starts = [i for i, r in enumerate(rows) if r["kernel"].startswith("row_frames_kernel")] + [len(rows)]
gemm_bytes, gemm_n = 0.0, 0
for si, tag in zip(range(len(starts) - 1), "abcdef"):
    step = rows[starts[si]:starts[si + 1]]
    if not any(r["kernel"].startswith("final_ln_pool") for r in step):
        continue  # truncated step at the end of the capture
    agg = OrderedDict()
    for r in step:
        a = agg.setdefault(r["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["gpu__time_duration.sum"] / 1e3
        a[2] += r["dram__bytes_read.sum"] / 1e6
        a[3] += r["dram__bytes_write.sum"] / 1e6
        if r["kernel"].startswith("gemm_tc2_kernel"):
            gemm_bytes += r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]
            gemm_n += 1
    tot = [sum(a[j] for a in agg.values()) for j in range(4)]
    grids = sorted({r["grid"] for r in step if r["kernel"].startswith("attention")})
    with open(f"{prefix}_step_shares_{tag}.csv", "w") as f:
        f.write(f"# ncu launch list, one step of bench.py (cold-cache, serialised); attention grids in this step: {', '.join(grids)}\n")
        f.write(f"# command: {CMD}\n# code state: {note}\n")
        f.write("kernel,launches,total_us,share,dram_read_MB,dram_write_MB\n")
        for k, a in agg.items():
            f.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot[1]:.4f},{a[2]:.1f},{a[3]:.1f}\n")
        f.write(f"TOTAL,{tot[0]},{tot[1]:.1f},1.0,{tot[2]:.1f},{tot[3]:.1f}\n")
if gemm_n:
    json.dump({"source": f"{prefix}_step_shares_*.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum over the {gemm_n} "
                         "gemm_tc2_kernel launches of the captured steps)",
               "gemm_launches": gemm_n, "dram_bytes_per_launch": gemm_bytes / gemm_n},
              open(prefix + "_gemm_traffic.json", "w"), indent=1)
print("steps:", len(starts) - 1, "gemm launches:", gemm_n)
