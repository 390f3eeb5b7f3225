"""Summarise an `ncu --set full` report into the per-launch table kept under profiles/ (one row per captured launch).

    python tools/ncu_full_summary.py gpurun_out/r03j_full.ncu-rep profiles/r03j_ncu_full_summary.csv "header note"
"""
import csv, io, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
COLS = [
    ("duration_us", "gpu__time_duration.sum", 1e-3 if "ns" in units[hdr.index("gpu__time_duration.sum")] else (1.0 if "us" in units[hdr.index("gpu__time_duration.sum")] else 1e3)),
    ("tensor_pipe_active_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("issue_active_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("lsu_smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    ("tmem_busy_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("regs_per_thread", "launch__registers_per_thread", 1),
    ("dram_read_MB", "dram__bytes_read.sum", None),
    ("dram_write_MB", "dram__bytes_write.sum", None),
    ("dram_throughput_pct", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("lts_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1),
]


def short(name):
    name = name.replace("(int)", "").replace("(EpilogueKind)", "")
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("unnamed>::", "").replace("loco::", "").replace("<", "", 1) if name.startswith("<") else name.replace("unnamed>::", "").replace("loco::", "")
    return re.sub(r"^void ", "", name)


def val(r, metric, scale):
    if metric not in hdr:
        return ""
    i = hdr.index(metric)
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return ""
    if scale is None:          # bytes with a unit column
        u = units[i]
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    else:
        v *= scale
    return f"{v:.2f}"


with open(out, "w") as f:
    f.write(f"# {note}\n# source: ncu --set full --clock-control none ({rep.split('/')[-1]}); per-launch times are cold-cache and serialised\n")
    f.write("kernel,grid," + ",".join(c[0] for c in COLS) + "\n")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = short(r[hdr.index("Kernel Name")])
        f.write(f'{k},"{r[hdr.index("launch__grid_size")]}",' + ",".join(val(r, m, s) for _, m, s in COLS) + "\n")
print("wrote", out)
