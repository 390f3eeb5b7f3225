"""Warp-stall samples of one profiled kernel by SOURCE LINE of its own .cu file.

ncu's `--page source --csv` lists SASS instructions with their stall samples but without line numbers; `nvdisasm -g` lists the same
instructions in the same order with `//## File ..., line N` markers.  This joins the two by instruction order and charges every
instruction to the innermost line of the kernel's own file in its inline chain (helpers from common.cuh are charged to their call site).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object file> <source file name> [top N]
e.g. python tools/ncu_lines.py gpurun_out/r02v_attn_tc_2999.ncu-rep attention_tc loco_asr_b200/csrc/build/attention_tc.o attention_tc.cu 40
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, kre, obj, src = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
i_src, i_n, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sass = []
for r in rows[hdr_i + 1:]:
    if len(r) <= i_n or r[0] == "Address":
        break
    sass.append((r[i_src].strip(), int(r[i_n] or 0), int(r[i_ex] or 0), {h: int(r[i] or 0) for h, i in stall_cols}))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kre, l))
lines, cur, group = [], 0, []
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        if lines:
            break
        continue
    if "//## File" in l:
        group += re.findall(r'"([^"]+)", line (\d+)', l)      # innermost first; "inlined at" parts follow on the same / next marker lines
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        own = [int(n) for f, n in group if f.endswith(src)]
        if own:
            cur = own[0]                                        # the innermost line of the kernel's own file
        group = []
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} in the report; joining by order anyway", file=sys.stderr)
by = {}
for (text, n, ex, st), ln in zip(sass, lines):
    d = by.setdefault(ln, {"n": 0, "ex": 0, "st": {}})
    d["n"] += n
    d["ex"] += ex
    for k, v in st.items():
        d["st"][k] = d["st"].get(k, 0) + v
total = sum(d["n"] for d in by.values())
srcpath = next((m.group(1) for l in dis for m in [re.search(r'//## File "([^"]+)"', l)] if m and m.group(1).endswith(src)), None)
text = open(srcpath).read().splitlines() if srcpath and os.path.exists(srcpath) else []
print(f"# {rep}: {kre}: {total} samples, {len(sass)} instructions")
print("line,samples,share,instructions_executed,top_stalls,source")
for ln, d in sorted(by.items(), key=lambda kv: -kv[1]["n"])[:top]:
    st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:3]
    s = " ".join(f"{k[6:]}={v}" for k, v in st if v)
    code = text[ln - 1].strip()[:90] if 0 < ln <= len(text) else ""
    print(f"{ln},{d['n']},{d['n'] / max(total, 1):.3f},{d['ex']},{s},\"{code}\"")
