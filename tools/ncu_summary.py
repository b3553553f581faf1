#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key metrics + stall breakdown + per-function SASS
sample shares.  usage: tools/ncu_summary.py report.ncu-rep [out.md]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    v, u = M.get(k, ("", ""))
    return "%s %s" % (v, u)
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = ["# ncu summary: %s" % rep, "", "| metric | value |", "|---|---|"]
for k in keys:
    out.append("| %s | %s |" % (k, g(k)))
out += ["", "## warp stall reasons (pc samples)", "", "| reason | samples | share |", "|---|---|---|"]
st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v) for h, (v, u) in M.items()
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v}
tot = sum(st.values()) or 1
for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
    if v: out.append("| %s | %d | %.1f%% |" % (k, v, 100 * v / tot))
# per-function shares from the SASS source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
fn = "kernel body"
agg = {}
for r in srows:
    if r and r[0] == "Address":
        h = {name: i for i, name in enumerate(r)}; continue
    if h is None or len(r) < len(h): continue
    text = r[h["Source"]]
    samples = float(r[h["# Samples"]] or 0); inst = float(r[h["Instructions Executed"]] or 0)
    tinst = float(r[h["Thread Instructions Executed"]] or 0)
    a = agg.setdefault(fn, [0, 0, 0, 0]); a[0] += samples; a[1] += inst; a[2] += tinst; a[3] += 1
    if re.match(r"\s*RET", text): fn = "fn#%d" % (len(agg))
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
out += ["", "## SASS regions split at RET (call order: kernel body, then the non-inlined device functions)", "",
        "| region | static instr | samples | share | warp instr executed | share | avg active lanes |", "|---|---|---|---|---|---|---|"]
for k, a in agg.items():
    out.append("| %s | %d | %d | %.1f%% | %.3g | %.1f%% | %.1f |" % (k, a[3], a[0], 100 * a[0] / ts, a[1], 100 * a[1] / ti, a[2] / a[1] if a[1] else 0))
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2: open(sys.argv[2], "w").write(txt)
print(txt)
