#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, and per-source-line instruction / stall totals.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_symbols]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
nsym = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__inst_executed_op_branch.sum", "sm__cycles_elapsed.max",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        v, u = m[k]
        extra = ""
        if nsym and u in ("inst", "", "byte", "Gbyte", "Mbyte") and k.endswith(".sum"):
            try:
                f = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6}.get(u, 1)
                extra = f"   [{f / nsym:.1f} per symbol]"
            except ValueError:
                pass
        print(f"{k:75s} {v} {u}{extra}")
st = sorted(((float(v[0]), k) for k, v in m.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and "pct" not in k), reverse=True)
tot = sum(x for x, _ in st) or 1
print("\nstall samples:")
for x, k in st[:12]:
    print(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {x:10.0f} {100 * x / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
rows = [r for r in rows if len(r) > 6]
if rows:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x == name:
                return i
        return None
    ci, cs, cx, csm = col("Source"), col("# Samples") or col("Sampling Data (All)"), col("Instructions Executed"), col("Address")
    # aggregate by opcode
    byop = collections.Counter(); samp = collections.Counter()
    for r in rows[1:]:
        try:
            op = r[ci].split()[0] if not r[ci].startswith("@") else r[ci].split()[1]
            op = op.split(".")[0]
            n = float(r[cx] or 0); s = float(r[cs] or 0)
        except Exception:
            continue
        byop[op] += n; samp[op] += s
    T = sum(byop.values()) or 1; S = sum(samp.values()) or 1
    print("\nexecuted warp-instructions by opcode (top 25):")
    for op, n in byop.most_common(25):
        extra = f"  [{n / nsym:.1f}/symbol]" if nsym else ""
        print(f"  {op:10s} {n:14.0f} {100 * n / T:5.1f}%   samples {100 * samp[op] / S:5.1f}%{extra}")
