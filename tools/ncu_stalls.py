#!/usr/bin/env python
"""Stall breakdown + hottest source lines of one .ncu-rep: python tools/ncu_stalls.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:60], "us", d.get("gpu__time_duration.sum"))
    items = [(k, d[k]) for k in hdr if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
    items = [(k, float(v.replace(",", ""))) for k, v in items if v not in ("", "n/a")]
    tot = sum(v for _, v in items) or 1
    print("  stalls:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for k, v in sorted(items, key=lambda kv: -kv[1])[:8]))
    for k in ["sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__occupancy_limit_shared_mem",
              "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
        if k in d: print("  ", k, d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        cur = r[1][:60]
        hdr = rows[i + 1]
        si = hdr.index("# Samples")
        body = []
        for q in rows[i + 2:]:
            if q and q[0] == "Kernel Name": break
            if len(q) > si: body.append(q)
        val = lambda q: float(q[si]) if q[si].replace(".", "").isdigit() else 0.0
        tot = sum(val(q) for q in body) or 1
        print("== SASS hot spots of", cur, "(share of samples, cumulative position)")
        idx = {id(q): k for k, q in enumerate(body)}
        for q in sorted(body, key=val, reverse=True)[:nl]:
            print(f"   {100 * val(q) / tot:5.1f}%  #{idx[id(q)]:5d}  {q[1].strip()[:110]}")
