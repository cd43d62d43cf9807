#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (text): launch list shares and the key raw metrics of the
captured kernels. Usage: python tools/ncu_summary.py <launches.csv> <report.ncu-rep> <out-prefix>"""
import collections
import csv
import subprocess
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0][:70], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# {'kernel':70s} {'n':>4s} {'avg_us':>10s} {'total_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"  {k:70s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e3:10.1f} {100*sum(v)/tot:6.1f}%\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum"]


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one launch per kernel)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']]}  grid {r[idx.get('launch__grid_size', 0)]}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"   {w:78s} {r[idx[w]]:>18s} {units[idx[w]]}\n")
            tens = [h for h in hdr if "tensor" in h and "dmma" in h.lower()]
            for h in tens[:6]:
                if h not in WANT:
                    f.write(f"   {h:78s} {r[idx[h]]:>18s} {units[idx[h]]}\n")


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    raw(sys.argv[2], sys.argv[3] + "_kernels.txt")
