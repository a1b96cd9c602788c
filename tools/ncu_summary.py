#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the few numbers the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def col(r, name):
    best = ("n/a", "")
    for h, i in idx.items():
        if h == name or h.endswith("." + name):
            if r[i] not in ("", "no data", "n/a"):
                return r[i], units[i]
            best = (r[i] or "n/a", units[i])
    return best
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active"]
print(f"# ncu summary of `{rep}`\n")
for r in rows[2:]:
    print(f"## {r[idx['Kernel Name']][:90]}  (id {r[idx['ID']]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        v, u = col(r, m)
        print(f"| {m} | {v} | {u} |")
    print()
