#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu -i REP --page raw --csv) into the few counters DESIGN.md argues from.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/summary.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA (inst) %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe FMA (cycles) %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe FMA-heavy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "pipe FMA-heavy (cycles) %"),
    ("sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "pipe FMA-lite %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU (MUFU) %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe FP64 %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]
stall_keys = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
print(f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none)\n")
for d in data:
    print(f"## {d[idx['Kernel Name']]}\n")
    print("| counter | value |\n|---|---|")
    for key, label in WANT:
        if key in idx and d[idx[key]] not in ("", "n/a"):
            print(f"| {label} (`{key}`) | {d[idx[key]]} {units[idx[key]]} |")
    stalls = sorted(((float(d[idx[k]].replace(",", "")), k) for k in stall_keys if d[idx[k]] not in ("", "n/a")), reverse=True)
    top = ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls[:6])
    print(f"| top stall reasons (warps per issue-active cycle) | {top} |\n")
