#!/usr/bin/env python
"""Condense an .ncu-rep (or an ncu launch-list csv) into the text summaries committed under profiles/.

    python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep   > profiles/r01_<what>.txt
    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/r01_launches.txt

Runs on the CPU box (ncu -i needs no GPU).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__cluster_max_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.max.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL_PREFIX = "smsp__pcsamp_warps_issue_stalled_"


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {r[name_i][:110]}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"   {k:75s} {d[k]:>16s} {u[k]}")
        stalls = [(h[len(STALL_PREFIX):], float(d[h].replace(",", ""))) for h in hdr
                  if h.startswith(STALL_PREFIX) and not h.endswith("_not_issued") and d[h] not in ("", "n/a")]
        tot = sum(v for _, v in stalls) or 1.0
        top = sorted(stalls, key=lambda kv: -kv[1])[:8]
        print("   warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top))
        print()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows:
        d.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")) / 1e3)
    total = sum(sum(v) for v in d.values())
    print(f"{'kernel':72s} {'launches':>8s} {'mean us':>9s} {'min us':>9s} {'share':>7s}   (ncu gpu__time_duration, serialised, cold cache)")
    for k, v in d.items():
        print(f"{k[:72]:72s} {len(v):8d} {sum(v) / len(v):9.1f} {min(v):9.1f} {100 * sum(v) / total:6.1f}%")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])
