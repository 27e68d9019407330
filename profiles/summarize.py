#!/usr/bin/env python3
"""Turns raw ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/<x>_launches.csv  > profiles/<x>_launches.txt
  python profiles/summarize.py full     gpurun_out/<x>.ncu-rep       > profiles/<x>_full.txt

`launches`: the `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list -> per-kernel count,
total device time and SHARE of the profiled region (cold-cache, serialised: shares are what is comparable).
`full`: one `ncu --set full` capture -> the metrics the roofline argument rests on (integer-pipe issue utilisation,
DRAM bytes, occupancy, local-memory traffic, top stall reasons).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(io.StringIO("".join(lines)))
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        rows.append((name, ns, r[ix["Grid Size"]], r[ix["Block Size"]]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms of device time (serialised under ncu)")
    print(f"{'kernel':<34}{'launches':>9}{'total ms':>12}{'avg us':>12}{'share':>8}   last grid x block")
    for name, (cnt, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:<34}{cnt:>9}{ns / 1e6:>12.3f}{ns / cnt / 1e3:>12.1f}{100 * ns / total:>7.1f}%   {grid} x {block}")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.max.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.max.pct_of_peak_sustained_elapsed", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__average_warp_latency_issue_stalled",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        vals = dict(zip(hdr, row))
        print(f"## kernel {vals.get('Kernel Name', '?')}  (id {vals.get('ID', '?')})")
        for h, u in zip(hdr, units):
            if any(h == k or h.startswith(k) for k in KEYS):
                print(f"  {h:<78}{vals[h]:>18} {u}")
        stalls = [(float(vals[h].replace(",", "") or 0), h) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[h]]
        stalls.sort(reverse=True)
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for v, h in stalls[:6]:
            print(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:<40}{v:>10.3f}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
