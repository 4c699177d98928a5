#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (one row per profiled launch) into the handful of metrics the roofline
discussion in DESIGN.md uses: duration, tensor / XU / FMA / ALU pipe activity, issue-slot use, DRAM / L2 / shared-memory
traffic, registers, occupancy, the top warp-stall reasons.

    ncu -i gpurun_out/att_r1c.ncu-rep --page raw --csv > gpurun_out/att_r1c_raw.csv
    python tools/ncu_summary.py gpurun_out/att_r1c_raw.csv [--group]      # --group: average launches of the same kernel
"""
import csv
import sys
from collections import OrderedDict, defaultdict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (LSU)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem wavefronts (tensor)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp insts"),
])
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
         "nsecond": 1e-3, "second": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    return hdr, units, [r for r in data if len(r) == len(hdr)]


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    path = sys.argv[1]
    group = "--group" in sys.argv
    hdr, units, data = load(path)
    col = {h: i for i, h in enumerate(hdr)}
    ki = col["Kernel Name"]
    stall_cols = [(h, i) for h, i in col.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    groups = defaultdict(list)
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "")
        groups[name if group else f"{r[0]}:{name}"].append(r)
    for name, rs in groups.items():
        print(f"== {name}  ({len(rs)} launch{'es' if len(rs) > 1 else ''})")
        for m, label in METRICS.items():
            if m not in col:
                continue
            i = col[m]
            vals = [v for v in (fnum(r[i]) for r in rs) if v is not None]
            if not vals:
                continue
            u = units[i]
            v = sum(vals) / len(vals)
            if u in SCALE:
                v *= SCALE[u]
                u = "B" if "byte" in u else "us"
            sv = f"{v:,.1f}" if abs(v) < 1e6 else f"{v:,.0f}"
            print(f"   {label:26s} {sv:>18s} {u:8s} {m}")
        stalls = []
        for h, i in stall_cols:
            vals = [v for v in (fnum(r[i]) for r in rs) if v is not None]
            if vals:
                stalls.append((sum(vals) / len(vals),
                               h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        stalls.sort(reverse=True)
        if stalls:
            print("   top stalls (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main()
