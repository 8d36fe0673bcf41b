#!/usr/bin/env python
"""Turns the .ncu-rep captures in gpurun_out/ into the small text summaries committed
under profiles/ (and profiles/ncu_traffic.json, which bench.py reads for roofline.traffic).
Run here (no GPU needed):  python profiles/summarize.py r01"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def raw(rep):
    p = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    v = float(val)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    traffic = {}
    for name in sorted(os.listdir(OUT)):
        if not (name.startswith(tag + "_") and name.endswith(".ncu-rep")):
            continue
        hdr, units, rows = raw(os.path.join(OUT, name))
        idx = {h: i for i, h in enumerate(hdr)}
        lines = [f"# {name}: ncu --set full --clock-control none (per-launch, cold-cache, serialised)"]
        for r in rows:
            lines.append("")
            lines.append("kernel: " + r[idx["Kernel Name"]])
            for k in KEYS:
                if k in idx:
                    lines.append(f"  {k} = {r[idx[k]]} {units[idx[k]]}")
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            lines.append(f"  dram traffic per launch = {rd + wr:.0f} bytes")
            traffic.setdefault(name.replace(".ncu-rep", ""), []).append(
                {"kernel": r[idx["Kernel Name"]], "dram_bytes": rd + wr, "duration": r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]]})
        with open(os.path.join(PROF, name.replace(".ncu-rep", ".summary.txt")), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", name.replace(".ncu-rep", ".summary.txt"))
    for name in sorted(os.listdir(OUT)):
        if name.startswith(tag + "_launches") and name.endswith(".csv"):
            rows = [r for r in csv.reader(open(os.path.join(OUT, name))) if len(r) > 5]
            hdr = rows[0]
            ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
            agg = {}
            for r in rows[1:]:
                k = r[ki].split("(")[0]
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1
                a[1] += float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1e-3)
            tot = sum(a[1] for a in agg.values())
            with open(os.path.join(PROF, name.replace(".csv", ".summary.txt")), "w") as f:
                f.write(f"# {name}: ncu --metrics gpu__time_duration.sum --clock-control none; share of listed launches\n")
                for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                    f.write(f"{100 * a[1] / tot:6.2f} %  {a[1]:12.1f} us  {a[0]:4d} launches  {k}\n")
            os.system(f"cp {os.path.join(OUT, name)} {os.path.join(PROF, name)}")
            print("wrote", name)
    # merge into the committed table: gpurun_out/ only holds the captures of the latest calls
    path = os.path.join(PROF, f"{tag}_ncu_traffic_raw.json")
    merged = json.load(open(path)) if os.path.exists(path) else {}
    merged.update(traffic)
    with open(path, "w") as f:
        json.dump(merged, f, indent=1)


if __name__ == "__main__":
    main()
