#!/usr/bin/env python
"""Per-source-line roll-up of an ncu capture (kernels compiled with -lineinfo): the SASS page of the report is joined with
nvdisasm --print-line-info of the same cubin by instruction order, then instructions executed, average active lanes and
stall samples are summed per source file, per function-sized line range and per hottest line.
  python profiles/src_summary.py <report.ncu-rep> <cubin> <mangled kernel name substring> [top_n]
cubins: cuobjdump -xelf all mind-fcl_b200/libfclb200.so"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_rows(rep):
    p = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                       stderr=subprocess.DEVNULL, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    hdr, out, name = None, [], ""
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = r[1]
        elif r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and len(r) >= len(hdr):
            out.append(r)
    return name, hdr, out


def line_map(cubin, kernel_sub):
    """[(opcode text, file, line)] for the kernel's .text section, in address order"""
    p = subprocess.run(["nvdisasm", "-c", "--print-line-info", cubin], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    inside, cur, out = False, ("?", 0), []
    for ln in p.stdout.splitlines():
        if ln.startswith("//--------------------- .text."):
            inside = kernel_sub in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), cur[0], cur[1]))
    return out


def main():
    rep, cubin, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    name, hdr, rows = sass_rows(rep)
    lm = line_map(cubin, ksub)
    print(f"# {rep.split('/')[-1]}  kernel: {name[:140]}")
    print(f"# SASS rows in the report {len(rows)}, instructions in the cubin's section {len(lm)}")
    n = min(len(rows), len(lm))
    by_file = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    by_line = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    ci, ti, si, ni = hdr["Instructions Executed"], hdr["Thread Instructions Executed"], hdr["# Samples"], hdr.get("stall_no_inst")
    mismatched = 0
    for k in range(n):
        r = rows[k]
        op_ncu = r[hdr["Source"]].split()[0] if r[hdr["Source"]].split() else ""
        op_dis = lm[k][0].split()[0] if lm[k][0].split() else ""
        if op_ncu.lstrip("@!P0123456789T") != op_dis.lstrip("@!P0123456789T") and op_ncu != op_dis:
            mismatched += 1
        v = [float(r[ci] or 0), float(r[ti] or 0), float(r[si] or 0), float(r[ni] or 0) if ni is not None else 0.0]
        for d, key in ((by_file, lm[k][1]), (by_line, (lm[k][1], lm[k][2]))):
            for j in range(4):
                d[key][j] += v[j]
    tot_i = sum(v[0] for v in by_file.values()) or 1
    tot_t = sum(v[1] for v in by_file.values())
    tot_s = sum(v[2] for v in by_file.values()) or 1
    print(f"# order check: {mismatched} of {n} opcodes differ between the two listings")
    print(f"warp instructions executed {tot_i:.4e}; average active lanes {tot_t / tot_i:.1f}; stall samples {tot_s:.0f}")
    print("--- by source file: % of instructions | avg lanes | % of stall samples | stall_no_inst share of its samples")
    for k, v in sorted(by_file.items(), key=lambda x: -x[1][0]):
        if v[0] / tot_i >= 0.002:
            print(f"{100 * v[0] / tot_i:6.2f} %  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / tot_s:6.2f} %  no_inst {100 * v[3] / max(v[2], 1):5.1f} %  {k}")
    # ranges of 25 source lines approximate functions
    by_range = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for (f, l), v in by_line.items():
        for j in range(4):
            by_range[(f, l // 25 * 25)][j] += v[j]
    print("--- by 25-line source range")
    for k, v in sorted(by_range.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{100 * v[0] / tot_i:6.2f} %  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / tot_s:6.2f} %  no_inst {100 * v[3] / max(v[2], 1):5.1f} %  {k[0]}:{k[1]}-{k[1] + 24}")


if __name__ == "__main__":
    main()
