#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of the default bench command and
# one --set full capture per top kernel.  Outputs go to gpurun_out/ (scratch); the
# summaries committed under profiles/ are produced from them by profiles/summarize.py.
#   gpurun --timeout 1800 -- 'bash profiles/collect.sh r01'
set -x
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:distanceGjkKernel -s 6 -c 2 -o $OUT/${TAG}_gjk_distance \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:distanceClosedKernel -s 3 -c 1 -o $OUT/${TAG}_closed_distance \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_c1b_convex.csv \
    python bench.py --workload c1b_convex --steps 1 --warmup 3 --queries 200000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:epaKernel -s 2 -c 1 -o $OUT/${TAG}_epa_convex \
    python bench.py --workload c1b_convex --steps 1 --warmup 3 --queries 200000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:collideClosedKernel -s 3 -c 1 -o $OUT/${TAG}_boxbox \
    python bench.py --workload c1a --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT
