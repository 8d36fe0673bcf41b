#!/bin/bash
# Second capture set of round 1 (mesh-shape, heightmap-shape, broadphase, C3/C4/C5 launch lists).
#   gpurun --timeout 1800 -- 'bash profiles/collect_r01b.sh'
set -x
TAG=r01
OUT=gpurun_out
mkdir -p $OUT
for W in c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_${W}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:bvhShapeCollideKernel -s 3 -c 1 -o $OUT/${TAG}_mesh_shape \
    python bench.py --workload c4 --steps 1 --warmup 3 --queries 30000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:heightmapShapeKernel -s 3 -c 1 -o $OUT/${TAG}_heightmap_shape \
    python bench.py --workload c4 --steps 1 --warmup 3 --queries 30000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bpQueryKernel -s 24 -c 1 -o $OUT/${TAG}_broadphase_query \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bpHierarchyKernel -s 24 -c 1 -o $OUT/${TAG}_broadphase_hierarchy \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT | tail -20
