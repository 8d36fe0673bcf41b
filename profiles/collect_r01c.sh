#!/bin/bash
# Final launch lists of round 1 (same commands as the bench lines), one per workload.
#   gpurun --timeout 1800 -- 'bash profiles/collect_r01c.sh'
set -x
TAG=r01
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_c2.log 2>&1
for W in c1a c1b c1b_convex c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_${W}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:octreeShapeKernel -c 1 -o $OUT/${TAG}_octree_shape \
    python -m pytest tests/test_octree_gpu.py -x -q > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvhShapeCollideKernel -s 3 -c 1 -o $OUT/${TAG}_mesh_shape \
    python bench.py --workload c4 --steps 1 --warmup 3 --queries 30000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:heightmapShapeKernel -s 3 -c 1 -o $OUT/${TAG}_heightmap_shape \
    python bench.py --workload c4 --steps 1 --warmup 3 --queries 30000 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bpQueryKernel -s 24 -c 1 -o $OUT/${TAG}_broadphase_query \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT | tail -12
