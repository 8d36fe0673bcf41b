#!/bin/bash
# Round-2 captures of the kernels VERDICT r01 names: ncu --set full with source, one launch each.
#   gpurun --timeout 1500 -- 'bash profiles/collect_r02a.sh'
set -x
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-workloads --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:epaKernel -c 1 -o $OUT/r02_epa_convex \
    $B --workload c1b_convex --queries 200000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:distanceGjkBinnedKernel -c 1 -o $OUT/r02_gjk_binned \
    $B --workload c2 --queries 3000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvhShapeCollideKernel -c 1 -o $OUT/r02_mesh_shape \
    $B --workload c4 --queries 20000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:heightmapShapeKernel -c 1 -o $OUT/r02_heightmap_shape \
    $B --workload c4 --queries 20000 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
