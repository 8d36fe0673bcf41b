#!/bin/bash
# Round-2 collection after the kernel work of the round: default bench line + reference arm, launch lists of every
# workload (same commands under ncu's duration-only pass), ncu --set full with source of the changed kernels.
#   gpurun --timeout 2400 -- 'bash profiles/collect_r02b.sh'
set -x
TAG=r02
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python bench.py --dtype f64 --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_bench_c2_f64.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_launches_c2.log 2>&1
for W in c1a c1b c1b_convex c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_launches_${W}.log 2>&1
done
B="python bench.py --no-cpu-baseline --no-workloads --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:distanceGjkBinnedKernel -c 1 -o $OUT/${TAG}_gjk_cta_pool \
    $B --workload c2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:epaKernel -c 1 -o $OUT/${TAG}_epa_convex_b \
    $B --workload c1b_convex --queries 200000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:epaKernel -c 1 -o $OUT/${TAG}_epa_box \
    $B --workload c1b > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvhShapeCollideKernel -c 1 -o $OUT/${TAG}_mesh_shape_b \
    $B --workload c4 --queries 30000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:heightmapShapeKernel -c 1 -o $OUT/${TAG}_heightmap_shape_b \
    $B --workload c4 --queries 30000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvhCollideKernel -c 1 -o $OUT/${TAG}_bvh_collide \
    $B --workload c3 --queries 200000 > /dev/null 2>&1
ls -la $OUT | tail -30
