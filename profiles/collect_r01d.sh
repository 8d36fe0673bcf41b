#!/bin/bash
# Round-1 final collection (after the EPA parallel-cone work and the scene-pair kernels):
# bench lines of every workload, launch lists of the same commands, ncu --set full of the new kernels.
#   gpurun --timeout 2400 -- 'bash profiles/collect_r01d.sh'
set -x
TAG=r01
OUT=gpurun_out
mkdir -p $OUT
bash profiles/final_bench.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_c2.log 2>&1
for W in c1a c1b c1b_convex c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_${W}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:scenePairKernel -s 2 -c 1 -o $OUT/${TAG}_scene_pair \
    python -m pytest tests/test_scene_pair_gpu.py -x -q -k "test_scene_pairs and float32" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:epaKernel -c 1 -o $OUT/${TAG}_epa_box_tile16 \
    python bench.py --workload c1b --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT | tail -30
