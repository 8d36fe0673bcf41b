#!/bin/bash
# Full GPU suite, then the C1a bench line, launch list and a --set full capture of the two-phase box-box kernel.
#   gpurun --timeout 1500 -- 'bash profiles/refresh_c1a.sh'
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_full.log 2>&1; tail -5 $OUT/pytest_gpu_full.log
python bench.py --workload c1a --steps 5 --warmup 3 > $OUT/bench_c1a_f32.json 2> $OUT/bench_c1a_f32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r01_launches_c1a.csv \
    python bench.py --workload c1a --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r01_launches_c1a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:boxBoxCollideKernel -s 6 -c 2 -o $OUT/r01_boxbox_two_phase \
    python bench.py --workload c1a --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat $OUT/bench_c1a_f32.json
