#!/bin/bash
# Round-2 final collection: the full GPU test suite with a fresh parity log, smoke(), default bench + reference arm,
# launch lists of every workload.   gpurun --timeout 2400 -- 'bash profiles/collect_r02c.sh'
set -x
TAG=r02
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_r02.jsonl
python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu_final.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $OUT/${TAG}_smoke_final.log 2>&1; tail -2 $OUT/${TAG}_smoke_final.log
python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python bench.py --dtype f64 --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_bench_c2_f64.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_launches_c2.log 2>&1
for W in c1a c1b c1b_convex c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $OUT/${TAG}_launches_${W}.log 2>&1
done
ls -la $OUT | tail -12
