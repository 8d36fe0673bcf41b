set -x
OUT=gpurun_out; mkdir -p $OUT
for W in c3 c4; do python bench.py --workload $W --steps 5 --warmup 3 > $OUT/bench_${W}_f32.json 2> $OUT/bench_${W}_f32.err; done
python bench.py --workload c3 --dtype f64 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_f64.json 2>/dev/null
python bench.py --workload c4 --dtype f64 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_c4_f64.json 2>/dev/null
for W in c3 c4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r01_launches_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r01_launches_${W}.log 2>&1
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
