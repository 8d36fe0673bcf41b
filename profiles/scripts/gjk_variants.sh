#!/bin/bash
# A/B of GJK pool builds on one box: FCLB_LIB selects the shared library (built by /tmp/build_gjk_variant.sh)
for v in "" _w8p480 _w6p360; do
  for dt in f32 f64; do
    FCLB_LIB=$PWD/mind-fcl_b200/libfclb200$v.so timeout 200 python bench.py --no-cpu-baseline --no-workloads --dtype $dt --steps 8 --warmup 3 2>&1 | python profiles/scripts/bench_line.py "variant=${v:-base} $dt"
  done
done
FCLB_GJK_BINNED=1 timeout 200 python bench.py --no-cpu-baseline --no-workloads --steps 8 --warmup 3 2>&1 | python profiles/scripts/bench_line.py "per-warp pool f32"
