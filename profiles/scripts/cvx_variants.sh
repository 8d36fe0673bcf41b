#!/bin/bash
# A/B of the Convex support variants (libraries built with -DFCLB_CVX_EAGER / -DFCLB_CVX_UNROLL)
for v in _old _plain _oldvec _old _plain; do
  for w in c4 c1b_convex; do
    FCLB_LIB=$PWD/mind-fcl_b200/libfclb200$v.so timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-workloads 2>&1 | python profiles/scripts/bench_line.py "variant=${v:-base} $w"
  done
done
