#!/bin/bash
# A/B of the eager leaf-stage threshold of the mesh-shape traversal (-DFCLB_EAGER_LEAF_MIN)
for v in "" _lm4 _lm8 _lm16; do
  FCLB_LIB=$PWD/mind-fcl_b200/libfclb200$v.so timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-workloads 2>&1 | python profiles/scripts/bench_line.py "variant=${v:-base} c4"
done
