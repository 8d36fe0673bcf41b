#!/bin/bash
# EPA tile width in double precision after the unroll-1 change
for t in 4 8 16; do
  for w in c1b c1b_convex; do
    FCLB_EPA_TILE=$t timeout 300 python bench.py --workload $w --dtype f64 --no-workloads --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('f64 tile %-3s %-10s device %.3f ms  %.3e q/s' % ('$t', '$w', d['ms_per_step'], d['value']))"
  done
done
