#!/bin/bash
# A/B of EPA kernel builds on one box: FCLB_LIB selects the shared library (built by /tmp/build_epa_variant.sh)
for v in "" _s128 _s256 _s512; do
  for w in c1b c1b_convex; do
    FCLB_LIB=$PWD/mind-fcl_b200/libfclb200$v.so python bench.py --workload $w --steps 5 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant=%-6s %-11s %.3e q/s  %.3f ms' % ('$v' or 'base', '$w', d['value'], d['ms_per_step']))"
  done
done
