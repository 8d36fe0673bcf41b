#!/bin/bash
# EPA launch knobs re-swept after the unroll-1 change (the kernel is a quarter smaller): tile width, tier-1 pool, resident CTAs
for cfg in "default X=0" "tile4 FCLB_EPA_TILE=4" "tile8 FCLB_EPA_TILE=8" "tile16 FCLB_EPA_TILE=16" "faces32x5 FCLB_EPA_TIER1_FACES=32 FCLB_EPA_BLOCKS_PER_SM=5" "faces24x6 FCLB_EPA_TIER1_FACES=24 FCLB_EPA_BLOCKS_PER_SM=6" "faces48x3 FCLB_EPA_TIER1_FACES=48 FCLB_EPA_BLOCKS_PER_SM=3" "iters16 FCLB_EPA_TIER1_ITERS=16" "iters64 FCLB_EPA_TIER1_ITERS=64" "default2 X=0"; do
  set -- $cfg; name=$1; shift
  for w in c1b c1b_convex; do
    env "$@" timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-10s %-10s device %.3f ms  %.3e q/s' % ('$name', '$w', d['ms_per_step'], d['value']))"
  done
done
