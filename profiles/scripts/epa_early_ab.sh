#!/bin/bash
# A/B of the tier-1 iteration cap (FCLB_EPA_TIER1_ITERS) and the early tier-2 consumers (FCLB_EPA_EARLY_TIER2 = CTAs
# on the second stream; 0 = off)
for cfg in "255 0" "32 0" "32 37" "24 37" "48 37" "32 148"; do
  set -- $cfg
  for w in c1b c1b_convex; do
    FCLB_EPA_TIER1_ITERS=$1 FCLB_EPA_EARLY_TIER2=$2 timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-workloads 2>&1 | python profiles/scripts/bench_line.py "tier1_iters=$1 early_ctas=$2 $w"
  done
done
