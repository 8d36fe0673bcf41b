#!/bin/bash
# early tier-2 consumers beside tier 1, after the tile change: Convex pairs and double precision
for cfg in "default X=0" "early37 FCLB_EPA_EARLY_TIER2=37" "early74 FCLB_EPA_EARLY_TIER2=74" "early148 FCLB_EPA_EARLY_TIER2=148" "iters24early148 FCLB_EPA_TIER1_ITERS=24 FCLB_EPA_EARLY_TIER2=148"; do
  set -- $cfg; name=$1; shift
  for w in "c1b_convex f32" "c1b f64" "c1b_convex f64"; do
    set -- "$@"; ww=${w% *}; dt=${w#* }
    env "$@" timeout 300 python bench.py --workload $ww --dtype $dt --no-workloads --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-15s %-10s %s device %.3f ms  %.3e q/s' % ('$name', '$ww', '$dt', d['ms_per_step'], d['value']))"
  done
done
