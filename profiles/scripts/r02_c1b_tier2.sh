#!/bin/bash
# c1b after the tile change: where the step goes (launch list) and the early tier-2 consumers again
OUT=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02_launches_c1b_end.csv python bench.py --workload c1b --no-workloads --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_c1b_end.csv")) if len(r) > 5]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]
    if "epaKernel" in k or "convexBool" in k:
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1e-3))
for k, v in agg.items(): print("%-50s %d launches, us each: %s" % (k, len(v), " ".join("%.0f" % x for x in v[:8])))
PY
for cfg in "default X=0" "early37 FCLB_EPA_EARLY_TIER2=37" "early74 FCLB_EPA_EARLY_TIER2=74" "early148 FCLB_EPA_EARLY_TIER2=148" "iters24 FCLB_EPA_TIER1_ITERS=24" "iters24early74 FCLB_EPA_TIER1_ITERS=24 FCLB_EPA_EARLY_TIER2=74" "default2 X=0"; do
  set -- $cfg; name=$1; shift
  for w in c1b; do
    env "$@" timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-14s %-10s device %.3f ms  %.3e q/s' % ('$name', '$w', d['ms_per_step'], d['value']))"
  done
done
