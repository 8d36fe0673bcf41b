#!/bin/bash
# CTA-wide GJK pool: slots per CTA against resident CTAs per SM (240 = 4 CTAs, 208 = 4, 192 = 5, 160 = 6), one box
OUT=gpurun_out
for rep in 1 2; do
  for s in 60 52 48 40; do
    lib=mind-fcl_b200/libfclb200_s$s.so; [ $s = 60 ] && lib=mind-fcl_b200/libfclb200.so
    for dt in f32 f64; do
      FCLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 --dtype $dt --no-workloads --no-cpu-baseline --steps 10 --warmup 3 > $OUT/slots_${s}_${dt}_$rep.json 2> $OUT/slots_${s}_${dt}_$rep.err
      python - <<PY
import json
d = json.loads(open("$OUT/slots_${s}_${dt}_$rep.json").read().strip().splitlines()[-1])
print("slots/warp=$s (x4 per CTA) $dt rep $rep: value %.4e  ms/step %.3f" % (d["value"], d["ms_per_step"]))
PY
    done
  done
done
