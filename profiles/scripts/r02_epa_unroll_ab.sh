#!/bin/bash
# EPA kernels with '#pragma unroll 1' on every pool / list loop of fclb_epa.cuh (12.0k -> 8.9k SASS instructions): A/B on one box
OUT=gpurun_out
FCLB_LIB=$PWD/mind-fcl_b200/libfclb200_u1.so timeout 600 python -m pytest tests/test_collide_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -2
for rep in 1 2; do
  for v in base u1; do
    lib=mind-fcl_b200/libfclb200.so; [ $v = u1 ] && lib=mind-fcl_b200/libfclb200_u1.so
    for w in c1b c1b_convex; do
      FCLB_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 5 --warmup 3 > $OUT/epa_${v}_${w}_$rep.json 2> $OUT/epa_${v}_${w}_$rep.err
      python - <<PY
import json
d = json.loads(open("$OUT/epa_${v}_${w}_$rep.json").read().strip().splitlines()[-1])
print("$v $w rep $rep: value %.4e  ms/step %.3f" % (d["value"], d["ms_per_step"]))
PY
    done
  done
done
