#!/bin/bash
# '#pragma unroll 1' on every variable-trip loop of the device code (all kernels): parity with the variant, then A/B per workload
OUT=gpurun_out
V=$PWD/mind-fcl_b200/libfclb200_u1all.so
FCLB_LIB=$V timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/u1all_pytest.log 2>&1; echo "suite with the variant rc $?: $(tail -1 $OUT/u1all_pytest.log)"
for rep in 1 2; do
  for v in base u1all; do
    lib=$PWD/mind-fcl_b200/libfclb200.so; [ $v = u1all ] && lib=$V
    FCLB_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > $OUT/u1all_${v}_$rep.json 2> $OUT/u1all_${v}_$rep.err
    FCLB_LIB=$lib timeout 300 python bench.py --dtype f64 --no-workloads --no-cpu-baseline --steps 5 --warmup 3 > $OUT/u1all_${v}_f64_$rep.json 2> /dev/null
    python - <<PY
import json
d = json.loads(open("$OUT/u1all_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep:", "  ".join("%s %.3f ms" % (w["config"]["workload"].split()[0] + ("x" if "convex-convex" in w["config"]["workload"] else ""), w["ms_per_step"]) for w in [d] + d.get("workloads", [])))
d = json.loads(open("$OUT/u1all_${v}_f64_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep: C2 f64 %.3f ms" % d["ms_per_step"])
PY
  done
done
