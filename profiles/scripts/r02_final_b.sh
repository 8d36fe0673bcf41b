#!/bin/bash
# Last check of round 2 (early tier-2 consumers on by default): whole GPU suite, smoke, default bench line + reference arm
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest_gpu_end.log 2>&1; echo "suite rc $?: $(tail -1 $OUT/r02_pytest_gpu_end.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke_end.log 2>&1; echo "smoke rc $?: $(tail -1 $OUT/r02_smoke_end.log)"
timeout 300 python bench.py > $OUT/bench_default_end.json 2> $OUT/bench_default_end.err; echo "default bench rc $?"
timeout 200 python bench.py --impl reference > $OUT/bench_reference_arm_end.json 2> $OUT/bench_reference_arm_end.err; echo "reference arm rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_end.json").read().strip().splitlines()[-1])
print("C2 %.3f ms %.3e q/s  e2e %.3e (%.2f ms)  frac %.4f launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("gpu_launches")))
for w in d.get("workloads", []):
    print("%-44s %.3f ms %.3e  e2e %.3e  cpu %.3e" % (w["config"]["workload"][:44], w["ms_per_step"], w["value"], w["e2e"]["value"], w.get("cpu_baseline", {}).get("value", 0)))
r = json.loads(open("gpurun_out/bench_reference_arm_end.json").read().strip().splitlines()[-1])
print("reference arm %.3e q/s" % r["value"])
PY
