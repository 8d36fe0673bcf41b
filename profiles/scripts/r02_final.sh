#!/bin/bash
# End of round 2, final tree: whole GPU suite, smoke, default bench line + reference arm, C2 launch list, --set full captures
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest_gpu_end.log 2>&1; echo "suite rc $?: $(tail -1 $OUT/r02_pytest_gpu_end.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke_end.log 2>&1; echo "smoke rc $?: $(tail -1 $OUT/r02_smoke_end.log)"
S=$(date +%s); timeout 900 python bench.py > $OUT/bench_default_end.json 2> $OUT/bench_default_end.err; echo "default bench rc $? in $(( $(date +%s) - S )) s"
S=$(date +%s); timeout 900 python bench.py --impl reference > $OUT/bench_reference_arm_end.json 2> $OUT/bench_reference_arm_end.err; echo "reference arm rc $? in $(( $(date +%s) - S )) s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches_c2_end.csv python bench.py --no-workloads --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:distanceGjkBinned -s 4 -c 1 -o $OUT/r02_gjk_end -f python bench.py --no-workloads --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:epaKernel -s 2 -c 1 -o $OUT/r02_epa_convex_end -f python bench.py --workload c1b_convex --no-workloads --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bvhShapeCollideKernel -s 2 -c 1 -o $OUT/r02_mesh_shape_end -f python bench.py --workload c4 --no-workloads --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $OUT/*_end*
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_end.json").read().strip().splitlines()[-1])
print("C2 %.3f ms %.3e q/s  e2e %.3e (%.2f ms)  frac %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
for w in d.get("workloads", []):
    print("%-44s %.3f ms %.3e  e2e %.3e  cpu %.3e" % (w["config"]["workload"][:44], w["ms_per_step"], w["value"], w["e2e"]["value"], w.get("cpu_baseline", {}).get("value", 0)))
r = json.loads(open("gpurun_out/bench_reference_arm_end.json").read().strip().splitlines()[-1])
print("reference arm %.3e q/s" % r["value"])
PY
