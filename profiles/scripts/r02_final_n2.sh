#!/bin/bash
# End of round 2, final tree on 2 GPUs: torchrun bench line (weak + strong), reference arm, single-process 2-GPU check, new test
OUT=gpurun_out
timeout 300 python -m pytest tests/test_distance_gpu.py -m gpu -x -q -k "stage_schedule or dev_entry" > $OUT/r02_stage_schedule.log 2>&1; echo "stage schedule test rc $?: $(tail -1 $OUT/r02_stage_schedule.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-workloads > $OUT/bench_c2_n2_end.json 2> $OUT/bench_c2_n2_end.err; echo "n2 rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference > $OUT/bench_reference_arm_n2_end.json 2> $OUT/bench_reference_arm_n2_end.err; echo "n2 reference rc $?"
timeout 300 python tests/multi_device_check.py 2 > $OUT/multi_device_check_n2_end.log 2>&1; echo "multi-device check rc $?: $(tail -1 $OUT/multi_device_check_n2_end.log)"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2_n2_end.json").read().strip().splitlines()[-1])
print("N=2 weak %.3e q/s  e2e %.3e | strong %s" % (d["value"], d["e2e"]["value"], json.dumps({k: d.get("strong_scaling", {}).get(k) for k in ("value", "ms_per_step")})))
print("  strong e2e", d.get("strong_scaling", {}).get("e2e", {}).get("value"))
r = json.loads(open("gpurun_out/bench_reference_arm_n2_end.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"))
PY
