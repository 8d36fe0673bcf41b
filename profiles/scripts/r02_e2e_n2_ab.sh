#!/bin/bash
# 2 GPUs, one box: e2e of C2 with the old host path (histogram copy, equal stages) against the new one (mapped histogram, tapered tail)
OUT=gpurun_out
port=29520
for rep in 1 2; do
for cfg in "old FCLB_HIST_COPY=1 FCLB_HOST_TAPER=0" "mapped FCLB_HOST_TAPER=0" "mapped+taper FCLB_HOST_TAPER=524288"; do
  set -- $cfg; name=$1; shift
  port=$((port + 1))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --no-workloads --no-cpu-baseline --steps 6 2>/dev/null | tail -1 > $OUT/n2ab_${name}_$rep.json
  python - <<PY
import json
d = json.loads(open("$OUT/n2ab_${name}_$rep.json").read().strip().splitlines()[-1])
e = d["e2e"]; s = d.get("strong_scaling", {}).get("e2e", {})
print("rep $rep %-13s weak e2e %.3e (%.2f ms, h2d %.1f GB/s per rank)  12S %.3e | strong e2e %.3e  12S %.3e | device %.3e" % ("$name", e["value"], e["ms_per_step"], e["h2d_gbs"], e["with_12S_poses"]["value"], s.get("value", 0), s.get("with_12S_poses", {}).get("value", 0), d["value"]))
PY
done
done
