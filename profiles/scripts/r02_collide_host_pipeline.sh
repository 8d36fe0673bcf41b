#!/bin/bash
# chunked three-stage host pipeline of fclb_collide_batch_host / fclb_gjk_epa_batch_host: parity + e2e of c1a / c1b, and a chunk sweep
OUT=gpurun_out
timeout 600 python -m pytest tests/test_collide_gpu.py tests/test_penetration_gpu.py tests/test_golden_gpu.py tests/test_host_api_gpu.py -x -q 2>&1 | tail -2
for chunk in 16777216 2097152 1048576; do
  for w in c1a c1b; do
    FCLB_HOST_CHUNK=$chunk timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 10 --warmup 3 > $OUT/pipe_${w}_$chunk.json 2> $OUT/pipe_${w}_$chunk.err
    python - <<PY
import json
d = json.loads(open("$OUT/pipe_${w}_$chunk.json").read().strip().splitlines()[-1])
print("$w host_chunk=$chunk (collide stage = chunk/8): value %.3e  e2e %.3e  ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
  done
done
