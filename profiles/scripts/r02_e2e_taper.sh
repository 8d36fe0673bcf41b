#!/bin/bash
# e2e of the C2 host path: per-stage timeline (FCLB_TRACE_HOST) and a tapered tail of the stage schedule (FCLB_HOST_TAPER)
OUT=gpurun_out
timeout 600 python -m pytest tests/test_distance_gpu.py tests/test_collide_gpu.py tests/test_multi_device_gpu.py -m gpu -x -q > $OUT/taper_pytest.log 2>&1; echo "distance + collide parity rc $?: $(tail -1 $OUT/taper_pytest.log)"
FCLB_TRACE_HOST=1 python bench.py --no-workloads --no-cpu-baseline --steps 3 --warmup 3 2> $OUT/taper_trace_base.err | tail -1 > $OUT/taper_trace_base.json
echo "--- timeline, equal 2M stages (last call)"; grep "fclb trace" $OUT/taper_trace_base.err | tail -30 | head -5; grep "fclb trace" $OUT/taper_trace_base.err | tail -5
FCLB_TRACE_HOST=1 FCLB_HOST_TAPER=262144 python bench.py --no-workloads --no-cpu-baseline --steps 3 --warmup 3 2> $OUT/taper_trace_t256k.err | tail -1 > $OUT/taper_trace_t256k.json
echo "--- timeline, taper to 256k (last call)"; grep "fclb trace" $OUT/taper_trace_t256k.err | tail -42 | head -7; grep "fclb trace" $OUT/taper_trace_t256k.err | tail -7
for rep in 1; do
for cfg in "2097152 0" "2097152 1048576" "2097152 524288" "2097152 262144" "2097152 131072" "1048576 262144" "4194304 262144" "1048576 131072"; do
  set -- $cfg
  FCLB_HOST_CHUNK=$1 FCLB_HOST_TAPER=$2 python bench.py --no-workloads --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d['e2e']
print('rep $rep chunk=%-8s taper=%-8s e2e(QT7) %.3e q/s %.2f ms | 12S %.3e q/s %.2f ms | device %.2f ms' % ('$1', '$2', e['value'], e['ms_per_step'], e['with_12S_poses']['value'], e['with_12S_poses']['ms_per_step'], d['ms_per_step']))"
done
done
