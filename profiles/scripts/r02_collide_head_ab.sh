#!/bin/bash
# Copy-out-bound collide host path (c1a): short first stage so that the result copies start early
OUT=gpurun_out
FCLB_HOST_HEAD=1024 FCLB_HOST_CHUNK=16384 timeout 600 python -m pytest tests/test_collide_gpu.py tests/test_penetration_gpu.py -m gpu -x -q > $OUT/chead_pytest_small.log 2>&1; echo "collide tests with 1k-query first stage / 2k stages rc $?: $(tail -1 $OUT/chead_pytest_small.log)"
timeout 600 python -m pytest tests/test_collide_gpu.py tests/test_full_size_gpu.py tests/test_mesh_shape_gpu.py tests/test_heightmap_gpu.py tests/test_octree_gpu.py -m gpu -x -q > $OUT/chead_pytest.log 2>&1; echo "collide + full size + scene-shape (one stage again) rc $?: $(tail -1 $OUT/chead_pytest.log)"
for rep in 1 2; do
for cfg in "equal-stages FCLB_HOST_HEAD=0" "head64k FCLB_HOST_HEAD=65536" "head16k FCLB_HOST_HEAD=16384" "head128k FCLB_HOST_HEAD=131072"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python bench.py --workload c1a --no-workloads --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rep $rep %-13s c1a  device %.3f ms  e2e %.3e q/s %.3f ms' % ('$name', d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0)))"
done
done
