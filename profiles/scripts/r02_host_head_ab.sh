#!/bin/bash
# Staged host pipelines of the compute-bound entry points (mesh-mesh: short first stage; mesh / heightmap / octree vs shape: staged at all)
OUT=gpurun_out
T="tests/test_mesh_shape_gpu.py tests/test_heightmap_gpu.py tests/test_octree_gpu.py tests/test_octree_pruned_gpu.py tests/test_bvh_gpu.py tests/test_distance_gpu.py tests/test_collide_gpu.py tests/test_scene_gjk_epa_gpu.py tests/test_scene_penetration_gpu.py"
FCLB_HOST_HEAD=1024 FCLB_HOST_CHUNK=16384 FCLB_HOST_TAPER=4096 timeout 900 python -m pytest $T -m gpu -x -q > $OUT/head_pytest_small.log 2>&1; echo "subset with 1k / 4k-query stages rc $?: $(tail -1 $OUT/head_pytest_small.log)"
timeout 900 python -m pytest $T tests/test_full_size_gpu.py tests/test_host_api_gpu.py tests/test_multi_device_gpu.py -m gpu -x -q > $OUT/head_pytest.log 2>&1; echo "subset + full size, default stages rc $?: $(tail -1 $OUT/head_pytest.log)"
for rep in 1 2; do
for cfg in "single-stage FCLB_HOST_HEAD=0 FCLB_HOST_CHUNK=67108864" "equal-stages FCLB_HOST_HEAD=0" "ramp FCLB_HOST_HEAD=65536" "ramp16k FCLB_HOST_HEAD=16384"; do
  set -- $cfg; name=$1; shift
  for w in c3 c4; do
    env "$@" timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rep $rep %-13s $w  device %.3f ms  e2e %.3e q/s %.3f ms' % ('$name', d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0)))"
  done
done
done
