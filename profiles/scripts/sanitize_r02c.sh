#!/bin/bash
# compute-sanitizer memcheck over the host pipelines changed at the end of round 2 (histogram stored into mapped host memory by
# scanKernel, short-stage schedules), every batch cut into 1k - 16k-query stages
OUT=gpurun_out
export FCLB_HOST_HEAD=1024 FCLB_HOST_CHUNK=16384 FCLB_HOST_TAPER=4096 FCLB_SCENE_HOST_STAGED=1
run() {  # log name, pytest args...
  local name=$1; shift
  timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "$@" -x -q > $OUT/sanitize_$name.log 2>&1
  echo "memcheck $* -> rc $? : $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/sanitize_$name.log | tr '\n' ' ')"
}
run host_distance tests/test_distance_gpu.py -k "(dev_entry_point or empty_and_single or (closed_form and float32))"
run host_collide tests/test_collide_gpu.py -k "closed_form_collide and float32"
run host_bvh tests/test_bvh_gpu.py -k "boolean_and_counts and float32"
run host_mesh_shape tests/test_mesh_shape_gpu.py -k "edge_cases"
