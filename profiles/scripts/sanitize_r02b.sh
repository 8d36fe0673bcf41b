#!/bin/bash
# compute-sanitizer: memcheck on the remaining round-2 kernels, racecheck on the shared-memory schedulers (GJK CTA pool, EPA tiles)
OUT=gpurun_out
run() {  # tool, log name, pytest args...
  local tool=$1 name=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest "$@" -x -q > $OUT/sanitize_$name.log 2>&1
  echo "$tool $* -> rc $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/sanitize_$name.log | tr '\n' ' ')"
}
run memcheck refit_bottomup tests/test_bvh_refit_gpu.py -k "bottomup and float32"
run memcheck distance_closed tests/test_distance_gpu.py -k "closed_form and float32"
run memcheck mesh_mpr tests/test_bvh_gpu.py -k "mpr_penetration and float32"
run racecheck gjk_cta_pool tests/test_distance_gpu.py -k "dev_entry_point"
run racecheck epa_tiles tests/test_collide_gpu.py -k "gjk_boolean_and_epa and float32"
