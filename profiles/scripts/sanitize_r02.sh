#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (small f32 cases)
OUT=gpurun_out
for T in "tests/test_ccd_mesh_gpu.py -k float32" "tests/test_ccd_scene_gpu.py -k float32" "tests/test_bvh_build_gpu.py -k float32" \
         "tests/test_bvh_refit_gpu.py -k bottomup_matches_reference_and_float32" "tests/test_ccd_gpu.py -k float32" \
         "tests/test_distance_gpu.py -k closed_form_and_float32"; do
  name=$(echo $T | sed 's#tests/##; s#[ /.-]#_#g')
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest $T -x -q > $OUT/sanitize_$name.log 2>&1
  echo "$T -> rc $? : $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/sanitize_$name.log | tr '\n' ' ')"
done
