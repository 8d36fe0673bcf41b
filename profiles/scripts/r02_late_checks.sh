#!/bin/bash
# late round 2: the whole GPU suite, then memcheck over the new CCD kernels (float cases)
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest_gpu_late.log 2>&1; echo "suite rc $?: $(tail -1 $OUT/r02_pytest_gpu_late.log)"
for f in test_ccd_scene_mesh_gpu test_ccd_scene_pair_gpu; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/$f.py -x -q -k "float32 and (heightmap or hm1-oc2 or oc1-oc2 or pair1 or pair3)" > $OUT/memcheck_$f.log 2>&1
  echo "$f memcheck rc $?: $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/memcheck_$f.log | tr '\n' ' ')"
done
