#!/bin/bash
# new scene-mesh CCD parity + the octree-shape count flake seen once under memcheck: repeat under memcheck / initcheck / racecheck
OUT=gpurun_out
timeout 900 python -m pytest tests/test_ccd_scene_mesh_gpu.py -x -q -s > $OUT/ccd_scene_mesh.log 2>&1; echo "ccd_scene_mesh rc $?: $(tail -1 $OUT/ccd_scene_mesh.log)"
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_octree_gpu.py -x -q -k float32 > $OUT/octree_plain_$i.log 2>&1; echo "plain $i rc $?: $(tail -1 $OUT/octree_plain_$i.log)"
done
for tool in memcheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_octree_gpu.py -x -q -k float32 > $OUT/octree_$tool.log 2>&1
  echo "$tool rc $?: $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/octree_$tool.log | tr '\n' ' ')"
done
