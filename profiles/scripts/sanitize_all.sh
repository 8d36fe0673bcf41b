#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite except the full-size configs (file by file, so one slow file cannot eat the budget)
OUT=gpurun_out
for f in tests/test_*_gpu.py; do
  case $f in *full_size*|*multi_device*|*host_api*) continue;; esac
  name=$(basename $f .py)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 3 python -m pytest $f -x -q -k "not large and not timing" > $OUT/memcheck_$name.log 2>&1
  echo "$f -> rc $? : $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/memcheck_$name.log | tr '\n' ' ')"
done
