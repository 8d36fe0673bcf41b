#!/bin/bash
# C2 GJK kernel: early finishers of a trip draw the remaining FULL units (FCLB_GJK_SHARE_UNITS=2) against one unit per warp per trip
OUT=gpurun_out
V=$PWD/mind-fcl_b200/libfclb200_su2.so
FCLB_LIB=$V timeout 600 python -m pytest tests/test_distance_gpu.py -m gpu -x -q > $OUT/su2_pytest.log 2>&1; echo "distance parity with the variant rc $?: $(tail -1 $OUT/su2_pytest.log)"
for rep in 1 2; do
  for v in base su2; do
    lib=$PWD/mind-fcl_b200/libfclb200.so; [ $v = su2 ] && lib=$V
    for dt in f32 f64; do
      FCLB_LIB=$lib timeout 300 python bench.py --dtype $dt --no-workloads --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep $dt: C2 %.3f ms  ' % d['ms_per_step'] + '  '.join('%s %.3f' % (k['kernel'], k['avg_ms']) for k in d['roofline']['kernels']))"
    done
  done
done
