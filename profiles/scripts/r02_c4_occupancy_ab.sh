#!/bin/bash
# C4 traversal kernels after the unroll-1 change: resident CTAs per SM (launch bounds) of the mesh-shape and heightmap-shape kernels
for rep in 1 2; do
  for v in base mb4 mb2 hm3; do
    lib=$PWD/mind-fcl_b200/libfclb200.so; [ $v != base ] && lib=$PWD/mind-fcl_b200/libfclb200_$v.so
    FCLB_LIB=$lib timeout 300 python bench.py --workload c4 --no-workloads --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: C4 %.3f ms  ' % d['ms_per_step'] + '  '.join('%s %.3f' % (k['kernel'], k['avg_ms']) for k in d['roofline']['kernels']))"
  done
done
