#!/bin/bash
# e2e of the C2 host path against the pipeline chunk size (FCLB_HOST_CHUNK queries per stage)
for c in 262144 524288 1048576 2097152 4194304 16777216; do
  FCLB_HOST_CHUNK=$c python bench.py --no-workloads --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d['e2e']
print('chunk=%-9s e2e(QT7) %.3e q/s %.2f ms | 12S %.3e q/s %.2f ms | device %.2f ms' % ('$c', e['value'], e['ms_per_step'], e['with_12S_poses']['value'], e['with_12S_poses']['ms_per_step'], d['ms_per_step']))"
done
