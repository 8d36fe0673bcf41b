#!/bin/bash
# Late round-2 collection on the final tree: smoke(), default bench (all workloads), reference arm, parity log -> JSON.
#   gpurun --timeout 1500 -- 'bash profiles/collect_r02d.sh'
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $OUT/r02_smoke_late.log 2>&1; tail -2 $OUT/r02_smoke_late.log
python bench.py > $OUT/r02_bench_default_late.json 2> $OUT/r02_bench_default_late.err; echo "bench rc $?"
python bench.py --impl reference > $OUT/r02_bench_reference_late.json 2> $OUT/r02_bench_reference_late.err; echo "reference rc $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default_late.json').read().strip().splitlines()[-1])
for w in [d] + d.get('workloads', []):
    r = w.get('roofline', {})
    print(w['config']['workload'][:28], 'value %.4g' % w['value'], 'ms %.3f' % w['ms_per_step'], 'e2e %.4g' % w['e2e']['value'], r.get('bound'), 'frac %.4f' % r.get('frac', 0), 'cpu %.3g' % (w.get('cpu_baseline') or {}).get('value', 0))
r = json.loads(open('gpurun_out/r02_bench_reference_late.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('unit'), r.get('cpu_baseline', {}).get('cores'))
PY
