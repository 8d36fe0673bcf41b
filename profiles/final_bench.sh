set -x
for W in c2 c1a c1b c1b_convex c3 c4 c5; do python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${W}_f32.json 2> gpurun_out/bench_${W}_f32.err; done
python bench.py --workload c2 --dtype f64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_f64.json 2>/dev/null
python bench.py --workload c3 --dtype f64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_f64.json 2>/dev/null
python bench.py --workload c4 --dtype f64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_f64.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_f*.json')):
    try:
        d=json.load(open(f)); cb=d.get('cpu_baseline') or {}
        print(f.split('/')[-1], 'value %.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'], 'cpu %.3g (%s cores)'%(cb.get('value') or 0, cb.get('cores')), 'roof %.3f'%d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
