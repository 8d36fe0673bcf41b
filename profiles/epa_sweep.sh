# tier-1 pool size (faces) x CTAs per SM for the two EPA workloads; prints q/s and ms per step
for W in c1b c1b_convex; do for F in 24 32 40 48 64 80 96 128; do for B in 2 3 4 6; do
  echo -n "$W faces=$F blocks=$B: "
  FCLB_EPA_TIER1_FACES=$F FCLB_EPA_BLOCKS_PER_SM=$B python bench.py --workload $W --steps 3 --warmup 3 --queries 400000 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g q/s %.3f ms' % (d['value'], d['ms_per_step']))"
done; done; done
