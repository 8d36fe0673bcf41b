# first-tier tile width (lanes per query) x pool size (faces) x CTAs per SM for the two EPA workloads
for W in c1b c1b_convex; do for T in 8 4 16; do for F in 24 32 40; do for B in 2 3 4; do
  echo -n "$W tile=$T faces=$F blocks=$B: "
  FCLB_EPA_TILE=$T FCLB_EPA_TIER1_FACES=$F FCLB_EPA_BLOCKS_PER_SM=$B python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g q/s %.3f ms' % (d['value'], d['ms_per_step']))"
done; done; done; done
