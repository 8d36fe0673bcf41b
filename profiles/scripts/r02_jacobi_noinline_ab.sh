#!/bin/bash
# jacobi3 out of line (one copy per kernel) against inlined at every OBB-fit site: parity of the kernels that fit OBBs, then C4 A/B
OUT=gpurun_out
B=$PWD/mind-fcl_b200/libfclb200_base.so
timeout 900 python -m pytest tests/test_mesh_shape_gpu.py tests/test_bvh_build_gpu.py tests/test_bvh_refit_gpu.py tests/test_octree_gpu.py tests/test_scene_gjk_epa_gpu.py tests/test_ccd_mesh_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/jac_pytest.log 2>&1; echo "subset with noinline rc $?: $(tail -1 $OUT/jac_pytest.log)"
for rep in 1 2; do
  for v in base noinl; do
    lib=$PWD/mind-fcl_b200/libfclb200.so; [ $v = base ] && lib=$B
    FCLB_LIB=$lib timeout 300 python bench.py --workload c4 --no-workloads --no-cpu-baseline --steps 5 --warmup 3 > $OUT/jac_${v}_$rep.json 2> $OUT/jac_${v}_$rep.err
    python - <<PY
import json
d = json.loads(open("$OUT/jac_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep: C4 %.3f ms  %s" % (d["ms_per_step"], json.dumps(d.get("roofline", {}).get("kernels", ""))[:400]))
PY
  done
done
