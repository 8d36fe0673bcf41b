"""Scene-pair kernels at scale (not a bench.py workload: BASELINE.json has no such config): a 1024^2 heightmap
(10 layers, C4's terrain cloud) against a 10k-triangle mesh, a 64^3-cell octree and a 128^2 heightmap, and a
256^3-cell octree against the mesh and the small octree, 2000 poses each,
checked against the reference on every query and timed beside it (reference: all host threads)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mind-fcl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scenes, oracle_py, fclb200 as fclb
from test_scene_pair_gpu import blob_points
from test_octree_gpu import octree_points

fclb.init(0)
ref = oracle_py.RefOracle()
threads = ref.hardware_threads()
dtype, st = np.float32, fclb.F32
pts = scenes.c4_heightmap_points()
hid = ref.heightmap_create(pts, 0.004, 512)
hm = fclb.heightmap_build_points_host(pts, 0.004, 512, st)
v, t = scenes.noisy_uv_sphere(n_lat=51, n_lon=100, radius=0.3, noise=0.03)
mid = ref.bvh_create(v, t)
obb, child, tri = ref.bvh_export(mid, dtype)
mesh = fclb.bvh_upload(obb, child, tri, st)
oid = ref.octree_create(blob_points(12, r=0.3, n=60000), 0.01, 32)
octree = fclb.octree_upload(*ref.octree_export(oid, dtype))
hid2 = ref.heightmap_create(blob_points(11, r=0.3, n=60000, upper_half=True), 0.005, 64)
h2, up2 = ref.heightmap_export(hid2, dtype, 64)
hm2 = fclb.heightmap_upload(h2, 0.005, up2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
H, O, M = fclb.SCENE_HEIGHTMAP, fclb.SCENE_OCTREE, fclb.SCENE_BVH
# a large octree: the terrain cloud voxelised at 1 cm (256^3 cells) -- many fully occupied inner nodes are impossible
# for a surface, so add a filled slab
slab = octree_points()
big_pts = np.ascontiguousarray(np.concatenate([pts[::2] * np.array([0.6, 0.6, 1.0]), slab]))
oidA = ref.octree_create(big_pts, 0.01, 128)
octA = fclb.octree_upload(*ref.octree_export(oidA, dtype))
cases = [("heightmap(1024^2)-mesh(10k tris)", H, hid, hm, M, mid, mesh), ("heightmap(1024^2)-octree", H, hid, hm, O, oid, octree),
         ("heightmap(1024^2)-heightmap(128^2)", H, hid, hm, H, hid2, hm2),
         ("octree(256^3 cells)-mesh(10k tris)", O, oidA, octA, M, mid, mesh), ("octree(256^3 cells)-octree", O, oidA, octA, O, oid, octree)]
for name, k1, r1, d1, k2, r2, d2 in cases:
    p1, p2 = scenes.heightmap_query_poses(n, dtype, 1.6 if k1 == H else 1.0, -0.3, 1.2 if k1 == H else 0.6, seed=61)
    for mc in (1, 2**31 - 1):
        req = fclb.make_request(max_contacts=mc)
        fclb.scene_pair_collide_batch_host(k1, d1, k2, d2, p1, p2, st, req)  # warm-up
        t0 = time.perf_counter()
        c, _, _ = fclb.scene_pair_collide_batch_host(k1, d1, k2, d2, p1, p2, st, req)
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        e, _, _ = ref.scene_pair_collide_batch({M: 0, H: 1, O: 2}[k1], r1, {M: 0, H: 1, O: 2}[k2], r2, p1, p2, 1, threads=threads, max_contacts=mc)
        t_cpu = time.perf_counter() - t0
        nn, nl = fclb.scene_last_visit_counts()
        diff = c.astype(np.int64) - e.astype(np.int64)
        print(f"   contacts ours - reference: +{int(diff[diff > 0].sum())} / {int(diff[diff < 0].sum())} over {int((diff != 0).sum())} queries")
        print(f"{name} max_contacts={mc}: colliding {int((e > 0).sum())}/{n}, contacts {int(e.sum())}, count mismatches "
              f"{int((c != e).sum())}; device call {1e3 * t_gpu:.2f} ms ({n / t_gpu:.3g} q/s, {nn / n:.0f} node pairs + {nl / n:.0f} leaf pairs "
              f"per query), reference on {threads} threads {1e3 * t_cpu:.1f} ms ({n / t_cpu:.3g} q/s)")
