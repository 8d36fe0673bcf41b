"""Probe for the octree-shape count difference that shows only under compute-sanitizer memcheck (box, max_contacts=1)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "mind-fcl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import fclb200 as fclb, oracle_py, scenes
from test_octree_gpu import setup_scene, PRIMS

fclb.init(0)
ref = oracle_py.RefOracle()
for dtype in (np.float32, np.float64):
    st, oid, oct_h, slots, rslots, sizes = setup_scene(fclb, ref, dtype)
    shapes = [PRIMS["box"]]
    table = fclb.shapes_upload(shapes)
    n = 2000
    p_oct, p_sh = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=4700)
    ids = np.zeros(n, np.uint32)
    e1, _ = ref.octree_shape_collide_batch(oid, shapes, ids, p_oct, p_sh, threads=8, max_contacts=1)
    eall, _ = ref.octree_shape_collide_batch(oid, shapes, ids, p_oct, p_sh, threads=8, max_contacts=2**31 - 1)
    for rep in range(3):
        c, node = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, fclb.make_request(max_contacts=1), want_node=True)
        m = np.nonzero(c != e1)[0]
        print(dtype.__name__, "mc=1 rep", rep, "mismatches", m.tolist(), "nodes", node[m].tolist(), flush=True)
    c, node = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, fclb.make_request(max_contacts=2**31 - 1), want_node=True)
    m = np.nonzero(c != eall)[0]
    print(dtype.__name__, "mc=all mismatches", m.tolist(), c[m].tolist(), eall[m].tolist(), "expected count of q0", int(eall[0]), flush=True)
    # reversed order: does the difference follow the data or the index?
    r = np.arange(n)[::-1].copy()
    c, _ = fclb.octree_shape_collide_batch_host(oct_h, table, ids, np.ascontiguousarray(p_oct[r]), np.ascontiguousarray(p_sh[r]), st, fclb.make_request(max_contacts=1))
    m = np.nonzero(c != e1[r])[0]
    print(dtype.__name__, "reversed: mismatching positions", m.tolist(), "= original queries", r[m].tolist(), flush=True)
    for k in (1, 2, 33, 500):
        c, _ = fclb.octree_shape_collide_batch_host(oct_h, table, ids[:k], np.ascontiguousarray(p_oct[:k]), np.ascontiguousarray(p_sh[:k]), st, fclb.make_request(max_contacts=1))
        m = np.nonzero(c != e1[:k])[0]
        print(dtype.__name__, f"first {k}: mismatches", m.tolist(), flush=True)
    # every query duplicated 4x: which copies differ
    d = np.repeat(np.arange(n), 4)
    c, _ = fclb.octree_shape_collide_batch_host(oct_h, table, np.zeros(4 * n, np.uint32), np.ascontiguousarray(p_oct[d]), np.ascontiguousarray(p_sh[d]), st, fclb.make_request(max_contacts=1))
    m = np.nonzero(c != e1[d])[0]
    print(dtype.__name__, "4x duplicated: mismatching positions", m.tolist(), "queries", d[m].tolist(), flush=True)
