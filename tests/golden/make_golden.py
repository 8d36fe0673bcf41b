#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by running the REFERENCE itself
(oracle/_ref/libfclref.so = the unmodified mind-fcl headers compiled against
oracle/eigen_shim) on small seeded inputs.  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py
The .npz files are committed; tests compare the oracle port (CPU) and the CUDA path
(GPU) against them, so parity stays pinned on machines without the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "mind-fcl_b200"), os.path.join(ROOT, "oracle")]
import oracle_py  # noqa: E402
import scenes  # noqa: E402

B, S, E, C, K, Y, V = range(7)


def shapes_arr(shapes):
    return np.array([[t, g] + list(p) + [0.0] * (3 - len(p)) for (t, g, p) in shapes], np.float64)


def mixed(n, dtype, combos, extent, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = scenes.random_poses(rng, n, extent, dtype)
    p2 = scenes.random_poses(rng, n, extent, dtype)
    idx = np.arange(n) % len(combos)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32), np.array([combos[i][1] for i in idx], np.uint32))
    return pairs, p1, p2


def octree_golden(ref):
    """Octree construction, two cumulative prunes and the consolidated tree of the second, as the reference builds them
    (Octree::rebuildTree, pruneOctreeByOBB, rebuildAccordingToPruneInfo): tests/test_octree_build.py::test_golden."""
    rng = np.random.Generator(np.random.PCG64(31))
    res, half = 0.04, 16
    g = np.arange(-10, 10) * res + res / 2
    X, Y, Z = np.meshgrid(g, g, np.arange(-6, 6) * res + res / 2, indexing="ij")
    keep = Z < 0.08 * np.sin(5 * X) * np.cos(4 * Y)
    pts = np.ascontiguousarray(np.concatenate([np.stack([X[keep], Y[keep], Z[keep]], 1), rng.uniform(-0.7, 0.7, size=(600, 3))]))
    e = np.array([0.3, 0.2, 0.5])
    boxes = [(scenes.euler_to_matrix(e[:1], e[1:2], e[2:3])[0], np.array([0.05, -0.1, -0.05]), np.array([0.22, 0.15, 0.12])),
             (np.eye(3), np.array([-0.24, 0.2, 0.0]), np.array([0.12, 0.12, 0.3]))]
    out = {"points": pts, "resolution": res, "half_shape": half,
           "obb": np.array([np.concatenate([a.reshape(9), c, x]) for a, c, x in boxes])}
    oid = ref.octree_create(pts, res, half)
    p1 = ref.octree_prune(oid, *boxes[0])
    p2 = ref.octree_prune(p1, *boxes[1])
    r2 = ref.octree_prune_rebuild(p1, *boxes[1])
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for name, gid in (("tree", oid), ("prune1", p1), ("prune2", p2), ("rebuilt", r2)):
            ch, full, leaf, root, layers = ref.octree_export(gid, dtype)
            out[f"{name}_children_{tag}"], out[f"{name}_full_{tag}"], out[f"{name}_leaf_{tag}"] = ch, full, leaf
            out[f"root_{tag}"], out["layers"] = root, layers
            pr = ref.octree_export_pruned(gid, dtype, len(full))
            if pr is not None:
                out[f"{name}_pruned_{tag}"] = pr
    np.savez_compressed(os.path.join(HERE, "octree.npz"), **out)
    print("octree.npz:", len(pts), "points,", len(out["tree_full_f32"]), "inner /", len(out["tree_leaf_f32"]), "leaf nodes,",
          int(out["prune2_pruned_f32"].sum()), "pruned,", len(out["rebuilt_full_f32"]), "inner after consolidation")


def main():
    ref = oracle_py.RefOracle()
    if sys.argv[1:] == ["octree"]:
        return octree_golden(ref)
    octree_golden(ref)
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        # distance: C2 mix + every closed-form specialisation
        shapes, pairs, p1, p2 = scenes.config_c2(3000, dtype, seed=77)
        d, w1, w2, ok = ref.distance_batch(shapes, pairs, p1, p2)
        np.savez_compressed(os.path.join(HERE, f"distance_c2_{tag}.npz"), shapes=shapes_arr(shapes),
                            pairs=pairs.view(np.uint32).reshape(-1, 2), poses1=p1, poses2=p2, dist=d, p1=w1, p2=w2, ok=ok)
        shapes = [(S, 0, (0.07,)), (B, 0, (0.3, 0.2, 0.1)), (C, 0, (0.05, 0.25)), (Y, 0, (0.08, 0.2)), (S, 0, (0.11,)),
                  (C, 0, (0.04, 0.3))]
        combos = [(0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (3, 0), (0, 4), (2, 5)]
        pairs, p1, p2 = mixed(2400, dtype, combos, 0.4, 78)
        d, w1, w2, ok = ref.distance_batch(shapes, pairs, p1, p2)
        np.savez_compressed(os.path.join(HERE, f"distance_closed_{tag}.npz"), shapes=shapes_arr(shapes),
                            pairs=pairs.view(np.uint32).reshape(-1, 2), poses1=p1, poses2=p2, dist=d, p1=w1, p2=w2, ok=ok)
        # collide: closed forms (incl. box-box) and generic pairs, with contacts
        shapes = [(S, 0, (0.15,)), (B, 0, (0.4, 0.3, 0.2)), (C, 0, (0.1, 0.4)), (Y, 0, (0.12, 0.3)), (S, 0, (0.2,)),
                  (B, 0, (0.3, 0.5, 0.25)), (E, 0, (0.3, 0.2, 0.25)), (K, 0, (0.25, 0.6))]
        combos = [(0, 4), (0, 2), (2, 0), (0, 1), (1, 0), (0, 3), (3, 0), (1, 5), (2, 1), (3, 1), (6, 1), (7, 2), (6, 6)]
        pairs, p1, p2 = mixed(2600, dtype, combos, 0.4, 79)
        out = {}
        for pen in (0, 1):
            for mc in (1, 4):
                c, ct = ref.collide_batch(shapes, pairs, p1, p2, max_keep=4, max_contacts=mc, penetration_mode=pen)
                out[f"counts_p{pen}_m{mc}"] = c
                out[f"contacts_p{pen}_m{mc}"] = ct
        np.savez_compressed(os.path.join(HERE, f"collide_{tag}.npz"), shapes=shapes_arr(shapes),
                            pairs=pairs.view(np.uint32).reshape(-1, 2), poses1=p1, poses2=p2, **out)
        # direct GJK + EPA (cvx_collide path), box-box as in test_epa2_with_gjk2.cpp:76-162
        shapes, pairs, p1, p2 = scenes.config_c1_boxes(3000, dtype, seed=80)
        g, e, m, geom, it = ref.gjk_epa_batch(shapes, pairs, p1, p2, mode=1)
        np.savez_compressed(os.path.join(HERE, f"gjk_epa_boxes_{tag}.npz"), shapes=shapes_arr(shapes),
                            pairs=pairs.view(np.uint32).reshape(-1, 2), poses1=p1, poses2=p2, gjk=g, epa=e, mpr=m, geom=geom,
                            iters=it)
        # mesh-mesh on small meshes, with the exported reference BVH
        m1 = scenes.noisy_uv_sphere(n_lat=9, n_lon=16)
        m2 = scenes.noisy_torus(n_major=16, n_minor=8)
        i1, i2 = ref.bvh_create(*m1), ref.bvh_create(*m2)
        o1, c1, t1 = ref.bvh_export(i1, dtype)
        o2, c2, t2 = ref.bvh_export(i2, dtype)
        q1, q2 = scenes.config_c3_poses(1500, dtype, seed=81, extent=1.3)
        cb, _ = ref.bvh_collide_batch(i1, i2, q1, q2, max_contacts=1)
        ca, _ = ref.bvh_collide_batch(i1, i2, q1, q2, max_contacts=2**31 - 1, want_pair=False)
        np.savez_compressed(os.path.join(HERE, f"mesh_{tag}.npz"), obb1=o1, child1=c1, tri1=t1, obb2=o2, child2=c2, tri2=t2,
                            poses1=q1, poses2=q2, counts_bool=cb, counts_all=ca)
    print("golden vectors written to", HERE)
    os.system(f"ls -la {HERE}")


if __name__ == "__main__":
    main()
