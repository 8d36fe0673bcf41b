"""Pruned octrees: Octree2CollisionGeometry::pruneBy(OBB) of the reference (pruneOctreeByOBB,
geometry/octree2/octree_prune-inl.h:10-103) yields prune_internal_nodes plus replacement leaf bitmasks and
fully-occupied flags (OctreePruneInfo, octree_node.h:56-72).  The device kernels take them as the optional
`pruned` mask of fclb_octree_upload; this test pins their handling (pruned inner nodes are skipped, the
new leaf / full arrays are the ones consulted) for the octree-shape kernel and the octree pair kernels:
counts and contact ids identical to the reference on the pruned geometry, float and double."""
import numpy as np
import pytest

import scenes
from test_octree_gpu import HALF, PRIMS, RES, octree_points
from test_scene_pair_gpu import blob_points

pytestmark = pytest.mark.gpu


def pruned_octree(fclb, ref_oracle, dtype):
    oid = ref_oracle.octree_create(octree_points(), RES, HALF)
    axis = scenes.euler_to_matrix(np.array([0.3]), np.array([0.2]), np.array([0.5]))[0]
    pid = ref_oracle.octree_prune(oid, axis, (0.05, -0.1, -0.05), (0.22, 0.15, 0.12))
    pid = ref_oracle.octree_prune(pid, np.eye(3), (-0.25, 0.2, 0.0), (0.1, 0.1, 0.3))  # a second, cumulative prune
    ch0, full0, leaf0, _, _ = ref_oracle.octree_export(oid, dtype)
    ch, full, leaf, root, n_layers = ref_oracle.octree_export(pid, dtype)
    pruned = ref_oracle.octree_export_pruned(pid, dtype, len(full))
    assert pruned is not None and pruned.any() and np.array_equal(ch, ch0)
    assert (leaf != leaf0).any() and (full != full0).any()
    # the same two prunes through our host mirror of pruneOctreeByOBB: identical prune info
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    m_pr, m_full, m_leaf = fclb.octree_prune_host(ch0, full0, leaf0, root, n_layers, axis, (0.05, -0.1, -0.05), (0.22, 0.15, 0.12), st)
    m_pr, m_full, m_leaf = fclb.octree_prune_host(ch0, m_full, m_leaf, root, n_layers, np.eye(3), (-0.25, 0.2, 0.0), (0.1, 0.1, 0.3),
                                                  st, pruned=m_pr)
    assert np.array_equal(m_pr, pruned) and np.array_equal(m_full, full) and np.array_equal(m_leaf, leaf)
    print(f"pruned octree: {int(pruned.sum())} of {len(full)} inner nodes pruned, {int((leaf != leaf0).sum())} leaf masks "
          f"changed, fully occupied inner nodes {int(full0.sum())} -> {int(full.sum())}")
    return oid, pid, fclb.octree_upload(ch, full, leaf, root, n_layers, pruned=pruned)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pruned_octree(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    oid, pid, oct_h = pruned_octree(fclb, ref_oracle, dtype)
    n = 1500
    # octree-shape
    shapes = list(PRIMS.values())
    table = fclb.shapes_upload(shapes)
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    p_oct, p_sh = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=4901)
    changed = 0
    for mc in (1, 2**31 - 1):
        req = fclb.make_request(max_contacts=mc)
        counts, _ = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, req, want_node=True)
        e_counts, _ = ref_oracle.octree_shape_collide_batch(pid, shapes, ids, p_oct, p_sh, threads=8, max_contacts=mc)
        u_counts, _ = ref_oracle.octree_shape_collide_batch(oid, shapes, ids, p_oct, p_sh, threads=8, max_contacts=mc)
        assert np.array_equal(counts, e_counts), (mc, np.nonzero(counts != e_counts)[0][:8])
        changed += int((e_counts != u_counts).sum())
        print(f"[pruned octree-shape {np.dtype(dtype).name} max_contacts={mc}] contacts {int(e_counts.sum())} "
              f"(unpruned {int(u_counts.sum())}), mismatches 0")
    assert changed > 0  # the prune must matter for these queries
    fclb.release(table)
    # octree pairs: pruned octree vs mesh, pruned octree vs a small octree
    v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24, radius=0.12, noise=0.02)
    mid = ref_oracle.bvh_create(v, t)
    obb, child, tri = ref_oracle.bvh_export(mid, dtype)
    mesh = fclb.bvh_upload(obb, child, tri, st)
    oidB = ref_oracle.octree_create(blob_points(12), RES, 16)
    octB = fclb.octree_upload(*ref_oracle.octree_export(oidB, dtype))
    keep, m = 4096, 300
    p1, p2 = scenes.heightmap_query_poses(m, dtype, 0.4, -0.45, 0.45, seed=4902)
    for name, k2, r2, d2 in (("octree-mesh", fclb.SCENE_BVH, mid, mesh), ("octree-octree", fclb.SCENE_OCTREE, oidB, octB)):
        req = fclb.make_request(max_contacts=2**31 - 1)
        counts, b1, b2 = fclb.scene_pair_collide_batch_host(fclb.SCENE_OCTREE, oct_h, k2, d2, p1, p2, st, req, keep)
        e_counts, e_b1, e_b2 = ref_oracle.scene_pair_collide_batch(2, pid, {fclb.SCENE_BVH: 0, fclb.SCENE_OCTREE: 2}[k2], r2, p1, p2,
                                                                   keep, threads=8, max_contacts=2**31 - 1)
        u_counts, _, _ = ref_oracle.scene_pair_collide_batch(2, oid, {fclb.SCENE_BVH: 0, fclb.SCENE_OCTREE: 2}[k2], r2, p1, p2, 1,
                                                             threads=8, max_contacts=2**31 - 1)
        assert np.array_equal(counts, e_counts), (name, np.nonzero(counts != e_counts)[0][:8])
        assert (e_counts != u_counts).any()
        for q in np.nonzero((counts > 0) & (counts <= keep))[0]:
            got = sorted(zip(b1[q, :counts[q]].tolist(), b2[q, :counts[q]].tolist()))
            exp = sorted(zip(e_b1[q, :counts[q]].tolist(), e_b2[q, :counts[q]].tolist()))
            assert got == exp, (name, q)
        print(f"[pruned {name} {np.dtype(dtype).name}] contacts {int(e_counts.sum())} (unpruned {int(u_counts.sum())}), ids identical")
    fclb.octree_release(oct_h)
    fclb.octree_release(octB)
    fclb.bvh_release(mesh)
