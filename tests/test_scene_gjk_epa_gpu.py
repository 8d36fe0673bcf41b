"""Contact generation with request.useDefaultPenetration() (GJK + EPA / closed-form contacts per leaf) for every query
that touches a scene geometry, against fcl::collide of the reference:
  mesh - shape         ShapeSimplexIntersect -> shapeTriangleIntersect with contacts   gjk_solver-inl.h:479-531
  heightmap / octree   ShapeIntersect<Box, Shape> per pixel / voxel box               heightmap_solver_leaf-inl.h:10-31,
                       - shape                                                         octree2_solver_leaf-inl.h:22-44
  the five scene pairs ShapeIntersect<Box, Box> (boxBox2) / ShapeSimplexIntersect<Box> heightmap_solver_leaf-inl.h:33-88,
                                                                                       octree2_solver_leaf-inl.h:46-404
Contact counts must be identical and the multiset of contact records (b1[, b2], normal, position, depth) bit-identical:
the order of a contact list follows each implementation's traversal and is not compared."""
import numpy as np
import pytest

import parity_util
import scenes
from test_octree_gpu import octree_points
from test_scene_pair_gpu import RES, blob_points, upload_heightmap, upload_octree

pytestmark = pytest.mark.gpu
ALL = 2**31 - 1


def multiset_compare(name, dtype, counts, ids, contacts, e_counts, e_ids, e_contacts, keep, e_keep):
    """ids: list of id arrays [n, keep].  Count mismatches are listed: a contact only one side reports must be within
    EPS of touching (|depth| <= EPS): the candidate traversal evaluates every leaf pair under its own (slack-guarded) node
    culls, the reference only those under ITS node culls, and GJK's tolerance turns a touching pair into a contact."""
    eps = parity_util.eps_touch(dtype)
    listed, unexplained = [], []
    n_c = n_bad = 0
    worst = 0.0
    fits = (np.maximum(counts, e_counts) > 0) & (np.maximum(counts, e_counts) <= min(keep, e_keep))
    for q in np.nonzero(fits)[0]:
        got = sorted(tuple(int(a[q, j]) for a in ids) + tuple(contacts[q, j].tolist()) for j in range(int(counts[q])))
        exp = sorted(tuple(int(a[q, j]) for a in e_ids) + tuple(e_contacts[q, j].tolist()) for j in range(int(e_counts[q])))
        if counts[q] != e_counts[q]:
            nk = len(ids)
            gk, xk = [g[:nk] for g in got], [x[:nk] for x in exp]
            for side, recs, other in (("ours only", got, xk), ("reference only", exp, gk)):
                pool = list(other)
                for r in recs:
                    if r[:nk] in pool:
                        pool.remove(r[:nk])
                        continue
                    item = {"query": int(q), "what": "contact reported by one side", "side": side, "ids": list(r[:nk]),
                            "depth": float(r[-1])}
                    item["class"] = "within eps of touching" if abs(r[-1]) <= eps else "UNEXPLAINED"
                    listed.append(item)
                    if abs(r[-1]) > eps:
                        unexplained.append(item)
            continue
        n_c += len(got)
        if got != exp:
            for g, x in zip(got, exp):
                if g != x:
                    n_bad += 1
                    assert g[:len(ids)] == x[:len(ids)], (name, q, g, x)
                    worst = max(worst, float(np.abs(np.asarray(g[len(ids):]) - np.asarray(x[len(ids):])).max()))
    mism_unfit = np.nonzero((counts != e_counts) & ~fits)[0]
    parity_util.record("test_scene_gjk_epa", name, dtype, len(counts), "contact counts, contact record multisets", listed,
                       {"colliding": int((e_counts > 0).sum()), "contacts": int(e_counts.sum()), "records_compared": n_c,
                        "records_not_bit_identical": n_bad, "worst_component_diff": worst, "unexplained": len(unexplained)})
    assert not unexplained, unexplained[:5]
    assert len(mism_unfit) == 0, mism_unfit[:10]
    assert n_c > 0
    assert n_bad == 0, (name, n_bad, worst)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_shape_default_penetration(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    keep = 256
    hull = scenes.ellipsoid_mesh(0.05, 0.075, 0.1)
    shapes = [(scenes.BOX, 0, (0.12, 0.08, 0.1)), (scenes.SPHERE, 0, (0.06,)), (scenes.CAPSULE, 0, (0.03, 0.12)),
              (scenes.CONVEX, fclb.convex_upload(*hull), ()), (scenes.CYLINDER, 0, (0.04, 0.1))]
    rshapes = shapes[:3] + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())] + shapes[4:]
    table = fclb.shapes_upload(shapes)
    n = 500
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    v, t = scenes.c4_scene_mesh(grid=40, n_boxes=20)
    v = v * 0.25
    mid = ref_oracle.bvh_create(v, t)
    bvh = fclb.bvh_build(v, t, st)
    rng = np.random.Generator(np.random.PCG64(5))
    p_mesh = scenes.random_poses(rng, n, 0.2, dtype)
    local = scenes.random_poses(rng, n, 0.4, np.float64)
    local[:, 11] = rng.uniform(0.0, 0.2, size=n)
    p_shape = scenes.compose_poses(p_mesh, local, dtype)
    pts = scenes.terrain_points(40_000, 0.64)
    hid = ref_oracle.heightmap_create(pts, 0.01, 64)
    heights, upper = ref_oracle.heightmap_export(hid, dtype, 64)
    hm = fclb.heightmap_upload(heights, 0.01, upper)
    p_hm, p_hs = scenes.heightmap_query_poses(n, dtype, 0.64, -0.05, 0.4, seed=11)
    oid = ref_oracle.octree_create(octree_points(), 0.01, 64)
    ch, full, leaf, root, n_layers = ref_oracle.octree_export(oid, dtype)
    octree = fclb.octree_upload(ch, full, leaf, root, n_layers)
    p_oc, p_os = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=12)
    cases = [("mesh-shape", fclb.SCENE_BVH, bvh, 0, mid, p_mesh, p_shape), ("heightmap-shape", fclb.SCENE_HEIGHTMAP, hm, 1, hid, p_hm, p_hs),
             ("octree-shape", fclb.SCENE_OCTREE, octree, 2, oid, p_oc, p_os)]
    for name, kind, handle, rkind, rid, ps, psh in cases:
        req = fclb.make_request(max_contacts=ALL, penetration_mode=1)
        counts, b1, contacts = fclb.scene_shape_contacts_batch_host(kind, handle, table, ids, ps, psh, st, req, keep)
        e_counts, e_b1, e_contacts = ref_oracle.scene_shape_contacts_batch(rkind, rid, rshapes, ids, ps, psh, 2048, threads=8,
                                                                           max_contacts=ALL, penetration_mode=1)
        assert int(e_counts.max()) <= 2048
        multiset_compare(f"{name} DefaultGJK_EPA", dtype, counts, [b1], contacts, e_counts, [e_b1], e_contacts, keep, 2048)
        # capped request: numContacts follows max_contacts (the free-space clipping of ShapeIntersect)
        for cap in (1, 3):
            req = fclb.make_request(max_contacts=cap, penetration_mode=1)
            c2, _, _ = fclb.scene_shape_contacts_batch_host(kind, handle, table, ids, ps, psh, st, req, 4)
            assert np.array_equal(c2, np.minimum(e_counts, cap)), (name, cap)
    # the counting entry points accept the request too (numContacts of the contact path)
    req = fclb.make_request(max_contacts=ALL, penetration_mode=1)
    c_m, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, p_mesh, p_shape, st, req)
    e_m, _, _ = ref_oracle.scene_shape_contacts_batch(0, mid, rshapes, ids, p_mesh, p_shape, 8, threads=8, max_contacts=ALL,
                                                      penetration_mode=1)
    assert np.array_equal(c_m, e_m)
    fclb.bvh_release(bvh)
    fclb.heightmap_release(hm)
    fclb.octree_release(octree)
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_pair_default_penetration(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    n, keep = 100, 4096
    hidA, hmA = upload_heightmap(fclb, ref_oracle, scenes.terrain_points(40_000, 64 * RES), 64, dtype)
    hidB, hmB = upload_heightmap(fclb, ref_oracle, blob_points(11, upper_half=True), 16, dtype)
    oidA, octA = upload_octree(fclb, ref_oracle, octree_points(), 64, dtype)
    oidB, octB = upload_octree(fclb, ref_oracle, blob_points(12), 16, dtype)
    v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24, radius=0.12, noise=0.02)
    mid = ref_oracle.bvh_create(v, t)
    obb, child, tri_verts = ref_oracle.bvh_export(mid, dtype)
    mesh = fclb.bvh_upload(obb, child, tri_verts, st)
    H, O, M = fclb.SCENE_HEIGHTMAP, fclb.SCENE_OCTREE, fclb.SCENE_BVH
    cases = [
        ("heightmap-heightmap", H, hidA, hmA, H, hidB, hmB, 0.05, 0.5, 5201),
        ("heightmap-mesh", H, hidA, hmA, M, mid, mesh, 0.0, 0.6, 5202),
        ("heightmap-octree", H, hidA, hmA, O, oidB, octB, 0.05, 0.6, 5203),
        ("octree-mesh", O, oidA, octA, M, mid, mesh, -0.1, 0.45, 5204),
        ("octree-octree", O, oidA, octA, O, oidB, octB, 0.0, 0.45, 5205),
    ]
    for name, k1, r1, d1, k2, r2, d2, zlo, zhi, seed in cases:
        p1, p2 = scenes.heightmap_query_poses(n, dtype, 0.4, zlo, zhi, seed=seed)
        req = fclb.make_request(max_contacts=ALL, penetration_mode=1)
        counts, b1, b2, contacts = fclb.scene_pair_contacts_batch_host(k1, d1, k2, d2, p1, p2, st, req, keep)
        e_counts, e_b1, e_b2, e_contacts = ref_oracle.scene_pair_collide_batch(
            k1, r1, k2, r2, p1, p2, keep, threads=8, want_contacts=True, max_contacts=ALL, penetration_mode=1)
        multiset_compare(f"{name} DefaultGJK_EPA", dtype, counts, [b1, b2], contacts, e_counts, [e_b1, e_b2], e_contacts, keep, keep)
    for h in (hmA, hmB):
        fclb.heightmap_release(h)
    for h in (octA, octB):
        fclb.octree_release(h)
    fclb.bvh_release(mesh)
