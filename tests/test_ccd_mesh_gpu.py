"""Translational continuous collision, shape vs mesh: fclb_translational_ccd_mesh_batch_host against the reference's
fcl::translational_ccd(shape, BVHModel<OBB<S>>) (detail/ccd/bvh_ccd_solver-inl.h RunSweptBV) on the same seeded inputs.
Bar: contact counts, triangle ids IN THE REFERENCE'S ORDER and time-of-collision intervals bit-identical."""
import numpy as np
import pytest

import parity_util
import scenes

pytestmark = pytest.mark.gpu


def make_inputs(n, dtype, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    ps = scenes.random_poses(rng, n, 1.0, dtype)
    pm = scenes.random_poses(rng, n, 0.2, dtype)
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    disp = np.concatenate([ax, rng.uniform(0.05, 1.2, size=(n, 1))], axis=1).astype(dtype)
    return ps, pm, disp


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_shape_mesh_ccd(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24, radius=0.5, noise=0.05)
    hull = scenes.ellipsoid_mesh(0.2, 0.3, 0.25)
    shapes = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.SPHERE, 0, (0.2,)), (scenes.CAPSULE, 0, (0.1, 0.3)),
              (scenes.CYLINDER, 0, (0.15, 0.3)), (scenes.CONE, 0, (0.2, 0.35)), (scenes.ELLIPSOID, 0, (0.2, 0.1, 0.3)),
              (scenes.CONVEX, fclb.convex_upload(*hull), ())]
    rshapes = shapes[:6] + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())]
    table = fclb.shapes_upload(shapes)
    bvh = fclb.bvh_build(v, t, st)
    oid = ref_oracle.bvh_obb_create(v, t)
    n, keep = 7000, 48
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    ps, pm, disp = make_inputs(n, dtype, 5)
    for request_type in (0, 1, 2):
        for max_contacts, mesh_moves in ((1, False), (4, False), (1000, False), (3, True), (1000, True)):
            c, prim, toc = fclb.translational_ccd_mesh_batch_host(bvh, table, ids, ps, pm, disp, st, request_type=request_type,
                                                                  max_contacts=max_contacts, mesh_moves=mesh_moves, max_keep=keep)
            ec, eprim, etoc = ref_oracle.translational_ccd_mesh_batch(oid, rshapes, ids, ps, pm, disp, request_type=request_type,
                                                                      max_contacts=max_contacts, mesh_moves=mesh_moves, keep=keep,
                                                                      threads=8)
            bad = np.nonzero(c != ec)[0]
            listed = [{"query": int(q), "ours": int(c[q]), "reference": int(ec[q])} for q in bad[:20]]
            same_ids = np.array_equal(prim, eprim)
            same_toc = np.array_equal(toc, etoc)
            parity_util.record("test_shape_mesh_ccd", f"7 shape kinds vs 576-triangle mesh, request {request_type}, max_contacts "
                               f"{max_contacts}, {'mesh' if mesh_moves else 'shape'} moves", dtype, n,
                               "contact counts, triangle ids in the reference's order, toc intervals", listed,
                               {"queries_with_contacts": int((ec > 0).sum()), "contacts": int(ec.sum()),
                                "count_mismatches": int(bad.size), "ids_identical": bool(same_ids), "toc_identical": bool(same_toc)})
            assert bad.size == 0, listed[:5]
            assert same_ids, np.argwhere(prim != eprim)[:5]
            assert same_toc, (np.argwhere(toc != etoc)[:5], np.abs(toc - etoc).max())
        assert int((ec > 0).sum()) > 500
    fclb.release(table)
    fclb.bvh_release(bvh)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mesh_pair_ccd(fclb, ref_oracle, dtype):
    """fclb_translational_ccd_mesh_pair_batch_host against fcl::translational_ccd(BVHModel<OBB>, BVHModel<OBB>)
    (bvh_ccd_solver-inl.h:425-551): counts, (b1, b2) in the reference's order, toc bit-identical."""
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    v1, t1 = scenes.noisy_uv_sphere(n_lat=9, n_lon=14, radius=0.4, noise=0.05)
    v2, t2 = scenes.noisy_torus(n_major=20, n_minor=10) if hasattr(scenes, "noisy_torus") else scenes.noisy_uv_sphere(n_lat=11, n_lon=16)
    b1 = fclb.bvh_build(v1, t1, st)
    b2 = fclb.bvh_build(v2, t2, st)
    o1 = ref_oracle.bvh_obb_create(v1, t1)
    o2 = ref_oracle.bvh_obb_create(v2, t2)
    n, keep = 1500, 64
    rng = np.random.Generator(np.random.PCG64(21))
    p1 = scenes.random_poses(rng, n, 1.2, dtype)
    p2 = scenes.random_poses(rng, n, 0.3, dtype)
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    disp = np.concatenate([ax, rng.uniform(0.05, 1.5, size=(n, 1))], axis=1).astype(dtype)
    for request_type in (0, 1, 2):
        for max_contacts in (1, 5, 100000):
            c, prim, toc = fclb.translational_ccd_mesh_pair_batch_host(b1, b2, p1, p2, disp, st, request_type=request_type,
                                                                       max_contacts=max_contacts, max_keep=keep)
            ec, eprim, etoc = ref_oracle.translational_ccd_mesh_pair_batch(o1, o2, p1, p2, disp, request_type=request_type,
                                                                           max_contacts=max_contacts, keep=keep, threads=8)
            bad = np.nonzero(c != ec)[0]
            listed = [{"query": int(q), "ours": int(c[q]), "reference": int(ec[q])} for q in bad[:20]]
            same_ids, same_toc = np.array_equal(prim, eprim), np.array_equal(toc, etoc)
            parity_util.record("test_mesh_pair_ccd", f"two meshes ({len(t1)} / {len(t2)} triangles), request {request_type}, "
                               f"max_contacts {max_contacts}", dtype, n, "contact counts, (b1, b2) in the reference's order, toc intervals",
                               listed, {"queries_with_contacts": int((ec > 0).sum()), "contacts": int(ec.sum()),
                                        "count_mismatches": int(bad.size), "ids_identical": bool(same_ids),
                                        "toc_identical": bool(same_toc)})
            assert bad.size == 0, listed[:5]
            assert same_ids, np.argwhere(prim != eprim)[:5]
            assert same_toc, (np.argwhere(toc != etoc)[:5], np.abs(toc - etoc).max())
        assert int((ec > 0).sum()) > 100
    fclb.bvh_release(b1)
    fclb.bvh_release(b2)


def test_ccd_edge_cases(fclb, ref_oracle):
    """empty batches, a displacement of zero length, a sweep that starts in contact, refused inputs"""
    st, dtype = fclb.F64, np.float64
    v, t = scenes.noisy_uv_sphere(n_lat=9, n_lon=14, radius=0.5, noise=0.0)
    bvh = fclb.bvh_build(v, t, st)
    oid = ref_oracle.bvh_obb_create(v, t)
    shapes = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.SPHERE, 0, (0.2,))]
    table = fclb.shapes_upload(shapes)
    ident = np.tile(np.concatenate([np.eye(3).reshape(9), np.zeros(3)]), (4, 1)).astype(dtype)
    # empty batch
    c, prim, toc = fclb.translational_ccd_mesh_batch_host(bvh, table, np.zeros(0, np.uint32), ident[:0], ident[:0],
                                                          np.zeros((0, 4), dtype), st)
    assert c.size == 0
    # 0: far away, no motion; 1: inside the mesh's surface shell, no motion; 2: starts in contact, moves away; 3: grazing sweep
    ps = ident.copy()
    ps[0, 9:] = (3, 0, 0)
    ps[1, 9:] = (0.5, 0, 0)
    ps[2, 9:] = (0.5, 0, 0)
    ps[3, 9:] = (-2, 0.69, 0)
    disp = np.array([[1, 0, 0, 0.0], [1, 0, 0, 0.0], [1, 0, 0, 2.0], [1, 0, 0, 4.0]], dtype)
    ids = np.array([0, 1, 0, 1], np.uint32)
    for request_type in (0, 1, 2):
        c, prim, toc = fclb.translational_ccd_mesh_batch_host(bvh, table, ids, ps, ident, disp, st, request_type=request_type,
                                                              max_contacts=1000, max_keep=64)
        ec, eprim, etoc = ref_oracle.translational_ccd_mesh_batch(oid, shapes, ids, ps, ident, disp, request_type=request_type,
                                                                  max_contacts=1000, keep=64)
        assert np.array_equal(c, ec) and np.array_equal(prim, eprim) and np.array_equal(toc, etoc)
        assert c[0] == 0 and c[1] > 0 and c[2] > 0
    # a Convex with six vertices takes the reference's fit6 box: refused, loudly
    octa_v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64) * 0.2
    octa_f = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    faces = np.array([x for f in octa_f for x in (3,) + f], np.int32)  # the reference's encoding: count, then the indices
    slot = fclb.convex_upload(octa_v, faces, len(octa_f))
    t6 = fclb.shapes_upload([(scenes.CONVEX, slot, ())])
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_mesh_batch_host(bvh, t6, np.zeros(1, np.uint32), ps[:1], ident[:1], disp[:1], st)
    # a tree of the other scalar type is refused
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_mesh_batch_host(bvh, table, ids, ps.astype(np.float32), ident.astype(np.float32),
                                               disp.astype(np.float32), fclb.F32)
    fclb.release(table)
    fclb.release(t6)
    fclb.bvh_release(bvh)
