"""Parity of the mesh-shape traversal (fclb_bvh_shape_collide_batch_*) against
fcl::collide(BVHModel<OBBRSS>, tf1, Shape, tf2) of the reference
(OrientedNodeBVHSolver::MeshShapeIntersect, traversal/collision/bvh_solver-inl.h:8-72) on
the same tree: boolean result (max_contacts=1) and contact counts (all contacts, capped)
must be identical for every shape type, float and double."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

PRIMS = {
    "box": (scenes.BOX, 0, (0.3, 0.2, 0.25)),
    "sphere": (scenes.SPHERE, 0, (0.15,)),
    "ellipsoid": (scenes.ELLIPSOID, 0, (0.2, 0.12, 0.16)),
    "capsule": (scenes.CAPSULE, 0, (0.08, 0.3)),
    "cone": (scenes.CONE, 0, (0.12, 0.3)),
    "cylinder": (scenes.CYLINDER, 0, (0.1, 0.3)),
}


def setup_scene(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    mesh = scenes.noisy_torus()
    mid = ref_oracle.bvh_create(*mesh)
    bvh = fclb.bvh_build(mesh[0], mesh[1], st)
    hulls = [scenes.ellipsoid_mesh(0.2, 0.3, 0.4), scenes.random_hull16()]
    slots = [fclb.convex_upload(*m) for m in hulls]
    rslots = [ref_oracle.register_convex(*m) for m in hulls]
    return st, mid, bvh, slots, rslots


def poses_for(n, dtype, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return scenes.random_poses(rng, n, 0.6, dtype), scenes.random_poses(rng, n, 0.9, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mesh_shape_every_type(fclb, ref_oracle, dtype):
    st, mid, bvh, slots, rslots = setup_scene(fclb, ref_oracle, dtype)
    n = 3000
    cases = [(k, [v], [v]) for k, v in PRIMS.items()]
    cases.append(("convex58", [(scenes.CONVEX, slots[0], ())], [(scenes.CONVEX, rslots[0], ())]))
    cases.append(("convex16", [(scenes.CONVEX, slots[1], ())], [(scenes.CONVEX, rslots[1], ())]))
    mixed = list(PRIMS.values())
    cases.append(("mixed", mixed + [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())],
                  mixed + [(scenes.CONVEX, rslots[0], ()), (scenes.CONVEX, rslots[1], ())]))
    for ci, (name, shapes, rshapes) in enumerate(cases):
        table = fclb.shapes_upload(shapes)
        pm, ps = poses_for(n, dtype, 100 + ci)
        ids = (np.arange(n) % len(shapes)).astype(np.uint32)
        for mc in (1, 2**31 - 1, 3):
            req = fclb.make_request(max_contacts=mc)
            counts, tri = fclb.bvh_shape_collide_batch_host(bvh, table, ids, pm, ps, st, req, want_tri=True)
            e_counts, e_tri = ref_oracle.mesh_shape_collide_batch(mid, rshapes, ids, pm, ps, threads=8, max_contacts=mc)
            mism = np.nonzero(counts != e_counts)[0]
            if mc == 1:
                n_node, n_leaf = fclb.scene_last_visit_counts()
                print(f"[mesh-{name} {np.dtype(dtype).name}] n={n} colliding={int((e_counts > 0).sum())} "
                      f"mismatches={len(mism)} {mism[:8].tolist()}; node tests/query {n_node / n:.1f}, "
                      f"leaf tests/query {n_leaf / n:.1f}")
            assert len(mism) == 0, (name, mc, mism[:10], counts[mism[:10]], e_counts[mism[:10]])
            assert ((tri >= 0) == (e_counts > 0)).all()
        fclb.release(table)
    fclb.bvh_release(bvh)


def test_mesh_shape_edge_cases(fclb, ref_oracle):
    st, mid, bvh, slots, rslots = setup_scene(fclb, ref_oracle, np.float64)
    table = fclb.shapes_upload([PRIMS["box"]])
    pm, ps = poses_for(16, np.float64, 7)
    ids = np.zeros(16, np.uint32)
    c, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids[:0], pm[:0], ps[:0], st, fclb.make_request())
    assert c.size == 0
    c, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, pm, ps, st, fclb.make_request(max_contacts=0))
    assert not c.any()
    with pytest.raises(fclb.FclbError):
        fclb.bvh_shape_collide_batch_host(bvh, table, ids + 5, pm, ps, st, fclb.make_request())
    # a penetration request through the counting entry point: numContacts of the contact path (tests/test_scene_gjk_epa_gpu.py)
    c_pen, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, pm, ps, st, fclb.make_request(max_contacts=2**31 - 1, penetration_mode=1))
    c_bool, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, pm, ps, st, fclb.make_request(max_contacts=2**31 - 1))
    assert ((c_pen > 0) == (c_bool > 0)).mean() > 0.99
    fclb.release(table)
    fclb.bvh_release(bvh)
