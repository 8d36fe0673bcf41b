"""Parity of the heightmap-shape scan (fclb_heightmap_shape_collide_batch_*) against
fcl::collide(HeightMapCollisionGeometry, tf_hm, Shape, tf_shape) of the reference
(heightmap_solver_traverse-inl.h:23-118) on the same bottom layer: boolean result
(max_contacts=1) and contact counts (all contacts, capped) for every shape type."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

HALF, RES = 64, 0.01
PRIMS = {
    "box": (scenes.BOX, 0, (0.12, 0.08, 0.1)),
    "sphere": (scenes.SPHERE, 0, (0.06,)),
    "ellipsoid": (scenes.ELLIPSOID, 0, (0.08, 0.05, 0.06)),
    "capsule": (scenes.CAPSULE, 0, (0.03, 0.12)),
    "cone": (scenes.CONE, 0, (0.05, 0.12)),
    "cylinder": (scenes.CYLINDER, 0, (0.04, 0.12)),
}


def setup_scene(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = scenes.terrain_points(40_000, HALF * RES)
    hid = ref_oracle.heightmap_create(pts, RES, HALF)
    ref_h, upper = ref_oracle.heightmap_export(hid, dtype, HALF)
    hm = fclb.heightmap_upload(ref_h, RES, upper)
    hull = scenes.ellipsoid_mesh(0.05, 0.075, 0.1)
    hull16 = scenes.random_hull16(scale=(0.08, 0.06, 0.1))
    slots = [fclb.convex_upload(*hull), fclb.convex_upload(*hull16)]
    rslots = [ref_oracle.register_convex(*hull), ref_oracle.register_convex(*hull16)]
    return st, hid, hm, slots, rslots


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_heightmap_shape_every_type(fclb, ref_oracle, dtype):
    st, hid, hm, slots, rslots = setup_scene(fclb, ref_oracle, dtype)
    n = 2000
    cases = [(k, [v], [v]) for k, v in PRIMS.items()]
    cases.append(("convex58", [(scenes.CONVEX, slots[0], ())], [(scenes.CONVEX, rslots[0], ())]))
    cases.append(("convex16", [(scenes.CONVEX, slots[1], ())], [(scenes.CONVEX, rslots[1], ())]))
    mixed = list(PRIMS.values())
    cases.append(("mixed", mixed + [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())],
                  mixed + [(scenes.CONVEX, rslots[0], ()), (scenes.CONVEX, rslots[1], ())]))
    for ci, (name, shapes, rshapes) in enumerate(cases):
        table = fclb.shapes_upload(shapes)
        p_hm, p_sh = scenes.heightmap_query_poses(n, dtype, HALF * RES, -0.05, 0.4, seed=4300 + ci)
        ids = (np.arange(n) % len(shapes)).astype(np.uint32)
        for mc in (1, 2**31 - 1, 3):
            req = fclb.make_request(max_contacts=mc)
            counts, pix = fclb.heightmap_shape_collide_batch_host(hm, table, ids, p_hm, p_sh, st, req, want_pixel=True)
            e_counts, e_pix = ref_oracle.heightmap_shape_collide_batch(hid, rshapes, ids, p_hm, p_sh, threads=8,
                                                                       max_contacts=mc)
            mism = np.nonzero(counts != e_counts)[0]
            if mc != 3:
                n_pix, n_leaf = fclb.scene_last_visit_counts()
                print(f"[heightmap-{name} {np.dtype(dtype).name} max_contacts={mc}] n={n} "
                      f"colliding={int((e_counts > 0).sum())} contacts={int(e_counts.sum())} mismatches={len(mism)} "
                      f"{mism[:8].tolist()}; pixels read/query {n_pix / n:.1f}, boxes tested/query {n_leaf / n:.1f}")
            assert len(mism) == 0, (name, mc, mism[:10], counts[mism[:10]], e_counts[mism[:10]])
            assert ((pix >= 0) == (e_counts > 0)).all()
        fclb.release(table)
    fclb.heightmap_release(hm)


def test_heightmap_large_roi_and_edges(fclb, ref_oracle):
    """A shape covering more than 1/8 of the map: the reference switches to the layer-pyramid
    traversal (heightmap_solver_traverse-inl.h:43-63); the pixel set must still agree."""
    st, hid, hm, slots, rslots = setup_scene(fclb, ref_oracle, np.float64)
    shapes = [(scenes.BOX, 0, (0.9, 0.7, 0.2)), (scenes.SPHERE, 0, (0.45,)), (scenes.CYLINDER, 0, (0.4, 0.15))]
    table = fclb.shapes_upload(shapes)
    n = 60
    p_hm, p_sh = scenes.heightmap_query_poses(n, np.float64, 0.3, 0.0, 0.5, seed=99)
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    req = fclb.make_request(max_contacts=2**31 - 1)
    counts, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, p_hm, p_sh, st, req)
    e_counts, _ = ref_oracle.heightmap_shape_collide_batch(hid, shapes, ids, p_hm, p_sh, threads=8, max_contacts=2**31 - 1)
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[heightmap large ROI] contacts ours={int(counts.sum())} ref={int(e_counts.sum())} mismatching queries={len(mism)}")
    assert len(mism) == 0, (mism[:10], counts[mism[:10]], e_counts[mism[:10]])
    # empty batch, zero max_contacts, far away
    c, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids[:0], p_hm[:0], p_sh[:0], st, req)
    assert c.size == 0
    c, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, p_hm, p_sh, st, fclb.make_request(max_contacts=0))
    assert not c.any()
    far = p_sh.copy()
    far[:, 9:] += 50.0
    c, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, p_hm, far, st, req)
    assert not c.any()
    fclb.release(table)
    fclb.heightmap_release(hm)
