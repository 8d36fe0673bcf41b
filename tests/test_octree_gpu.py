"""Parity of the octree-shape traversal (fclb_octree_shape_collide_batch_*) against
fcl::collide(Octree2CollisionGeometry, tf_octree, Shape, tf_shape) of the reference
(octree2_solver_traverse-inl.h:12-136) on the same node arrays: boolean result (max_contacts=1)
and contact counts (all contacts, capped) for every shape type, float and double.  Fully occupied
inner nodes count as ONE box, so the counts also pin the traversal's use of inner_nodes_fully_occupied."""
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


def octree_points(seed=7):
    """A thick terrain slab (many fully occupied 2x2x2 cells and inner nodes) plus scattered points."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.arange(-40, 40) * RES + RES / 2
    X, Y, Z = np.meshgrid(g, g, np.arange(-16, 16) * RES + RES / 2, indexing="ij")
    surf = 0.1 * np.sin(4 * X) * np.cos(3 * Y)
    keep = Z < surf
    slab = np.stack([X[keep], Y[keep], Z[keep]], axis=1)
    extra = rng.uniform(-0.7, 0.7, size=(4000, 3))
    return np.ascontiguousarray(np.concatenate([slab, extra]))


def setup_scene(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    oid = ref_oracle.octree_create(octree_points(), RES, HALF)
    ch, full, leaf, root, n_layers = ref_oracle.octree_export(oid, dtype)
    assert full.any() and (leaf == 255).any() and (leaf != 255).any()
    oct_h = fclb.octree_upload(ch, full, leaf, root, n_layers)
    hulls = [scenes.ellipsoid_mesh(0.05, 0.075, 0.1), scenes.random_hull16(scale=(0.08, 0.06, 0.1))]
    slots = [fclb.convex_upload(*m) for m in hulls]
    rslots = [ref_oracle.register_convex(*m) for m in hulls]
    return st, oid, oct_h, slots, rslots, (len(full), len(leaf), n_layers)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_octree_shape_every_type(fclb, ref_oracle, dtype):
    st, oid, oct_h, slots, rslots, sizes = setup_scene(fclb, ref_oracle, dtype)
    print(f"octree: {sizes[0]} inner nodes, {sizes[1]} leaf nodes, {sizes[2]} layers")
    n = 2000
    cases = [(k, [v], [v]) for k, v in PRIMS.items()]
    cases.append(("convex58", [(scenes.CONVEX, slots[0], ())], [(scenes.CONVEX, rslots[0], ())]))
    cases.append(("convex16", [(scenes.CONVEX, slots[1], ())], [(scenes.CONVEX, rslots[1], ())]))
    mixed = list(PRIMS.values())
    cases.append(("mixed", mixed + [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())],
                  mixed + [(scenes.CONVEX, rslots[0], ()), (scenes.CONVEX, rslots[1], ())]))
    for ci, (name, shapes, rshapes) in enumerate(cases):
        table = fclb.shapes_upload(shapes)
        p_oct, p_sh = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=4700 + ci)
        ids = (np.arange(n) % len(shapes)).astype(np.uint32)
        for mc in (1, 2**31 - 1, 3):
            req = fclb.make_request(max_contacts=mc)
            counts, node = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, req, want_node=True)
            e_counts, e_node = ref_oracle.octree_shape_collide_batch(oid, rshapes, ids, p_oct, p_sh, threads=8, max_contacts=mc)
            mism = np.nonzero(counts != e_counts)[0]
            if mc != 3:
                n_node, n_leaf = fclb.scene_last_visit_counts()
                print(f"[octree-{name} {np.dtype(dtype).name} max_contacts={mc}] n={n} colliding={int((e_counts > 0).sum())} "
                      f"contacts={int(e_counts.sum())} mismatches={len(mism)} {mism[:8].tolist()}; node boxes/query "
                      f"{n_node / n:.1f}, voxel boxes/query {n_leaf / n:.1f}")
            assert len(mism) == 0, (name, mc, mism[:10], counts[mism[:10]], e_counts[mism[:10]])
            assert ((node >= 0) == (e_counts > 0)).all()
        fclb.release(table)
    # edge cases
    table = fclb.shapes_upload([PRIMS["box"]])
    p_oct, p_sh = scenes.heightmap_query_poses(8, dtype, 0.4, -0.25, 0.25, seed=1)
    ids = np.zeros(8, np.uint32)
    c, _ = fclb.octree_shape_collide_batch_host(oct_h, table, ids[:0], p_oct[:0], p_sh[:0], st, fclb.make_request())
    assert c.size == 0
    c, _ = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, fclb.make_request(max_contacts=0))
    assert not c.any()
    fclb.release(table)
    fclb.octree_release(oct_h)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_octree_built_here(fclb, ref_oracle, dtype):
    """fclb_octree_build (our host mirror of Octree<S>::rebuildTree + upload) instead of the reference's arrays:
    counts AND the reported node ids (the reference's node numbering) equal the reference's."""
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = octree_points(11)
    oid = ref_oracle.octree_create(pts, RES, HALF)
    oct_h = fclb.octree_build(pts, RES, HALF, st)
    shapes = list(PRIMS.values())
    table = fclb.shapes_upload(shapes)
    n = 4000
    p_oct, p_sh = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=5100)
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    for mc in (1, 2**31 - 1):
        req = fclb.make_request(max_contacts=mc)
        counts, node = fclb.octree_shape_collide_batch_host(oct_h, table, ids, p_oct, p_sh, st, req, want_node=True)
        e_counts, e_node = ref_oracle.octree_shape_collide_batch(oid, shapes, ids, p_oct, p_sh, threads=8, max_contacts=mc)
        assert np.array_equal(counts, e_counts)
        assert ((node >= 0) == (e_counts > 0)).all()
        single = e_counts == 1  # one contact: the id is determined
        if mc > 1:
            assert single.any() and np.array_equal(node[single], e_node[single])
    print(f"[octree built here {np.dtype(dtype).name}] n={n} colliding={int((e_counts > 0).sum())} contacts={int(e_counts.sum())}")
    fclb.release(table)
    fclb.octree_release(oct_h)
