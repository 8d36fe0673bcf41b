"""MPR penetration of scene contacts (fclb_scene_shape_contacts_batch_*) against fcl::collide of the
reference with DirectedPenetration / IncrementalMinimumPenetration requests on a mesh, a heightmap and
an octree (collisionPenetrationMPR, narrowphase/collision_penetration-inl.h:189-252).
The contact lists are compared as sets keyed by Contact::b1 (the order of a list follows each
implementation's traversal): for every contact we store, the reference holds a contact with the same b1
and bit-identical normal / position / depth."""
import numpy as np
import pytest

import scenes
from test_octree_gpu import octree_points

pytestmark = pytest.mark.gpu
KEEP = 96


def compare(name, counts, b1, contacts, e_counts, e_b1, e_contacts):
    assert np.array_equal(counts, e_counts), name
    n_cmp = n_same = 0
    for q in np.nonzero(counts)[0]:
        k = int(min(counts[q], KEEP))
        ref = {int(e_b1[q, j]): e_contacts[q, j] for j in range(int(min(e_counts[q], e_b1.shape[1])))}
        for j in range(k):
            key = int(b1[q, j])
            assert key in ref, (name, q, key)
            n_cmp += 1
            n_same += int(np.array_equal(contacts[q, j], ref[key]))
    print(f"[{name}] queries with contacts {int((counts > 0).sum())}, contacts compared {n_cmp}, bit-identical {n_same}")
    assert n_cmp > 0 and n_same == n_cmp


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_contact_penetration(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    hull = scenes.ellipsoid_mesh(0.05, 0.075, 0.1)
    shapes = [(scenes.BOX, 0, (0.12, 0.08, 0.1)), (scenes.SPHERE, 0, (0.06,)), (scenes.CAPSULE, 0, (0.03, 0.12)),
              (scenes.CONVEX, fclb.convex_upload(*hull), ())]
    rshapes = shapes[:3] + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())]
    table = fclb.shapes_upload(shapes)
    n = 600
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    # mesh: a small terrain patch
    v, t = scenes.c4_scene_mesh(grid=40, n_boxes=20)
    v = v * 0.25
    mid = ref_oracle.bvh_create(v, t)
    bvh = fclb.bvh_build(v, t, st)
    rng = np.random.Generator(np.random.PCG64(5))
    p_mesh = scenes.random_poses(rng, n, 0.2, dtype)
    local = scenes.random_poses(rng, n, 0.4, np.float64)
    local[:, 11] = rng.uniform(0.0, 0.2, size=n)
    p_shape = scenes.compose_poses(p_mesh, local, dtype)
    # heightmap and octree
    pts = scenes.terrain_points(40_000, 0.64)
    hid = ref_oracle.heightmap_create(pts, 0.01, 64)
    heights, upper = ref_oracle.heightmap_export(hid, dtype, 64)
    hm = fclb.heightmap_upload(heights, 0.01, upper)
    p_hm, p_hs = scenes.heightmap_query_poses(n, dtype, 0.64, -0.05, 0.4, seed=11)
    oid = ref_oracle.octree_create(octree_points(), 0.01, 64)
    ch, full, leaf, root, n_layers = ref_oracle.octree_export(oid, dtype)
    octree = fclb.octree_upload(ch, full, leaf, root, n_layers)
    p_oc, p_os = scenes.heightmap_query_poses(n, dtype, 0.4, -0.25, 0.25, seed=12)
    cases = [("mesh", fclb.SCENE_BVH, bvh, 0, mid, p_mesh, p_shape), ("heightmap", fclb.SCENE_HEIGHTMAP, hm, 1, hid, p_hm, p_hs),
             ("octree", fclb.SCENE_OCTREE, octree, 2, oid, p_oc, p_os)]
    for name, kind, handle, rkind, rid, ps, psh in cases:
        for mode, direction in ((2, (0.0, 0.0, 1.0)), (3, (0.6, 0.0, 0.8))):
            req = fclb.make_request(max_contacts=2**31 - 1, penetration_mode=mode, direction=direction)
            counts, b1, contacts = fclb.scene_shape_contacts_batch_host(kind, handle, table, ids, ps, psh, st, req, KEEP)
            e_counts, e_b1, e_contacts = ref_oracle.scene_shape_contacts_batch(rkind, rid, rshapes, ids, ps, psh, 1024, threads=8,
                                                                               max_contacts=2**31 - 1, penetration_mode=mode,
                                                                               direction=direction)
            assert int(e_counts.max()) <= 1024
            compare(f"{name} mode={mode} {np.dtype(dtype).name}", counts, b1, contacts, e_counts, e_b1, e_contacts)
        # capped request: counts follow max_contacts
        req = fclb.make_request(max_contacts=2, penetration_mode=2, direction=(0.0, 0.0, 1.0))
        c2, _, _ = fclb.scene_shape_contacts_batch_host(kind, handle, table, ids, ps, psh, st, req, 4)
        assert np.array_equal(c2, np.minimum(e_counts, 2))
    # a boolean request through the same entry point: the ids of every kept contact, no contact geometry
    cb, bb, ctb = fclb.scene_shape_contacts_batch_host(fclb.SCENE_BVH, bvh, table, ids, p_mesh, p_shape, st,
                                                       fclb.make_request(max_contacts=2**31 - 1), KEEP)
    eb, _ = ref_oracle.mesh_shape_collide_batch(mid, rshapes, ids, p_mesh, p_shape, threads=8, max_contacts=2**31 - 1)
    assert np.array_equal(cb, eb) and not ctb.any() and (bb[cb > 0, 0] >= 0).all()
    fclb.bvh_release(bvh)
    fclb.heightmap_release(hm)
    fclb.octree_release(octree)
    fclb.release(table)
