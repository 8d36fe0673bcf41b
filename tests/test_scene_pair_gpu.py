"""Parity of the scene-pair traversal (fclb_scene_pair_collide_batch_*) against fcl::collide of the
reference on the same structures, for the five non-convex pairs of the collision matrix
(collision_func_matrix-inl.h:774-857):
    heightmap-heightmap   heightMapPairIntersect   (heightmap_solver_traverse-inl.h:189)
    heightmap-mesh        heightMapBVHIntersect    (:298)
    heightmap-octree      heightMapOctreeIntersect (:406)
    octree-mesh           octreeBVHIntersect       (octree2_solver_traverse-inl.h:138)
    octree-octree         octreePairIntersect      (:290)
Checked per query: the boolean (max_contacts = 1), the capped count (max_contacts = 3), the contact
count with all contacts requested, and -- for every query whose contacts fit the kept list -- the multiset of
(b1, b2) contact ids (encodePixel / encodeOctree2Node / triangle id).  Contact order follows the device
traversal and is not compared.  float and double."""
import numpy as np
import pytest

import scenes
from test_octree_gpu import octree_points

pytestmark = pytest.mark.gpu

RES = 0.01
KEEP = 4096


def blob_points(seed, r=0.12, n=6000, upper_half=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.normal(size=(n, 3))
    p /= np.linalg.norm(p, axis=1)[:, None]
    p *= r * rng.uniform(0.2, 1.0, size=(n, 1)) ** (1 / 3)
    if upper_half:
        p[:, 2] = np.abs(p[:, 2])
    return np.ascontiguousarray(p)


def upload_heightmap(fclb, ref_oracle, pts, half, dtype):
    hid = ref_oracle.heightmap_create(pts, RES, half)
    heights, upper = ref_oracle.heightmap_export(hid, dtype, half)
    return hid, fclb.heightmap_upload(heights, RES, upper)


def upload_octree(fclb, ref_oracle, pts, half, dtype):
    oid = ref_oracle.octree_create(pts, RES, half)
    ch, full, leaf, root, n_layers = ref_oracle.octree_export(oid, dtype)
    return oid, fclb.octree_upload(ch, full, leaf, root, n_layers)


def check_pair(fclb, ref_oracle, name, dtype, k1, ref1, dev1, k2, ref2, dev2, p1, p2):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    n = len(p1)
    for mc in (1, 3, 2**31 - 1):
        keep = KEEP if mc > 3 else 4
        counts, b1, b2 = fclb.scene_pair_collide_batch_host(k1, dev1, k2, dev2, p1, p2, st, fclb.make_request(max_contacts=mc), keep)
        e_counts, e_b1, e_b2 = ref_oracle.scene_pair_collide_batch(k1, ref1, k2, ref2, p1, p2, keep, threads=8, max_contacts=mc)
        mism = np.nonzero(counts != e_counts)[0]
        if mc != 3:
            n_node, n_leaf = fclb.scene_last_visit_counts()
            print(f"[{name} {np.dtype(dtype).name} max_contacts={mc}] n={n} colliding={int((e_counts > 0).sum())} "
                  f"contacts={int(e_counts.sum())} mismatches={len(mism)} {mism[:8].tolist()}; node pairs/query "
                  f"{n_node / n:.1f}, leaf pairs/query {n_leaf / n:.1f}")
        assert len(mism) == 0, (name, mc, mism[:10], counts[mism[:10]], e_counts[mism[:10]])
        # stored ids: valid for the first min(count, keep) slots, -1 after
        kept = np.minimum(counts, keep)
        slot = np.arange(keep)[None, :]
        assert ((b1 >= 0) == (slot < kept[:, None])).all() and ((b2 >= 0) == (slot < kept[:, None])).all()
        if mc > 3:
            n_sets = 0
            for q in np.nonzero((counts > 0) & (counts <= keep))[0]:
                # (a heightmap-octree contact names the octree side by its node index only, so the voxels of one
                # partial leaf repeat an id pair: compare multisets)
                got = sorted(zip(b1[q, :counts[q]].tolist(), b2[q, :counts[q]].tolist()))
                exp = sorted(zip(e_b1[q, :counts[q]].tolist(), e_b2[q, :counts[q]].tolist()))
                assert got == exp, (name, q, sorted(set(got) ^ set(exp))[:6])
                n_sets += 1
            print(f"    contact-id sets identical for {n_sets} queries")
            assert n_sets > 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_pairs(fclb, ref_oracle, dtype):
    n = 300
    hidA, hmA = upload_heightmap(fclb, ref_oracle, scenes.terrain_points(40_000, 64 * RES), 64, dtype)
    hidB, hmB = upload_heightmap(fclb, ref_oracle, blob_points(11, upper_half=True), 16, dtype)
    oidA, octA = upload_octree(fclb, ref_oracle, octree_points(), 64, dtype)
    oidB, octB = upload_octree(fclb, ref_oracle, blob_points(12), 16, dtype)
    v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24, radius=0.12, noise=0.02)
    mid = ref_oracle.bvh_create(v, t)
    obb, child, tri_verts = ref_oracle.bvh_export(mid, dtype)
    mesh = fclb.bvh_upload(obb, child, tri_verts, fclb.F32 if dtype == np.float32 else fclb.F64)
    H, O, M = fclb.SCENE_HEIGHTMAP, fclb.SCENE_OCTREE, fclb.SCENE_BVH
    cases = [
        ("heightmap-heightmap", H, hidA, hmA, H, hidB, hmB, -0.2, 0.5, 5101),
        ("heightmap-mesh", H, hidA, hmA, M, mid, mesh, -0.2, 0.6, 5102),
        ("heightmap-octree", H, hidA, hmA, O, oidB, octB, -0.2, 0.6, 5103),
        ("octree-mesh", O, oidA, octA, M, mid, mesh, -0.45, 0.45, 5104),
        ("octree-octree", O, oidA, octA, O, oidB, octB, -0.45, 0.45, 5105),
    ]
    for name, k1, r1, d1, k2, r2, d2, zlo, zhi, seed in cases:
        p1, p2 = scenes.heightmap_query_poses(n, dtype, 0.4, zlo, zhi, seed=seed)
        check_pair(fclb, ref_oracle, name, dtype, k1, r1, d1, k2, r2, d2, p1, p2)
    # edge cases: empty batch, max_contacts = 0, counts only, unsupported order
    p1, p2 = scenes.heightmap_query_poses(8, dtype, 0.4, -0.1, 0.3, seed=1)
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    c, _, _ = fclb.scene_pair_collide_batch_host(H, hmA, M, mesh, p1[:0], p2[:0], st, fclb.make_request())
    assert c.size == 0
    c, _, _ = fclb.scene_pair_collide_batch_host(H, hmA, M, mesh, p1, p2, st, fclb.make_request(max_contacts=0))
    assert not c.any()
    c, b1, b2 = fclb.scene_pair_collide_batch_host(H, hmA, O, octB, p1, p2, st, fclb.make_request(max_contacts=5))
    assert b1 is None and b2 is None and (c <= 5).all()
    with pytest.raises(fclb.FclbError):
        fclb.scene_pair_collide_batch_host(O, octA, H, hmA, p1, p2, st, fclb.make_request())
    with pytest.raises(fclb.FclbError):
        fclb.scene_pair_collide_batch_host(M, mesh, M, mesh, p1, p2, st, fclb.make_request())
    for h in (hmA, hmB):
        fclb.heightmap_release(h)
    for h in (octA, octB):
        fclb.octree_release(h)
    fclb.bvh_release(mesh)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_pair_penetration(fclb, ref_oracle, dtype):
    """DirectedPenetration / IncrementalMinimumPenetration requests on the five pair kinds
    (collisionPenetrationMPR, collision_penetration-inl.h:189-252): counts identical, and the multiset of
    contact records (b1, b2, normal, position, depth) bit-identical to the reference's for every query whose
    contacts fit the kept list."""
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    n, keep = 120, 2048
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
        for mode, direction in ((2, (0.0, 0.0, 1.0)), (3, (0.6, 0.0, 0.8))):
            req = fclb.make_request(max_contacts=2**31 - 1, penetration_mode=mode, direction=direction)
            counts, b1, b2, contacts = fclb.scene_pair_contacts_batch_host(k1, d1, k2, d2, p1, p2, st, req, keep)
            e_counts, e_b1, e_b2, e_contacts = ref_oracle.scene_pair_collide_batch(
                k1, r1, k2, r2, p1, p2, keep, threads=8, want_contacts=True, max_contacts=2**31 - 1, penetration_mode=mode,
                direction=direction)
            assert np.array_equal(counts, e_counts), (name, mode)
            n_q = n_c = 0
            for q in np.nonzero((counts > 0) & (counts <= keep))[0]:
                k = int(counts[q])
                got = sorted((int(b1[q, j]), int(b2[q, j])) + tuple(contacts[q, j].tolist()) for j in range(k))
                exp = sorted((int(e_b1[q, j]), int(e_b2[q, j])) + tuple(e_contacts[q, j].tolist()) for j in range(k))
                assert got == exp, (name, mode, q)
                n_q += 1
                n_c += k
            print(f"[{name} mode={mode} {np.dtype(dtype).name}] colliding {int((counts > 0).sum())}, contacts "
                  f"{int(counts.sum())}; records bit-identical for {n_c} contacts of {n_q} queries")
            assert n_c > 0
    req = fclb.make_request(max_contacts=2, penetration_mode=2, direction=(0.0, 0.0, 1.0))
    p1, p2 = scenes.heightmap_query_poses(n, dtype, 0.4, 0.0, 0.4, seed=5301)
    c2, _, _, _ = fclb.scene_pair_contacts_batch_host(H, hmA, M, mesh, p1, p2, st, req, 4)
    assert (c2 <= 2).all() and c2.any()
    with pytest.raises(fclb.FclbError):
        fclb.scene_pair_contacts_batch_host(H, hmA, M, mesh, p1, p2, st, fclb.make_request(), 4)
    for h in (hmA, hmB):
        fclb.heightmap_release(h)
    for h in (octA, octB):
        fclb.octree_release(h)
    fclb.bvh_release(mesh)
