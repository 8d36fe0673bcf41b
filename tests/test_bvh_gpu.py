"""Parity of the warp-per-query mesh-mesh traversal (fclb_bvh_collide_batch_*) against
fcl::collide(BVHModel<OBBRSS>, BVHModel<OBBRSS>) of the reference on the SAME trees
(built by the reference's own BVH builder and exported through the oracle, which is
what an integration would upload).  Boolean result and contact counts are exact:
the reference's own mesh test demands set equality of contact pairs
(test/test_fcl_collision.cpp:348-351)."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def make_meshes(fclb, ref_oracle, dtype, small=False):
    if small:
        m1 = scenes.noisy_uv_sphere(n_lat=13, n_lon=24)
        m2 = scenes.noisy_torus(n_major=24, n_minor=12)
    else:
        m1 = scenes.noisy_uv_sphere()
        m2 = scenes.noisy_torus()
    ids = [ref_oracle.bvh_create(*m1), ref_oracle.bvh_create(*m2)]
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    handles = []
    for i in ids:
        obb, child, tri = ref_oracle.bvh_export(i, dtype)
        handles.append(fclb.bvh_upload(obb, child, tri, st))
    return ids, handles, st


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mesh_mesh_boolean_and_counts(fclb, ref_oracle, dtype):
    ids, handles, st = make_meshes(fclb, ref_oracle, dtype)
    n = 20_000
    poses1, poses2 = scenes.config_c3_poses(n, dtype, extent=2.0)
    # boolean, early exit
    req = fclb.make_request(max_contacts=1)
    counts, pair = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1, poses2, st, req, want_pair=True)
    n_bv, n_leaf = fclb.bvh_last_visit_counts()
    e_counts, e_pair = ref_oracle.bvh_collide_batch(ids[0], ids[1], poses1, poses2, threads=8, max_contacts=1)
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[mesh-mesh bool {np.dtype(dtype).name}] n={n} colliding={int(e_counts.sum())} mismatches={len(mism)} "
          f"{mism[:10].tolist()}; device BV tests {n_bv}, leaf tests {n_leaf}")
    assert len(mism) == 0
    same_pair = (pair == e_pair).all(axis=1)[e_counts > 0].mean()
    print(f"   first reported pair equals the reference's DFS-first pair in {same_pair:.3f} of colliding queries (order-dependent output)")
    # all contacts: count parity
    m = 4_000
    req = fclb.make_request(max_contacts=2**31 - 1)
    counts, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1[:m], poses2[:m], st, req)
    n_bv, n_leaf = fclb.bvh_last_visit_counts()
    e_counts, _ = ref_oracle.bvh_collide_batch(ids[0], ids[1], poses1[:m], poses2[:m], threads=8, want_pair=False,
                                               max_contacts=2**31 - 1)
    r_bv, r_leaf, r_hit = ref_oracle.bvh_visit_counts(ids[0], ids[1], poses1[:m], poses2[:m], threads=8)
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[mesh-mesh count {np.dtype(dtype).name}] n={m} total contacts ours={int(counts.sum())} ref={int(e_counts.sum())} "
          f"mismatches={len(mism)}; BV tests ours={n_bv} ref={int(r_bv.sum())}; leaf tests ours={n_leaf} ref={int(r_leaf.sum())}")
    assert len(mism) == 0
    assert np.array_equal(r_hit, e_counts)
    # all-contacts traversal visits exactly the reference's node pairs (same tree, same predicate)
    assert n_bv == int(r_bv.sum()) and n_leaf == int(r_leaf.sum())
    # capped count
    req = fclb.make_request(max_contacts=5)
    counts5, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1[:m], poses2[:m], st, req)
    assert np.array_equal(counts5, np.minimum(e_counts, 5))
    for h in handles:
        fclb.bvh_release(h)


def test_mesh_mesh_small_and_edge_cases(fclb, ref_oracle):
    ids, handles, st = make_meshes(fclb, ref_oracle, np.float64, small=True)
    n = 3_000
    poses1, poses2 = scenes.config_c3_poses(n, np.float64, extent=1.2, seed=5)
    req = fclb.make_request(max_contacts=1)
    counts, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1, poses2, st, req)
    e_counts, _ = ref_oracle.bvh_collide_batch(ids[0], ids[1], poses1, poses2, threads=4, max_contacts=1)
    assert np.array_equal(counts, e_counts)
    # self pair, identical poses: deep overlap everywhere (stress for the stack)
    req = fclb.make_request(max_contacts=2**31 - 1)
    c, _ = fclb.bvh_collide_batch_host(handles[0], handles[0], poses2[:4], poses2[:4], st, req)
    e, _ = ref_oracle.bvh_collide_batch(ids[0], ids[0], poses2[:4], poses2[:4], max_contacts=2**31 - 1)
    print("self-collide contact counts", c.tolist(), e.tolist())
    assert np.array_equal(c, e)
    # empty batch, zero max_contacts
    c, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1[:0], poses2[:0], st, req)
    assert c.size == 0
    c, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], poses1[:10], poses2[:10], st, fclb.make_request(max_contacts=0))
    assert not c.any()
    for h in handles:
        fclb.bvh_release(h)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mesh_mesh_contact_generation(fclb, ref_oracle, dtype):
    """request.useDefaultPenetration(): contact points / normal / depth of Intersect::intersect_Triangle
    (intersect-inl.h:795-888).  Counts identical; contacts compared per triangle pair (b1, b2), bit-exact."""
    ids, handles, st = make_meshes(fclb, ref_oracle, dtype, small=True)
    n = 1500
    poses1, poses2 = scenes.config_c3_poses(n, dtype, extent=1.0, seed=8)
    rng = np.random.Generator(np.random.PCG64(4))
    poses2 = scenes.random_poses(rng, n, 0.2, dtype)  # both meshes posed: the contact frame quirk (tf2) shows
    keep = 128
    req = fclb.make_request(max_contacts=2**31 - 1, penetration_mode=1)
    counts, cid, contacts = fclb.bvh_collide_contacts_batch_host(handles[0], handles[1], poses1, poses2, st, req, keep)
    e_counts, e_id, e_contacts = ref_oracle.bvh_collide_contacts_batch(ids[0], ids[1], poses1, poses2, 4096, threads=8,
                                                                       max_contacts=2**31 - 1, penetration_mode=1)
    assert int(e_counts.max()) <= 4096
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[mesh-mesh contacts {np.dtype(dtype).name}] n={n} colliding={int((e_counts > 0).sum())} "
          f"contacts={int(e_counts.sum())} count mismatches={len(mism)}")
    assert len(mism) == 0
    n_cmp = n_same = 0
    for q in np.nonzero(counts)[0]:
        ref = {}
        for j in range(int(e_counts[q])):
            ref.setdefault((int(e_id[q, j, 0]), int(e_id[q, j, 1])), []).append(e_contacts[q, j])
        seen = {}
        for j in range(int(min(counts[q], keep))):
            key = (int(cid[q, j, 0]), int(cid[q, j, 1]))
            k = seen.get(key, 0)
            seen[key] = k + 1
            assert key in ref and k < len(ref[key]), (q, key)
            n_cmp += 1
            n_same += int(np.array_equal(contacts[q, j], ref[key][k]))
    print(f"   contacts compared {n_cmp}, bit-identical {n_same}")
    assert n_cmp > 0 and n_same == n_cmp
    # capped
    req = fclb.make_request(max_contacts=3, penetration_mode=1)
    c3, _, _ = fclb.bvh_collide_contacts_batch_host(handles[0], handles[1], poses1, poses2, st, req, 4)
    assert np.array_equal(c3, np.minimum(e_counts, 3))
    for h in handles:
        fclb.bvh_release(h)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [2, 3])
def test_mesh_mesh_mpr_penetration(fclb, ref_oracle, dtype, mode):
    """request.useDirectedPenetration(dir) / useIncrementalMinimumDistancePenetration(dir) on mesh pairs:
    collisionPenetrationMPR (collision_penetration-inl.h:189-252) = boolean collide, then computePenetrationMPR on the
    (triangle b1, triangle b2) of every contact.  Counts identical; contacts compared per triangle pair, bit-exact."""
    ids, handles, st = make_meshes(fclb, ref_oracle, dtype, small=True)
    n = 1200
    poses1, _ = scenes.config_c3_poses(n, dtype, extent=1.0, seed=18)
    rng = np.random.Generator(np.random.PCG64(14))
    poses2 = scenes.random_poses(rng, n, 0.2, dtype)
    keep = 4096
    d = (0.0, -0.6, 0.8)  # unit, as CollisionRequest::useDirectedPenetration stores it (it normalises in S)
    req = fclb.make_request(max_contacts=2**31 - 1, penetration_mode=mode, direction=d)
    counts, cid, contacts = fclb.bvh_collide_contacts_batch_host(handles[0], handles[1], poses1, poses2, st, req, keep)
    e_counts, e_id, e_contacts = ref_oracle.bvh_collide_contacts_batch(ids[0], ids[1], poses1, poses2, keep, threads=8,
                                                                       max_contacts=2**31 - 1, penetration_mode=mode, direction=d)
    assert int(e_counts.max()) <= keep
    assert np.array_equal(counts, e_counts)
    n_cmp = n_same = 0
    for q in np.nonzero(counts)[0]:
        ref = {(int(e_id[q, j, 0]), int(e_id[q, j, 1])): e_contacts[q, j] for j in range(int(e_counts[q]))}
        for j in range(int(counts[q])):
            key = (int(cid[q, j, 0]), int(cid[q, j, 1]))
            assert key in ref, (q, key)
            n_cmp += 1
            n_same += int(np.array_equal(contacts[q, j], ref[key]))
    print(f"[mesh-mesh MPR penetration mode {mode} {np.dtype(dtype).name}] colliding={int((e_counts > 0).sum())} contacts compared "
          f"{n_cmp}, bit-identical {n_same}")
    assert n_cmp > 1000 and n_same == n_cmp
    for h in handles:
        fclb.bvh_release(h)
