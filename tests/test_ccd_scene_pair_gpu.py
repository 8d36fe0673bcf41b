"""Translational continuous collision between two scene geometries (heightmap / octree vs heightmap / octree):
fclb_translational_ccd_scene_pair_batch_host against the reference's fcl::translational_ccd (RunHeightMapPair,
RunHeightMapOctree / RunOctreeHeightMap: heightmap_ccd_solver-inl.h:369-779; RunOctreePair: octree2_ccd_solver-inl.h:467-922)
on the same seeded inputs.
Bar: contact counts, (code 1, code 2) IN THE REFERENCE'S ORDER, toc intervals and both boxes bit-identical."""
import numpy as np
import pytest

import parity_util
import scenes
from test_ccd_scene_gpu import cloud

pytestmark = pytest.mark.gpu


def make_inputs(n, dtype, seed, spread):
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = scenes.random_poses(rng, n, spread, dtype)
    p2 = scenes.random_poses(rng, n, 0.1, dtype)
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    disp = np.concatenate([ax, rng.uniform(0.05, 1.0, size=(n, 1))], axis=1).astype(dtype)
    return p1, p2, disp


@pytest.fixture(scope="module")
def geometries(fclb, ref_oracle):
    """Two heightmaps and two octrees of different resolution on both sides, per scalar type."""
    clouds = {"hm1": (cloud(7), 0.025), "hm2": (cloud(11, 3000) * np.array([0.5, 0.5, 1.0]), 0.0125)}
    pts = cloud(9, 6000)
    pts[:, 2] -= 0.2
    clouds["oc1"] = (pts, 0.025)
    clouds["oc2"] = (cloud(13, 3000) * 0.6, 0.0125)
    out = {}
    for dtype in (np.float32, np.float64):
        st = fclb.F32 if dtype == np.float32 else fclb.F64
        g = {}
        for name, (p, res) in clouds.items():
            if name.startswith("hm"):
                ours = fclb.heightmap_upload(fclb.heightmap_build_host(p, res, 32, st), res)
                g[name] = (fclb.SCENE_HEIGHTMAP, ours, 1, ref_oracle.heightmap_create(p, res, 32))
            else:
                ours = fclb.octree_upload(*fclb.octree_build_host(p, res, 32, st))
                g[name] = (fclb.SCENE_OCTREE, ours, 2, ref_oracle.octree_create(p, res, 32))
        out[dtype] = g
    return out


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("pair", [("hm1", "hm2"), ("hm1", "oc2"), ("oc1", "hm2"), ("oc1", "oc2"), ("oc2", "oc1")])
def test_scene_pair_ccd(fclb, ref_oracle, geometries, dtype, pair):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    k1, h1, rk1, r1 = geometries[dtype][pair[0]]
    k2, h2, rk2, r2 = geometries[dtype][pair[1]]
    n, keep = 300, 96
    p1, p2, disp = make_inputs(n, dtype, 51, 1.2)
    for request_type, max_contacts in ((0, 1), (0, 7), (0, 10**9), (2, 10**9), (1, 40)):
        c, ids, toc, box = fclb.translational_ccd_scene_pair_batch_host(k1, h1, k2, h2, p1, p2, disp, st, request_type=request_type,
                                                                        max_contacts=max_contacts, max_keep=keep)
        ec, eids, etoc, ebox = ref_oracle.translational_ccd_scene_pair_batch(rk1, r1, rk2, r2, p1, p2, disp,
                                                                             request_type=request_type, max_contacts=max_contacts,
                                                                             keep=keep, threads=8)
        bad = np.nonzero(c != ec)[0]
        listed = [{"query": int(q), "ours": int(c[q]), "reference": int(ec[q])} for q in bad[:20]]
        same = {"ids": bool(np.array_equal(ids, eids)), "toc": bool(np.array_equal(toc, etoc)), "boxes": bool(np.array_equal(box, ebox))}
        parity_util.record("test_ccd_scene_pair", f"{pair[0]} (moving) vs {pair[1]}, request {request_type}, max_contacts {max_contacts}",
                           dtype, n, "contact counts, (code 1, code 2) in the reference's order, toc intervals, both boxes", listed,
                           {"queries_with_contacts": int((ec > 0).sum()), "contacts": int(ec.sum()), "count_mismatches": int(bad.size),
                            **{k + "_identical": v for k, v in same.items()}})
        assert bad.size == 0, listed[:5]
        assert same["ids"], np.argwhere(ids != eids)[:5]
        assert same["toc"], (np.argwhere(toc != etoc)[:5], np.abs(toc - etoc).max())
        assert same["boxes"], np.argwhere(box != ebox)[:5]
        assert int((ec > 0).sum()) > 50


def test_scene_ccd_edge_cases(fclb, ref_oracle, geometries):
    """Empty batches, zero-length displacements, geometries that start in contact and refused inputs of the two entry points
    for continuous collision between scene geometries (the same cases the reference is run on)."""
    dtype, st = np.float64, fclb.F64
    k1, h1, rk1, r1 = geometries[dtype]["hm1"]
    k2, h2, rk2, r2 = geometries[dtype]["oc2"]
    empty = np.zeros((0, 12), dtype)
    c, ids, toc, box = fclb.translational_ccd_scene_pair_batch_host(k1, h1, k2, h2, empty, empty, np.zeros((0, 4), dtype), st)
    assert c.shape == (0,) and ids.shape == (0, 8, 2)
    ident = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype), (3, 1))
    p1 = ident.copy()
    p1[1, 9:] = (0.05, 0.0, 2.0)   # far above: nothing, whatever the sweep
    p1[2, 9:] = (0.0, 0.0, 0.6)    # above the octree's points, swept down through them
    disp = np.array([[0, 0, -1, 0.0], [0, 0, -1, 0.0], [0, 0, -1, 1.5]], dtype)
    for request_type in (0, 2):
        c, ids, toc, box = fclb.translational_ccd_scene_pair_batch_host(k1, h1, k2, h2, p1, ident, disp, st, request_type=request_type,
                                                                        max_contacts=10**6, max_keep=32)
        ec, eids, etoc, ebox = ref_oracle.translational_ccd_scene_pair_batch(rk1, r1, rk2, r2, p1, ident, disp,
                                                                             request_type=request_type, max_contacts=10**6, keep=32)
        assert np.array_equal(c, ec) and np.array_equal(ids, eids) and np.array_equal(toc, etoc) and np.array_equal(box, ebox)
        # zero displacement: the maps overlap where they stand (the whole interval), the lifted one touches nothing
        assert c[0] > 0 and c[1] == 0 and c[2] > 0
        assert np.all(toc[0, : min(int(c[0]), 32), 0] == 0) and np.all(toc[0, : min(int(c[0]), 32), 1] == 1)
    # max_keep = 0: counts only
    c0, _, _, _ = fclb.translational_ccd_scene_pair_batch_host(k1, h1, k2, h2, p1, ident, disp, st, max_contacts=5, max_keep=0)
    assert np.array_equal(c0, np.minimum(ec, 5))
    # refused: a mesh handle where a heightmap / octree is expected, an unknown kind, an unknown request type
    v, t = scenes.noisy_uv_sphere(n_lat=5, n_lon=8, radius=0.2, noise=0.0)
    bvh = fclb.bvh_build(v, t, st)
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_scene_pair_batch_host(k1, bvh, k2, h2, p1, ident, disp, st)
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_scene_pair_batch_host(0, h1, k2, h2, p1, ident, disp, st)
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_scene_pair_batch_host(k1, h1, k2, h2, p1, ident, disp, st, request_type=3)
    # scene vs mesh: empty batch, a mesh of the other scalar type, a scene handle of the wrong kind
    c, ids, toc, box = fclb.translational_ccd_scene_mesh_batch_host(k1, h1, bvh, empty, empty, np.zeros((0, 4), dtype), st)
    assert c.shape == (0,)
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_scene_mesh_batch_host(k1, h1, bvh, p1.astype(np.float32), ident.astype(np.float32),
                                                     disp.astype(np.float32), fclb.F32)
    with pytest.raises(fclb.FclbError):
        fclb.translational_ccd_scene_mesh_batch_host(k2, h1, bvh, p1, ident, disp, st)
    fclb.bvh_release(bvh)
