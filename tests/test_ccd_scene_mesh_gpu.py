"""Translational continuous collision, heightmap / octree vs mesh: fclb_translational_ccd_scene_mesh_batch_host against the
reference's fcl::translational_ccd(scene, BVHModel<OBB<S>>) and its mesh-first entry (heightmap_ccd_solver-inl.h:168-366,
octree2_ccd_solver-inl.h:225-470) on the same seeded inputs.
Bar: contact counts, (pixel / node code, triangle id) IN THE REFERENCE'S ORDER, toc intervals and scene boxes bit-identical."""
import numpy as np
import pytest

import parity_util
import scenes
from test_ccd_scene_gpu import cloud

pytestmark = pytest.mark.gpu


def make_inputs(n, dtype, seed, spread):
    rng = np.random.Generator(np.random.PCG64(seed))
    pm = scenes.random_poses(rng, n, spread, dtype)
    pg = scenes.random_poses(rng, n, 0.1, dtype)
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    disp = np.concatenate([ax, rng.uniform(0.05, 1.0, size=(n, 1))], axis=1).astype(dtype)
    return pm, pg, disp


def run_case(fclb, ref_oracle, dtype, kind, scene, ref_kind, ref_scene, label):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    v, t = scenes.noisy_uv_sphere(n_lat=7, n_lon=10, radius=0.15, noise=0.02)
    bvh = fclb.bvh_build(v, t, st)
    oid = ref_oracle.bvh_obb_create(v, t)
    n, keep = 1200, 64
    pm, pg, disp = make_inputs(n, dtype, 41, 0.7)
    for request_type in (0, 1, 2):
        for max_contacts, mesh_moves in ((1, False), (5, False), (100000, False), (2, True), (100000, True)):
            c, ids, toc, box = fclb.translational_ccd_scene_mesh_batch_host(kind, scene, bvh, pg, pm, disp, st,
                                                                            request_type=request_type, max_contacts=max_contacts,
                                                                            mesh_moves=mesh_moves, max_keep=keep)
            ec, eids, etoc, ebox = ref_oracle.translational_ccd_scene_mesh_batch(ref_kind, ref_scene, oid, pg, pm, disp,
                                                                                 request_type=request_type,
                                                                                 max_contacts=max_contacts, mesh_moves=mesh_moves,
                                                                                 keep=keep, threads=8)
            # slots beyond counts[q] are zeros on our side, -1 / 0 on the harness side: compare the filled ones
            filled = np.arange(keep)[None, :] < np.minimum(ec, keep)[:, None]
            bad = np.nonzero(c != ec)[0]
            listed = [{"query": int(q), "ours": int(c[q]), "reference": int(ec[q])} for q in bad[:20]]
            same = {"ids": bool(np.array_equal(ids[filled], eids[filled])), "toc": bool(np.array_equal(toc[filled], etoc[filled])),
                    "boxes": bool(np.array_equal(box[filled], ebox[filled]))}
            parity_util.record("test_ccd_scene_mesh", f"{len(t)}-triangle mesh vs {label}, request {request_type}, max_contacts "
                               f"{max_contacts}, {'mesh' if mesh_moves else 'scene'} moves", dtype, n,
                               "contact counts, (code, triangle) in the reference's order, toc intervals, boxes", listed,
                               {"queries_with_contacts": int((ec > 0).sum()), "contacts": int(ec.sum()),
                                "count_mismatches": int(bad.size), **{k + "_identical": v for k, v in same.items()}})
            assert bad.size == 0, listed[:5]
            assert same["ids"], np.argwhere(ids != eids)[:5]
            assert same["toc"], (np.argwhere(toc != etoc)[:5], np.abs(toc - etoc).max())
            assert same["boxes"], np.argwhere(box != ebox)[:5]
        if request_type == 1:
            # kBoxApproximate pre-checks the leaf pair with the shapes' local AABBs, and the reference never computes the
            # TriangleP's (it stays the empty AABB it is constructed with), so the reference reports no contact at all
            # with this request; the device reproduces that arithmetic and must agree
            assert int(ec.sum()) == 0
        else:
            assert int((ec > 0).sum()) > 100
    fclb.bvh_release(bvh)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_heightmap_mesh_ccd(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = cloud(7)
    heights = fclb.heightmap_build_host(pts, 0.025, 32, st)
    hm = fclb.heightmap_upload(heights, 0.025)
    rhm = ref_oracle.heightmap_create(pts, 0.025, 32)
    run_case(fclb, ref_oracle, dtype, fclb.SCENE_HEIGHTMAP, hm, 1, rhm, "64 x 64 heightmap")
    fclb.heightmap_release(hm)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_octree_mesh_ccd(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = cloud(9, 6000)
    pts[:, 2] -= 0.2
    ch, full, leaf, root, layers = fclb.octree_build_host(pts, 0.025, 32, st)
    oc = fclb.octree_upload(ch, full, leaf, root, layers)
    roc = ref_oracle.octree_create(pts, 0.025, 32)
    run_case(fclb, ref_oracle, dtype, fclb.SCENE_OCTREE, oc, 2, roc, "octree of 6000 points")
    fclb.octree_release(oc)
