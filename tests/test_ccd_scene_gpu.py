"""Translational continuous collision, shape vs heightmap / octree: fclb_translational_ccd_scene_batch_host against the
reference's fcl::translational_ccd (heightmap_ccd_solver-inl.h, octree2_ccd_solver-inl.h) on the same seeded inputs.
Bar: contact counts, pixel / node codes IN THE REFERENCE'S ORDER, toc intervals and boxes bit-identical."""
import numpy as np
import pytest

import parity_util
import scenes

pytestmark = pytest.mark.gpu


def make_inputs(n, dtype, seed, spread):
    rng = np.random.Generator(np.random.PCG64(seed))
    ps = scenes.random_poses(rng, n, spread, dtype)
    pg = scenes.random_poses(rng, n, 0.1, dtype)
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    disp = np.concatenate([ax, rng.uniform(0.05, 1.0, size=(n, 1))], axis=1).astype(dtype)
    return ps, pg, disp


def cloud(seed, n=4000):
    rng = np.random.Generator(np.random.PCG64(seed))
    xy = rng.uniform(-0.75, 0.75, size=(n, 2))
    z = 0.25 + 0.15 * np.sin(4 * xy[:, 0]) * np.cos(3 * xy[:, 1]) + 0.02 * rng.normal(size=n)
    return np.concatenate([xy, np.abs(z)[:, None]], axis=1)


def run_case(fclb, ref_oracle, dtype, kind, scene, ref_kind, ref_scene, label):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    hull = scenes.ellipsoid_mesh(0.1, 0.15, 0.12)
    shapes = [(scenes.BOX, 0, (0.15, 0.1, 0.12)), (scenes.SPHERE, 0, (0.1,)), (scenes.CAPSULE, 0, (0.05, 0.15)),
              (scenes.CYLINDER, 0, (0.07, 0.15)), (scenes.CONE, 0, (0.1, 0.17)), (scenes.ELLIPSOID, 0, (0.1, 0.05, 0.15)),
              (scenes.CONVEX, fclb.convex_upload(*hull), ())]
    rshapes = shapes[:6] + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())]
    table = fclb.shapes_upload(shapes)
    n, keep = 3500, 48
    ids = (np.arange(n) % len(shapes)).astype(np.uint32)
    ps, pg, disp = make_inputs(n, dtype, 31, 0.7)
    for request_type in (0, 1, 2):
        for max_contacts, scene_moves in ((1, False), (5, False), (100000, False), (2, True), (100000, True)):
            c, code, toc, box = fclb.translational_ccd_scene_batch_host(kind, scene, table, ids, ps, pg, disp, st,
                                                                        request_type=request_type, max_contacts=max_contacts,
                                                                        scene_moves=scene_moves, max_keep=keep)
            ec, ecode, etoc, ebox = ref_oracle.translational_ccd_scene_batch(ref_kind, ref_scene, rshapes, ids, ps, pg, disp,
                                                                             request_type=request_type, max_contacts=max_contacts,
                                                                             scene_moves=scene_moves, keep=keep, threads=8)
            bad = np.nonzero(c != ec)[0]
            listed = [{"query": int(q), "ours": int(c[q]), "reference": int(ec[q])} for q in bad[:20]]
            same = {"codes": bool(np.array_equal(code, ecode)), "toc": bool(np.array_equal(toc, etoc)),
                    "boxes": bool(np.array_equal(box, ebox))}
            parity_util.record("test_ccd_scene", f"7 shape kinds vs {label}, request {request_type}, max_contacts {max_contacts}, "
                               f"{'scene' if scene_moves else 'shape'} moves", dtype, n,
                               "contact counts, codes in the reference's order, toc intervals, boxes", listed,
                               {"queries_with_contacts": int((ec > 0).sum()), "contacts": int(ec.sum()),
                                "count_mismatches": int(bad.size), **{k + "_identical": v for k, v in same.items()}})
            assert bad.size == 0, listed[:5]
            assert same["codes"], np.argwhere(code != ecode)[:5]
            assert same["toc"], (np.argwhere(toc != etoc)[:5], np.abs(toc - etoc).max())
            assert same["boxes"], np.argwhere(box != ebox)[:5]
        assert int((ec > 0).sum()) > 300
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_shape_heightmap_ccd(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = cloud(7)
    heights = fclb.heightmap_build_host(pts, 0.025, 32, st)
    hm = fclb.heightmap_upload(heights, 0.025)
    rhm = ref_oracle.heightmap_create(pts, 0.025, 32)
    run_case(fclb, ref_oracle, dtype, fclb.SCENE_HEIGHTMAP, hm, 1, rhm, "64 x 64 heightmap")
    fclb.heightmap_release(hm)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_shape_octree_ccd(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    pts = cloud(9, 6000)
    pts[:, 2] -= 0.2
    ch, full, leaf, root, layers = fclb.octree_build_host(pts, 0.025, 32, st)
    oc = fclb.octree_upload(ch, full, leaf, root, layers)
    roc = ref_oracle.octree_create(pts, 0.025, 32)
    run_case(fclb, ref_oracle, dtype, fclb.SCENE_OCTREE, oc, 2, roc, "octree of 6000 points")
    fclb.octree_release(oc)
