"""Device refit of BVHModel<OBBRSS> (fclb_bvh_refit_*) against the reference's beginReplaceModel / replaceSubModel /
endReplaceModel(refit = true, bottomup = false) (geometry/bvh/BVH_model-inl.h:318-375, refitTreeTopDown :624-637):
the node OBBs of the refitted tree must equal the reference's bit for bit, and queries on the refitted tree must give
the reference's answers."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def deform(v, seed, amount):
    rng = np.random.Generator(np.random.PCG64(seed))
    w = v * (1.0 + amount * np.sin(3.0 * v[:, [1, 2, 0]])) + rng.normal(scale=0.2 * amount, size=v.shape)
    return np.ascontiguousarray(w)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_refit_matches_reference(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    for name, (v, t) in (("sphere", scenes.noisy_uv_sphere(n_lat=21, n_lon=40)), ("torus", scenes.noisy_torus())):
        mid = ref_oracle.bvh_create(v, t)
        h = fclb.bvh_build(v, t, st)
        for step, amount in enumerate((0.05, 0.2, 0.0)):
            v2 = deform(v, 100 + step, amount)
            ref_oracle.bvh_refit(mid, v2, bottomup=False)
            e_obb, e_child, e_tri = ref_oracle.bvh_export(mid, dtype)
            tri_verts = np.ascontiguousarray(v2.astype(dtype)[t].reshape(len(t), 9))
            assert np.array_equal(tri_verts, e_tri)
            fclb.bvh_refit_host(h, tri_verts)
            obb, child, tri = fclb.bvh_export(h)
            assert np.array_equal(child, e_child)
            same = (obb == e_obb).all(axis=1)
            print(f"[refit {name} {np.dtype(dtype).name} step {step}] nodes {len(obb)}, bit-identical OBBs {int(same.sum())}")
            assert same.all(), np.nonzero(~same)[0][:10]
            assert np.array_equal(tri, e_tri)
        # queries on the refitted tree: mesh-shape counts against the reference's refitted model
        n = 3000
        rng = np.random.Generator(np.random.PCG64(9))
        shapes = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.SPHERE, 0, (0.2,))]
        table = fclb.shapes_upload(shapes)
        pm, ps = scenes.random_poses(rng, n, 0.3, dtype), scenes.random_poses(rng, n, 1.0, dtype)
        ids = (np.arange(n) % 2).astype(np.uint32)
        req = fclb.make_request(max_contacts=2**31 - 1)
        c, _ = fclb.bvh_shape_collide_batch_host(h, table, ids, pm, ps, st, req)
        e, _ = ref_oracle.mesh_shape_collide_batch(mid, shapes, ids, pm, ps, threads=8, max_contacts=2**31 - 1)
        assert np.array_equal(c, e) and e.any()
        fclb.release(table)
        fclb.bvh_release(h)


def test_refit_large_mesh_timing(fclb):
    """the C4 scene mesh (200k triangles, 399,999 nodes): refit on the device, reported time"""
    v, t = scenes.c4_scene_mesh()
    h = fclb.bvh_build(v, t, fclb.F32)
    tri_verts = np.ascontiguousarray(deform(v, 7, 0.02).astype(np.float32)[t].reshape(len(t), 9))
    fclb.bvh_refit_host(h, tri_verts)
    fclb.bvh_refit_host(h, tri_verts)
    ms = fclb.last_kernel_ms()
    obb, child, tri = fclb.bvh_export(h)
    assert np.isfinite(obb).all() and np.array_equal(tri, tri_verts)
    print(f"[refit] {len(t)} triangles / {len(obb)} nodes refitted on the device in {ms:.2f} ms (kernel)")
    fclb.bvh_release(h)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_refit_bottomup_matches_reference(fclb, ref_oracle, dtype):
    """endReplaceModel(refit = true, bottomup = true) -- the reference's default arguments (refitTreeBottomUp,
    BVH_model-inl.h:580-617): leaf boxes from the 3-point fit, inner boxes merged from their children
    (OBB::operator+, math/bv/OBB-inl.h:116-293).  Node-for-node identical OBBs, and the reference's answers on the tree."""
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    for name, (v, t) in (("sphere", scenes.noisy_uv_sphere(n_lat=21, n_lon=40)), ("torus", scenes.noisy_torus())):
        mid = ref_oracle.bvh_create(v, t)
        h = fclb.bvh_build(v, t, st)
        for step, amount in enumerate((0.05, 0.2)):
            v2 = deform(v, 200 + step, amount)
            ref_oracle.bvh_refit(mid, v2, bottomup=True)
            e_obb, e_child, e_tri = ref_oracle.bvh_export(mid, dtype)
            tri_verts = np.ascontiguousarray(v2.astype(dtype)[t].reshape(len(t), 9))
            fclb.bvh_refit_host(h, tri_verts, bottomup=True)
            obb, child, tri = fclb.bvh_export(h)
            assert np.array_equal(child, e_child)
            same = (obb == e_obb).all(axis=1)
            leaves = child < 0
            print(f"[bottom-up refit {name} {np.dtype(dtype).name} step {step}] nodes {len(obb)}, bit-identical OBBs "
                  f"{int(same.sum())} (leaves {int(same[leaves].sum())} of {int(leaves.sum())}), max diff "
                  f"{np.abs(obb - e_obb).max():.3g}")
            assert same.all(), (np.nonzero(~same)[0][:10], np.abs(obb - e_obb).max())
        n = 3000
        rng = np.random.Generator(np.random.PCG64(19))
        shapes = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.SPHERE, 0, (0.2,))]
        table = fclb.shapes_upload(shapes)
        pm, ps = scenes.random_poses(rng, n, 0.3, dtype), scenes.random_poses(rng, n, 1.0, dtype)
        ids = (np.arange(n) % 2).astype(np.uint32)
        req = fclb.make_request(max_contacts=2**31 - 1)
        c, _ = fclb.bvh_shape_collide_batch_host(h, table, ids, pm, ps, st, req)
        e, _ = ref_oracle.mesh_shape_collide_batch(mid, shapes, ids, pm, ps, threads=8, max_contacts=2**31 - 1)
        assert np.array_equal(c, e) and e.any()
        fclb.release(table)
        fclb.bvh_release(h)
