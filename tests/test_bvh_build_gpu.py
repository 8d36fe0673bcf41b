"""BVHModel<OBBRSS> built on the device (fclb_bvh_build_device) against the reference's own builder
(BVHModel::beginModel / addSubModel / endModel, BVH_model-inl.h:402-570): node OBBs, child links (hence node numbering
and the primitive order) identical, bit for bit."""
import time

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_build_matches_reference(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    for name, (v, t) in (("sphere", scenes.noisy_uv_sphere(n_lat=21, n_lon=40)), ("torus", scenes.noisy_torus()),
                         ("two triangles", (np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float64),
                                            np.array([[0, 1, 2], [0, 2, 3]], np.int32)))):
        mid = ref_oracle.bvh_create(v, t)
        e_obb, e_child, e_tri = ref_oracle.bvh_export(mid, dtype)
        h = fclb.bvh_build_device(v, t, st)
        obb, child, tri = fclb.bvh_export(h)
        assert np.array_equal(child, e_child), np.nonzero(child != e_child)[0][:10]
        same = (obb == e_obb).all(axis=1)
        print(f"[device build {name} {np.dtype(dtype).name}] nodes {len(obb)}, bit-identical OBBs {int(same.sum())}, "
              f"build launches {fclb.last_kernel_ms():.2f} ms")
        assert same.all(), np.nonzero(~same)[0][:10]
        assert np.array_equal(tri, e_tri)
        fclb.bvh_release(h)


def test_device_build_large_mesh(fclb):
    """the C4 scene mesh (200k triangles): the device build equals the host mirror's tree; reported times"""
    v, t = scenes.c4_scene_mesh()
    t0 = time.perf_counter()
    obb_h, child_h, tri_h = fclb.bvh_build_host(v, t, fclb.F32)
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = fclb.bvh_build_device(v, t, fclb.F32)
    t_dev = time.perf_counter() - t0
    ms = fclb.last_kernel_ms()
    obb, child, tri = fclb.bvh_export(h)
    print(f"[device build] {len(t)} triangles / {len(obb)} nodes: device launches {ms:.1f} ms (call {t_dev * 1e3:.0f} ms incl. "
          f"upload), host mirror {t_host * 1e3:.0f} ms")
    assert np.array_equal(child, child_h)
    assert np.array_equal(obb, obb_h)
    assert np.array_equal(tri, tri_h)
    fclb.bvh_release(h)
