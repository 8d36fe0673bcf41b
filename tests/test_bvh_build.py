"""The host BVH builder (fclb_bvh_build_host: the mirror of BVHModel<OBBRSS>::endModel,
reference geometry/bvh/BVH_model-inl.h:402-570) against the tree the reference's own
builder produces for the same triangle soup (exported from oracle/_ref): node count,
child links / primitive ids and every OBB (axis, To, extent) must be bit-identical.
Host-only: no GPU, no compute call."""
import numpy as np
import pytest

import scenes


@pytest.mark.parametrize("mesh", ["sphere10k", "torus10k", "sphere_small", "two_triangles"])
def test_builder_matches_reference_tree(ref_oracle, mesh):
    import fclb200 as fclb

    if mesh == "sphere10k":
        v, t = scenes.noisy_uv_sphere()
    elif mesh == "torus10k":
        v, t = scenes.noisy_torus()
    elif mesh == "sphere_small":
        v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24)
    else:
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], np.float64)
        t = np.array([[0, 1, 2], [1, 3, 2]], np.int32)
    mid = ref_oracle.bvh_create(v, t)
    for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
        r_obb, r_child, r_tri = ref_oracle.bvh_export(mid, dt)
        obb, child, tri = fclb.bvh_build_host(v, t, st)
        assert len(child) == len(r_child) == 2 * len(t) - 1
        assert np.array_equal(child, r_child)
        assert np.array_equal(obb, np.asarray(r_obb).reshape(-1, 15))
        assert np.array_equal(tri, np.asarray(r_tri).reshape(-1, 9))


def test_builder_rejects_bad_input():
    import fclb200 as fclb

    v = np.zeros((3, 3))
    with pytest.raises(fclb.FclbError):
        fclb.bvh_build_host(v, np.array([[0, 1, 5]], np.int32), fclb.F32)
