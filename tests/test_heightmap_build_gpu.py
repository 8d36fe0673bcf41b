"""Device construction of a LayeredHeightMap from a point cloud (fclb_heightmap_build_dev /
fclb_heightmap_build_points_host) against the reference's FlatHeightMap::updateHeightsByPointGenerationFunctor
(flat_heightmap-inl.h:249-272) and layer pyramid (layered_heightmap-inl.h:77-101): every layer bit-identical,
for float and double maps, incl. negative / out-of-range points, an empty cloud and a 1M-point cloud."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def pyramid(bottom):
    layers = [bottom]
    while layers[-1].shape[0] > 2 and layers[-1].shape[1] > 2:
        d = layers[-1]
        layers.append(d.reshape(d.shape[0] // 2, 2, d.shape[1] // 2, 2).max(axis=(1, 3)))
    return layers


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_heightmap_device_build(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    rng = np.random.Generator(np.random.PCG64(42))
    for half, res, n_pts in ((64, 0.01, 40_000), (512, 0.004, 1_000_000)):
        pts = scenes.terrain_points(n_pts, half * res)
        extra = rng.uniform(-1.3 * half * res, 1.3 * half * res, size=(2000, 3))  # out of range / negative z
        pts = np.ascontiguousarray(np.concatenate([pts, extra]))
        hid = ref_oracle.heightmap_create(pts, res, half)
        ref_h, upper = ref_oracle.heightmap_export(hid, dtype, half)
        hm = fclb.heightmap_build_points_host(pts, res, half, st)
        info = fclb.heightmap_info(hm)
        assert (info["full_x"], info["full_y"]) == (2 * half, 2 * half) and info["upper_bound_mm"] == int(ref_h.max())
        expect = pyramid(np.asarray(ref_h))
        assert info["n_layers"] == len(expect)
        for k, e in enumerate(expect):
            assert np.array_equal(fclb.heightmap_export(hm, k), e), (half, k)
        # the uploaded map (host pyramid) and the built one agree layer by layer
        up = fclb.heightmap_upload(ref_h, res, upper)
        for k in range(len(expect)):
            assert np.array_equal(fclb.heightmap_export(up, k), fclb.heightmap_export(hm, k))
        # and answer queries identically
        table = fclb.shapes_upload([(scenes.BOX, 0, (0.12, 0.08, 0.1))])
        p_hm, p_sh = scenes.heightmap_query_poses(500, dtype, half * res, -0.05, 0.4, seed=9)
        ids = np.zeros(500, np.uint32)
        req = fclb.make_request(max_contacts=2**31 - 1)
        c1, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, p_hm, p_sh, st, req)
        c2, _ = fclb.heightmap_shape_collide_batch_host(up, table, ids, p_hm, p_sh, st, req)
        assert np.array_equal(c1, c2) and c1.any()
        print(f"[heightmap build {np.dtype(dtype).name}] {len(pts)} points -> {2 * half}^2 map, {len(expect)} layers identical, "
              f"non-empty pixels {int((expect[0] > 0).sum())}")
        fclb.release(table)
        fclb.heightmap_release(hm)
        fclb.heightmap_release(up)
    empty = fclb.heightmap_build_points_host(np.zeros((0, 3)), 0.01, 8, st)
    assert not fclb.heightmap_export(empty, 0).any() and fclb.heightmap_info(empty)["upper_bound_mm"] == 0
    fclb.heightmap_release(empty)
    with pytest.raises(fclb.FclbError):
        fclb.heightmap_build_points_host(np.zeros((4, 3)), 0.01, 12, st)  # half shape not a power of two
