"""Device build of octree2::Octree<S> (fclb_octree_build_points_host / _dev) against the reference's rebuildTree
(geometry/octree2/octree_construction-inl.h:10-74,111-205) on the same point stream: inner children (the node NUMBERING the
octree kernels report in contact ids), fully-occupied flags, leaf masks, root box, layer count -- every array identical."""
import time

import numpy as np
import pytest

import scenes
from test_octree_gpu import octree_points

pytestmark = pytest.mark.gpu


def clouds():
    rng = np.random.Generator(np.random.PCG64(41))
    yield "terrain+noise", np.concatenate([scenes.terrain_points(30_000, 0.6), rng.uniform(-0.7, 0.7, size=(5000, 3))]), 0.01, 64
    yield "octree_points", octree_points(), 0.01, 64
    dense = np.stack(np.meshgrid(*[np.arange(-8, 8) + 0.5] * 3, indexing="ij"), -1).reshape(-1, 3) * 0.05
    yield "dense block (fully occupied nodes)", np.concatenate([dense, dense[::-1], rng.uniform(-0.9, 0.9, size=(300, 3))]), 0.05, 16
    yield "tiny grid", rng.uniform(-0.25, 0.25, size=(200, 3)), 0.1, 2
    yield "points on voxel faces and outside", np.concatenate([np.round(rng.uniform(-1.5, 1.5, size=(4000, 3)) * 20) / 20,
                                                                np.array([[np.nan, 0, 0], [1e30, 0, 0], [-1e30, 1, 1]])]), 0.05, 16
    yield "empty", np.zeros((0, 3)), 0.01, 64


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_octree_build_matches_reference(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    for name, pts, res, half in clouds():
        h = fclb.octree_build_points_host(pts, res, half, st)
        ch, full, leaf, root, layers = fclb.octree_export(h)
        finite = np.isfinite(pts).all(axis=1) & (np.abs(pts) < 1e20).all(axis=1)  # (the reference is fed what it can convert)
        e_ch, e_full, e_leaf, e_root, e_layers = fclb.octree_build_host(pts[finite], res, half, st)
        assert layers == e_layers and np.array_equal(root, e_root), name
        assert ch.shape == e_ch.shape and leaf.shape == e_leaf.shape, (name, ch.shape, e_ch.shape, leaf.shape, e_leaf.shape)
        assert np.array_equal(ch, e_ch), (name, np.nonzero((ch != e_ch).any(axis=1))[0][:10])
        assert np.array_equal(full, e_full) and np.array_equal(leaf, e_leaf), name
        if len(pts[finite]):
            oid = ref_oracle.octree_create(pts[finite], res, half)
            r_ch, r_full, r_leaf, r_root, r_layers = ref_oracle.octree_export(oid, dtype)
            assert np.array_equal(ch, r_ch) and np.array_equal(full, r_full) and np.array_equal(leaf, r_leaf) and layers == r_layers, name
        print(f"[device octree {np.dtype(dtype).name}] {name}: {len(pts)} points -> {len(full)} inner / {len(leaf)} leaf nodes, "
              f"{int(full.sum())} fully occupied: identical to the reference")
        fclb.octree_release(h)


def test_device_octree_build_queries_and_timing(fclb, ref_oracle):
    """a perception-cycle sized cloud: build on the device, query it, compare with the host mirror's tree"""
    st, dtype = fclb.F32, np.float32
    rng = np.random.Generator(np.random.PCG64(5))
    pts = np.concatenate([scenes.terrain_points(1_500_000, 2.5, z_max=0.8), rng.uniform(-2.5, 2.5, size=(500_000, 3))])
    t = time.perf_counter()
    h = fclb.octree_build_points_host(pts, 0.01, 256, st)
    dt_dev = time.perf_counter() - t
    ms_kernel = fclb.last_kernel_ms()
    t = time.perf_counter()
    e_ch, e_full, e_leaf, e_root, e_layers = fclb.octree_build_host(pts, 0.01, 256, st)
    dt_host = time.perf_counter() - t
    ch, full, leaf, root, layers = fclb.octree_export(h)
    assert np.array_equal(ch, e_ch) and np.array_equal(full, e_full) and np.array_equal(leaf, e_leaf)
    print(f"[device octree] {len(pts)} points -> {len(full)} inner / {len(leaf)} leaf nodes: device {ms_kernel:.1f} ms of kernels "
          f"({dt_dev * 1e3:.0f} ms with the upload), host mirror {dt_host * 1e3:.0f} ms")
    shapes = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.SPHERE, 0, (0.15,))]
    table = fclb.shapes_upload(shapes)
    n = 2000
    p_oc, p_sh = scenes.heightmap_query_poses(n, dtype, 1.5, -0.1, 0.9, seed=3)
    ids = (np.arange(n) % 2).astype(np.uint32)
    req = fclb.make_request(max_contacts=2**31 - 1)
    c, _ = fclb.octree_shape_collide_batch_host(h, table, ids, p_oc, p_sh, st, req)
    h2 = fclb.octree_upload(e_ch, e_full, e_leaf, e_root, e_layers)
    c2, _ = fclb.octree_shape_collide_batch_host(h2, table, ids, p_oc, p_sh, st, req)
    assert np.array_equal(c, c2) and c.any()
    for x in (h, h2):
        fclb.octree_release(x)
    fclb.release(table)
