"""Host mirror of octree2::Octree<S>::rebuildTree (fclb_octree_build_host; reference
geometry/octree2/octree-inl.h:15-142, octree_construction-inl.h:10-74,111-205) against the node arrays the
reference builds from the same point stream: inner children (the node NUMBERING, which the octree kernels report
in contact ids), fully-occupied flags, leaf bitmasks, root box and layer count, float and double.
Host-only: no GPU, no compute call."""
import numpy as np
import pytest

from test_octree_gpu import octree_points


def test_known_answers():
    import fclb200 as fclb

    # one point in voxel (2, 2, 2) of a 4^3 grid: root child 7 -> leaf node 0, bit 0
    ch, full, leaf, root, layers = fclb.octree_build_host(np.array([[0.5, 0.5, 0.5]]), 1.0, 2, fclb.F64)
    assert layers == 3 and ch.shape == (1, 8) and full.tolist() == [0] and leaf.tolist() == [1]
    assert ch[0].tolist() == [0xFFFFFFFF] * 7 + [0]
    assert root.tolist() == [-2, -2, -2, 2, 2, 2]
    # every voxel of the grid: all 8 leaf nodes full, hence the root; numbering follows first arrival
    g = np.arange(-2, 2) + 0.5
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    ch, full, leaf, root, layers = fclb.octree_build_host(pts, 1.0, 2, fclb.F32)
    assert full.tolist() == [1] and leaf.tolist() == [255] * 8
    assert ch[0].tolist() == [0, 4, 2, 6, 1, 5, 3, 7]  # z varies fastest in the stream: child 4 arrives second
    # out-of-range points (incl. the upper face, which floors to full_shape) are dropped; an empty cloud is a bare root
    ch, full, leaf, root, layers = fclb.octree_build_host(np.array([[2.0, 0, 0], [0, -2.1, 0], [9, 9, 9]]), 1.0, 2, fclb.F64)
    assert ch.shape == (1, 8) and (ch == 0xFFFFFFFF).all() and len(leaf) == 0 and full.tolist() == [0]
    ch, full, leaf, root, layers = fclb.octree_build_host(np.zeros((0, 3)), 0.5, 8, fclb.F32)
    assert layers == 5 and ch.shape == (1, 8) and len(leaf) == 0 and root.tolist() == [-4, -4, -4, 4, 4, 4]
    with pytest.raises(fclb.FclbError):
        fclb.octree_build_host(pts, 1.0, 3, fclb.F32)  # half shape must be a power of two >= 2
    with pytest.raises(fclb.FclbError):
        fclb.octree_build_host(pts, 1.0, 1, fclb.F32)


@pytest.mark.parametrize("half_shape,res,seed", [(64, 0.01, 7), (16, 0.04, 8), (256, 0.0025, 9), (2, 0.4, 10)])
def test_tree_matches_reference(ref_oracle, half_shape, res, seed):
    import fclb200 as fclb

    pts = octree_points(seed)
    if half_shape == 256:  # points on voxel faces and far outside the grid
        rng = np.random.Generator(np.random.PCG64(seed))
        lattice = rng.integers(-300, 300, size=(5000, 3)) * res
        pts = np.ascontiguousarray(np.concatenate([pts[::7], lattice]))
    oid = ref_oracle.octree_create(pts, res, half_shape)
    for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
        r_ch, r_full, r_leaf, r_root, r_layers = ref_oracle.octree_export(oid, dt)
        ch, full, leaf, root, layers = fclb.octree_build_host(pts, res, half_shape, st)
        assert layers == r_layers
        assert np.array_equal(root, r_root)
        assert ch.shape == r_ch.shape and np.array_equal(ch, r_ch)
        assert np.array_equal(full, r_full)
        assert np.array_equal(leaf, r_leaf)
        if half_shape == 64:
            assert full.any() and (leaf == 255).any() and (leaf != 255).any()
