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


@pytest.mark.parametrize("half_shape,res,seed", [(64, 0.01, 7), (16, 0.04, 8)])
def test_prune_matches_reference(ref_oracle, half_shape, res, seed):
    """fclb_octree_prune_host against Octree2CollisionGeometry::pruneBy(obb, rebuild=False) (pruneOctreeByOBB,
    octree_prune-inl.h:10-103): prune mask, new leaf masks and new fully-occupied flags identical for chains of
    cumulative prunes by random boxes -- incl. axis-aligned boxes whose faces lie ON voxel faces / centres, where
    the float SSE association of OBB<float>::overlap decides -- float and double."""
    import fclb200 as fclb
    import scenes

    pts = octree_points(seed)
    oid = ref_oracle.octree_create(pts, res, half_shape)
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    span = res * half_shape
    n_chain, per_chain = 12, 3
    total_pruned = total_leaf = 0
    for c in range(n_chain):
        ours = {}
        for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
            ch, full, leaf, root, layers = ref_oracle.octree_export(oid, dt)
            ours[dt] = [ch, root, layers, None, full, leaf]
        pid = oid
        for k in range(per_chain):
            if c % 3 == 0:  # lattice-aligned box: centre and extents multiples of half a voxel
                axis = np.eye(3)
                center = rng.integers(-half_shape, half_shape, 3) * res * 0.5
                extent = rng.integers(1, half_shape, 3) * res * 0.5
            else:
                e = rng.uniform(-np.pi, np.pi, 3)
                axis = scenes.euler_to_matrix(e[:1], e[1:2], e[2:3])[0]
                center = rng.uniform(-0.6 * span, 0.6 * span, 3)
                extent = rng.uniform(0.05 * span, 0.5 * span, 3)
            pid = ref_oracle.octree_prune(pid, axis, center, extent)
            for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
                ch, root, layers, pruned, full, leaf = ours[dt]
                pruned, full, leaf = fclb.octree_prune_host(ch, full, leaf, root, layers, axis, center, extent, st, pruned=pruned)
                ours[dt][3:] = [pruned, full, leaf]
                r_ch, r_full, r_leaf, _, _ = ref_oracle.octree_export(pid, dt)
                r_pruned = ref_oracle.octree_export_pruned(pid, dt, len(r_full))
                assert np.array_equal(pruned, r_pruned), (c, k, dt, np.nonzero(pruned != r_pruned)[0][:8])
                assert np.array_equal(leaf, r_leaf), (c, k, dt, np.nonzero(leaf != r_leaf)[0][:8])
                assert np.array_equal(full, r_full), (c, k, dt, np.nonzero(full != r_full)[0][:8])
        total_pruned += int(ours[np.float32][3].sum())
        total_leaf += int((ours[np.float32][5] != ref_oracle.octree_export(oid, np.float32)[2]).sum())
    assert total_pruned > 0 and total_leaf > 0
    print(f"prune chains: {total_pruned} inner nodes pruned, {total_leaf} leaf masks changed over {n_chain} chains")


@pytest.mark.parametrize("half_shape,res,seed", [(64, 0.01, 7), (16, 0.04, 8)])
def test_consolidate_matches_reference(ref_oracle, half_shape, res, seed):
    """fclb_octree_prune_host + fclb_octree_consolidate_host against pruneBy(obb, rebuild_octree=True)
    (Octree::rebuildAccordingToPruneInfo, octree_construction-inl.h:247-369): the renumbered inner / leaf arrays and
    the re-derived fully-occupied flags are identical, incl. a second prune of the consolidated tree and a prune
    that removes everything."""
    import fclb200 as fclb
    import scenes

    oid = ref_oracle.octree_create(octree_points(seed), res, half_shape)
    span = res * half_shape
    rng = np.random.Generator(np.random.PCG64(200 + seed))
    boxes = []
    for _ in range(6):
        e = rng.uniform(-np.pi, np.pi, 3)
        boxes.append((scenes.euler_to_matrix(e[:1], e[1:2], e[2:3])[0], rng.uniform(-0.5 * span, 0.5 * span, 3),
                      rng.uniform(0.1 * span, 0.45 * span, 3)))
    boxes.append((np.eye(3), np.zeros(3), np.full(3, 3 * span)))  # swallows the root
    for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
        dropped = 0
        for bi, (axis, center, extent) in enumerate(boxes):
            rid = ref_oracle.octree_prune_rebuild(oid, axis, center, extent)
            ch, full, leaf, root, layers = ref_oracle.octree_export(oid, dt)
            for again in range(2):  # the consolidated tree pruned once more by the next box
                pr, n_full, n_leaf = fclb.octree_prune_host(ch, full, leaf, root, layers, axis, center, extent, st)
                c_ch, c_full, c_leaf = fclb.octree_consolidate_host(ch, pr, n_leaf, layers)
                r_ch, r_full, r_leaf, r_root, r_layers = ref_oracle.octree_export(rid, dt)
                assert ref_oracle.octree_export_pruned(rid, dt, len(r_full)) is None
                assert r_layers == layers and np.array_equal(r_root, root)
                assert c_ch.shape == r_ch.shape and np.array_equal(c_ch, r_ch), (bi, again, dt)
                assert np.array_equal(c_full, r_full) and np.array_equal(c_leaf, r_leaf), (bi, again, dt)
                dropped += len(full) - len(c_full)
                if again == 0:
                    axis, center, extent = boxes[(bi + 1) % 6]
                    rid = ref_oracle.octree_prune_rebuild(rid, axis, center, extent)
                    ch, full, leaf = c_ch, c_full, c_leaf
            if bi == len(boxes) - 1:
                assert c_ch.shape == (1, 8) and (c_ch == 0xFFFFFFFF).all() and len(c_leaf) == 0
        assert dropped > 0


def test_malformed_links_are_rejected():
    import fclb200 as fclb

    pts = octree_points(7)
    ch, full, leaf, root, layers = fclb.octree_build_host(pts, 0.01, 64, fclb.F64)
    bad = ch.copy()
    bad[0, np.nonzero(ch[0] != 0xFFFFFFFF)[0][0]] = len(full)  # an inner link one past the inner array
    with pytest.raises(fclb.FclbError):
        fclb.octree_prune_host(bad, full, leaf, root, layers, np.eye(3), (0, 0, 0), (0.1, 0.1, 0.1), fclb.F64)
    with pytest.raises(fclb.FclbError):
        fclb.octree_consolidate_host(bad, np.zeros(len(full), np.uint8), leaf, layers)
    # a leaf-layer link past the leaf array (but inside the inner array's range)
    parents = np.nonzero((ch != 0xFFFFFFFF).any(1))[0]
    deep = parents[-1]  # the last inner node was created on the way to a leaf: its children are leaf links
    bad = ch.copy()
    bad[deep, np.nonzero(ch[deep] != 0xFFFFFFFF)[0][0]] = len(leaf)
    with pytest.raises(fclb.FclbError):
        fclb.octree_prune_host(bad, full, leaf, root, layers, np.eye(3), (0, 0, 0), (0.1, 0.1, 0.1), fclb.F32)


def test_cpp_octree_mirror_host_only(tmp_path):
    """include/fcl_b200/fcl.h octree2::Octree<S> (rebuildTree, isPointOccupied) compiled with g++ and run without a GPU."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "mind-fcl_b200")
    exe = str(tmp_path / "test_octree_host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "test_octree_host.cpp"),
                    "-L", lib_dir, "-lfclb200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout


def test_golden():
    """The committed fixture tests/golden/octree.npz (made by tests/golden/make_golden.py from the reference itself):
    build, two cumulative prunes and the consolidated tree, without the reference library at hand."""
    import os

    import fclb200 as fclb

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "octree.npz"))
    res, half, layers = float(g["resolution"]), int(g["half_shape"]), int(g["layers"])
    for tag, st in (("f32", fclb.F32), ("f64", fclb.F64)):
        ch, full, leaf, root, n_layers = fclb.octree_build_host(g["points"], res, half, st)
        assert n_layers == layers and np.array_equal(root, g[f"root_{tag}"])
        assert np.array_equal(ch, g[f"tree_children_{tag}"]) and np.array_equal(full, g[f"tree_full_{tag}"])
        assert np.array_equal(leaf, g[f"tree_leaf_{tag}"])
        pr = None
        for k, name in enumerate(("prune1", "prune2")):
            o = g["obb"][k]
            pr, full, leaf = fclb.octree_prune_host(ch, full, leaf, root, layers, o[:9].reshape(3, 3), o[9:12], o[12:15], st, pruned=pr)
            assert np.array_equal(pr, g[f"{name}_pruned_{tag}"]) and np.array_equal(full, g[f"{name}_full_{tag}"])
            assert np.array_equal(leaf, g[f"{name}_leaf_{tag}"])
        # pruneBy(second box, rebuild=True) on the once-pruned geometry = consolidation of the twice-pruned arrays
        c_ch, c_full, c_leaf = fclb.octree_consolidate_host(ch, pr, leaf, layers)
        assert np.array_equal(c_ch, g[f"rebuilt_children_{tag}"]) and np.array_equal(c_full, g[f"rebuilt_full_{tag}"])
        assert np.array_equal(c_leaf, g[f"rebuilt_leaf_{tag}"])
        assert pr.any() and len(c_full) < len(full)


def test_size_query_tree_is_not_reused_for_other_points():
    """The size query keeps its tree for the data call only when the point DATA is the same (full hash): a cloud rewritten in
    place between the two calls must be built afresh (round-1 advisor finding: 3 probe points let a stale tree through)."""
    import ctypes as C

    import fclb200 as fclb

    rng = np.random.Generator(np.random.PCG64(99))
    pts = np.ascontiguousarray(rng.uniform(-0.6, 0.6, size=(500, 3)))
    lib = fclb.load()
    ni, nl, layers = C.c_uint32(), C.c_uint32(), C.c_int()
    root = np.zeros(6, np.float64)
    P = fclb._ptr
    rc = lib.fclb_octree_build_host(P(pts), len(pts), 0.01, 64, fclb.F64, None, None, 0, C.byref(ni), None, 0, C.byref(nl), P(root),
                                    C.byref(layers))
    assert rc == 5
    pts[1:499] *= 0.5  # same address, same first / middle-ish / last point pattern is not enough: data changed
    cap_i, cap_l = 4096, 4096
    ch = np.zeros((cap_i, 8), np.uint32)
    full = np.zeros(cap_i, np.uint8)
    leaf = np.zeros(cap_l, np.uint8)
    rc = lib.fclb_octree_build_host(P(pts), len(pts), 0.01, 64, fclb.F64, P(ch), P(full), cap_i, C.byref(ni), P(leaf), cap_l,
                                    C.byref(nl), P(root), C.byref(layers))
    assert rc == 0
    fresh = fclb.octree_build_host(pts.copy(), 0.01, 64, fclb.F64)
    assert ni.value == len(fresh[1]) and nl.value == len(fresh[2])
    assert np.array_equal(ch[:ni.value], fresh[0]) and np.array_equal(leaf[:nl.value], fresh[2])
