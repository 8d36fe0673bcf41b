"""Parity of the device broadphase (fclb_broadphase_*, fclb_compute_aabb_batch_*) against the
reference's BinaryAABB_Tree (broadphase/binary_AABB_tree-inl.h) and CollisionObject::computeAABB:
world AABBs bit-identical; the SET of reported pairs identical for SelfCollision, TreeCollision,
SingleObjectCollision and after UpdateObjectAABB (the reference's own tests compare against a
brute-force pair set, test/broadphase/test_binary_AABB_tree_collision.cpp)."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def canon(pairs, ordered):
    p = np.asarray(pairs, np.uint64).reshape(-1, 2)
    if not ordered:
        p = np.sort(p, axis=1)
    return set(map(tuple, p.tolist()))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_compute_aabb_and_pair_sets(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    n = 20_000
    shapes, shape_ids, poses = scenes.config_c5_scene(n, dtype, seed=5001)
    hull = scenes.random_hull16(scale=(1.0, 0.8, 1.2))
    shapes = shapes + [(scenes.CAPSULE, 0, (0.5, 2.0)), (scenes.CONE, 0, (0.8, 2.0)), (scenes.ELLIPSOID, 0, (1.0, 0.5, 0.7))]
    rshapes = list(shapes) + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())]
    shapes = shapes + [(scenes.CONVEX, fclb.convex_upload(*hull), ())]
    shape_ids = (shape_ids + 3 * (np.arange(n) % 5 == 0) + (np.arange(n) % 7 == 0)).astype(np.uint32) % len(shapes)
    table = fclb.shapes_upload(shapes)
    boxes = fclb.compute_aabb_batch_host(table, shape_ids, poses, st)
    e_boxes = ref_oracle.compute_aabb_batch(rshapes, shape_ids, poses)
    assert np.array_equal(boxes, e_boxes), "world AABBs differ from CollisionObject::computeAABB"
    ids = (np.arange(n, dtype=np.uint64) * 3 + 11)
    tree = fclb.broadphase_build_host(boxes, ids, st)
    rtree = ref_oracle.broadphase_create(e_boxes, ids)
    ours = fclb.broadphase_self_pairs_host(tree)
    ref = ref_oracle.broadphase_self_pairs(rtree)
    print(f"[broadphase self {np.dtype(dtype).name}] n={n} pairs ours={len(ours)} ref={len(ref)} "
          f"box tests/object {fclb.broadphase_last_visits() / n:.1f}")
    assert len(ours) == len(ref)
    assert canon(ours, False) == canon(ref, False)
    assert len(canon(ours, False)) == len(ours), "a pair was reported twice"
    # second tree + TreeCollision
    m = 5_000
    shapes2, sid2, poses2 = scenes.config_c5_scene(m, dtype, seed=5002, neighbours=2.0)
    table2 = fclb.shapes_upload(shapes2)
    boxes2 = fclb.compute_aabb_batch_host(table2, sid2, poses2, st)
    ids2 = np.arange(m, dtype=np.uint64) + 1_000_000
    tree2 = fclb.broadphase_build_host(boxes2, ids2, st)
    rtree2 = ref_oracle.broadphase_create(boxes2, ids2)
    ours = fclb.broadphase_tree_pairs_host(tree, tree2)
    ref = ref_oracle.broadphase_tree_pairs(rtree, rtree2)
    print(f"[broadphase tree-tree] pairs ours={len(ours)} ref={len(ref)}")
    assert canon(ours, True) == canon(ref, True) and len(ours) == len(ref)
    # SingleObjectCollision for a batch of query boxes
    ours = fclb.broadphase_query_pairs_host(tree, boxes2[:500], ids2[:500], st)
    ref = ref_oracle.broadphase_query_pairs(rtree, boxes2[:500], ids2[:500])
    assert canon(ours, True) == canon(ref, True) and len(ours) == len(ref)
    # UpdateObjectAABB: move 10 % of the objects
    rng = np.random.Generator(np.random.PCG64(9))
    sel = rng.choice(n, n // 10, replace=False)
    moved = boxes[sel] + np.tile(rng.uniform(-3, 3, size=(len(sel), 3)).astype(dtype), 2)
    fclb.broadphase_update_host(tree, ids[sel], moved, st)
    assert ref_oracle.broadphase_update(rtree, ids[sel], moved) == 1
    ours = fclb.broadphase_self_pairs_host(tree)
    ref = ref_oracle.broadphase_self_pairs(rtree)
    print(f"[broadphase after update] pairs ours={len(ours)} ref={len(ref)}")
    assert canon(ours, False) == canon(ref, False) and len(ours) == len(ref)
    with pytest.raises(fclb.FclbError):
        fclb.broadphase_update_host(tree, np.array([5], np.uint64), moved[:1], st)  # unknown id
    for t in (tree, tree2):
        fclb.broadphase_release(t)
    fclb.release(table)
    fclb.release(table2)


def test_broadphase_edge_cases(fclb, ref_oracle):
    st = fclb.F64
    one = np.array([[0, 0, 0, 1, 1, 1]], np.float64)
    t1 = fclb.broadphase_build_host(one, np.array([7], np.uint64), st)
    assert len(fclb.broadphase_self_pairs_host(t1)) == 0
    q = fclb.broadphase_query_pairs_host(t1, np.array([[0.5, 0.5, 0.5, 2, 2, 2], [3, 3, 3, 4, 4, 4]], np.float64),
                                         np.array([1, 2], np.uint64), st)
    assert q.tolist() == [[7, 1]]
    # identical boxes (equal Morton codes) and touching boxes (overlap is inclusive)
    same = np.tile(one, (65, 1))
    same[64] = [1, 1, 1, 2, 2, 2]
    t2 = fclb.broadphase_build_host(same, np.arange(65, dtype=np.uint64), st)
    pairs = fclb.broadphase_self_pairs_host(t2)
    assert len(pairs) == 65 * 64 // 2
    rt = ref_oracle.broadphase_create(same, np.arange(65, dtype=np.uint64))
    assert canon(pairs, False) == canon(ref_oracle.broadphase_self_pairs(rt), False)
    fclb.broadphase_release(t1)
    fclb.broadphase_release(t2)
