"""The whole C5 scene step on the device (fclb_scene_self_collide_*: computeAABB, tree build,
SelfCollision, boolean collide per candidate) against the reference's CPU pipeline
(CollisionObject::computeAABB + BinaryAABB_Tree::Rebuild + SelfCollision with fcl::collide):
  * the candidate count equals the reference's;
  * for OUR ordered candidate list, numContacts of every pair equals fcl::collide on the same ordered pair;
  * the total number of colliding pairs is compared with the reference's own run (whose pair orientation
    follows its median-split tree) and any difference is printed -- it can only come from pairs whose
    boolean depends on the argument order, i.e. within rounding of touching."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scene_self_collide(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    n = 30_000
    shapes, shape_ids, poses = scenes.config_c5_scene(n, dtype, seed=5100)
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=1)
    cand, hits, pairs, counts = fclb.scene_self_collide(table, shape_ids, poses, n, st, req, host=True, want_pairs=True)
    r_hits, r_cand = ref_oracle.scene_self_collide(shapes, shape_ids, poses)
    print(f"[scene {np.dtype(dtype).name}] objects={n} candidates ours={cand} ref={r_cand}; colliding ours={hits} ref={r_hits}")
    assert cand == r_cand
    assert hits == int((counts > 0).sum())
    # exact narrowphase parity on our ordered pairs
    a, b = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    qpairs = scenes.make_pairs(shape_ids[a], shape_ids[b])
    e_counts, _ = ref_oracle.collide_batch(shapes, qpairs, np.ascontiguousarray(poses[a]), np.ascontiguousarray(poses[b]),
                                           max_keep=0, threads=8, want_contacts=False, max_contacts=1)
    mism = np.nonzero(counts != e_counts)[0]
    assert len(mism) == 0, (mism[:10], counts[mism[:10]], e_counts[mism[:10]])
    assert abs(hits - r_hits) <= max(2, r_hits // 100000), "colliding-pair totals differ beyond order-dependent knife edges"
    # device-pointer variant gives the same totals
    import torch

    d_ids = torch.from_numpy(shape_ids.view(np.int32)).cuda()
    d_pose = torch.from_numpy(poses).cuda()
    cand2, hits2 = fclb.scene_self_collide(table, d_ids, d_pose, n, st, req, host=False)
    assert (cand2, hits2) == (cand, hits)
    fclb.release(table)
