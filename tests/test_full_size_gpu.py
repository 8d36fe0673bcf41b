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

import parity_util
import scenes

pytestmark = pytest.mark.gpu
TOL = 1e-4


DTYPES = [np.float32, np.float64]


def st_of(fclb, dtype):
    return fclb.F32 if dtype == np.float32 else fclb.F64


@pytest.mark.parametrize("dtype", DTYPES)
def test_c2_full_size(fclb, ref_oracle, dtype):
    n = 10_000_000
    shapes, pairs, p1, p2 = scenes.config_c2(n, dtype)
    table = fclb.shapes_upload(shapes)
    r = fclb.distance_batch_host(table, pairs, p1, p2, st_of(fclb, dtype))
    exp = ref_oracle.distance_batch(shapes, pairs, p1, p2, threads=16)
    parity_util.check_distance(ref_oracle, "test_c2_full_size", "C2 10M mixed-primitive distance", dtype, shapes, pairs, p1, p2,
                               (r.dist, r.p1, r.p2, r.ok), exp)
    valid = (exp[3] != 0) & (r.ok == 1)
    gap = np.abs(np.linalg.norm(r.p1[valid].astype(np.float64) - r.p2[valid], axis=1) - r.dist[valid])
    assert gap.max() <= (1e-5 if dtype == np.float32 else 1e-12), "|p1 - p2| must equal the reported distance"
    fclb.release(table)


@pytest.mark.parametrize("dtype", DTYPES)
def test_c1a_full_size(fclb, ref_oracle, dtype):
    n = 1_000_000
    shapes, pairs, p1, p2 = scenes.config_c1_boxes(n, dtype)
    table = fclb.shapes_upload(shapes)
    kw = dict(max_contacts=4, penetration_mode=1)
    counts, contacts = fclb.collide_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), fclb.make_request(**kw), max_keep=4)
    ref = ref_oracle.collide_batch(shapes, pairs, p1, p2, max_keep=4, threads=16, **kw)
    parity_util.check_collide(ref_oracle, "test_c1a_full_size", "C1a 1M box-box fcl::collide (boxBox2), max_contacts=4", dtype,
                              shapes, pairs, p1, p2, (counts, contacts), ref, kw, 4)
    fclb.release(table)


@pytest.mark.parametrize("dtype", DTYPES)
def test_c1b_boxes_full_size(fclb, ref_oracle, dtype):
    """C1b (i): 1M box-box pairs through GJK(128, 1e-6) + EPA(256, 255, 1e-6), test_epa2_with_gjk2.cpp:76-162."""
    n = 1_000_000
    shapes, pairs, p1, p2 = scenes.config_c1_boxes(n, dtype, seed=1002)
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=1, penetration_mode=1)
    ours = fclb.gjk_epa_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req)
    e_gjk, e_epa, _, e_geom, _ = ref_oracle.gjk_epa_batch(shapes, pairs, p1, p2, threads=16)
    parity_util.check_gjk_epa(ref_oracle, "test_c1b_boxes_full_size", "C1b 1M box-box GJK+EPA", dtype, shapes, pairs, p1, p2, ours,
                              (e_gjk, e_epa, e_geom))
    fclb.release(table)


@pytest.mark.parametrize("dtype", DTYPES)
def test_c1b_convex_full_size(fclb, ref_oracle, dtype):
    """C1b (ii): 1M convex-convex pairs (58-vertex hill-climb hull vs 16-vertex scan hull) through the cvx_collide path
    driven directly AND through fcl::collide with DefaultGJK_EPA (gjk_solver-inl.h:100-129)."""
    n = 1_000_000
    convex, pairs, p1, p2 = scenes.config_c1_convex(n, dtype)
    shapes = [(scenes.CONVEX, fclb.convex_upload(*m), ()) for m in convex]
    rshapes = [(scenes.CONVEX, ref_oracle.register_convex(*m), ()) for m in convex]
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=1, penetration_mode=1)
    ours = fclb.gjk_epa_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req)
    e_gjk, e_epa, _, e_geom, _ = ref_oracle.gjk_epa_batch(rshapes, pairs, p1, p2, threads=16)
    parity_util.check_gjk_epa(ref_oracle, "test_c1b_convex_full_size", "C1b 1M convex-convex GJK+EPA", dtype, rshapes, pairs, p1,
                              p2, ours, (e_gjk, e_epa, e_geom))
    kw = dict(max_contacts=1, penetration_mode=1)
    counts, contacts = fclb.collide_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req, max_keep=1)
    ref = ref_oracle.collide_batch(rshapes, pairs, p1, p2, max_keep=1, threads=16, **kw)
    parity_util.check_collide(ref_oracle, "test_c1b_convex_full_size", "C1b 1M convex-convex fcl::collide DefaultGJK_EPA", dtype,
                              rshapes, pairs, p1, p2, (counts, contacts), ref, kw, 1)
    fclb.release(table)


@pytest.mark.parametrize("dtype", DTYPES)
def test_c3_full_size(fclb, ref_oracle, dtype):
    n = 1_000_000
    meshes = [scenes.noisy_uv_sphere(), scenes.noisy_torus()]
    st = st_of(fclb, dtype)
    handles = [fclb.bvh_build(v, t, st) for v, t in meshes]
    ids = [ref_oracle.bvh_create(v, t) for v, t in meshes]
    p1, p2 = scenes.config_c3_poses(n, dtype)
    counts, _ = fclb.bvh_collide_batch_host(handles[0], handles[1], p1, p2, st, fclb.make_request(max_contacts=1))
    e_counts, _ = ref_oracle.bvh_collide_batch(ids[0], ids[1], p1, p2, threads=16, want_pair=False, max_contacts=1)
    mism = np.nonzero(counts != e_counts)[0]
    listed = [{"query": int(q), "what": "boolean", "ours": int(counts[q]), "reference": int(e_counts[q]), "class": "UNEXPLAINED"}
              for q in mism]
    parity_util.record("test_c3_full_size", "C3 1M mesh-mesh OBBRSS boolean collide, 2 x 10k triangles", dtype, n, "booleans", listed,
                       {"colliding": int(e_counts.sum()), "unexplained": len(listed)})
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
    """One full 100k-object scene: the candidate-pair count equals the reference pipeline's, and fcl::collide of the reference
    on every candidate pair the device reports gives the device's boolean (mismatches listed + classified)."""
    n = 100_000
    dtype = np.float32
    shapes, shape_ids, poses = scenes.config_c5_scene(n, dtype, seed=5000)
    table = fclb.shapes_upload(shapes)
    kw = dict(max_contacts=1, penetration_mode=0)
    cand, hits, id_pairs, counts = fclb.scene_self_collide(table, shape_ids, poses, n, fclb.F32, fclb.make_request(max_contacts=1),
                                                           want_pairs=True)
    r_hits, r_cand = ref_oracle.scene_self_collide(shapes, shape_ids, poses)
    assert cand == r_cand, (cand, r_cand)
    a, b = id_pairs[:, 0].astype(np.int64), id_pairs[:, 1].astype(np.int64)
    pairs = scenes.make_pairs(shape_ids[a], shape_ids[b])
    pa, pb = np.ascontiguousarray(poses[a]), np.ascontiguousarray(poses[b])
    ref = ref_oracle.collide_batch(shapes, pairs, pa, pb, max_keep=1, threads=16, want_contacts=False, **kw)
    parity_util.check_collide(ref_oracle, "test_c5_full_size", "C5 one 100k-object scene: boolean collide on every candidate pair",
                              dtype, shapes, pairs, pa, pb, (counts, None), ref, kw, 1)
    # the reference's own pipeline visits a self pair in the order its median-split tree gives; fcl::collide(a, b) and
    # fcl::collide(b, a) agree except within rounding of touching, so its total must lie between the totals of the two orders
    ref_sw = ref_oracle.collide_batch(shapes, scenes.make_pairs(shape_ids[b], shape_ids[a]), pb, pa, max_keep=1, threads=16,
                                      want_contacts=False, **kw)[0]
    lo, hi = int(((ref[0] > 0) & (ref_sw > 0)).sum()), int(((ref[0] > 0) | (ref_sw > 0)).sum())
    print(f"[C5 full size] objects={n} candidates={cand}; colliding ours={hits} reference pipeline={r_hits} "
          f"(order-independent pairs {lo}, order-dependent {hi - lo})")
    assert lo <= r_hits <= hi
    fclb.release(table)
