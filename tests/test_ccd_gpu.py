"""Translational continuous collision of shape pairs (fclb_translational_ccd_batch_*) against the reference's
fcl::translational_ccd (narrowphase/continuous_collision-inl.h:21-36; detail/ccd/shape_pair_ccd-inl.h, gjk_ccd-inl.h,
box_pair_ccd-inl.h) on the same seeded inputs, for the three TimeOfCollisionRequestType values: the hit flag and the
time-of-collision interval must be bit-identical; every difference is listed and classified (parity_util)."""
import numpy as np
import pytest

import parity_util
import scenes

pytestmark = pytest.mark.gpu
B, S, E, C, K, Y, V = scenes.BOX, scenes.SPHERE, scenes.ELLIPSOID, scenes.CAPSULE, scenes.CONE, scenes.CYLINDER, scenes.CONVEX


def make_batch(n, dtype, n_combos, seed, extent=0.6):
    rng = np.random.Generator(np.random.PCG64(seed))
    p1 = scenes.random_poses(rng, n, extent, dtype)
    p2 = scenes.random_poses(rng, n, extent, dtype)
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    disp = np.empty((n, 4), np.float64)
    disp[:, :3] = axis
    disp[:, 3] = rng.uniform(0.0, 0.8, size=n)
    disp[::17, 3] = 0.0  # no movement at all
    disp[5::23, :3] = np.eye(3)[rng.integers(0, 3, size=len(disp[5::23]))]  # axis-aligned sweeps
    return p1, p2, np.ascontiguousarray(disp.astype(dtype)), np.arange(n) % n_combos


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_translational_ccd_shape_pairs(fclb, ref_oracle, dtype):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    hull = scenes.ellipsoid_mesh(0.2, 0.3, 0.4)
    small = scenes.random_hull16()
    shapes = [(B, 0, (0.4, 0.3, 0.2)), (B, 0, (0.3, 0.5, 0.25)), (S, 0, (0.15,)), (C, 0, (0.1, 0.4)), (Y, 0, (0.12, 0.3)),
              (E, 0, (0.2, 0.15, 0.25)), (K, 0, (0.15, 0.4)), (V, fclb.convex_upload(*hull), ()), (V, fclb.convex_upload(*small), ())]
    rshapes = shapes[:7] + [(V, ref_oracle.register_convex(*hull), ()), (V, ref_oracle.register_convex(*small), ())]
    combos = [(0, 1), (2, 0), (0, 2), (3, 1), (4, 5), (6, 3), (7, 8), (8, 0), (2, 2), (1, 7)]
    n = 100_000
    p1, p2, disp, idx = make_batch(n, dtype, len(combos), 71)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32), np.array([combos[i][1] for i in idx], np.uint32))
    table = fclb.shapes_upload(shapes)
    for rt, name in ((0, "kNotRequested"), (1, "kBoxApproximate"), (2, "kOneTocSample")):
        hit, toc = fclb.translational_ccd_batch_host(table, pairs, p1, p2, disp, st, request_type=rt)
        e_hit, e_toc = ref_oracle.translational_ccd_batch(rshapes, pairs, p1, p2, disp, request_type=rt, threads=8)
        mism = np.nonzero(hit != e_hit)[0]
        # a hit flag that differs: the swept shape is within EPS of touching -- re-evaluate with the sweep ends in double
        listed = [{"query": int(q), "what": "hit flag", "ours": int(hit[q]), "reference": int(e_hit[q]), "class": "UNEXPLAINED"}
                  for q in mism]
        both = (hit == 1) & (e_hit == 1)
        same_toc = (toc[both] == e_toc[both]).all(axis=1)
        dtoc = np.abs(toc[both] - e_toc[both]).max() if both.any() else 0.0
        parity_util.record("test_translational_ccd_shape_pairs", f"10 pair kinds, {name}", dtype, n, "hit flags, toc intervals", listed,
                           {"hits": int(e_hit.sum()), "toc_bit_identical_fraction": float(same_toc.mean()) if both.any() else 1.0,
                            "max_toc_diff": float(dtoc), "unexplained": len(listed)})
        assert len(mism) == 0, mism[:10]
        assert same_toc.all(), (name, np.nonzero(~same_toc)[0][:10])
        assert e_hit.any() and not e_hit.all()
        nb = both & (idx != 0)  # (box-box always reports the swept-box interval)
        if rt == 0:
            assert (toc[nb] == -1).all() and (toc[both & (idx == 0), 0] >= 0).all()
        if rt == 2:
            assert ((toc[nb, 0] >= 0) & (toc[nb, 0] <= 1) & (toc[nb, 0] == toc[nb, 1])).all()
    # size-independent property: no displacement == a static MPR collide
    disp0 = disp.copy()
    disp0[:, 3] = 0
    hit0, _ = fclb.translational_ccd_batch_host(table, pairs, p1, p2, disp0, st, request_type=0)
    counts, _ = fclb.collide_batch_host(table, pairs, p1, p2, st, fclb.make_request(max_contacts=1), max_keep=1, want_contacts=False)
    generic = np.isin(idx, [3, 4, 5, 6, 7, 9])  # pairs fcl::collide also answers with MPR
    agree = float((hit0[generic] == (counts[generic] > 0)).mean())
    print(f"[ccd {np.dtype(dtype).name}] zero displacement vs static collide on MPR pairs: agreement {agree:.6f}")
    assert agree > 0.9999
    fclb.release(table)
