"""The MPR penetration request modes on the device (fclb_collide_batch with
FCLB_PEN_DIRECTED / FCLB_PEN_INCREMENTAL_MIN) against fcl::collide of the reference with
CollisionRequest::useDirectedPenetration / useIncrementalMinimumDistancePenetration
(collision_interface-inl.h:22-30 -> collisionPenetrationMPR, collision_penetration-inl.h:189-252;
MPR::RunDirectedPenetration mpr.hpp:497; RunIncrementalMinimumPenetrationDistance
mpr_incremental_penetration.hpp:191).  Counts identical; depth / normal / position within TOL,
the fraction of bit-identical contact records is reported."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-4, np.float64: 1e-6}
B, S, E, C, K, Y, V = scenes.BOX, scenes.SPHERE, scenes.ELLIPSOID, scenes.CAPSULE, scenes.CONE, scenes.CYLINDER, scenes.CONVEX


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [2, 3])
def test_mpr_penetration_modes(fclb, ref_oracle, dtype, mode):
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    hulls = [scenes.ellipsoid_mesh(0.2, 0.3, 0.4), scenes.random_hull16()]
    slots = [fclb.convex_upload(*m) for m in hulls]
    rslots = [ref_oracle.register_convex(*m) for m in hulls]
    prims = [(B, 0, (0.5, 0.4, 0.3)), (S, 0, (0.25,)), (E, 0, (0.3, 0.2, 0.25)), (C, 0, (0.15, 0.4)), (K, 0, (0.2, 0.5)),
             (Y, 0, (0.2, 0.4))]
    shapes = prims + [(V, slots[0], ()), (V, slots[1], ())]
    rshapes = prims + [(V, rslots[0], ()), (V, rslots[1], ())]
    combos = [(0, 0), (0, 1), (1, 1), (0, 3), (3, 5), (2, 4), (6, 7), (7, 6), (6, 0), (5, 1), (1, 3), (4, 4)]
    n = 24_000
    rng = np.random.Generator(np.random.PCG64(60 + mode))
    p1 = scenes.random_poses(rng, n, 0.35, dtype)
    p2 = scenes.random_poses(rng, n, 0.35, dtype)
    idx = np.arange(n) % len(combos)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32), np.array([combos[i][1] for i in idx], np.uint32))
    table = fclb.shapes_upload(shapes)
    for direction in ((0.0, 0.0, 1.0), (0.6, -0.48, 0.64)):
        req = fclb.make_request(max_contacts=1, penetration_mode=mode, direction=direction)
        counts, contacts = fclb.collide_batch_host(table, pairs, p1, p2, st, req, max_keep=1)
        e_counts, e_contacts = ref_oracle.collide_batch(rshapes, pairs, p1, p2, max_keep=1, threads=8, max_contacts=1,
                                                        penetration_mode=mode, direction=direction)
        mism = np.nonzero(counts != e_counts)[0]
        hit = (counts > 0) & (e_counts > 0)
        g, e = contacts[hit, 0], e_contacts[hit, 0]
        same = float((g == e).all(axis=1).mean())
        failed = int((e[:, 8] < 0).sum())
        dd = np.abs(g[:, 8] - e[:, 8])
        dn = np.abs(g[:, 2:5] - e[:, 2:5]).max(axis=1)
        dp = np.abs(g[:, 5:8] - e[:, 5:8]).max(axis=1)
        print(f"[MPR penetration mode={mode} dir={direction} {np.dtype(dtype).name}] n={n} colliding={int(hit.sum())} "
              f"count mismatches={len(mism)}; records bit-identical {same:.5f}; depth<0 (reference reports failure) {failed}; "
              f"max diff depth {dd.max():.2e} normal {dn.max():.2e} pos {dp.max():.2e}")
        assert len(mism) == 0
        assert ((g[:, 8] < 0) == (e[:, 8] < 0)).all()
        assert np.quantile(dd, 0.999) <= TOL[dtype] and np.quantile(dn, 0.999) <= 10 * TOL[dtype]
        assert np.quantile(dp, 0.999) <= 10 * TOL[dtype]
        assert same > 0.99
    fclb.release(table)
