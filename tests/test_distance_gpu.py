"""Parity of the batched CUDA distance path (fclb_distance_batch_*) against the
reference's GJKSolver<S>::shapeDistance (gjk_solver-inl.h:801) run by the
reference oracle on the same seeded inputs.

Bar (BASELINE.json north_star): the separated/not-separated flag is bit-exact
except for pairs within EPS of touching, which are listed and counted; distance
and witness points within TOL.
"""
import numpy as np
import pytest

import parity_util
import scenes

pytestmark = pytest.mark.gpu

TOL = parity_util.TOL


def compare_distance(ref_oracle, shapes, pairs, poses1, poses2, got, exp, dtype, label, test):
    parity_util.check_distance(ref_oracle, test, label, dtype, shapes, pairs, poses1, poses2, got, exp)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c2_mixed_primitive_distance(fclb, ref_oracle, dtype):
    n = 300_000
    shapes, pairs, poses1, poses2 = scenes.config_c2(n, dtype)
    table = fclb.shapes_upload(shapes)
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    r = fclb.distance_batch_host(table, pairs, poses1, poses2, st)
    exp = ref_oracle.distance_batch(shapes, pairs, poses1, poses2, threads=8)
    compare_distance(ref_oracle, shapes, pairs, poses1, poses2, (r.dist, r.p1, r.p2, r.ok), exp, dtype, "C2 at 300k",
                     "test_c2_mixed_primitive_distance")
    for k, name in enumerate(("sphere-box", "capsule-box", "cylinder-box")):
        sel = np.arange(n) % 3 == k
        same = (r.dist[sel] == exp[0][sel]).mean()
        print(f"  {name}: bit-identical distance fraction {same:.6f}")
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_closed_form_pairs(fclb, ref_oracle, dtype):
    """Every closed-form distance specialisation (gjk_solver-inl.h:902-988), both argument orders."""
    S, B, C, Y = scenes.SPHERE, scenes.BOX, scenes.CAPSULE, scenes.CYLINDER
    shapes = [(S, 0, (0.07,)), (B, 0, (0.3, 0.2, 0.1)), (C, 0, (0.05, 0.25)), (Y, 0, (0.08, 0.2)), (S, 0, (0.11,)),
              (C, 0, (0.04, 0.3))]
    combos = [(0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (3, 0), (0, 4), (2, 5)]
    n = 40_000
    rng = np.random.Generator(np.random.PCG64(11))
    poses1 = scenes.random_poses(rng, n, 0.4, dtype)
    poses2 = scenes.random_poses(rng, n, 0.4, dtype)
    idx = np.arange(n) % len(combos)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32),
                              np.array([combos[i][1] for i in idx], np.uint32))
    table = fclb.shapes_upload(shapes)
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    r = fclb.distance_batch_host(table, pairs, poses1, poses2, st)
    exp = ref_oracle.distance_batch(shapes, pairs, poses1, poses2, threads=8)
    # capsule-capsule returns "true" with a possibly negative distance (capsule_capsule-inl.h:141-246)
    cc = idx == 7
    assert np.array_equal(r.ok[cc] != 0, exp[3][cc] != 0)
    assert np.abs(r.dist[cc] - exp[0][cc]).max() <= TOL[dtype]
    keep = ~cc
    compare_distance(ref_oracle, shapes, np.ascontiguousarray(pairs[keep]), np.ascontiguousarray(poses1[keep]),
                     np.ascontiguousarray(poses2[keep]), (r.dist[keep], r.p1[keep], r.p2[keep], r.ok[keep]),
                     tuple(a[keep] for a in exp), dtype, "closed-form distance pairs", "test_closed_form_pairs")
    fclb.release(table)


def test_dev_entry_point_matches_host(fclb):
    import torch

    n = 50_000
    shapes, pairs, poses1, poses2 = scenes.config_c2(n, np.float32)
    table = fclb.shapes_upload(shapes)
    h = fclb.distance_batch_host(table, pairs, poses1, poses2, fclb.F32)
    dev = torch.device("cuda:0")
    d_pairs = torch.from_numpy(pairs.view(np.uint32).reshape(n, 2).astype(np.int64)).to(torch.int32).to(dev) \
        if False else torch.from_numpy(pairs.view(np.uint32).reshape(n, 2).view(np.int32)).to(dev)
    d_p1 = torch.from_numpy(poses1).to(dev)
    d_p2 = torch.from_numpy(poses2).to(dev)
    dist = torch.empty(n, dtype=torch.float32, device=dev)
    w1 = torch.empty(n, 3, dtype=torch.float32, device=dev)
    w2 = torch.empty(n, 3, dtype=torch.float32, device=dev)
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    fclb.distance_batch_dev(table, d_pairs, d_p1, d_p2, n, fclb.F32, dist, w1, w2, ok)
    assert np.array_equal(dist.cpu().numpy(), h.dist)
    assert np.array_equal(ok.cpu().numpy(), h.ok)
    assert np.array_equal(w1.cpu().numpy(), h.p1)
    fclb.release(table)


def test_empty_and_single(fclb, ref_oracle):
    shapes, pairs, poses1, poses2 = scenes.config_c2(1, np.float64)
    table = fclb.shapes_upload(shapes)
    r = fclb.distance_batch_host(table, pairs[:0], poses1[:0], poses2[:0], fclb.F64)
    assert r.dist.size == 0
    r = fclb.distance_batch_host(table, pairs, poses1, poses2, fclb.F64)
    exp = ref_oracle.distance_batch(shapes, pairs, poses1, poses2)
    assert (r.ok[0] != 0) == (exp[3][0] != 0)
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_signed_distance(fclb, ref_oracle, dtype):
    """fclb_signed_distance_batch vs GJKSolver::shapeSignedDistance (gjk_solver-inl.h:810-880): generic GJK
    distance when separated, -(EPA depth) with the EPA witness points when penetrating."""
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    tol = 1e-4 if dtype == np.float32 else 1e-6
    hull = scenes.ellipsoid_mesh(0.2, 0.3, 0.4)
    shapes = [(scenes.SPHERE, 0, (0.25,)), (scenes.BOX, 0, (0.5, 0.4, 0.3)), (scenes.CAPSULE, 0, (0.15, 0.4)),
              (scenes.CYLINDER, 0, (0.2, 0.4)), (scenes.CONVEX, fclb.convex_upload(*hull), ())]
    rshapes = shapes[:4] + [(scenes.CONVEX, ref_oracle.register_convex(*hull), ())]
    combos = [(0, 1), (1, 1), (2, 1), (3, 2), (4, 1), (4, 4), (0, 0)]
    n = 21_000
    rng = np.random.Generator(np.random.PCG64(77))
    p1 = scenes.random_poses(rng, n, 0.45, dtype)
    p2 = scenes.random_poses(rng, n, 0.45, dtype)
    idx = np.arange(n) % len(combos)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32), np.array([combos[i][1] for i in idx], np.uint32))
    table = fclb.shapes_upload(shapes)
    r = fclb.signed_distance_batch_host(table, pairs, p1, p2, st)
    e_dist, e_p1, e_p2, e_ok = ref_oracle.signed_distance_batch(rshapes, pairs, p1, p2, threads=8)
    # ok flag: listed + classified (the flag flips when GJK neither separates nor intersects, or EPA fails: near touching)
    mism = np.nonzero((r.ok != 0) != (e_ok != 0))[0]
    listed, unexplained = parity_util.classify_touching(ref_oracle, rshapes, pairs, p1, p2, mism, dtype, r.ok, e_ok, "ok flag")
    both = (r.ok != 0) & (e_ok != 0)
    pen = both & (e_dist < 0)
    same = float(((r.dist[both] == e_dist[both]) & (r.p1[both] == e_p1[both]).all(axis=1)).mean())
    dd = np.where(both, np.abs(r.dist - e_dist), 0.0)
    bad = np.nonzero(dd > tol)[0]
    if bad.size:
        sub = lambda a: np.ascontiguousarray(a[bad])
        d64 = ref_oracle.signed_distance_batch(rshapes, sub(pairs), sub(p1).astype(np.float64), sub(p2).astype(np.float64))[0]
        l2, u2 = parity_util.classify_continuous(bad, r.dist[bad], e_dist[bad], {"in double": d64}, dtype, "signed distance")
        listed += l2
        unexplained += u2
    dp = np.where(both, np.maximum(np.abs(r.p1 - e_p1).max(axis=1), np.abs(r.p2 - e_p2).max(axis=1)), 0.0)
    parity_util.record("test_signed_distance", "7 pair kinds incl. convex", dtype, n, "ok flags, signed distances", listed,
                       {"separated": int((both & ~pen).sum()), "penetrating": int(pen.sum()), "records_bit_identical_fraction": same,
                        "max_distance_diff": float(dd.max()), "max_witness_diff": float(dp.max()), "unexplained": len(unexplained)})
    assert not unexplained, unexplained[:5]
    assert (r.dist[(r.ok == 0)] == -1).all()
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_qt7_pose_encoding(fclb, ref_oracle, dtype):
    """FCLB_POSE_QT7 host entry point: the device expands quaternion + translation with Eigen's toRotationMatrix arithmetic
    in S; results equal the 12-S entry point on the expanded poses bit for bit, and the reference's on the same poses."""
    n = 200_000
    st = fclb.F32 if dtype == np.float32 else fclb.F64
    shapes, pairs, qt1, qt2 = scenes.config_c2_qt(n, dtype)
    p1, p2 = scenes.expand_qt7(qt1), scenes.expand_qt7(qt2)
    table = fclb.shapes_upload(shapes)
    a = fclb.distance_batch_qt_host(table, pairs, qt1, qt2, st)
    b = fclb.distance_batch_host(table, pairs, p1, p2, st)
    for x, y in ((a.dist, b.dist), (a.p1, b.p1), (a.p2, b.p2), (a.ok, b.ok)):
        assert np.array_equal(x, y)
    exp = ref_oracle.distance_batch(shapes, pairs, p1, p2, threads=8)
    compare_distance(ref_oracle, shapes, pairs, p1, p2, (a.dist, a.p1, a.p2, a.ok), exp, dtype, "C2 200k, QT7 poses",
                     "test_qt7_pose_encoding")
    # the rotation the device builds is orthonormal to rounding
    R = p1[:1000, :9].reshape(-1, 3, 3).astype(np.float64)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < (1e-5 if dtype == np.float32 else 1e-13)
    fclb.release(table)


_STAGED = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import fclb200, scenes
fclb200.init(0)
n = int(sys.argv[4])
shapes, pairs, qt1, qt2 = scenes.config_c2_qt(n, np.float32)
table = fclb200.shapes_upload(shapes)
a = fclb200.distance_batch_qt_host(table, pairs, qt1, qt2, fclb200.F32)
b = fclb200.distance_batch_host(table, pairs, scenes.expand_qt7(qt1), scenes.expand_qt7(qt2), fclb200.F32)
np.savez(sys.argv[3], qd=a.dist, qp1=a.p1, qp2=a.p2, qok=a.ok, d=b.dist, p1=b.p1, p2=b.p2, ok=b.ok)
"""


def _stage_sizes(n, chunk, taper):
    """The schedule of distance_batch_host_fmt (fclb_engine.cu)."""
    out, b = [], 0
    while b < n:
        rem, m = n - b, chunk
        if taper and rem < 2 * chunk:
            m = max(taper, (rem // 2 + 4095) // 4096 * 4096)
            if rem < m + taper:
                m = rem
        m = min(m, rem)
        out.append(m)
        b += m
    return out


@pytest.mark.parametrize("chunk,taper,n", [(32768, 4096, 200_001), (16384, 0, 50_001), (4096, 65536, 50_001)])
def test_host_stage_schedule(fclb, tmp_path, chunk, taper, n):
    """The pipeline stages of fclb_distance_batch_*host (FCLB_HOST_CHUNK queries, tail tapering to FCLB_HOST_TAPER; read at
    engine start, hence the child process) cut the batch at different places -- ragged last stages included -- and
    every query's result equals the single-stage call's bit for bit."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "staged.npz")
    env = dict(os.environ, FCLB_HOST_CHUNK=str(chunk), FCLB_HOST_TAPER=str(taper), FCLB_TRACE_HOST="1")
    p = subprocess.run([sys.executable, "-c", _STAGED, os.path.join(root, "mind-fcl_b200"), os.path.join(root, "oracle"), out,
                        str(n)], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    stages = [int(l.split()[4]) for l in p.stderr.splitlines() if l.startswith("fclb trace: stage")]
    want = _stage_sizes(n, chunk, max(taper, 4096) if taper else 0)
    assert len(want) > 3 and stages == want + want, (stages, want)
    shapes, pairs, qt1, qt2 = scenes.config_c2_qt(n, np.float32)
    table = fclb.shapes_upload(shapes)
    one = fclb.distance_batch_host(table, pairs, scenes.expand_qt7(qt1), scenes.expand_qt7(qt2), fclb.F32)  # 2M stages: one
    z = np.load(out)
    for k, ref in (("d", one.dist), ("p1", one.p1), ("p2", one.p2), ("ok", one.ok)):
        assert np.array_equal(z[k], ref) and np.array_equal(z["q" + k], ref)
    fclb.release(table)
