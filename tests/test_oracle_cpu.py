"""CPU tests of the oracles (test infrastructure) -- no GPU needed.

  * known answers copied from the reference's own tests (file:line cited per case);
  * the committed golden vectors (tests/golden/, generated from the reference by
    make_golden.py) against the CPU restatement (oracle port), bit for bit;
  * when oracle/_ref is present (build container), the reference oracle against the
    same golden vectors, so the fixtures provably regenerate.
"""
import os

import numpy as np
import pytest

import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B, S, E, C, K, Y, V = range(7)
I = [1, 0, 0, 0, 1, 0, 0, 0, 1]


def pose(t=(0, 0, 0)):
    return np.array([I + list(t)], np.float64)


def shapes_from(arr):
    return [(int(r[0]), int(r[1]), tuple(r[2:5])) for r in arr]


def pairs_from(arr):
    return scenes.make_pairs(arr[:, 0].copy(), arr[:, 1].copy())


KNOWN_DISTANCE = [
    # (shape1, shape2, translation of shape2, expected distance or None for "not separated", tol, reference test)
    ((S, 0, (20,)), (B, 0, (5, 5, 5)), (0, 0, 0), None, 0, "test/test_fcl_geometric_shapes.cpp shapeDistance_boxsphere: identity"),
    ((S, 0, (20,)), (B, 0, (5, 5, 5)), (22.6, 0, 0), 0.1, 1e-3, "shapeDistance_boxsphere: x=22.6 -> 0.1"),
    ((S, 0, (20,)), (B, 0, (5, 5, 5)), (40, 0, 0), 17.5, 1e-3, "shapeDistance_boxsphere: x=40 -> 17.5"),
    ((S, 0, (20,)), (S, 0, (10,)), (0, 40, 0), 10.0, 1e-3, "test_fcl_geometric_shapes.cpp:3624-3640 spheresphere y=40 -> 10"),
    ((S, 0, (20,)), (S, 0, (10,)), (30.1, 0, 0), 0.1, 1e-3, "spheresphere x=30.1 -> 0.1"),
    ((S, 0, (20,)), (S, 0, (10,)), (29.9, 0, 0), None, 0, "spheresphere x=29.9 -> penetrating"),
    ((B, 0, (10, 10, 10)), (B, 0, (10, 10, 10)), (10.1, 0, 0), 0.1, 1e-3, "test_fcl_geometric_shapes.cpp:3717-3760 boxbox x=10.1 -> 0.1"),
    ((B, 0, (10, 10, 10)), (B, 0, (10, 10, 10)), (20.1, 0, 0), 10.1, 1e-3, "boxbox x=20.1 -> 10.1"),
    ((B, 0, (10, 10, 10)), (B, 0, (10, 10, 10)), (0, 20.2, 0), 10.2, 1e-3, "boxbox y=20.2 -> 10.2"),
    ((B, 0, (20, 40, 50)), (B, 0, (10, 10, 10)), (0, 0, 0), None, 0, "boxbox identity -> penetrating"),
    ((Y, 0, (5, 10)), (Y, 0, (5, 10)), (10.1, 0, 0), 0.1, 1e-3, "test_fcl_geometric_shapes.cpp:3870-3900 cylindercylinder x=10.1 -> 0.1"),
    ((Y, 0, (5, 10)), (Y, 0, (5, 10)), (0, 0, 0), None, 0, "cylindercylinder identity -> penetrating"),
]


@pytest.mark.parametrize("which", ["port", "ref"])
def test_known_answers_distance(which, port_oracle, request):
    oracle = port_oracle if which == "port" else request.getfixturevalue("ref_oracle")
    for s1, s2, t, expect, tol, note in KNOWN_DISTANCE:
        for dtype in (np.float32, np.float64):
            d, p1, p2, ok = oracle.distance_batch([s1, s2], scenes.make_pairs([0], [1]), pose().astype(dtype),
                                                  pose(t).astype(dtype))
            if expect is None:
                assert ok[0] == 0 and d[0] < 0, note
            else:
                assert ok[0] == 1 and abs(d[0] - expect) < tol, (note, d[0])


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", ["distance_c2", "distance_closed"])
def test_port_matches_golden_distance(port_oracle, name, tag):
    g = np.load(os.path.join(GOLD, f"{name}_{tag}.npz"))
    d, p1, p2, ok = port_oracle.distance_batch(shapes_from(g["shapes"]), pairs_from(g["pairs"]), g["poses1"], g["poses2"])
    assert np.array_equal(ok, g["ok"])
    sep = g["ok"] != 0
    assert np.array_equal(d[sep], g["dist"][sep]), "port oracle must be bit-identical to the reference"
    assert np.array_equal(p1[sep], g["p1"][sep]) and np.array_equal(p2[sep], g["p2"][sep])
    assert np.all(d[~sep] == -1)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_reference_regenerates_golden(ref_oracle, tag):
    g = np.load(os.path.join(GOLD, f"distance_c2_{tag}.npz"))
    d, p1, p2, ok = ref_oracle.distance_batch(shapes_from(g["shapes"]), pairs_from(g["pairs"]), g["poses1"], g["poses2"])
    assert np.array_equal(ok, g["ok"]) and np.array_equal(d, g["dist"])
    g = np.load(os.path.join(GOLD, f"gjk_epa_boxes_{tag}.npz"))
    gj, ep, mp, geom, it = ref_oracle.gjk_epa_batch(shapes_from(g["shapes"]), pairs_from(g["pairs"]), g["poses1"],
                                                    g["poses2"], mode=1)
    assert np.array_equal(gj, g["gjk"]) and np.array_equal(ep, g["epa"]) and np.array_equal(geom, g["geom"])
    g = np.load(os.path.join(GOLD, f"collide_{tag}.npz"))
    c, ct = ref_oracle.collide_batch(shapes_from(g["shapes"]), pairs_from(g["pairs"]), g["poses1"], g["poses2"],
                                     max_keep=4, max_contacts=4, penetration_mode=1)
    assert np.array_equal(c, g["counts_p1_m4"]) and np.array_equal(ct, g["contacts_p1_m4"])


def test_epa_sphere_known_answer(ref_oracle):
    """test/cvx_collide/test_epa2.cpp:21-71: two unit spheres overlapping by 0.2 -> EPA depth 0.2 +- 1e-4."""
    rng = np.random.Generator(np.random.PCG64(3))
    n = 2000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p1 = np.tile(np.array(I + [0, 0, 0], np.float64), (n, 1))
    p2 = p1.copy()
    p2[:, 9:] = d * 1.8
    shapes = [(S, 0, (1.0,)), (S, 0, (1.0,))]
    gj, ep, mp, geom, it = ref_oracle.gjk_epa_batch(shapes, scenes.make_pairs(np.zeros(n, np.uint32), np.ones(n, np.uint32)), p1, p2)
    assert np.all(gj == 0)
    okk = ep != 0
    assert okk.mean() > 0.999
    assert np.abs(geom[okk, 0] - 0.2).max() < 1e-4


def test_gjk_mpr_agree_with_boxbox_on_boxes(ref_oracle):
    """test/cvx_collide/test_gjk2.cpp:20-78 and test_mpr_primitive.cpp:106-146: GJK / MPR booleans agree with
    boxBox2 on random box pairs (here: the golden box batch; agreement up to knife-edge pairs)."""
    g = np.load(os.path.join(GOLD, "gjk_epa_boxes_f64.npz"))
    shapes, pairs = shapes_from(g["shapes"]), pairs_from(g["pairs"])
    c, _ = ref_oracle.collide_batch(shapes, pairs, g["poses1"], g["poses2"], max_contacts=4, penetration_mode=1, max_keep=4)
    bb = c > 0
    assert (bb != (g["gjk"] == 0)).sum() <= 2
    assert (bb != (g["mpr"] == 0)).sum() <= 2
    # EPA depth vs boxBox2 depth: one-sided, tolerance 1e-4 (test_epa2_with_gjk2.cpp:131)
    _, ct = ref_oracle.collide_batch(shapes, pairs, g["poses1"], g["poses2"], max_contacts=1, penetration_mode=1, max_keep=1)
    both = bb & (g["gjk"] == 0) & (g["epa"] != 0)
    assert np.all(g["geom"][both, 0] <= ct[both, 0, 8] + 1e-4)
