"""CUDA path against the committed golden vectors (generated from the reference by
tests/golden/make_golden.py).  Needs no oracle library at run time."""
import os

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shapes_from(arr):
    return [(int(r[0]), int(r[1]), tuple(r[2:5])) for r in arr]


def pairs_from(arr):
    return scenes.make_pairs(arr[:, 0].copy(), arr[:, 1].copy())


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", ["distance_c2", "distance_closed"])
def test_distance_golden(fclb, name, tag):
    g = np.load(os.path.join(GOLD, f"{name}_{tag}.npz"))
    st = fclb.F32 if tag == "f32" else fclb.F64
    table = fclb.shapes_upload(shapes_from(g["shapes"]))
    r = fclb.distance_batch_host(table, pairs_from(g["pairs"]), g["poses1"], g["poses2"], st)
    assert np.array_equal(r.ok != 0, g["ok"] != 0)
    sep = g["ok"] != 0
    tol = 1e-4 if tag == "f32" else 1e-6
    assert np.abs(r.dist[sep] - g["dist"][sep]).max() <= tol
    print(f"{name} {tag}: bit-identical distances {(r.dist[sep] == g['dist'][sep]).mean():.5f}")
    fclb.release(table)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_collide_golden(fclb, tag):
    g = np.load(os.path.join(GOLD, f"collide_{tag}.npz"))
    st = fclb.F32 if tag == "f32" else fclb.F64
    table = fclb.shapes_upload(shapes_from(g["shapes"]))
    pairs = pairs_from(g["pairs"])
    for pen in (0, 1):
        for mc in (1, 4):
            req = fclb.make_request(max_contacts=mc, penetration_mode=pen)
            counts, contacts = fclb.collide_batch_host(table, pairs, g["poses1"], g["poses2"], st, req, max_keep=4)
            assert np.array_equal(counts, g[f"counts_p{pen}_m{mc}"]), (pen, mc)
            if pen:
                tol = 1e-4 if tag == "f32" else 1e-6
                d = np.abs(contacts[..., 2:] - g[f"contacts_p{pen}_m{mc}"][..., 2:])
                assert np.quantile(d, 0.999) <= tol
    fclb.release(table)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gjk_epa_golden(fclb, tag):
    g = np.load(os.path.join(GOLD, f"gjk_epa_boxes_{tag}.npz"))
    st = fclb.F32 if tag == "f32" else fclb.F64
    table = fclb.shapes_upload(shapes_from(g["shapes"]))
    gjk, epa, geom = fclb.gjk_epa_batch_host(table, pairs_from(g["pairs"]), g["poses1"], g["poses2"], st,
                                             fclb.make_request(max_contacts=1, penetration_mode=1))
    assert np.array_equal(gjk, g["gjk"])
    hit = g["gjk"] == 0
    assert np.array_equal(epa[hit], g["epa"][hit])
    tol = 1e-4 if tag == "f32" else 1e-6
    assert np.abs(geom[hit, 0] - g["geom"][hit, 0]).max() <= tol
    fclb.release(table)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_mesh_golden(fclb, tag):
    g = np.load(os.path.join(GOLD, f"mesh_{tag}.npz"))
    st = fclb.F32 if tag == "f32" else fclb.F64
    h1 = fclb.bvh_upload(g["obb1"], g["child1"], g["tri1"], st)
    h2 = fclb.bvh_upload(g["obb2"], g["child2"], g["tri2"], st)
    c, _ = fclb.bvh_collide_batch_host(h1, h2, g["poses1"], g["poses2"], st, fclb.make_request(max_contacts=1))
    assert np.array_equal(c, g["counts_bool"])
    c, _ = fclb.bvh_collide_batch_host(h1, h2, g["poses1"], g["poses2"], st, fclb.make_request(max_contacts=2**31 - 1))
    assert np.array_equal(c, g["counts_all"])
    fclb.bvh_release(h1)
    fclb.bvh_release(h2)
