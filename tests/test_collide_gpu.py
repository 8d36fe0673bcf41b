"""Parity of the batched CUDA collide path against the reference:
  * fclb_collide_batch_*  vs fcl::collide(shape, shape)            (oracle collide_batch)
  * fclb_gjk_epa_batch_*  vs cvx_collide GJK + EPA driven directly  (oracle gjk_epa_batch),
    the way test/cvx_collide/test_epa2_with_gjk2.cpp:76-162 drives them.
Boolean results / contact counts / status codes must be identical except for listed
near-touching pairs; depths, normals and positions within TOL.
"""
import numpy as np
import pytest

import parity_util
import scenes

pytestmark = pytest.mark.gpu

B, S, E, C, K, Y, V = scenes.BOX, scenes.SPHERE, scenes.ELLIPSOID, scenes.CAPSULE, scenes.CONE, scenes.CYLINDER, scenes.CONVEX


def st_of(fclb, dtype):
    return fclb.F32 if dtype == np.float32 else fclb.F64


def mixed_batch(n, dtype, shapes, combos, extent, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    poses1 = scenes.random_poses(rng, n, extent, dtype)
    poses2 = scenes.random_poses(rng, n, extent, dtype)
    idx = np.arange(n) % len(combos)
    pairs = scenes.make_pairs(np.array([combos[i][0] for i in idx], np.uint32),
                              np.array([combos[i][1] for i in idx], np.uint32))
    return pairs, poses1, poses2, idx


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("penetration", [0, 1])
def test_closed_form_collide(fclb, ref_oracle, dtype, penetration):
    shapes = [(S, 0, (0.15,)), (B, 0, (0.4, 0.3, 0.2)), (C, 0, (0.1, 0.4)), (Y, 0, (0.12, 0.3)), (S, 0, (0.2,)),
              (B, 0, (0.3, 0.5, 0.25))]
    combos = [(0, 4), (0, 2), (2, 0), (0, 1), (1, 0), (0, 3), (3, 0), (1, 5)]
    n = 80_000
    pairs, p1, p2, idx = mixed_batch(n, dtype, shapes, combos, 0.35, 21)
    table = fclb.shapes_upload(shapes)
    for max_contacts in (1, 4):
        req = fclb.make_request(max_contacts=max_contacts, penetration_mode=penetration)
        counts, contacts = fclb.collide_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req, max_keep=4)
        kw = dict(max_contacts=max_contacts, penetration_mode=penetration)
        e_counts, e_contacts = ref_oracle.collide_batch(shapes, pairs, p1, p2, max_keep=4, threads=8, **kw)
        parity_util.check_collide(ref_oracle, "test_closed_form_collide", f"closed-form pairs pen={penetration} max={max_contacts}",
                                  dtype, shapes, pairs, p1, p2, (counts, contacts), (e_counts, e_contacts), kw, 4)
    fclb.release(table)


def convex_tables(fclb, oracles):
    (m0, m1) = (scenes.ellipsoid_mesh(0.2, 0.3, 0.4), scenes.random_hull16())
    slots = [fclb.convex_upload(*m0), fclb.convex_upload(*m1)]
    oslots = [[o.register_convex(*m0), o.register_convex(*m1)] for o in oracles]
    return slots, oslots


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gjk_boolean_and_epa(fclb, ref_oracle, dtype):
    """Direct cvx_collide path: GJK status identical, EPA status identical, depth / witness within TOL."""
    slots, oslots = convex_tables(fclb, [ref_oracle])
    cases = {
        "box-box": ([(B, 0, (2.0, 1.0, 0.5)), (B, 0, (1.0, 1.0, 1.0))], [(0, 1)], 2.0, 60_000),
        "capsule/cylinder-box": ([(C, 0, (0.3, 0.8)), (Y, 0, (0.3, 0.6)), (B, 0, (0.8, 0.6, 0.4))], [(0, 2), (1, 2), (2, 0)], 0.7, 40_000),
        "sphere/ellipsoid/cone mix": ([(S, 0, (0.3,)), (E, 0, (0.3, 0.2, 0.4)), (K, 0, (0.3, 0.7)), (B, 0, (0.5, 0.5, 0.5))],
                                      [(0, 3), (1, 3), (2, 3), (1, 2), (0, 1)], 0.6, 30_000),
    }
    for name, (shapes, combos, extent, n) in cases.items():
        pairs, p1, p2, idx = mixed_batch(n, dtype, shapes, combos, extent, 33)
        table = fclb.shapes_upload(shapes)
        req = fclb.make_request(max_contacts=1, penetration_mode=1)
        gjk, epa, geom = fclb.gjk_epa_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req)
        e_gjk, e_epa, _, e_geom, _ = ref_oracle.gjk_epa_batch(shapes, pairs, p1, p2, threads=8)
        parity_util.check_gjk_epa(ref_oracle, "test_gjk_boolean_and_epa", name, dtype, shapes, pairs, p1, p2, (gjk, epa, geom),
                                  (e_gjk, e_epa, e_geom))
        fclb.release(table)
    # convex-convex (58-vertex walk support vs 16-vertex scan support)
    shapes = [(V, slots[0], ()), (V, slots[1], ())]
    oshapes = [(V, oslots[0][0], ()), (V, oslots[0][1], ())]
    n = 40_000
    pairs, p1, p2, idx = mixed_batch(n, dtype, shapes, [(0, 1), (1, 0)], 0.3, 35)
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=1, penetration_mode=1)
    gjk, epa, geom = fclb.gjk_epa_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req)
    e_gjk, e_epa, _, e_geom, _ = ref_oracle.gjk_epa_batch(oshapes, pairs, p1, p2, threads=8)
    parity_util.check_gjk_epa(ref_oracle, "test_gjk_boolean_and_epa", "convex-convex", dtype, oshapes, pairs, p1, p2,
                              (gjk, epa, geom), (e_gjk, e_epa, e_geom))
    fclb.release(table)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("penetration", [0, 1])
def test_generic_collide_api(fclb, ref_oracle, dtype, penetration):
    """fcl::collide on pairs without a closed form: MPR (boolean) or GJK+EPA (contact)."""
    shapes = [(C, 0, (0.2, 0.6)), (Y, 0, (0.25, 0.5)), (B, 0, (0.6, 0.5, 0.4)), (E, 0, (0.3, 0.2, 0.25)), (K, 0, (0.25, 0.6))]
    combos = [(0, 2), (1, 2), (2, 1), (3, 2), (4, 0), (0, 1), (3, 3)]
    n = 70_000
    pairs, p1, p2, idx = mixed_batch(n, dtype, shapes, combos, 0.55, 41)
    table = fclb.shapes_upload(shapes)
    req = fclb.make_request(max_contacts=1, penetration_mode=penetration)
    counts, contacts = fclb.collide_batch_host(table, pairs, p1, p2, st_of(fclb, dtype), req, max_keep=1)
    e_counts, e_contacts = ref_oracle.collide_batch(shapes, pairs, p1, p2, max_keep=1, threads=8, max_contacts=1,
                                                    penetration_mode=penetration)
    parity_util.check_collide(ref_oracle, "test_generic_collide_api", f"generic pairs pen={penetration}", dtype, shapes, pairs,
                              p1, p2, (counts, contacts), (e_counts, e_contacts),
                              dict(max_contacts=1, penetration_mode=penetration), 1)
    fclb.release(table)
