"""Parity bookkeeping shared by the `-m gpu` tests (SURVEY.md 8(d) "Parity reporting").

Rule: every result that differs from the reference's is LISTED, re-evaluated by the
reference oracle (oracle/_ref) in double precision and CLASSIFIED; a test passes only when
each listed query is within EPS of touching (or, for a continuous quantity, when the
device answer is within TOL of one of the reference's own re-evaluations).  Anything else is an
unexplained mismatch and fails the test.  The lists are appended, one JSON object per
line, to $FCLB_PARITY_LOG (default gpurun_out/parity_r02.jsonl when that directory
exists); the committed copy is profiles/parity_r02.json.

  EPS_TOUCH  1e-4 * scale (float), 1e-6 * scale (double); scale = 1 (shape sizes are O(1) in every config)
  TOL        1e-4 (float), 1e-6 (double) absolute on distances / depths / points -- the reference's own
             test tolerances (test_epa2_with_gjk2.cpp:160, test_gjk2_distance.cpp)
"""
from __future__ import annotations

import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {np.float32: 1e-4, np.float64: 1e-6}
EPS_TOUCH = {np.float32: 1e-4, np.float64: 1e-6}


def tol(dtype):
    return TOL[np.dtype(dtype).type]


def eps_touch(dtype):
    return EPS_TOUCH[np.dtype(dtype).type]


def _log_path():
    p = os.environ.get("FCLB_PARITY_LOG")
    if p:
        return p
    d = os.path.join(ROOT, "gpurun_out")
    return os.path.join(d, "parity_r02.jsonl") if os.path.isdir(d) else None


def record(test, config, dtype, n, compared, mismatches, extra=None):
    """One line per (test, config, dtype): what was compared bit for bit and every listed mismatch."""
    entry = {"test": test, "config": config, "dtype": np.dtype(dtype).name, "n": int(n), "compared": compared,
             "n_mismatches": len(mismatches), "mismatches": mismatches[:200]}
    if extra:
        entry.update(extra)
    print("[parity] " + json.dumps({k: v for k, v in entry.items() if k != "mismatches"}))
    for m in mismatches[:20]:
        print("[parity]    listed: " + json.dumps(m))
    p = _log_path()
    if p:
        with open(p, "a") as f:
            f.write(json.dumps(entry) + "\n")
    return entry


def signed_distance_f64(ref_oracle, shapes, pairs, poses1, poses2, idx):
    """GJKSolver<double>::shapeSignedDistance of the listed queries (the reference in the wider type):
    > 0 separated by that much, < 0 penetrating by that much; nan when the reference's EPA fails."""
    idx = np.asarray(idx, np.int64)
    if idx.size == 0:
        return np.zeros(0)
    p1 = np.ascontiguousarray(poses1[idx].astype(np.float64))
    p2 = np.ascontiguousarray(poses2[idx].astype(np.float64))
    d, _, _, ok = ref_oracle.signed_distance_batch(shapes, np.ascontiguousarray(pairs[idx]), p1, p2, threads=1)
    d = d.astype(np.float64)
    d[ok == 0] = np.nan
    return d


def boolean_flips_under_perturbation(ref_oracle, shapes, pairs, poses1, poses2, idx, eps, evaluate):
    """Fallback classifier: the reference's own double-precision boolean evaluated with shape 2 shifted by +-eps along
    the three axes; a query whose answer changes inside that ball is within eps of touching."""
    out = []
    for q in np.asarray(idx, np.int64):
        p1 = np.ascontiguousarray(poses1[q:q + 1].astype(np.float64))
        answers = set()
        for ax in range(3):
            for sgn in (-1.0, 1.0):
                p2 = np.ascontiguousarray(poses2[q:q + 1].astype(np.float64))
                p2[0, 9 + ax] += sgn * eps
                answers.add(bool(evaluate(shapes, np.ascontiguousarray(pairs[q:q + 1]), p1, p2)))
        out.append(len(answers) > 1)
    return np.asarray(out, bool)


def classify_touching(ref_oracle, shapes, pairs, poses1, poses2, idx, dtype, ours, ref, what, evaluate=None):
    """List + classify the queries `idx` whose discrete result (boolean / count / status) differs.
    Returns (listed, unexplained): each listed item carries both answers and the double-precision signed distance."""
    eps = eps_touch(dtype)
    idx = np.asarray(idx, np.int64)
    sd = signed_distance_f64(ref_oracle, shapes, pairs, poses1, poses2, idx)
    listed, unexplained = [], []
    for k, q in enumerate(idx):
        item = {"query": int(q), "what": what, "ours": int(ours[q]), "reference": int(ref[q]),
                "signed_distance_f64": None if np.isnan(sd[k]) else float(sd[k])}
        near = (not np.isnan(sd[k])) and abs(sd[k]) <= eps
        if not near and evaluate is not None:
            near = bool(boolean_flips_under_perturbation(ref_oracle, shapes, pairs, poses1, poses2, [q], eps, evaluate)[0])
            item["flips_within_eps_ball"] = near
        item["class"] = "within eps of touching" if near else "UNEXPLAINED"
        listed.append(item)
        if not near:
            unexplained.append(item)
    return listed, unexplained


def classify_continuous(idx, ours, ref, alts, dtype, what, scale=1.0):
    """Continuous results (depth, distance, witness coordinate) that differ by more than TOL between the device and the
    reference in S.  `alts` = {label: values} are the reference's own re-evaluations of the same queries (in double, and /
    or with a perturbed tolerance): a listed query is explained when the device's answer is within TOL of at least one of
    them, i.e. the difference is the rounding / tolerance sensitivity of the reference's algorithm at that query (an
    ill-conditioned polytope), not a different algorithm."""
    t = tol(dtype) * scale
    listed, unexplained = [], []
    for k, q in enumerate(np.asarray(idx, np.int64)):
        o, r = float(ours[k]), float(ref[k])
        item = {"query": int(q), "what": what, "ours": o, "reference": r}
        best = None
        for label, vals in alts.items():
            item["reference_" + label] = float(vals[k])
            if abs(o - float(vals[k])) <= t:
                best = label
        item["class"] = (f"within TOL of the reference re-evaluated {best}" if best else "UNEXPLAINED")
        listed.append(item)
        if best is None:
            unexplained.append(item)
    return listed, unexplained


def check_gjk_epa(ref_oracle, test, config, dtype, rshapes, pairs, p1, p2, ours, ref):
    """cvx_collide GJK + EPA driven directly (test/cvx_collide/test_epa2_with_gjk2.cpp:76-162): GJK status, EPA status, depth
    and witness points of the device against the reference's, every difference listed and classified.  Returns the log entry."""
    gjk, epa, geom = ours
    e_gjk, e_epa, e_geom = ref
    n = len(gjk)
    t = tol(dtype)
    listed, unexplained = classify_touching(ref_oracle, rshapes, pairs, p1, p2, np.nonzero(gjk != e_gjk)[0], dtype, gjk, e_gjk,
                                            "GJK status")
    both = (gjk == 0) & (e_gjk == 0)
    ddepth = np.abs(geom[:, 0] - e_geom[:, 0])
    dwit = np.abs(geom[:, 1:] - e_geom[:, 1:]).max(axis=1)
    bad = np.nonzero(both & ((epa != e_epa) | (ddepth > t) | (dwit > t)))[0]
    if bad.size:
        sub = lambda a: np.ascontiguousarray(a[bad])
        alts = {}
        p1d, p2d = sub(p1).astype(np.float64), sub(p2).astype(np.float64)
        alts["in double"] = ref_oracle.gjk_epa_batch(rshapes, sub(pairs), p1d, p2d, threads=1)[3][:, 0]
        for label, dtol in (("with EPA tolerance 2e-6", 2e-6), ("with EPA tolerance 5e-7", 5e-7)):
            alts[label] = ref_oracle.gjk_epa_batch(rshapes, sub(pairs), sub(p1), sub(p2), threads=1, distance_tol=dtol)[3][:, 0]
        l2, u2 = classify_continuous(bad, geom[bad, 0], e_geom[bad, 0], alts, dtype, "EPA depth")
        for item, q in zip(l2, bad):
            item.update({"epa_status_ours": int(epa[q]), "epa_status_reference": int(e_epa[q]),
                         "max_witness_diff": float(dwit[q])})
        listed += l2
        unexplained += u2
    ok = both & (epa == e_epa)
    ident = float((geom[ok] == e_geom[ok]).all(axis=1).mean()) if ok.any() else 1.0
    entry = record(test, config, dtype, n, "GJK status, EPA status, depth, witness points", listed,
                   {"intersecting": int(both.sum()), "records_bit_identical_fraction": ident,
                    "max_depth_diff": float(ddepth[both].max()) if both.any() else 0.0,
                    "unexplained": len(unexplained)})
    assert not unexplained, unexplained[:5]
    return entry


def check_collide(ref_oracle, test, config, dtype, rshapes, pairs, p1, p2, ours, ref, req_kw, max_keep):
    """fcl::collide(shape, shape) per query: contact counts (bit-exact, mismatches classified as near-touching) and, when
    the request generates contacts, every stored contact record {normal, pos, depth} within TOL; records beyond TOL are
    listed and re-evaluated by the reference in double (a contact set is explained when each device contact is within TOL
    of a contact of the reference in S or in double -- boxBox2 may cull another subset of one contact polygon,
    DESIGN.md 3)."""
    counts, contacts = ours
    e_counts, e_contacts = ref
    n = len(counts)
    t = tol(dtype)

    def boolean(shapes, pr, a, b):
        c, _ = ref_oracle.collide_batch(shapes, pr, a, b, max_keep=1, threads=1, want_contacts=False, **req_kw)
        return c[0] > 0

    listed, unexplained = classify_touching(ref_oracle, rshapes, pairs, p1, p2, np.nonzero(counts != e_counts)[0], dtype,
                                            counts, e_counts, "contact count", evaluate=boolean)
    ident = None
    if contacts is not None and e_contacts is not None and req_kw.get("penetration_mode", 0):
        same = (counts == e_counts) & (e_counts > 0)
        slot = np.arange(max_keep)[None, :] < np.minimum(e_counts, max_keep)[:, None]
        diff = np.abs(contacts[..., 2:] - e_contacts[..., 2:]).max(axis=-1)
        diff = np.where(slot, diff, 0.0).max(axis=1)
        bad = np.nonzero(same & (diff > t))[0]
        m = same[:, None] & slot
        ident = float((contacts[m][:, 2:] == e_contacts[m][:, 2:]).all(axis=-1).mean()) if m.any() else 1.0
        if bad.size:
            sub = lambda a: np.ascontiguousarray(a[bad])
            c64, k64 = ref_oracle.collide_batch(rshapes, sub(pairs), sub(p1).astype(np.float64), sub(p2).astype(np.float64),
                                                max_keep=max_keep, threads=1, **req_kw)
            for j, q in enumerate(bad):
                k = int(min(counts[q], max_keep))
                pool = np.concatenate([e_contacts[q, :k, 2:].astype(np.float64), k64[j, :min(int(c64[j]), max_keep), 2:]])
                worst = max(float(np.abs(pool - contacts[q, i, 2:].astype(np.float64)).max(axis=1).min()) for i in range(k))
                item = {"query": int(q), "what": "contact records", "count": k, "max_slotwise_diff": float(diff[q]),
                        "worst_distance_to_a_reference_contact": worst, "count_reference_f64": int(c64[j])}
                item["class"] = ("every device contact within TOL of a contact of the reference (in S or in double)"
                                 if worst <= t else "UNEXPLAINED")
                listed.append(item)
                if worst > t:
                    unexplained.append(item)
    entry = record(test, config, dtype, n, "contact counts" + (", contact records" if ident is not None else ""), listed,
                   {"colliding": int((e_counts > 0).sum()), "contacts": int(e_counts.sum()),
                    "records_bit_identical_fraction": ident, "unexplained": len(unexplained)})
    assert not unexplained, unexplained[:5]
    return entry


def check_distance(ref_oracle, test, config, dtype, rshapes, pairs, p1, p2, ours, ref):
    """GJKSolver::shapeDistance per query: separated flags (mismatches classified), distances and witness points within
    TOL (listed and re-evaluated in double otherwise)."""
    g_dist, g_p1, g_p2, g_ok = ours
    e_dist, e_p1, e_p2, e_ok = ref
    n = len(e_ok)
    t = tol(dtype)
    g_sep, e_sep = g_ok != 0, e_ok != 0
    listed, unexplained = classify_touching(ref_oracle, rshapes, pairs, p1, p2, np.nonzero(g_sep != e_sep)[0], dtype,
                                            g_sep, e_sep, "separated flag")
    both = g_sep & e_sep
    dd = np.where(both, np.abs(g_dist - e_dist), 0.0)
    bad = np.nonzero(dd > t)[0]
    if bad.size:
        sub = lambda a: np.ascontiguousarray(a[bad])
        d64 = ref_oracle.distance_batch(rshapes, sub(pairs), sub(p1).astype(np.float64), sub(p2).astype(np.float64))[0]
        l2, u2 = classify_continuous(bad, g_dist[bad], e_dist[bad], {"in double": d64}, dtype, "distance")
        listed += l2
        unexplained += u2
    # witness points: where the reference's extraction is valid (ok == 1) they must realise the reported distance, and
    # agree coordinate-wise within TOL unless listed (a flat closest feature has no unique witness pair)
    valid = both & (g_ok == 1)
    wd = np.where(valid, np.maximum(np.abs(g_p1 - e_p1).max(axis=1), np.abs(g_p2 - e_p2).max(axis=1)), 0.0)
    wbad = np.nonzero(wd > t)[0]
    for q in wbad:
        realised = float(np.linalg.norm(g_p1[q].astype(np.float64) - g_p2[q]))
        ok = abs(realised - float(g_dist[q])) <= 10 * t
        item = {"query": int(q), "what": "witness points", "max_coordinate_diff": float(wd[q]), "ours_realised_distance": realised,
                "ours_distance": float(g_dist[q]),
                "class": "non-unique witness pair: the device's points realise its distance" if ok else "UNEXPLAINED"}
        listed.append(item)
        if not ok:
            unexplained.append(item)
    ident = float((g_dist[both] == e_dist[both]).mean()) if both.any() else 1.0
    entry = record(test, config, dtype, n, "separated flags, distances, witness points", listed,
                   {"separated": int(e_sep.sum()), "distances_bit_identical_fraction": ident,
                    "max_distance_diff": float(dd.max()) if n else 0.0, "max_witness_diff": float(wd.max()) if n else 0.0,
                    "witness_invalid_in_reference": int((both & (g_ok == 3)).sum()), "unexplained": len(unexplained)})
    assert not unexplained, unexplained[:5]
    not_sep = ~g_sep & ~e_sep
    assert np.all(g_dist[not_sep] == -1)
    return entry
