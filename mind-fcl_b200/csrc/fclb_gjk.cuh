// fclb_gjk.cuh -- per-thread GJK: boolean intersection + separation distance.
//
// Behavioural contract: include/fcl/cvx_collide/gjk.hpp (Evaluate, :11-146;
// simplexProjection2/3/4, :162-362) and gjk_distance.hpp (distance loop :11-106,
// computeMinDistanceAndUpdateSimplex2/3/4 :128-371, witness extraction :376-470).
//
// B200 design (not the reference's): one query per thread.  The reference keeps
// a GJKSimplex of four {vertex, direction} structs and copies whole simplices
// when it tries sub-simplices.  Here the four {vertex, direction} slots live in
// shared memory (thread-strided, bank-conflict free, natively indexable) and a
// simplex is just an ORDER WORD: byte k holds the slot id of the reference's
// vertices[k].  "Copy the simplex, drop a vertex, recurse" becomes integer
// shuffling of that word; no vertex is ever moved.  All floating-point
// expressions keep the reference's operation order (see fclb_math.cuh).
#pragma once
#include "fclb_shapes.cuh"

namespace fclb {

enum GjkStatus : int {  // == cvx_collide::GJK_Status (gjk.h:13-29)
  GJK_INTERSECT = 0,
  GJK_SEPARATED = 1,
  GJK_NO_PROGRESS = 2,
  GJK_ITER_LIMIT = 3,
  GJK_FAILED = 4
};

// Per-thread view of the 4-slot {vertex, direction} store in shared memory.
// Element (slot, c) of thread t sits at base[(slot*6 + c) * stride + t].
template <typename S>
struct SlotStore {
  S* base;
  int stride;
  FCLB_DI V3<S> vtx(int slot) const {
    const S* p = base + (slot * 6) * stride;
    return mk<S>(p[0], p[stride], p[2 * stride]);
  }
  FCLB_DI V3<S> dir(int slot) const {
    const S* p = base + (slot * 6 + 3) * stride;
    return mk<S>(p[0], p[stride], p[2 * stride]);
  }
  FCLB_DI void put(int slot, const V3<S>& v, const V3<S>& d) {
    S* p = base + (slot * 6) * stride;
    p[0] = v.x;
    p[stride] = v.y;
    p[2 * stride] = v.z;
    p[3 * stride] = d.x;
    p[4 * stride] = d.y;
    p[5 * stride] = d.z;
  }
};

// Simplex = rank + order word (byte k = slot of reference vertices[k]).
struct Simp {
  uint32_t ord;
  int rank;
};
FCLB_DI int slotOf(const Simp& s, int k) { return (s.ord >> (8 * k)) & 0xff; }
FCLB_DI uint32_t ord1(int a) { return uint32_t(a); }
FCLB_DI uint32_t ord2(int a, int b) { return uint32_t(a) | (uint32_t(b) << 8); }
FCLB_DI uint32_t ord3(int a, int b, int c) { return uint32_t(a) | (uint32_t(b) << 8) | (uint32_t(c) << 16); }
FCLB_DI int freeSlot(const Simp& s) {  // branch-free: lowest slot not named by the first `rank` bytes
  uint32_t used = (s.rank > 0) ? (1u << (s.ord & 3u)) : 0u;
  used |= (s.rank > 1) ? (1u << ((s.ord >> 8) & 3u)) : 0u;
  used |= (s.rank > 2) ? (1u << ((s.ord >> 16) & 3u)) : 0u;
  used |= (s.rank > 3) ? (1u << ((s.ord >> 24) & 3u)) : 0u;
  return __ffs(~used) - 1;
}
template <typename S>
FCLB_DI void addVertex(SlotStore<S>& st, Simp& s, const V3<S>& v, const V3<S>& d) {
  if (s.rank < 0) s.rank = 0;
  const int slot = freeSlot(s);
  st.put(slot, v, d);
  s.ord = (s.ord & ~(0xffu << (8 * s.rank))) | (uint32_t(slot) << (8 * s.rank));
  s.rank += 1;
}

enum ProjStatus : int { PROJ_FAILED = 0, PROJ_CONTINUE = 1, PROJ_INTERSECT = 2, PROJ_ZERO_VOLUME = 3 };

// gjk.hpp:162-201
template <typename S>
FCLB_DI int simplexProjection2(const SlotStore<S>& st, Simp& s, V3<S>& direction) {
  const int sB = slotOf(s, 0), sA = slotOf(s, 1);
  const V3<S> a = st.vtx(sA), b = st.vtx(sB);
  const V3<S> a_to_b = b - a;
  const S ao_dot_ab = -dot(a_to_b, a);
  const V3<S> ab_cross_ao = cross(a, a_to_b);
  if (ao_dot_ab > 0 && sqnorm(ab_cross_ao) <= S(0)) return PROJ_INTERSECT;
  if (ao_dot_ab <= 0) {
    s.ord = ord1(sA);
    s.rank = 1;
    direction = -normalized(a);
    return PROJ_CONTINUE;
  }
  direction = normalized(cross(ab_cross_ao, a_to_b));
  return PROJ_CONTINUE;
}

// gjk.hpp:204-286
template <typename S>
FCLB_DI int simplexProjection3(const SlotStore<S>& st, Simp& s, V3<S>& direction, S tol) {
  const int sC = slotOf(s, 0), sB = slotOf(s, 1), sA = slotOf(s, 2);
  const V3<S> a = st.vtx(sA), b = st.vtx(sB), c = st.vtx(sC);
  const V3<S> a_to_b = b - a;
  const V3<S> a_to_c = c - a;
  const bool a_sep_b = dot(a, a_to_b) >= 0;
  const bool a_sep_c = dot(a, a_to_c) >= 0;
  if (a_sep_b && a_sep_c) {
    s.ord = ord1(sA);
    s.rank = 1;
    direction = normalized(-a);
    return PROJ_CONTINUE;
  }
  const V3<S> abc_normal = cross(a_to_b, a_to_c);
  const V3<S> ac_normal_in_abc = cross(abc_normal, a_to_c);
  if (dot(a, ac_normal_in_abc) <= 0) {
    s.ord = ord2(sC, sA);
    s.rank = 2;
    const V3<S> ac_cross_ao = cross(a, a_to_c);
    direction = normalized(cross(ac_cross_ao, a_to_c));
    return PROJ_CONTINUE;
  }
  const V3<S> ab_normal_in_abc = cross(a_to_b, abc_normal);
  if (dot(a, ab_normal_in_abc) <= 0) {
    s.ord = ord2(sB, sA);
    s.rank = 2;
    const V3<S> ab_cross_ao = cross(a, a_to_b);
    direction = normalized(cross(ab_cross_ao, a_to_b));
    return PROJ_CONTINUE;
  }
  const S area = norm(abc_normal);
  if (area < tol * tol) return PROJ_FAILED;
  const V3<S> n_unit = abc_normal / area;
  const S n_dot_oa = dot(n_unit, a);
  if (fabs_(n_dot_oa) < tol) return PROJ_INTERSECT;
  direction = (n_dot_oa <= 0) ? n_unit : -n_unit;
  return PROJ_CONTINUE;
}

// gjk.hpp:289-362 (tetrahedron case) falling into :204-286 (triangle case).
// One function so the triangle code exists once in the SASS.
template <typename S>
FCLB_DI int simplexProjection(const SlotStore<S>& st, Simp& s, V3<S>& direction, S tol) {
  if (s.rank == 2) return simplexProjection2(st, s, direction);
  if (s.rank == 4) {
    const int sD = slotOf(s, 0), sC = slotOf(s, 1), sB = slotOf(s, 2), sA = slotOf(s, 3);
    const V3<S> a = st.vtx(sA), b = st.vtx(sB), c = st.vtx(sC), d = st.vtx(sD);
    const V3<S> ab = b - a, ac = c - a, ad = d - a;
    V3<S> abc_n = cross(ab, ac);
    V3<S> acd_n = cross(ac, ad);
    V3<S> abd_n = cross(ab, ad);
    const S abc_dot_ad = dot(abc_n, ad);
    const S acd_dot_ab = dot(acd_n, ab);
    const S abd_dot_ac = dot(abd_n, ac);
    if (fabs_(abc_dot_ad) <= S(0)) return PROJ_ZERO_VOLUME;
    if (abc_dot_ad > 0) abc_n = abc_n * S(-1);
    if (acd_dot_ab > 0) acd_n = acd_n * S(-1);
    if (abd_dot_ac > 0) abd_n = abd_n * S(-1);
    const bool d_side = dot(a, abc_n) > 0;
    const bool c_side = dot(a, abd_n) > 0;
    const bool b_side = dot(a, acd_n) > 0;
    if (d_side && c_side && b_side) return PROJ_INTERSECT;
    if (!b_side) {
      s.ord = ord3(sD, sC, sA);  // remove b
    } else if (!c_side) {
      s.ord = ord3(sD, sB, sA);  // remove c
    } else {
      s.ord = ord3(sC, sB, sA);  // remove d
    }
    s.rank = 3;
  }
  return simplexProjection3(st, s, direction, tol);
}

// ---------------------------------------------------------------------------
// Separation distance: sub-simplex closest point.
// gjk_distance.hpp:128-153.  (ord, rank) in/out; returns the closest point.
template <typename S>
FCLB_DI V3<S> minDist2(const SlotStore<S>& st, Simp& s, S tol) {
  const int slot_new = slotOf(s, 1), slot_old = slotOf(s, 0);
  const V3<S> s1 = st.vtx(slot_new), s2 = st.vtx(slot_old);
  const V3<S> s1_to_s2 = s2 - s1;
  const S sq_len = sqnorm(s1_to_s2);
  const S t = -dot(s1, s1_to_s2);
  if (t <= 0 || sq_len <= tol * tol) {
    s.ord = ord1(slot_new);
    s.rank = 1;
    return s1;
  } else if (t >= sq_len) {
    s.ord = ord1(slot_old);
    s.rank = 1;
    return s2;
  }
  const S w2 = t / sq_len;
  return w2 * s2 + (S(1.0) - w2) * s1;
}

// gjk_distance.hpp:156-289 (triangle) and :128-153 (segment) as ONE routine.
// The reference evaluates up to three edge sub-simplices of a triangle through
// separate code paths; here the candidate edges are collected into a mask and
// evaluated by one loop body (same order, same strict "<" selection).  A
// rank-2 input is the degenerate case "one candidate edge, no triangle".
// Keeping a single copy of each piece keeps the kernel inside the 32 KB L1.5
// instruction cache.
template <typename S>
FCLB_DI V3<S> subSimplexClosest(const SlotStore<S>& st, Simp& s, S tol) {
  int k0, k1, k2, mask;
  V3<S> s1 = zero3<S>(), n = zero3<S>();
  S area_sq = S(0);
  if (s.rank == 2) {
    k1 = slotOf(s, 0);
    k2 = slotOf(s, 1);
    k0 = k1;
    mask = 1;
  } else {
    k0 = slotOf(s, 0);
    k1 = slotOf(s, 1);
    k2 = slotOf(s, 2);
    s1 = st.vtx(k2);
    const V3<S> s2 = st.vtx(k1), s3 = st.vtx(k0);
    const V3<S> s1_to_s2 = s2 - s1;
    const V3<S> s1_to_s3 = s3 - s1;
    const bool s1_sep_s2 = dot(s1, s1_to_s2) >= 0;
    const bool s1_sep_s3 = dot(s1, s1_to_s3) >= 0;
    if (s1_sep_s2 && s1_sep_s3) {
      s.ord = ord1(k2);
      s.rank = 1;
      return s1;
    }
    n = cross(s1_to_s2, s1_to_s3);
    area_sq = sqnorm(n);
    const bool zero_area = area_sq <= S(0);
    const bool s12_sep = dot(s1, cross(n, s1_to_s2)) > 0;
    const bool s13_sep = dot(s1, cross(n, s1_to_s3)) < 0;
    const bool s23_sep = dot(s2, cross(n, s3 - s2)) > 0;
    // candidate edges, bit 0: [k1,k2]  bit 1: [k0,k2]  bit 2: [k0,k1]
    if (!s1_sep_s2 && s12_sep) {
      mask = 1;  // decided: segment s1s2 (:186-193)
    } else if (!s1_sep_s3 && s13_sep) {
      mask = 2;  // decided: segment s1s3 (:200-206)
    } else {
      mask = ((zero_area || s12_sep) ? 1 : 0) | ((zero_area || s13_sep) ? 2 : 0) | ((zero_area || s23_sep) ? 4 : 0);
    }
  }
  S best_sq = S(-1);
  V3<S> best_pt = zero3<S>();
  Simp best_s = s;
#pragma unroll 1
  for (int e = 0; e < 3; e++) {
    if (!(mask & (1 << e))) continue;
    Simp c;
    c.rank = 2;
    c.ord = (e == 0) ? ord2(k1, k2) : ((e == 1) ? ord2(k0, k2) : ord2(k0, k1));
    const V3<S> p = minDist2(st, c, tol);
    const S d2 = sqnorm(p);
    if (best_sq < 0 || d2 < best_sq) {
      best_sq = d2;
      best_s = c;
      best_pt = p;
    }
  }
  if (best_sq < 0) {  // only reachable from the triangle case: interior projection
    const S d = dot(s1, n);
    return n * (d / area_sq);
  }
  s = best_s;
  return best_pt;
}
template <typename S>
FCLB_DI V3<S> minDist3(const SlotStore<S>& st, Simp& s, S tol) {
  return subSimplexClosest(st, s, tol);
}

// gjk_distance.hpp:109-126 with :292-371 (tetrahedron = best of the three faces
// through the newest vertex).  status: 0 NoImprovement, 1 OK, 2 Failed
template <typename S>
FCLB_DI int minDistUpdate(const SlotStore<S>& st, Simp& s, V3<S>& out, S tol) {
  if (s.rank == 1) {
    out = st.vtx(slotOf(s, 0));
    return 1;
  }
  if (s.rank < 1 || s.rank > 4) return 2;
  const bool tetra = (s.rank == 4);
  const int k0 = slotOf(s, 0), k1 = slotOf(s, 1), k2 = slotOf(s, 2), k3 = slotOf(s, 3);
  const int n_cand = tetra ? 3 : 1;
  S best_sq = S(-1);
  V3<S> best_pt = zero3<S>();
  Simp best_s = s;
#pragma unroll 1
  for (int f = 0; f < n_cand; f++) {
    Simp c = s;
    if (tetra) {
      c.rank = 3;
      c.ord = (f == 0) ? ord3(k1, k2, k3) : ((f == 1) ? ord3(k0, k2, k3) : ord3(k0, k1, k3));
    }
    const V3<S> p = subSimplexClosest(st, c, tol);
    const S d2 = sqnorm(p);
    if (best_sq < 0 || d2 < best_sq) {
      best_sq = d2;
      best_pt = p;
      best_s = c;
    }
  }
  if (best_sq < 0) return 0;
  s = best_s;
  out = best_pt;
  return 1;
}

// gjk_distance.hpp:376-470 (extractSeparationPointNoSubSimplex)
template <typename S, typename MD>
FCLB_DI bool extractSeparationPoint(const MD& shape, const SlotStore<S>& st, const Simp& s, V3<S>& p0, V3<S>& p1) {
  constexpr S bary_tol = S(1e-3);  // gjk.h:131
  if (s.rank == 4 || s.rank <= 0) return false;
  if (s.rank == 1) {
    const V3<S> d = st.dir(slotOf(s, 0));
    p0 = shape.support0(d);
    p1 = shape.support1(-d);
    return true;
  } else if (s.rank == 2) {
    const int ka = slotOf(s, 0), kb = slotOf(s, 1);
    const V3<S> s1 = st.vtx(ka), s2 = st.vtx(kb);
    const V3<S> d1 = st.dir(ka), d2 = st.dir(kb);
    const V3<S> s1_to_s2 = s2 - s1;
    const S sq_len = sqnorm(s1_to_s2);
    if (sq_len <= S(0)) {
      p0 = shape.support0(d1);
      p1 = shape.support1(-d1);
      return true;
    }
    const S t = -dot(s1, s1_to_s2);
    const S w2 = t / sq_len;
    if (w2 > 1 + bary_tol) {
      p0 = shape.support0(d2);
      p1 = shape.support1(-d2);
      return false;
    } else if (w2 < -bary_tol) {
      p0 = shape.support0(d1);
      p1 = shape.support1(-d1);
      return false;
    }
    const S w1 = S(1.0) - w2;
    p0 = shape.support0(d1) * w1 + shape.support0(d2) * w2;
    p1 = shape.support1(-d1) * w1 + shape.support1(-d2) * w2;
    return true;
  }
  const int ka = slotOf(s, 0), kb = slotOf(s, 1), kc = slotOf(s, 2);
  const V3<S> s1 = st.vtx(ka), s2 = st.vtx(kb), s3 = st.vtx(kc);
  const V3<S> s1_to_s2 = s2 - s1;
  const V3<S> s1_to_s3 = s3 - s1;
  const V3<S> n = cross(s1_to_s2, s1_to_s3);
  const S area_sq = sqnorm(n);
  if (area_sq <= S(0)) return false;
  const S d = dot(s1, n);
  const V3<S> o_proj = n * (d / area_sq);
  const S area = fsqrt(area_sq);
  const S w2 = norm(cross(s1_to_s3, s1 - o_proj)) / area;
  const S w3 = norm(cross(s1_to_s2, s1 - o_proj)) / area;
  const S w1 = S(1.0) - w2 - w3;
  if (w1 < -bary_tol || w2 < -bary_tol || w3 < -bary_tol) return false;
  const V3<S> d1 = st.dir(ka), d2 = st.dir(kb), d3 = st.dir(kc);
  p0 = (shape.support0(d1) * w1 + shape.support0(d2) * w2) + shape.support0(d3) * w3;
  p1 = (shape.support1(-d1) * w1 + shape.support1(-d2) * w2) + shape.support1(-d3) * w3;
  return true;
}

// Witness extraction split in two for the state-machine kernel: the PLAN (which
// stored directions to support along, with which weights; gjk_distance.hpp
// :376-470) is pure simplex arithmetic, the supports themselves are evaluated by
// the kernel's single shared support site.
//   n == 0         : nothing to evaluate, result invalid (:381-384,:438,:451-455)
//   weighted==false: one support pair along slot[0], p = support directly
//   weighted==true : p = sum_k support(slot[k]) * w[k], accumulated left to right
template <typename S>
struct ExtractPlan {
  int n;
  bool weighted;
  bool valid;
  uint32_t slots;  // byte k = slot of the k-th direction
  S w0, w1, w2;
};
template <typename S>
FCLB_DI ExtractPlan<S> planExtraction(const SlotStore<S>& st, const Simp& s) {
  constexpr S bary_tol = S(1e-3);  // gjk.h:131
  ExtractPlan<S> p;
  p.n = 0;
  p.weighted = false;
  p.valid = false;
  p.slots = 0;
  p.w0 = p.w1 = p.w2 = S(0);
  if (s.rank == 4 || s.rank <= 0) return p;
  if (s.rank == 1) {
    p.n = 1;
    p.valid = true;
    p.slots = ord1(slotOf(s, 0));
    return p;
  }
  if (s.rank == 2) {
    const int ka = slotOf(s, 0), kb = slotOf(s, 1);
    const V3<S> s1 = st.vtx(ka), s2 = st.vtx(kb);
    const V3<S> s1_to_s2 = s2 - s1;
    const S sq_len = sqnorm(s1_to_s2);
    p.n = 1;
    if (sq_len <= S(0)) {
      p.valid = true;
      p.slots = ord1(ka);
      return p;
    }
    const S t = -dot(s1, s1_to_s2);
    const S w2 = t / sq_len;
    if (w2 > 1 + bary_tol) {
      p.slots = ord1(kb);  // written but reported invalid
      return p;
    } else if (w2 < -bary_tol) {
      p.slots = ord1(ka);
      return p;
    }
    p.n = 2;
    p.weighted = true;
    p.valid = true;
    p.slots = ord2(ka, kb);
    p.w0 = S(1.0) - w2;
    p.w1 = w2;
    return p;
  }
  const int ka = slotOf(s, 0), kb = slotOf(s, 1), kc = slotOf(s, 2);
  const V3<S> s1 = st.vtx(ka), s2 = st.vtx(kb), s3 = st.vtx(kc);
  const V3<S> s1_to_s2 = s2 - s1;
  const V3<S> s1_to_s3 = s3 - s1;
  const V3<S> n = cross(s1_to_s2, s1_to_s3);
  const S area_sq = sqnorm(n);
  if (area_sq <= S(0)) return p;
  const S d = dot(s1, n);
  const V3<S> o_proj = n * (d / area_sq);
  const S area = fsqrt(area_sq);
  const S w2 = norm(cross(s1_to_s3, s1 - o_proj)) / area;
  const S w3 = norm(cross(s1_to_s2, s1 - o_proj)) / area;
  const S w1 = S(1.0) - w2 - w3;
  if (w1 < -bary_tol || w2 < -bary_tol || w3 < -bary_tol) return p;
  p.n = 3;
  p.weighted = true;
  p.valid = true;
  p.slots = ord3(ka, kb, kc);
  p.w0 = w1;
  p.w1 = w2;
  p.w2 = w3;
  return p;
}

// gjk_distance.hpp:11-106.  The simplex holds the one separating vertex.
template <typename S, typename MD>
FCLB_DI bool gjkMinDistance(const MD& shape, SlotStore<S>& st, Simp& s, S tol, int max_iter, V3<S>& p0, V3<S>& p1,
                            uint32_t* n_support) {
  if (s.rank != 1) return false;
  V3<S> cur;
  if (minDistUpdate(st, s, cur, tol) != 1) return false;
  S book = norm(cur);
  if (book <= tol) return extractSeparationPoint(shape, st, s, p0, p1);
  V3<S> next_dir = (-cur) / book;
  const S tol_sq = tol * tol;
  int it = 0;
  while (it < max_iter) {
    it += 1;
    const V3<S> nv = shape.support(next_dir);
    if (n_support) *n_support += 2;
    const S delta = dot(next_dir, nv - cur);
    if (delta < tol) return extractSeparationPoint(shape, st, s, p0, p1);
    #pragma unroll 1
    for (int j = 0; j < s.rank; j++) {
      if (sqnorm(st.vtx(slotOf(s, j)) - nv) < tol_sq) return extractSeparationPoint(shape, st, s, p0, p1);
    }
    addVertex(st, s, nv, next_dir);
    const int us = minDistUpdate(st, s, cur, tol);
    if (us == 0) {
      return extractSeparationPoint(shape, st, s, p0, p1);
    } else if (us == 1) {
      const S nd = norm(cur);
      const S improvement = book - nd;
      if (improvement < tol) {
        return extractSeparationPoint(shape, st, s, p0, p1);
      } else if (nd < tol) {
        return extractSeparationPoint(shape, st, s, p0, p1);
      }
      book = nd;
      next_dir = (-cur) / book;
    } else {
      return false;
    }
  }
  return false;
}

template <typename S>
struct GjkDistOut {
  bool valid;  // is_separation_point_valid
  V3<S> p0, p1;
};

// gjk.hpp:11-146.  `guess` is the initial search direction (callers pass
// -solver_guess = (-1,0,0), gjk_solver-inl.h:77,103).  When dist != nullptr the
// separated branch continues into the distance loop.
template <typename S, typename MD>
FCLB_DI int gjkEvaluate(const MD& shape, SlotStore<S>& st, Simp& s, V3<S> guess, S tol, int max_iter,
                        GjkDistOut<S>* dist, uint32_t* n_support) {
  V3<S> direction = guess;
  if (sqnorm(direction) <= S(0)) direction = mk<S>(S(1), S(0), S(0));
  direction = normalized(direction);

  V3<S> v = shape.support(direction);
  if (n_support) *n_support += 2;
  s.rank = -1;
  s.ord = 0;
  addVertex(st, s, v, direction);
  const S tol_sq = tol * tol;

  bool separated = false;
  int status = GJK_ITER_LIMIT;
  if (sqnorm(v) <= tol_sq) {
    return GJK_INTERSECT;
  } else if (dot(v, direction) < 0) {
    separated = true;
  } else {
    direction = direction * S(-1);
    int it = 0;
    while (it < max_iter) {
      it += 1;
      v = shape.support(direction);
      if (n_support) *n_support += 2;
      if (dot(v, direction) < 0) {
        separated = true;
        break;
      }
      bool dup = false;
      #pragma unroll 1
      for (int j = 0; j < s.rank; j++) {
        if (sqnorm(st.vtx(slotOf(s, j)) - v) < tol_sq) dup = true;
      }
      if (dup) {
        status = GJK_NO_PROGRESS;
        break;
      }
      if (sqnorm(v) <= tol_sq) {
        addVertex(st, s, v, direction);
        return GJK_INTERSECT;
      }
      addVertex(st, s, v, direction);
      const int ps = simplexProjection(st, s, direction, tol);
      if (ps == PROJ_FAILED) return GJK_FAILED;
      if (ps == PROJ_INTERSECT) return GJK_INTERSECT;
      if (ps == PROJ_ZERO_VOLUME) {
        status = GJK_NO_PROGRESS;
        break;
      }
    }
  }
  if (!separated) return status;  // ConvergeNoProgress / IterationLimit

  if (dist != nullptr) {
    // process_separated_vertex (gjk.hpp:22-52)
    s.rank = -1;
    s.ord = 0;
    addVertex(st, s, v, direction);
    dist->valid = gjkMinDistance(shape, st, s, tol, max_iter, dist->p0, dist->p1, n_support);
  }
  return GJK_SEPARATED;
}

}  // namespace fclb
