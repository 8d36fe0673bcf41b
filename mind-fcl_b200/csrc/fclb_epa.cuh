// fclb_epa.cuh -- Expanding Polytope Algorithm, ONE WARP PER QUERY, polytope in
// shared memory.
//
// Behavioural contract (results, not structure):
//   include/fcl/cvx_collide/epa.hpp                 Evaluate :240, main loop :137-237,
//                                                   findNextSupportDirection :11-113,
//                                                   checkTerminateCondition :269-298,
//                                                   assignPenetrationPair* :341-497
//   include/fcl/cvx_collide/epa_simplex2polytope.hpp simplexToPolytope{,2,3,4} :46-380
//   include/fcl/cvx_collide/epa_polytope.hpp        AddNew* :185-290, ComputeMinDistanceToOrigin :295-345,
//                                                   ComputeFaceNormalPointingOutward :348-410
//   include/fcl/cvx_collide/epa_polytope_expand.hpp ExpandPolytope :33-91, computeVisiblePatch :121-180,
//                                                   removeAccordingToVisibility :244-281
//   include/fcl/cvx_collide/epa_polytope_utils.h    pointTo{Segment,Triangle}SquaredDistance, pointTo{Line,Plane}Distance
//
// The reference keeps three pointer-linked, push-front lists in heap-allocated
// pools (3 std::vector resizes per query) and walks them serially.  What its
// RESULTS depend on is only
//   (1) the selection order of ComputeMinDistanceToOrigin: strict "<" while
//       scanning vertices, then edges, then faces, each newest-first;
//   (2) the order in which ExpandPolytope creates new edges/faces (it walks the
//       edge list newest-first), because that fixes (1) for later iterations.
// Pool slot identity, free-list order and the flood-fill visiting order are
// unobservable.  So here every element carries an insertion SEQUENCE NUMBER,
// the pools are flat index arrays (u16 links) in shared memory, and
//   * the O(V+E+F) nearest-feature scan is a warp-parallel arg-min with the key
//     (distance, class, newest-first) -- exactly the sequential tie-break;
//   * face visibility is evaluated for all faces at once and the visible patch
//     is grown by warp-parallel label propagation (same set as the DFS);
//   * the new cone of faces is built from the border edges ranked by sequence
//     number, with their distance records computed one element per lane.
// Control flow between those sections is warp-uniform (all lanes execute the
// same scalar code on broadcast shared-memory reads).
#pragma once
#include "fclb_gjk.cuh"

// Out-of-line on purpose: the EPA kernel calls these from dozens of sites; fully
// inlined it was 74k SASS lines (1.2 MB) and instruction-fetch bound (ncu:
// stall_no_inst 40 % of samples).  See DESIGN.md 4.4.
#define FCLB_DN __device__ __noinline__

namespace fclb {

enum EpaStatus : int { EPA_FAILED = 0, EPA_OK = 1, EPA_TOUCHING = 2, EPA_ITER_LIMIT = 3, EPA_MALLOC_FAILED = 4 };  // epa.h:14

constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kNil = 0xffff;

template <typename S>
struct MinDist {  // MinDistanceToSimplex, epa_polytope_utils.h:15-20
  bool in_simplex;
  S dist_sq;
  V3<S> witness;
};

// epa_polytope_utils.h:22-52
template <typename S>
FCLB_DN MinDist<S> pointToSegment(V3<S> p, V3<S> x0, V3<S> x1) {
  MinDist<S> r;
  const V3<S> d = x1 - x0;
  const V3<S> a = x0 - p;
  const S len_sq = sqnorm(d);
  const S t = S(-1.0) * dot(a, d) / len_sq;
  if (t <= S(0)) {
    r.witness = x0;
    r.in_simplex = false;
    r.dist_sq = sqnorm(x0 - p);
  } else if (t >= S(1.0)) {
    r.witness = x1;
    r.in_simplex = false;
    r.dist_sq = sqnorm(x1 - p);
  } else {
    const V3<S> w = x0 + t * d;
    r.witness = w;
    r.in_simplex = true;
    r.dist_sq = sqnorm(w);
  }
  return r;
}
// epa_polytope_utils.h:64-76
template <typename S>
FCLB_DN S pointToLineDistance(V3<S> p, V3<S> a, V3<S> b) {
  const V3<S> ab = b - a;
  const S len = norm(ab);
  if (len <= S(0)) return norm(p - a);
  return norm(cross(p - a, ab)) / len;
}
// epa_polytope_utils.h:78-93
template <typename S>
FCLB_DN S pointToPlaneDistance(V3<S> p, V3<S> a, V3<S> b, V3<S> c) {
  const V3<S> n = cross(a - b, b - c);
  const S len = norm(n);
  if (len <= S(0)) return pointToLineDistance(p, a, b);
  const V3<S> p_to_a = a - p;
  const V3<S> un = n / len;
  return fabs_(dot(un, p_to_a));
}
// epa_polytope_utils.h:95-165
template <typename S>
FCLB_DN MinDist<S> pointToTriangle(V3<S> p, V3<S> a, V3<S> b, V3<S> c) {
  MinDist<S> r;
  const V3<S> dl0 = a - b, dl1 = b - c, dl2 = c - a;
  const V3<S> n = cross(dl0, dl1);
  const S n_sq = sqnorm(n);
  const S area = fsqrt(n_sq);
  bool in_tri = false;
  if (!(fabs_(area) <= S(1e-16))) {
    const S d = dot(a - p, n);
    const V3<S> p_to_proj = n * (d / n_sq);
    const V3<S> proj = p + p_to_proj;
    const S w0 = norm(cross(dl1, b - proj)) / area;
    const S w1 = norm(cross(dl2, c - proj)) / area;
    const S w2 = S(1.0) - w0 - w1;
    in_tri = true;
    if (w0 < S(0) || w0 > S(1)) {
      in_tri = false;
    } else if (w1 < S(0) || w1 > S(1)) {
      in_tri = false;
    } else if (w2 < S(0) || w2 > S(1)) {
      in_tri = false;
    }
    const S w2_check = norm(cross(dl0, a - proj)) / area;
    if (fabs_(w2_check - w2) > S(1e-3)) in_tri = false;
    if (in_tri) {
      r.in_simplex = true;
      r.dist_sq = sqnorm(p_to_proj);
      r.witness = a * w0 + b * w1 + c * w2;
      return r;
    }
  }
  // process_distance_in_sub_simplex: best of the three edges, first minimum wins
  MinDist<S> best = pointToSegment(p, a, b);
  {
    const MinDist<S> e1 = pointToSegment(p, b, c);
    if (e1.dist_sq < best.dist_sq) best = e1;
    const MinDist<S> e2 = pointToSegment(p, c, a);
    if (e2.dist_sq < best.dist_sq) best = e2;
  }
  r.in_simplex = false;
  r.dist_sq = best.dist_sq;
  r.witness = best.witness;
  return r;
}

// Out-of-line support evaluation: the Minkowski difference is read through a
// pointer (it lives in local memory, L1-resident), everything else stays in registers.
template <typename S, typename MD>
FCLB_DN void epaSupportBoth(const MD* shape, V3<S> d, S* out) {
  const V3<S> s0 = shape->support0(d);
  const V3<S> s1 = shape->support1(-d);
  out[0] = s0.x; out[1] = s0.y; out[2] = s0.z;
  out[3] = s1.x; out[4] = s1.y; out[5] = s1.z;
}

// ---------------------------------------------------------------------------
// Shared-memory polytope of one warp.  Capacities as the reference:
// faces = max_faces, edges = vertices = floor(1.51 * max_faces) (epa_polytope.hpp:88-96).
template <typename S>
struct PolyStore {
  int vcap, ecap, fcap;
  // vertices
  S *vx, *vy, *vz, *dx, *dy, *dz, *vd;
  uint16_t *v_seq, *v_newedge;
  uint8_t *v_alive, *v_rm;
  // edges
  uint16_t *e_v0, *e_v1, *e_f0, *e_f1, *e_seq;
  S* e_d;
  uint8_t *e_alive, *e_in, *e_vis;  // vis: 0 unknown, 1 border, 2 internal
  // faces
  uint16_t *f_e0, *f_e1, *f_e2, *f_a, *f_b, *f_c, *f_seq;
  S* f_d;
  uint8_t *f_alive, *f_in, *f_vis;  // vis: 0 unknown, 1 visible(in patch), 2 hidden, 3 "outside" but not reached

  // scratch of the parallel cone construction (expand): three lists of edge slots / vertices
  uint16_t *e_tmp0, *e_tmp1, *e_tmp2;

  // vertices: the reference's pool has floor(1.51 F) slots, but a closed triangulated polytope with E <= 1.51 F
  // edges has V = (E + 6) / 3 <= F / 2 + 2 vertices.  The first-tier pool (compact_v) is sized for that; running out
  // of it reports MallocFailed, which defers the query to the full-capacity tier like any other exhausted pool.
  static __host__ __device__ int vertexCap(int max_faces, bool compact_v) {
    const int full = int(1.51 * double(max_faces));
    const int small = max_faces / 2 + 4;
    return (compact_v && small < full) ? small : full;
  }
  static __host__ __device__ size_t bytes(int max_faces, bool compact_v) {
    const size_t f = size_t(max_faces), v = size_t(vertexCap(max_faces, compact_v)), e = size_t(1.51 * double(max_faces));
    size_t b = 0;
    b += 7 * v * sizeof(S) + e * sizeof(S) + f * sizeof(S);
    b += (2 * v + 8 * e + 7 * f) * sizeof(uint16_t);
    b += 2 * v + 3 * e + 3 * f;
    return (b + 15) / 16 * 16;
  }
  FCLB_DI void bind(unsigned char* base, int max_faces, bool compact_v) {
    fcap = max_faces;
    vcap = vertexCap(max_faces, compact_v);
    ecap = int(1.51 * double(max_faces));
    S* ps = reinterpret_cast<S*>(base);
    vx = ps; ps += vcap;
    vy = ps; ps += vcap;
    vz = ps; ps += vcap;
    dx = ps; ps += vcap;
    dy = ps; ps += vcap;
    dz = ps; ps += vcap;
    vd = ps; ps += vcap;
    e_d = ps; ps += ecap;
    f_d = ps; ps += fcap;
    uint16_t* p16 = reinterpret_cast<uint16_t*>(ps);
    v_seq = p16; p16 += vcap;
    v_newedge = p16; p16 += vcap;
    e_v0 = p16; p16 += ecap;
    e_v1 = p16; p16 += ecap;
    e_f0 = p16; p16 += ecap;
    e_f1 = p16; p16 += ecap;
    e_seq = p16; p16 += ecap;
    e_tmp0 = p16; p16 += ecap;
    e_tmp1 = p16; p16 += ecap;
    e_tmp2 = p16; p16 += ecap;
    f_e0 = p16; p16 += fcap;
    f_e1 = p16; p16 += fcap;
    f_e2 = p16; p16 += fcap;
    f_a = p16; p16 += fcap;
    f_b = p16; p16 += fcap;
    f_c = p16; p16 += fcap;
    f_seq = p16; p16 += fcap;
    uint8_t* p8 = reinterpret_cast<uint8_t*>(p16);
    v_alive = p8; p8 += vcap;
    v_rm = p8; p8 += vcap;
    e_alive = p8; p8 += ecap;
    e_in = p8; p8 += ecap;
    e_vis = p8; p8 += ecap;
    f_alive = p8; p8 += fcap;
    f_in = p8; p8 += fcap;
    f_vis = p8; p8 += fcap;
  }
  FCLB_DI V3<S> vloc(int i) const { return mk<S>(vx[i], vy[i], vz[i]); }
  FCLB_DI V3<S> vdir(int i) const { return mk<S>(dx[i], dy[i], dz[i]); }
};

struct Feature {  // nearest feature: cls 0 vertex, 1 edge, 2 face; idx = slot; -1 none
  int cls;
  int idx;
};

// T = lanes cooperating on one query (a "tile": 8, 16 or 32 consecutive lanes of a
// warp).  Small polytopes (boxes, hulls: < 30 faces at the 99th percentile) leave
// most of a 32-lane warp idle in the parallel sections and replicate the uniform
// code 32x, so the first pass runs FOUR queries per warp (T = 8) with a small pool;
// queries that outgrow it are re-run by a T = 32 pass with the reference's full
// capacity (fclb_epa_launch.cuh).
template <typename S, typename MD, int T>
struct EpaWarp {
  PolyStore<S> P;
  const MD& shape;
  const int lane;      // lane within the tile
  const unsigned tmask;  // this tile's lanes within the warp
  const int tshift;
  int v_hw, e_hw, f_hw;     // high-water marks (slots ever used)
  int v_n, e_n, f_n;        // alive counts
  int v_sq, e_sq, f_sq;     // next sequence numbers
  uint32_t* n_support;

  FCLB_DI EpaWarp(const MD& sh, unsigned char* smem, int max_faces, bool compact_v, int warp_lane, uint32_t* ns)
      : shape(sh),
        lane(warp_lane % T),
        tmask(T == 32 ? 0xffffffffu : (((1u << T) - 1u) << ((warp_lane / T) * T))),
        tshift((warp_lane / T) * T),
        n_support(ns) {
    P.bind(smem, max_faces, compact_v);
  }
  // tile-scoped collectives
  FCLB_DI void sync() const { __syncwarp(tmask); }
  FCLB_DI unsigned ballot(bool p) const { return (__ballot_sync(tmask, p) >> tshift) & (T == 32 ? 0xffffffffu : ((1u << T) - 1u)); }
  FCLB_DI bool any(bool p) const { return __any_sync(tmask, p) != 0; }
  template <typename V>
  FCLB_DI V shflXor(V v, int off) const {
    return __shfl_xor_sync(tmask, v, off, T);
  }
  template <typename V>
  FCLB_DI V shfl(V v, int src) const {
    return __shfl_sync(tmask, v, src, T);
  }

  // the ONE place the shape support mappings are instantiated in this kernel
  FCLB_DI void supportBoth(const V3<S>& d, V3<S>& s0, V3<S>& s1) const {
    if (n_support) *n_support += 2;
    S out[6];
    epaSupportBoth<S, MD>(&shape, d, out);
    s0 = mk<S>(out[0], out[1], out[2]);
    s1 = mk<S>(out[3], out[4], out[5]);
  }
  FCLB_DI V3<S> support(const V3<S>& d) const {
    V3<S> s0, s1;
    supportBoth(d, s0, s1);
    return s0 - s1;
  }

  // Polytope::Reset (epa_polytope.hpp:78-103)
  FCLB_DI void reset() {
    #pragma unroll 1
    for (int i = lane; i < P.vcap; i += T) P.v_alive[i] = 0;
    #pragma unroll 1
    for (int i = lane; i < P.ecap; i += T) P.e_alive[i] = 0;
    #pragma unroll 1
    for (int i = lane; i < P.fcap; i += T) P.f_alive[i] = 0;
    v_hw = e_hw = f_hw = 0;
    v_n = e_n = f_n = 0;
    v_sq = e_sq = f_sq = 0;
    sync();
  }

  // a dead slot of a pool (uniform result), recycled ones first (see gatherFree); -1 if the pool is full
  FCLB_DI int allocSlot(const uint8_t* alive, int cap, int& hw, int n_alive) const {
    if (n_alive < hw) {
      #pragma unroll 1
      for (int base = 0; base < hw; base += T) {
        const int i = base + lane;
        const unsigned m = ballot(i < hw && !alive[i]);
        if (m) return base + __ffs(m) - 1;
      }
    }
    if (hw < cap) return hw++;
    return -1;
  }

  // `need` free slots of a pool into out[0 .. need): dead slots below the high-water mark first (ascending), then
  // fresh ones -- slot identity is unobservable, and recycling first keeps the high-water marks (the trip counts of
  // every per-element loop) at the live size of the polytope instead of the pool capacity.  The caller has checked
  // that the pool has that many free slots.
  FCLB_DI void gatherFree(const uint8_t* alive, int cap, int& hw, int n_alive, int need, uint16_t* out) const {
    int got = 0;
    const unsigned lt = (1u << lane) - 1u;
    if (n_alive < hw) {
      #pragma unroll 1
      for (int base = 0; base < hw && got < need; base += T) {
        const int i = base + lane;
        const bool pred = i < hw && !alive[i];
        const unsigned mask = ballot(pred);
        const int pos = got + __popc(mask & lt);
        if (pred && pos < need) out[pos] = uint16_t(i);
        got += __popc(mask);
      }
      if (got > need) got = need;
    }
    const int fresh = need - got;
    #pragma unroll 1
    for (int j = lane; j < fresh; j += T) out[got + j] = uint16_t(hw + j);
    hw += fresh;
    (void)cap;
  }

  // AddNewVertex (epa_polytope.hpp:185-209)
  FCLB_DI int addVertex(const V3<S>& v, const V3<S>& d) {
    const int s = allocSlot(P.v_alive, P.vcap, v_hw, v_n);
    if (s < 0) return -1;
    if (lane == 0) {
      P.vx[s] = v.x; P.vy[s] = v.y; P.vz[s] = v.z;
      P.dx[s] = d.x; P.dy[s] = d.y; P.dz[s] = d.z;
      P.vd[s] = sqnorm(v);
      P.v_seq[s] = uint16_t(v_sq);
      P.v_alive[s] = 1;
    }
    v_sq++;
    v_n++;
    sync();
    return s;
  }
  // topology part of AddNewEdge (:212-240); the distance record is filled by fillEdge
  FCLB_DI int addEdgeTopo(int v1, int v2) {
    if (v1 < 0 || v2 < 0) return -1;
    const int s = allocSlot(P.e_alive, P.ecap, e_hw, e_n);
    if (s < 0) return -1;
    if (lane == 0) {
      P.e_v0[s] = uint16_t(v1);
      P.e_v1[s] = uint16_t(v2);
      P.e_f0[s] = kNil;
      P.e_f1[s] = kNil;
      P.e_seq[s] = uint16_t(e_sq);
      P.e_alive[s] = 1;
      P.e_vis[s] = 0;
    }
    e_sq++;
    e_n++;
    sync();
    return s;
  }
  FCLB_DI void fillEdge(int s) {  // any single lane
    const MinDist<S> md = pointToSegment(zero3<S>(), P.vloc(P.e_v0[s]), P.vloc(P.e_v1[s]));
    P.e_d[s] = md.dist_sq;
    P.e_in[s] = md.in_simplex ? 1 : 0;
  }
  // topology part of AddNewFace (:243-290). Returns slot, or -1 (malloc / "wrong edge").
  FCLB_DI int addFaceTopo(int e1, int e2, int e3) {
    if (e1 < 0 || e2 < 0 || e3 < 0) return -1;
    const int s = allocSlot(P.f_alive, P.fcap, f_hw, f_n);
    if (s < 0) return -1;
    bool ok = true;
    if (lane == 0) {
      const int a = P.e_v0[e1], b = P.e_v1[e1];
      const int c = (P.e_v0[e2] != a && P.e_v0[e2] != b) ? P.e_v0[e2] : P.e_v1[e2];
      P.f_e0[s] = uint16_t(e1);
      P.f_e1[s] = uint16_t(e2);
      P.f_e2[s] = uint16_t(e3);
      P.f_a[s] = uint16_t(a);
      P.f_b[s] = uint16_t(b);
      P.f_c[s] = uint16_t(c);
      P.f_seq[s] = uint16_t(f_sq);
      P.f_alive[s] = 1;
      P.f_vis[s] = 0;
      const int es[3] = {e1, e2, e3};
      #pragma unroll 1
      for (int i = 0; i < 3 && ok; i++) {
        if (P.e_f0[es[i]] == kNil) {
          P.e_f0[es[i]] = uint16_t(s);
        } else if (P.e_f1[es[i]] != kNil) {
          ok = false;  // wrong connection at degeneration (:279-283)
        } else {
          P.e_f1[es[i]] = uint16_t(s);
        }
      }
    }
    ok = shfl(ok ? 1 : 0, 0) != 0;
    f_sq++;
    f_n++;
    sync();
    return ok ? s : -1;
  }
  FCLB_DI void fillFace(int s) {  // any single lane
    const MinDist<S> md = pointToTriangle(zero3<S>(), P.vloc(P.f_a[s]), P.vloc(P.f_b[s]), P.vloc(P.f_c[s]));
    P.f_d[s] = md.dist_sq;
    P.f_in[s] = md.in_simplex ? 1 : 0;
  }

  // The two initial polytopes (formNewTetrahedronPolytope, epa_simplex2polytope.hpp:177-215, and the hexahedron
  // of simplexToPolytope2, :300-380) are fixed sequences of AddNewVertex / AddNewEdge / AddNewFace on EMPTY pools,
  // so the k-th vertex / edge / face lands in slot k with sequence number k, an edge's faces_of_edge[0 / 1] are the
  // first / second face of the sequence that names it, and a face's vertices follow from its first two edges
  // (epa_polytope.hpp:243-290).  The tiles write the whole structure in one parallel pass from these tables
  // (4 bits per entry) instead of 14 / 26 serial allocations.
  //   ev0 / ev1: vertices of edge k;  fe0 / fe1 / fe2: edges of face k
  template <int NV>
  FCLB_DI bool buildFromTables(const V3<S> (&v)[NV], const V3<S> (&d)[NV], int ne, int nf, unsigned long long ev0,
                               unsigned long long ev1, unsigned long long fe0, unsigned long long fe1,
                               unsigned long long fe2) {
    // a pool too small for the sequence fails one of the reference's allocations => "Failed"
    if (P.vcap < NV || P.ecap < ne || P.fcap < nf) return false;
    auto nib = [](unsigned long long t, int k) { return int((t >> (4 * k)) & 0xfull); };
    #pragma unroll 1
    for (int k = lane; k < NV; k += T) {
      V3<S> vv = v[0], dd = d[0];
#pragma unroll
      for (int j = 1; j < NV; j++)
        if (j == k) {
          vv = v[j];
          dd = d[j];
        }
      P.vx[k] = vv.x; P.vy[k] = vv.y; P.vz[k] = vv.z;
      P.dx[k] = dd.x; P.dy[k] = dd.y; P.dz[k] = dd.z;
      P.vd[k] = sqnorm(vv);
      P.v_seq[k] = uint16_t(k);
      P.v_alive[k] = 1;
    }
    #pragma unroll 1
    for (int k = lane; k < ne; k += T) {
      P.e_v0[k] = uint16_t(nib(ev0, k));
      P.e_v1[k] = uint16_t(nib(ev1, k));
      int f0 = kNil, f1 = kNil;
      #pragma unroll 1
      for (int f = nf - 1; f >= 0; f--)
        if (nib(fe0, f) == k || nib(fe1, f) == k || nib(fe2, f) == k) {
          f1 = f0;
          f0 = f;
        }
      P.e_f0[k] = uint16_t(f0);
      P.e_f1[k] = uint16_t(f1);
      P.e_seq[k] = uint16_t(k);
      P.e_alive[k] = 1;
      P.e_vis[k] = 0;
    }
    #pragma unroll 1
    for (int k = lane; k < nf; k += T) {
      const int e1 = nib(fe0, k), e2 = nib(fe1, k), e3 = nib(fe2, k);
      const int a = nib(ev0, e1), b = nib(ev1, e1);
      const int c = (nib(ev0, e2) != a && nib(ev0, e2) != b) ? nib(ev0, e2) : nib(ev1, e2);
      P.f_e0[k] = uint16_t(e1);
      P.f_e1[k] = uint16_t(e2);
      P.f_e2[k] = uint16_t(e3);
      P.f_a[k] = uint16_t(a);
      P.f_b[k] = uint16_t(b);
      P.f_c[k] = uint16_t(c);
      P.f_seq[k] = uint16_t(k);
      P.f_alive[k] = 1;
      P.f_vis[k] = 0;
    }
    v_hw = v_n = v_sq = NV;
    e_hw = e_n = e_sq = ne;
    f_hw = f_n = f_sq = nf;
    sync();
    #pragma unroll 1
    for (int k = lane; k < ne + nf; k += T) {
      if (k < ne)
        fillEdge(k);
      else
        fillFace(k - ne);
    }
    sync();
    return true;
  }
  // formNewTetrahedronPolytope (epa_simplex2polytope.hpp:177-215): edges (0,1) (1,2) (2,0) (3,0) (3,1) (3,2),
  // faces (e0,e1,e2) (e3,e4,e0) (e4,e5,e1) (e5,e3,e2)
  FCLB_DI bool formTetrahedron(const V3<S> (&v)[4], const V3<S> (&d)[4]) {
    return buildFromTables<4>(v, d, 6, 4, 0x333210ull, 0x210021ull, 0x5430ull, 0x3541ull, 0x2102ull);
  }

  // extractTouchingPoint (epa_simplex2polytope.hpp:11-43)
  FCLB_DI void touchingPoint(const V3<S>& dir, V3<S>& p0, V3<S>& p1) {
    V3<S> a0, a1;
    supportBoth(dir, a0, a1);
    const V3<S> mid = (a0 + a1) / S(2);
    p0 = mid;
    p1 = mid;
  }

  // simplexToPolytope3 (epa_simplex2polytope.hpp:135-175): 1 Touching, 2 Failed, 3 = the tetrahedron to form is
  // (v[0..2], fourth vertex written to v[3] / d[3])
  FCLB_DI int simplexToPolytope3(V3<S> (&v)[4], V3<S> (&d)[4], S thr, V3<S>& p0, V3<S>& p1) {
    const V3<S> a = v[0], b = v[1], c = v[2];
    const V3<S> ab = b - a, ac = c - a;
    V3<S> n = cross(ab, ac);
    if (sqnorm(n) <= S(0)) return 2;
    n = normalized(n);
    const V3<S> dir0 = n;
    const V3<S> d0 = support(dir0);
    const S d0_to_plane = pointToPlaneDistance(d0, a, b, c);
    const V3<S> dir1 = -n;
    const V3<S> d1 = support(dir1);
    const S d1_to_plane = pointToPlaneDistance(d1, a, b, c);
    if (d0_to_plane <= thr) {
      touchingPoint(dir0, p0, p1);
      return 1;
    } else if (d1_to_plane <= thr) {
      touchingPoint(dir1, p0, p1);
      return 1;
    }
    if (d0_to_plane > d1_to_plane) {
      v[3] = d0;
      d[3] = dir0;
    } else {
      v[3] = d1;
      d[3] = dir1;
    }
    return 3;
  }

  // simplexToPolytope2 (epa_simplex2polytope.hpp:218-380)
  FCLB_DI int simplexToPolytope2(const V3<S>& a, const V3<S>& da, const V3<S>& b, const V3<S>& db, S thr, V3<S>& p0,
                                 V3<S>& p1) {
    const S thr_sq = thr * thr;
    const V3<S> a_to_b = b - a;
    if (sqnorm(a_to_b) < thr_sq) return 2;
    const V3<S> u = normalized(a_to_b);
    V3<S> d_init = mk<S>(u.y, -u.x, S(0));
    bool valid = sqnorm(d_init) > thr_sq;
    if (!valid) {
      d_init = mk<S>(u.z, S(0), -u.x);
      valid = sqnorm(d_init) > thr_sq;
    }
    if (!valid) {
      d_init = mk<S>(S(0), u.z, -u.y);
      valid = sqnorm(d_init) > thr_sq;
    }
    if (!valid) return 2;
    d_init = normalized(d_init);
    V3<S> d_cur = d_init, d_not = zero3<S>(), v_not = zero3<S>();
    bool found = false;
    for (int i = 0; i < 72; i++) {
      const V3<S> v_cur = support(d_cur);
      const S dist = pointToLineDistance(v_cur, a, b);
      if (dist > thr) {
        v_not = v_cur;
        d_not = d_cur;
        found = true;
        break;
      }
      const S pi_value = S(3.145926);  // sic (epa_simplex2polytope.hpp:270)
      const S delta_len = S(2.0) * pi_value / S(72);
      const V3<S> delta = cross(u, v_cur);
      d_cur = normalized(d_cur + delta_len * delta);
    }
    if (!found) {
      touchingPoint(d_init, p0, p1);
      return 1;
    }
    const V3<S> v0 = v_not, dir0 = d_not;
    const V3<S> dir1 = -dir0;
    const V3<S> v1 = support(dir1);
    if (pointToLineDistance(v1, a, b) < thr) {
      touchingPoint(dir1, p0, p1);
      return 1;
    }
    const V3<S> nrm = cross(v0 - a, v1 - a);
    if (sqnorm(nrm) < thr_sq) return 2;
    const V3<S> dir2 = normalized(nrm);
    const V3<S> v2 = support(dir2);
    if (pointToLineDistance(v2, a, b) < thr) {
      touchingPoint(dir2, p0, p1);
      return 1;
    }
    const V3<S> dir3 = -dir2;
    const V3<S> v3 = support(dir3);
    if (pointToLineDistance(v3, a, b) < thr) {
      touchingPoint(dir3, p0, p1);
      return 1;
    }
    // vertices a, v0, b, v1, v2, v3; edges (0,1) (1,2) (2,3) (3,0) (4,0) (4,1) (4,2) (4,3) (5,0) (5,1) (5,2) (5,3);
    // faces (e4,e5,e0) (e5,e6,e1) (e6,e7,e2) (e7,e4,e3) (e8,e9,e0) (e9,e10,e1) (e10,e11,e2) (e11,e8,e3)  (:300-380)
    const V3<S> pv[6] = {a, v0, b, v1, v2, v3};
    const V3<S> pd[6] = {da, dir0, db, dir1, dir2, dir3};
    return buildFromTables<6>(pv, pd, 12, 8, 0x555544443210ull, 0x321032100321ull, 0xba987654ull, 0x8ba94765ull,
                              0x32103210ull)
               ? 0
               : 2;
  }

  // simplexToPolytope (epa_simplex2polytope.hpp:46-133)
  FCLB_DI int simplexToPolytope(const SlotStore<S>& st, const Simp& sx, S thr, V3<S>& p0, V3<S>& p1) {
    const S thr_sq = thr * thr;
    V3<S> v[4], d[4];
    for (int i = 0; i < 4; i++) {
      v[i] = zero3<S>();
      d[i] = zero3<S>();
    }
    #pragma unroll 1
    for (int i = 0; i < sx.rank; i++) {
      v[i] = st.vtx(slotOf(sx, i));
      d[i] = st.dir(slotOf(sx, i));
    }
    #pragma unroll 1
    for (int i = 0; i < sx.rank; i++) {
      if (sqnorm(v[i]) <= thr_sq) {
        touchingPoint(d[i], p0, p1);
        return 1;
      }
    }
    int mode = sx.rank;  // 4: tetrahedron in v/d, 3: triangle in v[0..2], 2: segment, 1: point
    if (sx.rank == 4) {  // simplexToPolytope4 :96-133
      // the four faces are tried in the order abc, acd, abd, bcd; the first one
      // whose plane passes (almost) through the origin reduces to the triangle case
      const V3<S> o = zero3<S>();
      int sel = -1;
      const int tri[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 1, 3}, {1, 2, 3}};
#pragma unroll 1
      for (int k = 0; k < 4 && sel < 0; k++)
        if (pointToPlaneDistance(o, v[tri[k][0]], v[tri[k][1]], v[tri[k][2]]) < thr) sel = k;
      if (sel >= 0) {
        const int i0 = tri[sel][0], i1 = tri[sel][1], i2 = tri[sel][2];
        const V3<S> ta = v[i0], tb = v[i1], tc = v[i2], tda = d[i0], tdb = d[i1], tdc = d[i2];
        v[0] = ta; v[1] = tb; v[2] = tc;
        d[0] = tda; d[1] = tdb; d[2] = tdc;
        mode = 3;
      }
    }
    if (mode == 3) {
      const int r = simplexToPolytope3(v, d, thr, p0, p1);
      if (r != 3) return r;
      mode = 4;
    }
    if (mode == 4) return formTetrahedron(v, d) ? 0 : 2;
    if (mode == 2) return simplexToPolytope2(v[0], d[0], v[1], d[1], thr, p0, p1);
    supportBoth(d[0], p0, p1);
    return 1;
  }

  // ---- nearest feature ------------------------------------------------------
  // ComputeMinDistanceToOrigin (epa_polytope.hpp:295-345): sequential scan with a
  // strict "<" over vertices, edges, faces (each newest first).  Equivalent key:
  // smaller distance, then smaller class, then larger sequence number.
  static FCLB_DI bool better(S d, int cls, int seq, S bd, int bcls, int bseq) {
    if (d < bd) return true;
    if (d > bd) return false;
    if (cls != bcls) return cls < bcls;
    return seq > bseq;
  }
  FCLB_DI Feature nearest(bool exclude_vertex) const {
    S bd = S(INFINITY);
    int bcls = 3, bseq = -1, bidx = -1;
    if (!exclude_vertex) {
      #pragma unroll 1
      for (int i = lane; i < v_hw; i += T) {
        if (!P.v_alive[i]) continue;
        const S d = P.vd[i];
        if (d < S(INFINITY) && better(d, 0, P.v_seq[i], bd, bcls, bseq)) {
          bd = d; bcls = 0; bseq = P.v_seq[i]; bidx = i;
        }
      }
    }
    #pragma unroll 1
    for (int i = lane; i < e_hw; i += T) {
      if (!P.e_alive[i]) continue;
      if (!exclude_vertex && !P.e_in[i]) continue;
      const S d = P.e_d[i];
      if (d < S(INFINITY) && better(d, 1, P.e_seq[i], bd, bcls, bseq)) {
        bd = d; bcls = 1; bseq = P.e_seq[i]; bidx = i;
      }
    }
    #pragma unroll 1
    for (int i = lane; i < f_hw; i += T) {
      if (!P.f_alive[i] || !P.f_in[i]) continue;
      const S d = P.f_d[i];
      if (d < S(INFINITY) && better(d, 2, P.f_seq[i], bd, bcls, bseq)) {
        bd = d; bcls = 2; bseq = P.f_seq[i]; bidx = i;
      }
    }
#pragma unroll
    for (int off = T / 2; off > 0; off >>= 1) {
      const S od = shflXor(bd, off);
      const int ocls = shflXor(bcls, off);
      const int oseq = shflXor(bseq, off);
      const int oidx = shflXor(bidx, off);
      if (oidx >= 0 && (bidx < 0 || better(od, ocls, oseq, bd, bcls, bseq))) {
        bd = od; bcls = ocls; bseq = oseq; bidx = oidx;
      }
    }
    Feature f;
    f.cls = bcls;
    f.idx = bidx;
    return f;
  }

  // ComputeFaceNormalPointingOutward (epa_polytope.hpp:348-410).  `par`: the
  // caller is warp-uniform and the rare all-vertex pass may use all lanes.
  FCLB_DI bool faceNormal(int f, V3<S>& normal, S* area, bool par) const {
    const V3<S> a = P.vloc(P.f_a[f]), b = P.vloc(P.f_b[f]), c = P.vloc(P.f_c[f]);
    const V3<S> e1 = a - b, e2 = b - c;
    const V3<S> cr = cross(e1, e2);
    const S nrm = norm(cr);
    if (area) *area = S(0.5) * nrm;
    if (nrm <= S(0)) return false;
    const V3<S> dir = cr / nrm;
    const S o_to_a_dot_n = dot(a, dir);
    if (fabs_(o_to_a_dot_n) >= S(1e-4)) {
      normal = (o_to_a_dot_n > 0) ? dir : -dir;
      return true;
    }
    S max_pos = S(0), min_neg = S(0);
    if (par) {
      #pragma unroll 1
      for (int i = lane; i < v_hw; i += T) {
        if (!P.v_alive[i]) continue;
        const S dv = dot(P.vloc(i), dir);
        if (dv > max_pos) max_pos = dv;
        if (dv < min_neg) min_neg = dv;
      }
#pragma unroll
      for (int off = T / 2; off > 0; off >>= 1) {
        const S om = shflXor(max_pos, off);
        const S on = shflXor(min_neg, off);
        if (om > max_pos) max_pos = om;
        if (on < min_neg) min_neg = on;
      }
    } else {
      #pragma unroll 1
      for (int i = 0; i < v_hw; i++) {
        if (!P.v_alive[i]) continue;
        const S dv = dot(P.vloc(i), dir);
        if (dv > max_pos) max_pos = dv;
        if (dv < min_neg) min_neg = dv;
      }
    }
    normal = (fabs_(max_pos) > fabs_(min_neg)) ? -dir : dir;
    return true;
  }
  // IsPointOutsidePolytopeFace (epa_polytope_expand.hpp:13-31), threshold 0
  FCLB_DI bool pointOutsideFace(int f, const V3<S>& pt, bool par) const {
    V3<S> n;
    S area = S(0);
    if (!faceNormal(f, n, &area, par)) return area <= S(0);
    const V3<S> on_face = P.vloc(P.f_a[f]);
    return dot(n, pt - on_face) >= S(0);
  }

  // findNextSupportDirection (epa.hpp:11-113): 0 OK, 1 Failed, 2 Converge
  FCLB_DI int faceCandidate(int f, bool try_witness, const V3<S>& witness, S dist_sq, S tol, V3<S>& next_d, V3<S>& next_v,
                            int& start_face) {
    const S outer_thr = S(1e-3) * S(1e-3);
    V3<S> fn;
    if (!faceNormal(f, fn, nullptr, true)) return 1;
    const V3<S> on_face = P.vloc(P.f_a[f]);
    if (try_witness && dist_sq > outer_thr) {
      next_d = normalized(witness);
      next_v = support(next_d);
      const S delta = dot(fn, next_v - on_face);
      if (delta >= S(0)) {
        start_face = f;
        return 0;
      }
    }
    next_d = fn;
    next_v = support(next_d);
    const S delta = dot(fn, next_v - on_face);
    if (delta > tol) {
      start_face = f;
      return 0;
    }
    return 2;
  }

  // ExpandPolytope (epa_polytope_expand.hpp:33-91): 0 OK, 1 Failed, 2 MallocFailed
  FCLB_DI int expand(const V3<S>& nv, const V3<S>& nd, int start_face) {
    // initVisibilityCacheVariables + visibility predicate of EVERY face
    // (cached_vertex2new_v_edge and the edges' cached_visibility are written before they are read below)
    #pragma unroll 1
    for (int i = lane; i < v_hw; i += T) P.v_rm[i] = 1;
    #pragma unroll 1
    for (int i = lane; i < f_hw; i += T) {
      if (!P.f_alive[i]) continue;
      P.f_vis[i] = pointOutsideFace(i, nv, false) ? 3 : 2;  // 3 = outside (not reached yet), 2 = hidden
    }
    sync();
    if (lane == 0) P.f_vis[start_face] = 1;  // the start face is visible by construction (:133)
    sync();
    // grow the visible patch: a face joins when it is "outside" and shares an
    // edge with a patch face (computeVisiblePatch, :121-180).  Lanes read f_vis while others relabel 3 -> 1 inside a round
    // (compute-sanitizer racecheck reports it as a warning): labels only ever move 3 -> 1, a stale read postpones a face
    // to the next round, and rounds repeat until nothing changes -- the fixed point, hence the result, is the same.
    bool broken = false;
    while (true) {
      bool changed = false;
      #pragma unroll 1
      for (int i = lane; i < f_hw; i += T) {
        if (!P.f_alive[i] || P.f_vis[i] != 1) continue;
        const int es[3] = {P.f_e0[i], P.f_e1[i], P.f_e2[i]};
        for (int k = 0; k < 3; k++) {
          const int e = es[k];
          const int g = (P.e_f0[e] == i) ? P.e_f1[e] : P.e_f0[e];
          if (g == kNil) {
            broken = true;
            continue;
          }
          if (P.f_vis[g] == 3) {
            P.f_vis[g] = 1;
            changed = true;
          }
        }
      }
      sync();
      if (!any(changed)) break;
    }
    if (any(broken)) return 1;
    // edge classification + vertex keep flags (updateVertexRemoveFlag, :183-205)
    #pragma unroll 1
    for (int i = lane; i < e_hw; i += T) {
      if (!P.e_alive[i]) continue;
      const int f0 = P.e_f0[i], f1 = P.e_f1[i];
      const bool in0 = (f0 != kNil) && P.f_vis[f0] == 1;
      const bool in1 = (f1 != kNil) && P.f_vis[f1] == 1;
      const uint8_t vis = (in0 && in1) ? 2 : ((in0 || in1) ? 1 : 0);
      P.e_vis[i] = vis;
      if (vis != 2) {
        P.v_rm[P.e_v0[i]] = 0;
        P.v_rm[P.e_v1[i]] = 0;
      }
    }
    sync();
    // removeAccordingToVisibility (:244-281)
    int rm_f = 0, rm_e = 0, rm_v = 0;
    #pragma unroll 1
    for (int i = lane; i < e_hw; i += T) {
      if (!P.e_alive[i]) continue;
      if (P.e_vis[i] == 2) {
        P.e_alive[i] = 0;
        rm_e++;
      } else if (P.e_vis[i] == 1) {
        // drop the reference to the removed face, keep the hidden one in f0
        const int f0 = P.e_f0[i], f1 = P.e_f1[i];
        const bool in0 = (f0 != kNil) && P.f_vis[f0] == 1;
        P.e_f0[i] = in0 ? uint16_t(f1) : uint16_t(f0);
        P.e_f1[i] = kNil;
      }
    }
    sync();
    #pragma unroll 1
    for (int i = lane; i < f_hw; i += T) {
      if (P.f_alive[i] && P.f_vis[i] == 1) {
        P.f_alive[i] = 0;
        rm_f++;
      }
    }
    #pragma unroll 1
    for (int i = lane; i < v_hw; i += T) {
      if (P.v_alive[i] && P.v_rm[i]) {
        P.v_alive[i] = 0;
        rm_v++;
      }
    }
#pragma unroll
    for (int off = T / 2; off > 0; off >>= 1) {
      rm_f += shflXor(rm_f, off);
      rm_e += shflXor(rm_e, off);
      rm_v += shflXor(rm_v, off);
    }
    f_n -= rm_f;
    e_n -= rm_e;
    v_n -= rm_v;
    sync();

    const int new_v = addVertex(nv, nd);
    if (new_v < 0) return 2;

    // ---- the new cone (epa_polytope_expand.hpp:52-85), built in parallel ----
    // The reference walks the edge list newest-first; for every border edge it creates the side edges
    // (new_v, v_i) on first use of v_i and then the face (edge, side(v_0), side(v_1)).  Results depend on the
    // creation ORDER (sequence numbers) and on which face is an edge's faces_of_edge[0 / 1], not on slots:
    //   rank r of a border edge       = number of border edges with a larger sequence number
    //   new face r                    = (border edge r, side(v0), side(v1)), face sequence f_sq + r
    //   "first use" of a vertex       = smallest key 2 r + k over the border-edge ends (r, k) that touch it;
    //   side edges are numbered in first-use order, their faces [0 / 1] are the two incident faces by rank.
    // A vertex with more than two border-edge ends would give its side edge a third face ("wrong edge",
    // epa_polytope.hpp:279-283), an exhausted edge / face pool fails an allocation: both end the evaluation with
    // MallocFailed (epa.hpp:221-226) whatever the order, so they are detected up front.
    const unsigned lt = (1u << lane) - 1u;
    int m = 0;
    #pragma unroll 1
    for (int base = 0; base < e_hw; base += T) {
      const int i = base + lane;
      const bool pred = i < e_hw && P.e_alive[i] && P.e_vis[i] == 1;
      const unsigned mask = ballot(pred);
      if (pred) P.e_tmp0[m + __popc(mask & lt)] = uint16_t(i);
      m += __popc(mask);
    }
    sync();
    #pragma unroll 1
    for (int j = lane; j < m; j += T) {
      const int ej = P.e_tmp0[j];
      const int sj = P.e_seq[ej];
      int r = 0;
      #pragma unroll 1
      for (int k = 0; k < m; k++) r += (int(P.e_seq[P.e_tmp0[k]]) > sj) ? 1 : 0;
      P.e_tmp1[r] = uint16_t(ej);
    }
    sync();
    // first uses, in key order: e_tmp0[ordinal of the side edge] = vertex (the unsorted list is dead by now)
    int n_new = 0;
    bool wrong = false;
    #pragma unroll 1
    for (int base = 0; base < 2 * m; base += T) {
      const int t = base + lane;
      bool first = false;
      int vtx = 0;
      if (t < 2 * m) {
        const int e = P.e_tmp1[t >> 1];
        vtx = (t & 1) ? P.e_v1[e] : P.e_v0[e];
        first = true;
        int occ = 0;
        #pragma unroll 1
        for (int u = 0; u < 2 * m; u++) {
          const int eu = P.e_tmp1[u >> 1];
          const int vu = (u & 1) ? P.e_v1[eu] : P.e_v0[eu];
          if (vu == vtx) {
            occ++;
            if (u < t) first = false;
          }
        }
        if (occ > 2) wrong = true;
      }
      const unsigned mask = ballot(first);
      if (first) P.e_tmp0[n_new + __popc(mask & lt)] = uint16_t(vtx);
      n_new += __popc(mask);
    }
    if (any(wrong)) return 2;
    if (n_new > P.ecap - e_n || m > P.fcap - f_n) return 2;
    sync();
    gatherFree(P.e_alive, P.ecap, e_hw, e_n, n_new, P.e_tmp2);
    sync();
    #pragma unroll 1
    for (int o = lane; o < n_new; o += T) {
      const int sl = P.e_tmp2[o], vi = P.e_tmp0[o];
      P.e_v0[sl] = uint16_t(new_v);
      P.e_v1[sl] = uint16_t(vi);
      P.e_f0[sl] = kNil;
      P.e_f1[sl] = kNil;
      P.e_seq[sl] = uint16_t(e_sq + o);
      P.e_alive[sl] = 1;
      P.e_vis[sl] = 0;
      P.v_newedge[vi] = uint16_t(sl);
      fillEdge(sl);
    }
    e_sq += n_new;
    e_n += n_new;
    sync();
    gatherFree(P.f_alive, P.fcap, f_hw, f_n, m, P.e_tmp2);
    sync();
    #pragma unroll 1
    for (int r = lane; r < m; r += T) {
      const int sl = P.e_tmp2[r], edge = P.e_tmp1[r];
      const int va = P.e_v0[edge], vb = P.e_v1[edge];
      P.f_e0[sl] = uint16_t(edge);
      P.f_e1[sl] = P.v_newedge[va];
      P.f_e2[sl] = P.v_newedge[vb];
      P.f_a[sl] = uint16_t(va);
      P.f_b[sl] = uint16_t(vb);
      P.f_c[sl] = uint16_t(new_v);
      P.f_seq[sl] = uint16_t(f_sq + r);
      P.f_alive[sl] = 1;
      P.f_vis[sl] = 0;
      P.e_f1[edge] = uint16_t(sl);  // removeAccordingToVisibility left the hidden face in slot 0
      fillFace(sl);
    }
    #pragma unroll 1
    for (int o = lane; o < n_new; o += T) {  // faces of the side edges, by rank
      const int vi = P.e_tmp0[o], sl = P.v_newedge[vi];
      int f0 = kNil, f1 = kNil;
      #pragma unroll 1
      for (int r = m - 1; r >= 0; r--) {
        const int edge = P.e_tmp1[r];
        if (P.e_v0[edge] == vi || P.e_v1[edge] == vi) {
          f1 = f0;
          f0 = P.e_tmp2[r];
        }
      }
      P.e_f0[sl] = uint16_t(f0);
      P.e_f1[sl] = uint16_t(f1);
    }
    f_sq += m;
    f_n += m;
    sync();
    return 0;
  }

  struct RawFeature {  // epa.h:70-74
    int cls;
    MinDist<S> md;
    V3<S> a, b, c, da, db, dc;
  };

  // assignPenetrationPairFromSegment (epa.hpp:377-433)
  FCLB_DI bool pairFromSegment(const RawFeature& f, S& depth, V3<S>& p0, V3<S>& p1) {
    depth = fsqrt(f.md.dist_sq);
    const V3<S> p = f.md.witness;
    const V3<S> a_to_b = f.b - f.a;
    int mx = 0;
    S mxd = fabs_(a_to_b.x);
    if (fabs_(a_to_b.y) > mxd) {
      mx = 1;
      mxd = fabs_(a_to_b.y);
    }
    if (fabs_(a_to_b.z) > mxd) {
      mx = 2;
      mxd = fabs_(a_to_b.z);
    }
    if (double(mxd) <= 1e-10) {
      supportBoth(f.da, p0, p1);
      return true;
    }
    const S s = (comp(p, mx) - comp(f.a, mx)) / (comp(f.b, mx) - comp(f.a, mx));
    if (s < 0 || s > S(1.0)) return false;
    const S aw = S(1.0) - s, bw = s;
    V3<S> a0, a1, b0, b1;
    supportBoth(f.da, a0, a1);
    supportBoth(f.db, b0, b1);
    p0 = aw * a0 + bw * b0;
    p1 = aw * a1 + bw * b1;
    return true;
  }
  // assignPenetrationPairFromFace (epa.hpp:436-497)
  FCLB_DI bool pairFromFace(const RawFeature& f, S& depth, V3<S>& p0, V3<S>& p1) {
    depth = fsqrt(f.md.dist_sq);
    const V3<S> p = f.md.witness;
    const V3<S> dl0 = f.a - f.b, dl1 = f.b - f.c, dl2 = f.c - f.a;
    const V3<S> n = cross(dl0, dl1);
    const S area = fsqrt(sqnorm(n));
    const S w0 = norm(cross(dl1, f.b - p)) / area;
    const S w1 = norm(cross(dl2, f.c - p)) / area;
    const S w2_check = norm(cross(dl0, f.a - p)) / area;
    const S w2 = S(1.0) - w0 - w1;
    if (fabs_(w2_check - w2) > S(0.01)) return false;
    V3<S> a0, a1, b0, b1, c0, c1;
    supportBoth(f.da, a0, a1);
    supportBoth(f.db, b0, b1);
    supportBoth(f.dc, c0, c1);
    p0 = w0 * a0 + w1 * b0 + w2 * c0;
    p1 = w0 * a1 + w1 * b1 + w2 * c1;
    return true;
  }
  // assignPenetrationPair (epa.hpp:341-375)
  FCLB_DI void assignPair(const V3<S>& cand_d, const RawFeature& f, S& depth, V3<S>& p0, V3<S>& p1) {
    if (f.cls == 0) {
      supportBoth(f.da, p0, p1);
      depth = norm(f.a);
      return;
    } else if (f.cls == 1) {
      if (pairFromSegment(f, depth, p0, p1)) return;
    } else if (f.cls == 2) {
      if (pairFromFace(f, depth, p0, p1)) return;
    }
    supportBoth(cand_d, p0, p1);
    depth = norm(p0 - p1);
  }

  // checkTerminateCondition (epa.hpp:269-298)
  FCLB_DI bool converged(const RawFeature& f, const V3<S>& new_v, S tol) const {
    S delta_sq;
    if (f.cls == 1) {
      delta_sq = pointToSegment(new_v, f.a, f.b).dist_sq;
    } else {
      const S d = pointToPlaneDistance(new_v, f.a, f.b, f.c);
      delta_sq = d * d;
    }
    return delta_sq < tol * tol;
  }

  // evaluateFromInitializedPolytope (epa.hpp:137-237)
  // One iteration of evaluateFromInitializedPolytope's loop (epa.hpp:137-237).  Returns kEpaContinue or the
  // final status.  Kept separate so that the tiles of a warp can take their iterations in lockstep.
  static constexpr int kEpaContinue = -1;
  int iteration = 0;
  FCLB_DI int step(int max_iterations, S tol, S& depth, V3<S>& p0, V3<S>& p1) {
    {
      Feature nf = nearest(false);
      if (nf.idx < 0) return EPA_FAILED;
      if (nf.cls == 0) {
        const S vdist = P.vd[nf.idx];
        const S ratio = S(1) - S(1e-3);
        const Feature other = nearest(true);
        if (other.idx < 0) return EPA_FAILED;
        if (other.cls == 1) {
          if (ratio * P.e_d[other.idx] < vdist) nf = other;
        } else if (other.cls == 2) {
          if (ratio * P.f_d[other.idx] < vdist) nf = other;
        }
      }
      if (nf.cls != 1 && nf.cls != 2) return EPA_FAILED;

      // transformToRawFeature (epa.hpp:301-338); the witness is recomputed from the
      // stored vertex order (a pure function of it)
      RawFeature raw;
      raw.cls = nf.cls;
      raw.c = zero3<S>();
      raw.dc = zero3<S>();
      if (nf.cls == 1) {
        const int a = P.e_v0[nf.idx], b = P.e_v1[nf.idx];
        raw.a = P.vloc(a);
        raw.b = P.vloc(b);
        raw.da = P.vdir(a);
        raw.db = P.vdir(b);
        raw.md = pointToSegment(zero3<S>(), raw.a, raw.b);
      } else {
        const int a = P.f_a[nf.idx], b = P.f_b[nf.idx], c = P.f_c[nf.idx];
        raw.a = P.vloc(a);
        raw.b = P.vloc(b);
        raw.c = P.vloc(c);
        raw.da = P.vdir(a);
        raw.db = P.vdir(b);
        raw.dc = P.vdir(c);
        raw.md = pointToTriangle(zero3<S>(), raw.a, raw.b, raw.c);
      }

      V3<S> next_d = zero3<S>(), next_v = zero3<S>();
      int start_face = -1;
      int ds;
      if (nf.cls == 2) {
        ds = faceCandidate(nf.idx, true, raw.md.witness, raw.md.dist_sq, tol, next_d, next_v, start_face);
      } else {
        const S outer_thr = S(1e-3) * S(1e-3);
        const int f0 = P.e_f0[nf.idx], f1 = P.e_f1[nf.idx];
        bool decided = false;
        ds = 1;
        if (raw.md.dist_sq > outer_thr) {
          next_d = normalized(raw.md.witness);
          next_v = support(next_d);
          if (f0 != kNil && pointOutsideFace(f0, next_v, true)) {
            start_face = f0;
            ds = 0;
            decided = true;
          } else if (f1 != kNil && pointOutsideFace(f1, next_v, true)) {
            start_face = f1;
            ds = 0;
            decided = true;
          }
        }
        if (!decided) {
          if (f0 == kNil || f1 == kNil) return EPA_FAILED;
          ds = faceCandidate(f0, false, raw.md.witness, raw.md.dist_sq, tol, next_d, next_v, start_face);
          if (ds != 0) ds = faceCandidate(f1, false, raw.md.witness, raw.md.dist_sq, tol, next_d, next_v, start_face);
        }
      }
      if (ds == 1) return EPA_FAILED;
      if (ds == 2) {
        assignPair(next_d, raw, depth, p0, p1);
        return EPA_OK;
      }
      if (start_face < 0) return EPA_FAILED;
      if (converged(raw, next_v, tol)) {
        assignPair(next_d, raw, depth, p0, p1);
        return EPA_OK;
      }
      const int es = expand(next_v, next_d, start_face);
      if (es == 1) return EPA_FAILED;
      if (es == 2) {
        assignPair(next_d, raw, depth, p0, p1);
        return EPA_MALLOC_FAILED;
      }
      iteration += 1;
      if (iteration >= max_iterations) {
        assignPair(next_d, raw, depth, p0, p1);
        return EPA_ITER_LIMIT;
      }
    }
    return kEpaContinue;
  }
  FCLB_DI int run(int max_iterations, S tol, S& depth, V3<S>& p0, V3<S>& p1) {
    iteration = 0;
    while (true) {
      const int r = step(max_iterations, tol, depth, p0, p1);
      if (r != kEpaContinue) return r;
    }
  }
  // reset + simplexToPolytope: kEpaContinue when the iteration loop has to run
  FCLB_DI int begin(const SlotStore<S>& st, const Simp& sx, S tol, S& depth, V3<S>& p0, V3<S>& p1) {
    reset();
    iteration = 0;
    const int s2p = simplexToPolytope(st, sx, tol, p0, p1);
    if (s2p == 2) return EPA_FAILED;
    if (s2p == 1) {
      depth = S(0);
      return EPA_TOUCHING;
    }
    return kEpaContinue;
  }

  // EPA::Evaluate (epa.hpp:240-256) via evaluateFromUnInitializedPolytope (:116-135)
  FCLB_DI int evaluate(const SlotStore<S>& st, const Simp& sx, int max_iterations, S tol, S& depth, V3<S>& p0,
                       V3<S>& p1) {
    reset();
    const int s2p = simplexToPolytope(st, sx, tol, p0, p1);
    if (s2p == 2) return EPA_FAILED;
    if (s2p == 1) {
      depth = S(0);
      return EPA_TOUCHING;
    }
    return run(max_iterations, tol, depth, p0, p1);
  }
};

}  // namespace fclb
