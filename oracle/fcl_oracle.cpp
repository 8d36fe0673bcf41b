// oracle/fcl_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement ("port" oracle)
// of the reference's convex narrowphase path, plain C++ without Eigen.
//
// Pinned against the reference itself: tests/test_oracle_cpu.py runs this file
// and oracle/_ref/libfclref.so (the unmodified reference headers) on the same
// seeded inputs and requires bit-identical flags and outputs.  It exists so the
// parity tests and bench.py's CPU leg have an oracle on machines where
// /root/reference (and hence oracle/_ref) is not available.  The product never
// links, loads or calls it.
//
// Every function cites the reference file:line it restates
// (paths relative to /root/reference/include/fcl/).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <thread>
#include <vector>

#include "fcl_oracle_math.h"

namespace orc {

enum { T_BOX = 0, T_SPHERE = 1, T_ELLIPSOID = 2, T_CAPSULE = 3, T_CONE = 4, T_CYLINDER = 5, T_CONVEX = 6 };

struct ShapeRec {
  uint32_t type;
  uint32_t geom;
  double p[3];
};
struct PairRec {
  uint32_t s1, s2;
};

// Work counters of the generic GJK path (the flop / query model is built from counts that are
// deterministic properties of the input, taken on the CPU restatement).  Active only inside fclport_gjk_work_counters.
struct WorkCounters {
  uint64_t queries = 0;
  uint64_t support_vertices = 0;  // MinkowskiDiff::supportVertex evaluations (two shape supports each)
  uint64_t extract_supports = 0;  // single-shape supports of the witness extraction (gjk_distance.hpp:376-470)
  uint64_t convex_dots = 0;       // d . v evaluations inside Convex::findExtremeVertex (both variants)
  uint64_t project[5] = {0, 0, 0, 0, 0};  // simplexProjection{2,3,4} calls by simplex rank at entry (boolean loop)
  uint64_t update[5] = {0, 0, 0, 0, 0};   // computeMinDistanceAndUpdateSimplex calls by rank at entry (distance loop)
  uint64_t separated = 0;
  void add(const WorkCounters& o) {
    queries += o.queries;
    support_vertices += o.support_vertices;
    extract_supports += o.extract_supports;
    convex_dots += o.convex_dots;
    for (int k = 0; k < 5; k++) {
      project[k] += o.project[k];
      update[k] += o.update[k];
    }
    separated += o.separated;
  }
};
static thread_local WorkCounters* t_cnt = nullptr;

// ---------------------------------------------------------------------------
// Convex<S>  (geometry/shape/convex-inl.h)
template <typename T>
struct ConvexData {
  std::vector<Vec3<T>> verts;
  std::vector<int> nbr;  // neighbors_ encoding, convex.h:219-242
  bool walk = false;     // find_extreme_via_neighbors_
  int seed[6];           // init_direction_vertex_cache_, convex-inl.h:153-200
  Vec3<T> interior;      // convex-inl.h:64-71

  // findExtremeVertexIndexNaive, convex-inl.h:133-150
  int extremeNaive(const Vec3<T>& d) const {
    int best = 0;
    T best_v = d.dot(verts[0]);
    if (t_cnt) t_cnt->convex_dots += verts.size();
    for (int i = 1; i < int(verts.size()); i++) {
      const T v = d.dot(verts[i]);
      if (v > best_v) {
        best = i;
        best_v = v;
      }
    }
    return best;
  }
  // findExtremeVertexViaNeighbours, convex-inl.h:202-281
  int extremeWalk(const Vec3<T>& d) const {
    static const T axes[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    int init = -1;
    T max_dot = 0;
    for (int k = 0; k < 6; k++) {
      const T dv = Vec3<T>(axes[k][0], axes[k][1], axes[k][2]).dot(d);
      if (init < 0 || dv > max_dot) {
        init = seed[k];
        max_dot = dv;
      }
    }
    int ext = init;
    T ext_v = d.dot(verts[ext]);
    if (t_cnt) t_cnt->convex_dots += 1;
    int parent0 = init, parent1 = init;
    bool keep = true;
    while (keep) {
      keep = false;
      const int start = nbr[ext], count = nbr[start], old_ext = ext;
      for (int k = start + 1; k <= start + count; k++) {
        const int nb = nbr[k];
        if (nb == parent0 || nb == parent1) continue;
        const T nv = d.dot(verts[nb]);
        if (t_cnt) t_cnt->convex_dots += 1;
        if (nv > ext_v) {
          parent1 = ext;
          keep = true;
          ext = nb;
          ext_v = nv;
        }
      }
      parent0 = old_ext;
    }
    return ext;
  }
  const Vec3<T>& findExtremeVertex(const Vec3<T>& d) const { return verts[walk ? extremeWalk(d) : extremeNaive(d)]; }
};

struct ConvexSrc {
  std::vector<double> verts;
  std::vector<int> faces;
  int num_faces;
};
static std::vector<ConvexSrc>& convexSources() {
  static std::vector<ConvexSrc> t;
  return t;
}

// Convex ctor + FindVertexNeighbors + ValidateTopology, convex-inl.h:52-75,293-407
template <typename T>
ConvexData<T> buildConvex(const ConvexSrc& src) {
  ConvexData<T> c;
  const int n = int(src.verts.size() / 3);
  for (int i = 0; i < n; i++) c.verts.emplace_back(T(src.verts[3 * i]), T(src.verts[3 * i + 1]), T(src.verts[3 * i + 2]));
  Vec3<T> sum;
  for (const auto& v : c.verts) sum = sum + v;
  c.interior = sum * (T)(1.0 / n);
  std::vector<std::set<int>> nb(n);
  std::map<std::pair<int, int>, int> edge_faces;
  int fi = 0;
  for (int f = 0; f < src.num_faces; f++) {
    const int cnt = src.faces[fi];
    int prev = src.faces[fi + cnt];
    for (int i = fi + 1; i <= fi + cnt; i++) {
      const int v = src.faces[i];
      nb[v].insert(prev);
      nb[prev].insert(v);
      edge_faces[std::make_pair(std::min(v, prev), std::max(v, prev))]++;
      prev = v;
    }
    fi += cnt + 1;
  }
  c.nbr.resize(n);
  bool connected = true;
  for (int v = 0; v < n; v++) {
    c.nbr[v] = int(c.nbr.size());
    c.nbr.push_back(int(nb[v].size()));
    c.nbr.insert(c.nbr.end(), nb[v].begin(), nb[v].end());
    if (nb[v].empty()) connected = false;
  }
  bool watertight = true;
  for (const auto& kv : edge_faces)
    if (kv.second != 2) watertight = false;
  c.walk = (n > 32) && watertight && connected;  // convex.h:259
  static const T axes[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  for (int k = 0; k < 6; k++) c.seed[k] = c.extremeNaive(Vec3<T>(axes[k][0], axes[k][1], axes[k][2]));
  return c;
}

// ---------------------------------------------------------------------------
// GJKGeometryData + support functions (cvx_collide/gjk_shape.hpp:36-179,
// narrowphase/detail/gjk_solver_cvx-inl.h:63-128)
template <typename T>
struct Geom {
  int type;
  Vec3<T> data;
  const ConvexData<T>* cvx = nullptr;
};

template <typename T>
Vec3<T> supportOf(const Geom<T>& g, const Vec3<T>& dir) {
  switch (g.type) {
    case T_BOX: {  // gjk_shape.hpp:44-52
      const Vec3<T>& side = g.data;
      return Vec3<T>((dir[0] > 0) ? (side[0] / 2) : (-side[0] / 2), (dir[1] > 0) ? (side[1] / 2) : (-side[1] / 2),
                     (dir[2] > 0) ? (side[2] / 2) : (-side[2] / 2));
    }
    case T_SPHERE:  // :65-72
      return dir * g.data[0];
    case T_ELLIPSOID: {  // :80-91
      const T a2 = g.data[0] * g.data[0], b2 = g.data[1] * g.data[1], c2 = g.data[2] * g.data[2];
      const Vec3<T> v(a2 * dir[0], b2 * dir[1], c2 * dir[2]);
      const T d = std::sqrt(v.dot(dir));
      return v / d;
    }
    case T_CAPSULE: {  // :104-121
      const T radius = g.data[0], lz = g.data[1];
      const T half_h = lz * T(0.5);
      Vec3<T> pos1(0, 0, half_h), pos2(0, 0, -half_h);
      const Vec3<T> v = dir * radius;
      pos1 = pos1 + v;
      pos2 = pos2 + v;
      return (dir.dot(pos1) > dir.dot(pos2)) ? pos1 : pos2;
    }
    case T_CONE: {  // :134-158
      const T radius = g.data[0], lz = g.data[1];
      T zdist = dir[0] * dir[0] + dir[1] * dir[1];
      T len = zdist + dir[2] * dir[2];
      zdist = std::sqrt(zdist);
      len = std::sqrt(len);
      const T half_h = lz * T(0.5);
      const T sin_a = radius / std::sqrt(radius * radius + 4 * half_h * half_h);
      if (dir[2] > len * sin_a) return Vec3<T>(0, 0, half_h);
      if (zdist > 0) {
        const T rad = radius / zdist;
        return Vec3<T>(rad * dir[0], rad * dir[1], -half_h);
      }
      return Vec3<T>(0, 0, -half_h);
    }
    case T_CYLINDER: {  // :171-185
      const T radius = g.data[0], lz = g.data[1];
      const T zdist = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1]);
      const T half_h = lz * T(0.5);
      if (zdist == 0.0) return Vec3<T>(0, 0, (dir[2] > 0) ? half_h : -half_h);
      const T d = radius / zdist;
      return Vec3<T>(d * dir[0], d * dir[1], (dir[2] > 0) ? half_h : -half_h);
    }
    case T_CONVEX:  // geometry/shape/shape_gjk_interface-inl.h:86-93
      return g.cvx->findExtremeVertex(dir);
    default:
      return Vec3<T>();
  }
}
template <typename T>
Vec3<T> interiorOf(const Geom<T>& g) {  // gjk_solver_cvx-inl.h:110-128
  if (g.type == T_CONVEX) return g.cvx->interior;
  return Vec3<T>();
}

// cvx_collide/minkowski_diff.h:16-51, .hpp:17-84
template <typename T>
struct MinkowskiDiff {
  Geom<T> shapes[2];
  Mat3<T> toshape1;
  Xform<T> toshape0;
  mutable uint32_t n_support = 0;
  Vec3<T> support0(const Vec3<T>& d) const {
    n_support++;
    return supportOf(shapes[0], d);
  }
  Vec3<T> support1(const Vec3<T>& d) const {
    n_support++;
    return toshape0 * supportOf(shapes[1], toshape1 * d);
  }
  Vec3<T> support(const Vec3<T>& d) const {
    if (t_cnt) t_cnt->support_vertices += 1;
    return support0(d) - support1(-d);
  }
  Vec3<T> interior() const { return interiorOf(shapes[0]) - toshape0 * interiorOf(shapes[1]); }
  // gjk_solver-inl.h:79-85
  void setPoses(const Xform<T>& tf1, const Xform<T>& tf2) {
    toshape1 = tf2.R.transpose() * tf1.R;
    toshape0 = tf1.inverse() * tf2;
  }
};

template <typename T>
struct MVertex {  // MinkowskiDiffVertex
  Vec3<T> vertex, direction;
};
template <typename T>
struct Simplex {  // GJKSimplex, gjk.h:31-47
  MVertex<T> vertices[4];
  int rank = -1;
  void reset() { rank = -1; }
  void add(const MVertex<T>& v) {
    if (rank < 0) rank = 0;
    vertices[rank] = v;
    rank += 1;
  }
};

enum GjkStatus { GJK_INTERSECT = 0, GJK_SEPARATED = 1, GJK_NO_PROGRESS = 2, GJK_ITER_LIMIT = 3, GJK_FAILED = 4 };

// cvx_collide/gjk.hpp + gjk_distance.hpp
template <typename T>
class Gjk {
 public:
  Gjk(size_t max_it, T tol) : max_iterations_(max_it), tolerance_(tol) {}
  struct DistOut {
    bool valid = false;
    Vec3<T> p0, p1;
  };

  // gjk.hpp:11-146
  GjkStatus evaluate(const MinkowskiDiff<T>& shape, Simplex<T>& simplex, const Vec3<T>& guess, DistOut* dist) const {
    Vec3<T> direction = guess;
    if (direction.squaredNorm() <= 0.0) direction = Vec3<T>(1, 0, 0);
    direction.normalize();
    MVertex<T> vertex{shape.support(direction), direction};
    simplex.reset();
    simplex.add(vertex);
    const T tol_sq = tolerance_ * tolerance_;
    if (vertex.vertex.squaredNorm() <= tol_sq) return GJK_INTERSECT;
    if (vertex.vertex.dot(direction) < 0) return separated(shape, simplex, vertex, dist);
    direction = direction * T(-1);
    size_t it = 0;
    while (it < max_iterations_) {
      it += 1;
      vertex.direction = direction;
      vertex.vertex = shape.support(direction);
      if (vertex.vertex.dot(direction) < 0) return separated(shape, simplex, vertex, dist);
      for (int j = 0; j < simplex.rank; j++)
        if ((simplex.vertices[j].vertex - vertex.vertex).squaredNorm() < tol_sq) return GJK_NO_PROGRESS;
      if (vertex.vertex.squaredNorm() <= tol_sq) {
        simplex.add(vertex);
        return GJK_INTERSECT;
      }
      simplex.add(vertex);
      const int ps = project(simplex, direction);
      if (ps == P_FAILED) return GJK_FAILED;
      if (ps == P_INTERSECT) return GJK_INTERSECT;
      if (ps == P_ZERO_VOLUME) return GJK_NO_PROGRESS;
    }
    return GJK_ITER_LIMIT;
  }

 private:
  const size_t max_iterations_;
  const T tolerance_;
  enum { P_FAILED, P_CONTINUE, P_INTERSECT, P_ZERO_VOLUME };

  // process_separated_vertex, gjk.hpp:22-52
  GjkStatus separated(const MinkowskiDiff<T>& shape, Simplex<T>& simplex, const MVertex<T>& v, DistOut* dist) const {
    if (!dist) return GJK_SEPARATED;
    simplex.reset();
    simplex.add(v);
    dist->valid = minDistance(shape, simplex, dist->p0, dist->p1);
    return GJK_SEPARATED;
  }

  int project(Simplex<T>& s, Vec3<T>& dir) const {
    if (t_cnt && s.rank >= 0 && s.rank < 5) t_cnt->project[s.rank] += 1;
    if (s.rank == 2) return project2(s, dir);
    if (s.rank == 3) return project3(s, dir);
    return project4(s, dir);
  }
  // gjk.hpp:162-201
  int project2(Simplex<T>& s, Vec3<T>& dir) const {
    const MVertex<T> B = s.vertices[0], A = s.vertices[1];
    const Vec3<T>& a = A.vertex;
    const Vec3<T> a_to_b = B.vertex - a;
    const T ao_dot_ab = -a_to_b.dot(a);
    const Vec3<T> ab_cross_ao = a.cross(a_to_b);
    if (ao_dot_ab > 0 && ab_cross_ao.squaredNorm() <= 0.0) return P_INTERSECT;
    if (ao_dot_ab <= 0) {
      s.reset();
      s.add(A);
      dir = -a.normalized();
      return P_CONTINUE;
    }
    dir = ab_cross_ao.cross(a_to_b);
    dir.normalize();
    return P_CONTINUE;
  }
  // gjk.hpp:204-286
  int project3(Simplex<T>& s, Vec3<T>& dir) const {
    const MVertex<T> C = s.vertices[0], B = s.vertices[1], A = s.vertices[2];
    const Vec3<T>& a = A.vertex;
    const Vec3<T> a_to_b = B.vertex - a, a_to_c = C.vertex - a;
    const bool sep_b = a.dot(a_to_b) >= 0, sep_c = a.dot(a_to_c) >= 0;
    if (sep_b && sep_c) {
      s.reset();
      s.add(A);
      dir = -a;
      dir.normalize();
      return P_CONTINUE;
    }
    const Vec3<T> n = a_to_b.cross(a_to_c);
    const Vec3<T> ac_n = n.cross(a_to_c);
    if (a.dot(ac_n) <= 0) {
      s.vertices[1] = A;
      s.rank = 2;
      dir = a.cross(a_to_c).cross(a_to_c);
      dir.normalize();
      return P_CONTINUE;
    }
    const Vec3<T> ab_n = a_to_b.cross(n);
    if (a.dot(ab_n) <= 0) {
      s.vertices[0] = B;
      s.vertices[1] = A;
      s.rank = 2;
      dir = a.cross(a_to_b).cross(a_to_b);
      dir.normalize();
      return P_CONTINUE;
    }
    const T area = n.norm();
    if (area < tolerance_ * tolerance_) return P_FAILED;
    const Vec3<T> nu = n / area;
    const T nd = nu.dot(a);
    if (std::abs(nd) < tolerance_) return P_INTERSECT;
    dir = (nd <= 0) ? nu : -nu;
    return P_CONTINUE;
  }
  // gjk.hpp:289-362
  int project4(Simplex<T>& s, Vec3<T>& dir) const {
    const MVertex<T> D = s.vertices[0], C = s.vertices[1], B = s.vertices[2], A = s.vertices[3];
    const Vec3<T>& a = A.vertex;
    const Vec3<T> ab = B.vertex - a, ac = C.vertex - a, ad = D.vertex - a;
    Vec3<T> abc = ab.cross(ac), acd = ac.cross(ad), abd = ab.cross(ad);
    const T abc_ad = abc.dot(ad), acd_ab = acd.dot(ab), abd_ac = abd.dot(ac);
    if (std::abs(abc_ad) <= 0.0) return P_ZERO_VOLUME;
    if (abc_ad > 0) abc = abc * T(-1);
    if (acd_ab > 0) acd = acd * T(-1);
    if (abd_ac > 0) abd = abd * T(-1);
    const bool d_side = a.dot(abc) > 0, c_side = a.dot(abd) > 0, b_side = a.dot(acd) > 0;
    if (d_side && c_side && b_side) return P_INTERSECT;
    if (!b_side) {
      s.vertices[2] = A;
    } else if (!c_side) {
      s.vertices[1] = B;
      s.vertices[2] = A;
    } else {
      s.vertices[0] = C;
      s.vertices[1] = B;
      s.vertices[2] = A;
    }
    s.rank = 3;
    return project3(s, dir);
  }

  // ---- separation distance, gjk_distance.hpp ----
  enum { U_NO_IMPROVEMENT, U_OK, U_FAILED };
  // :128-153
  void update2(Simplex<T>& s, Vec3<T>& out) const {
    const Vec3<T> s1 = s.vertices[1].vertex, s2 = s.vertices[0].vertex;
    const Vec3<T> d = s2 - s1;
    const T sq = d.squaredNorm();
    const T t = -s1.dot(d);
    if (t <= 0 || sq <= tolerance_ * tolerance_) {
      out = s1;
      s.vertices[0] = s.vertices[1];
      s.rank = 1;
    } else if (t >= sq) {
      out = s2;
      s.rank = 1;
    } else {
      const T w2 = t / sq;
      out = w2 * s2 + (T(1.0) - w2) * s1;
    }
  }
  // :156-289
  void update3(Simplex<T>& s, Vec3<T>& out) const {
    const Vec3<T> s1 = s.vertices[2].vertex, s2 = s.vertices[1].vertex, s3 = s.vertices[0].vertex;
    const Vec3<T> s12 = s2 - s1, s13 = s3 - s1;
    const bool sep2 = s1.dot(s12) >= 0, sep3 = s1.dot(s13) >= 0;
    if (sep2 && sep3) {
      out = s1;
      s.vertices[0] = s.vertices[2];
      s.rank = 1;
      return;
    }
    const Vec3<T> n = s12.cross(s13);
    const T area_sq = n.squaredNorm();
    const bool zero_area = area_sq <= T(0.0);
    const Vec3<T> n12 = n.cross(s12);
    const bool e12 = s1.dot(n12) > 0;
    if (!sep2 && e12) {
      s.vertices[0] = s.vertices[1];
      s.vertices[1] = s.vertices[2];
      s.rank = 2;
      update2(s, out);
      return;
    }
    const Vec3<T> n13 = n.cross(s13);
    const bool e13 = s1.dot(n13) < 0;
    if (!sep3 && e13) {
      s.vertices[1] = s.vertices[2];
      s.rank = 2;
      update2(s, out);
      return;
    }
    T best = -1;
    Vec3<T> best_pt, pt;
    Simplex<T> best_s, c;
    if (zero_area || e12) {
      c = s;
      c.vertices[0] = c.vertices[1];
      c.vertices[1] = c.vertices[2];
      c.rank = 2;
      update2(c, pt);
      const T d2 = pt.squaredNorm();
      if (best < 0 || d2 < best) {
        best = d2;
        best_s = c;
        best_pt = pt;
      }
    }
    if (zero_area || e13) {
      c = s;
      c.vertices[1] = c.vertices[2];
      c.rank = 2;
      update2(c, pt);
      const T d2 = pt.squaredNorm();
      if (best < 0 || d2 < best) {
        best = d2;
        best_s = c;
        best_pt = pt;
      }
    }
    const Vec3<T> n23 = n.cross(s3 - s2);
    const bool e23 = s2.dot(n23) > 0;
    if (zero_area || e23) {
      c = s;
      c.rank = 2;
      update2(c, pt);
      const T d2 = pt.squaredNorm();
      if (best < 0 || d2 < best) {
        best = d2;
        best_s = c;
        best_pt = pt;
      }
    }
    if (best < 0) {
      const T d = s1.dot(n);
      out = n * (d / area_sq);
    } else {
      s = best_s;
      out = best_pt;
    }
  }
  // :292-371
  int update4(Simplex<T>& s, Vec3<T>& out) const {
    T best = -1;
    Vec3<T> best_pt, pt;
    Simplex<T> best_s, c;
    for (int f = 0; f < 3; f++) {
      c = s;
      if (f == 0) {
        c.vertices[0] = c.vertices[1];
        c.vertices[1] = c.vertices[2];
        c.vertices[2] = c.vertices[3];
      } else if (f == 1) {
        c.vertices[1] = c.vertices[2];
        c.vertices[2] = c.vertices[3];
      } else {
        c.vertices[2] = c.vertices[3];
      }
      c.rank = 3;
      update3(c, pt);
      const T d2 = pt.squaredNorm();
      if (best < 0 || d2 < best) {
        best = d2;
        best_pt = pt;
        best_s = c;
      }
    }
    if (best < 0) return U_NO_IMPROVEMENT;
    s = best_s;
    out = best_pt;
    return U_OK;
  }
  // :109-126
  int update(Simplex<T>& s, Vec3<T>& out) const {
    if (t_cnt && s.rank >= 0 && s.rank < 5) t_cnt->update[s.rank] += 1;
    if (s.rank == 1) {
      out = s.vertices[0].vertex;
      return U_OK;
    }
    if (s.rank == 2) {
      update2(s, out);
      return U_OK;
    }
    if (s.rank == 3) {
      update3(s, out);
      return U_OK;
    }
    if (s.rank == 4) return update4(s, out);
    return U_FAILED;
  }
  // extractSeparationPointNoSubSimplex, :376-470
  bool extract(const MinkowskiDiff<T>& shape, const Simplex<T>& s, Vec3<T>& p0, Vec3<T>& p1) const {
    const T bary_tol = T(1e-3);
    if (t_cnt && s.rank >= 1 && s.rank <= 3) t_cnt->extract_supports += 2 * uint64_t(s.rank);
    if (s.rank == 4 || s.rank <= 0) return false;
    if (s.rank == 1) {
      const auto& v = s.vertices[0];
      p0 = shape.support0(v.direction);
      p1 = shape.support1(-v.direction);
      return true;
    }
    if (s.rank == 2) {
      const auto &v1 = s.vertices[0], &v2 = s.vertices[1];
      const Vec3<T> d = v2.vertex - v1.vertex;
      const T sq = d.squaredNorm();
      if (sq <= 0.0) {
        p0 = shape.support0(v1.direction);
        p1 = shape.support1(-v1.direction);
        return true;
      }
      const T t = -v1.vertex.dot(d);
      const T w2 = t / sq;
      if (w2 > 1 + bary_tol) {
        p0 = shape.support0(v2.direction);
        p1 = shape.support1(-v2.direction);
        return false;
      }
      if (w2 < -bary_tol) {
        p0 = shape.support0(v1.direction);
        p1 = shape.support1(-v1.direction);
        return false;
      }
      const T w1 = T(1.0) - w2;
      p0 = shape.support0(v1.direction) * w1 + shape.support0(v2.direction) * w2;
      p1 = shape.support1(-v1.direction) * w1 + shape.support1(-v2.direction) * w2;
      return true;
    }
    const auto &v1 = s.vertices[0], &v2 = s.vertices[1], &v3 = s.vertices[2];
    const Vec3<T>&s1 = v1.vertex, &s2 = v2.vertex, &s3 = v3.vertex;
    const Vec3<T> s12 = s2 - s1, s13 = s3 - s1;
    const Vec3<T> n = s12.cross(s13);
    const T area_sq = n.squaredNorm();
    if (area_sq <= 0.0) return false;
    const T d = s1.dot(n);
    const Vec3<T> proj = n * (d / area_sq);
    const T area = std::sqrt(area_sq);
    const T w2 = (s13.cross(s1 - proj)).norm() / area;
    const T w3 = (s12.cross(s1 - proj)).norm() / area;
    const T w1 = T(1.0) - w2 - w3;
    if (w1 < -bary_tol || w2 < -bary_tol || w3 < -bary_tol) return false;
    p0 = shape.support0(v1.direction) * w1 + shape.support0(v2.direction) * w2 + shape.support0(v3.direction) * w3;
    p1 = shape.support1(-v1.direction) * w1 + shape.support1(-v2.direction) * w2 + shape.support1(-v3.direction) * w3;
    return true;
  }
  // findMinimumDistancePointsWithSeparatedVertexInit, :11-106
  bool minDistance(const MinkowskiDiff<T>& shape, Simplex<T>& s, Vec3<T>& p0, Vec3<T>& p1) const {
    if (s.rank != 1) return false;
    Vec3<T> cur;
    if (update(s, cur) != U_OK) return false;
    T book = cur.norm();
    if (book <= tolerance_) return extract(shape, s, p0, p1);
    Vec3<T> dir = -cur / book;
    const T tol_sq = tolerance_ * tolerance_;
    size_t it = 0;
    while (it < max_iterations_) {
      it += 1;
      MVertex<T> nv{shape.support(dir), dir};
      const T delta = dir.dot(nv.vertex - cur);
      if (delta < tolerance_) return extract(shape, s, p0, p1);
      for (int j = 0; j < s.rank; j++)
        if ((s.vertices[j].vertex - nv.vertex).squaredNorm() < tol_sq) return extract(shape, s, p0, p1);
      s.add(nv);
      const int us = update(s, cur);
      if (us == U_NO_IMPROVEMENT) return extract(shape, s, p0, p1);
      if (us != U_OK) return false;
      const T nd = cur.norm();
      if (book - nd < tolerance_) return extract(shape, s, p0, p1);
      if (nd < tolerance_) return extract(shape, s, p0, p1);
      book = nd;
      dir = -cur / book;
    }
    return false;
  }
};

// ---------------------------------------------------------------------------
// closed-form distance routines (narrowphase/detail/primitive_shape_algorithm/)
template <typename T>
T eps78() {  // math/constants.h:169-172
  return T(std::pow(double(std::numeric_limits<T>::epsilon()), 7. / 8.));
}

// sphere_box-inl.h:59-81,167-205
template <typename T>
bool sphereBoxDistance(T r, const Xform<T>& X_FS, const Vec3<T>& side, const Xform<T>& X_FB, T* dist, Vec3<T>* p_FSb,
                       Vec3<T>* p_FBs) {
  const Vec3<T> p_BC = (X_FB.inverse() * X_FS).t;
  const Vec3<T> half = side / T(2);
  Vec3<T> p_BN;
  bool clamped = false;
  for (int i = 0; i < 3; i++) {
    p_BN[i] = p_BC[i];
    if (p_BC[i] < -half[i]) {
      clamped = true;
      p_BN[i] = -half[i];
    }
    if (p_BC[i] > half[i]) {
      clamped = true;
      p_BN[i] = half[i];
    }
  }
  if (clamped) {
    const Vec3<T> p_NC = p_BC - p_BN;
    const T sq = p_NC.squaredNorm();
    if (sq > r * r) {
      const T d = std::sqrt(sq);
      *dist = d - r;
      *p_FBs = X_FB * p_BN;
      *p_FSb = X_FB * ((p_NC / d) * (d - r) + p_BN);
      return true;
    }
  }
  *dist = -1;
  return false;
}
// sphere_capsule-inl.h:51-67,105-147
template <typename T>
bool sphereCapsuleDistance(T r1, const Xform<T>& tf1, T r2, T lz, const Xform<T>& tf2, T* dist, Vec3<T>* p1,
                           Vec3<T>* p2) {
  const Vec3<T> pos1(0, 0, T(0.5 * lz)), pos2(0, 0, T(-0.5 * lz));
  const Vec3<T> s_c = tf2.inverse() * tf1.t;
  const Vec3<T> v = pos2 - pos1, w = s_c - pos1;
  const T c1 = w.dot(v), c2 = v.dot(v);
  Vec3<T> seg;
  if (c1 <= 0)
    seg = pos1;
  else if (c2 <= c1)
    seg = pos2;
  else
    seg = pos1 + v * (c1 / c2);
  Vec3<T> diff = s_c - seg;
  const T distance = diff.norm() - r1 - r2;
  if (distance <= 0) {
    *dist = -1;
    return false;
  }
  *dist = distance;
  diff.normalize();
  *p1 = tf2 * (s_c - diff * r1);
  *p2 = tf2 * (seg + diff * r2);
  return true;
}
// sphere_cylinder-inl.h:62-93,206-244
template <typename T>
bool sphereCylinderDistance(T r_s, const Xform<T>& X_FS, T radius, T height, const Xform<T>& X_FC, T* dist,
                            Vec3<T>* p_FSc, Vec3<T>* p_FCs) {
  const Vec3<T> p_CS = (X_FC.inverse() * X_FS).t;
  Vec3<T> p_CN = p_CS;
  bool clamped = false;
  const T half_h = height / 2;
  if (p_CS[2] > half_h) {
    clamped = true;
    p_CN[2] = half_h;
  } else if (p_CS[2] < -half_h) {
    clamped = true;
    p_CN[2] = -half_h;
  }
  const T sq_xy = p_CS[0] * p_CS[0] + p_CS[1] * p_CS[1];
  if (sq_xy > radius * radius) {
    clamped = true;
    // unqualified sqrt() in the reference (:85) => double overload for T = float
    const T k = T(double(radius) / std::sqrt(double(sq_xy)));
    p_CN[0] = p_CS[0] * k;
    p_CN[1] = p_CS[1] * k;
  }
  if (clamped) {
    const Vec3<T> p_NS = p_CS - p_CN;
    const T sq = p_NS.squaredNorm();
    if (sq > r_s * r_s) {
      const T d = std::sqrt(sq);
      *dist = d - r_s;
      *p_FCs = X_FC * p_CN;
      *p_FSc = X_FC * (p_CS - (p_NS * r_s / d));
      return true;
    }
  }
  *dist = -1;
  return false;
}
// sphere_sphere-inl.h:72-89
template <typename T>
bool sphereSphereDistance(T r1, const Xform<T>& tf1, T r2, const Xform<T>& tf2, T* dist, Vec3<T>* p1, Vec3<T>* p2) {
  const Vec3<T> o1 = tf1.t, o2 = tf2.t, diff = o1 - o2;
  const T len = diff.norm();
  if (len > r1 + r2) {
    *dist = len - (r1 + r2);
    *p1 = o1 - diff * (r1 / len);
    *p2 = o2 + diff * (r2 / len);
    return true;
  }
  *dist = -1;
  return false;
}
template <typename T>
T clampT(T n, T lo, T hi) {
  if (n < lo) return lo;
  if (n > hi) return hi;
  return n;
}
// capsule_capsule-inl.h:60-139
template <typename T>
T closestPtSegmentSegment(const Vec3<T>& P1, const Vec3<T>& Q1, const Vec3<T>& P2, const Vec3<T>& Q2, Vec3<T>* C1,
                          Vec3<T>* C2) {
  const T kEps = eps78<T>(), kEpsSq = kEps * kEps;
  const Vec3<T> d1 = Q1 - P1, d2 = Q2 - P2, r = P1 - P2;
  const T a = d1.dot(d1), e = d2.dot(d2), f = d2.dot(r);
  T s, t;
  if (a <= kEpsSq && e <= kEpsSq) {
    *C1 = P1;
    *C2 = P2;
    return (*C1 - *C2).squaredNorm();
  }
  if (a <= kEpsSq) {
    s = 0;
    t = clampT(f / e, T(0), T(1));
  } else {
    const T c = d1.dot(r);
    if (e <= kEpsSq) {
      t = 0;
      s = clampT(-c / a, T(0), T(1));
    } else {
      const T b = d1.dot(d2);
      const T denom = std::max(T(0), a * e - b * b);
      s = (denom > kEpsSq) ? clampT((b * f - c * e) / denom, T(0), T(1)) : T(0);
      t = (b * s + f) / e;
      if (t < 0) {
        t = 0;
        s = clampT(-c / a, T(0), T(1));
      } else if (t > 1) {
        t = 1;
        s = clampT((b - c) / a, T(0), T(1));
      }
    }
  }
  *C1 = P1 + d1 * s;
  *C2 = P2 + d2 * t;
  return (*C1 - *C2).squaredNorm();
}
// capsule_capsule-inl.h:141-246
template <typename T>
bool capsuleCapsuleDistance(T r1, T lz1, const Xform<T>& X1, T r2, T lz2, const Xform<T>& X2, T* dist, Vec3<T>* W1,
                            Vec3<T>* W2) {
  const Vec3<T> o1 = X1.t, o2 = X2.t, z1 = X1.R.col(2), z2 = X2.R.col(2);
  const Vec3<T> arm1 = (lz1 / 2) * z1, arm2 = (lz2 / 2) * z2;
  Vec3<T> N1, N2;
  const T sq = closestPtSegmentSegment(o1 + arm1, o1 - arm1, o2 + arm2, o2 - arm2, &N1, &N2);
  const T seg = std::sqrt(sq);
  *dist = seg - r1 - r2;
  const T eps = eps78<T>();
  Vec3<T> vhat;
  if (seg > eps) {
    vhat = (N2 - N1) / seg;
  } else if (std::abs(z1.dot(z2)) < 1 - eps) {
    vhat = z1.cross(z2).normalized();
  } else {
    vhat = X1.R.col(0);
  }
  *W1 = N1 + vhat * r1;
  *W2 = N2 - vhat * r2;
  return true;
}

// ---------------------------------------------------------------------------
template <typename T>
struct ShapeTable {
  std::vector<Geom<T>> geoms;
  std::vector<ConvexData<T>> convex;
  ShapeTable(const ShapeRec* s, int n) {
    convex.reserve(convexSources().size());
    for (const auto& src : convexSources()) convex.push_back(buildConvex<T>(src));
    for (int i = 0; i < n; i++) {
      Geom<T> g;
      g.type = int(s[i].type);
      g.data = Vec3<T>(T(s[i].p[0]), T(s[i].p[1]), T(s[i].p[2]));
      if (g.type == T_CONVEX) g.cvx = &convex.at(s[i].geom);
      geoms.push_back(g);
    }
  }
};

template <typename F>
void parallelFor(size_t n, int n_threads, F&& f) {
  if (n_threads <= 1 || n < 2) {
    f(size_t(0), n);
    return;
  }
  std::vector<std::thread> ts;
  const size_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const size_t b = std::min(n, chunk * t), e = std::min(n, chunk * (t + 1));
    if (b >= e) break;
    ts.emplace_back([=, &f] { f(b, e); });
  }
  for (auto& t : ts) t.join();
}

// GJKSolver<S>::shapeDistance, narrowphase/detail/gjk_solver-inl.h:762-808,902-988
template <typename T>
bool shapeDistance(const Geom<T>& g1, const Xform<T>& tf1, const Geom<T>& g2, const Xform<T>& tf2, T gjk_tol,
                   size_t gjk_it, T* dist, Vec3<T>* p1, Vec3<T>* p2) {
  const int a = g1.type, b = g2.type;
  if (a == T_SPHERE && b == T_BOX) return sphereBoxDistance(g1.data[0], tf1, g2.data, tf2, dist, p1, p2);
  if (a == T_BOX && b == T_SPHERE) return sphereBoxDistance(g2.data[0], tf2, g1.data, tf1, dist, p2, p1);
  if (a == T_SPHERE && b == T_CAPSULE)
    return sphereCapsuleDistance(g1.data[0], tf1, g2.data[0], g2.data[1], tf2, dist, p1, p2);
  if (a == T_CAPSULE && b == T_SPHERE)
    return sphereCapsuleDistance(g2.data[0], tf2, g1.data[0], g1.data[1], tf1, dist, p2, p1);
  if (a == T_SPHERE && b == T_CYLINDER)
    return sphereCylinderDistance(g1.data[0], tf1, g2.data[0], g2.data[1], tf2, dist, p1, p2);
  if (a == T_CYLINDER && b == T_SPHERE)
    return sphereCylinderDistance(g2.data[0], tf2, g1.data[0], g1.data[1], tf1, dist, p2, p1);
  if (a == T_SPHERE && b == T_SPHERE) return sphereSphereDistance(g1.data[0], tf1, g2.data[0], tf2, dist, p1, p2);
  if (a == T_CAPSULE && b == T_CAPSULE)
    return capsuleCapsuleDistance(g1.data[0], g1.data[1], tf1, g2.data[0], g2.data[1], tf2, dist, p1, p2);
  // ShapeDistanceIndepImpl::run, :762-798
  MinkowskiDiff<T> shape;
  shape.shapes[0] = g1;
  shape.shapes[1] = g2;
  shape.setPoses(tf1, tf2);
  Gjk<T> gjk(gjk_it, gjk_tol);
  Simplex<T> simplex;
  typename Gjk<T>::DistOut out;
  const Vec3<T> guess(1, 0, 0);
  const GjkStatus st = gjk.evaluate(shape, simplex, -guess, &out);
  if (st == GJK_SEPARATED) {
    *p1 = tf1 * out.p0;
    *p2 = tf1 * out.p1;
    *dist = (out.p0 - out.p1).norm();
    return true;
  }
  *dist = -1;
  return false;
}

template <typename T>
int distanceBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const T* poses1, const T* poses2, size_t n,
                  double gjk_tol, uint32_t gjk_max_iter, T* dist, T* p1, T* p2, uint8_t* ok, int threads) {
  const ShapeTable<T> tab(shapes, n_shapes);
  const T tol = gjk_tol > 0 ? T(gjk_tol) : eps78<T>();  // gjk_solver-inl.h:1121-1130
  const size_t its = gjk_max_iter ? gjk_max_iter : 128;
  parallelFor(n, threads, [&](size_t b, size_t e) {
    for (size_t q = b; q < e; q++) {
      const Xform<T> tf1 = loadPose(poses1 + 12 * q), tf2 = loadPose(poses2 + 12 * q);
      T d = 0;
      Vec3<T> a, c;
      const bool r = shapeDistance(tab.geoms[pairs[q].s1], tf1, tab.geoms[pairs[q].s2], tf2, tol, its, &d, &a, &c);
      if (dist) dist[q] = d;
      if (ok) ok[q] = r ? 1 : 0;
      for (int k = 0; k < 3; k++) {
        if (p1) p1[3 * q + k] = a[k];
        if (p2) p2[3 * q + k] = c[k];
      }
    }
  });
  return 0;
}

// The generic GJK path of every query (closed forms bypassed), with the work counters on: boolean GJK with the solver's
// guess, plus the distance refinement + witness extraction when want_distance.
template <typename T>
int gjkWorkCounters(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const T* poses1, const T* poses2, size_t n,
                    double gjk_tol, uint32_t gjk_max_iter, int want_distance, uint64_t* out, int threads) {
  const ShapeTable<T> tab(shapes, n_shapes);
  const T tol = gjk_tol > 0 ? T(gjk_tol) : eps78<T>();
  const size_t its = gjk_max_iter ? gjk_max_iter : 128;
  std::vector<WorkCounters> per(size_t(std::max(1, threads)));
  std::atomic<int> next{0};
  parallelFor(n, threads, [&](size_t b, size_t e) {
    WorkCounters& c = per[size_t(next.fetch_add(1)) % per.size()];
    t_cnt = &c;
    for (size_t q = b; q < e; q++) {
      const Xform<T> tf1 = loadPose(poses1 + 12 * q), tf2 = loadPose(poses2 + 12 * q);
      MinkowskiDiff<T> shape;
      shape.shapes[0] = tab.geoms[pairs[q].s1];
      shape.shapes[1] = tab.geoms[pairs[q].s2];
      shape.setPoses(tf1, tf2);
      Gjk<T> gjk(its, tol);
      Simplex<T> simplex;
      typename Gjk<T>::DistOut d;
      const Vec3<T> guess(1, 0, 0);
      const GjkStatus st = gjk.evaluate(shape, simplex, -guess, want_distance ? &d : nullptr);
      c.queries += 1;
      if (st == GJK_SEPARATED) c.separated += 1;
    }
    t_cnt = nullptr;
  });
  WorkCounters tot;
  for (const auto& c : per) tot.add(c);
  out[0] = tot.queries;
  out[1] = tot.support_vertices;
  out[2] = tot.extract_supports;
  out[3] = tot.convex_dots;
  for (int k = 0; k < 5; k++) out[4 + k] = tot.project[k];
  for (int k = 0; k < 5; k++) out[9 + k] = tot.update[k];
  out[14] = tot.separated;
  out[15] = 0;
  return 0;
}

}  // namespace orc

extern "C" {

int fclport_register_convex(const double* verts, int n_verts, const int* faces, int faces_len, int num_faces) {
  orc::ConvexSrc c;
  c.verts.assign(verts, verts + 3 * size_t(n_verts));
  c.faces.assign(faces, faces + faces_len);
  c.num_faces = num_faces;
  orc::convexSources().push_back(std::move(c));
  return int(orc::convexSources().size()) - 1;
}

int fclport_distance_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs, const void* poses1,
                           const void* poses2, size_t n, double gjk_tol, uint32_t gjk_max_iter, void* dist, void* p1,
                           void* p2, uint8_t* ok, int n_threads) {
  using namespace orc;
  if (scalar_type == 0)
    return distanceBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                                (const float*)poses2, n, gjk_tol, gjk_max_iter, (float*)dist, (float*)p1, (float*)p2,
                                ok, n_threads);
  return distanceBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                               (const double*)poses2, n, gjk_tol, gjk_max_iter, (double*)dist, (double*)p1,
                               (double*)p2, ok, n_threads);
}

// out[16]: queries, supportVertex evaluations, extraction supports, convex dot products, project calls by rank [5],
// distance-update calls by rank [5], separated queries, 0
int fclport_gjk_work_counters(int scalar_type, const void* shapes, int n_shapes, const void* pairs, const void* poses1,
                              const void* poses2, size_t n, double gjk_tol, uint32_t gjk_max_iter, int want_distance,
                              uint64_t* out, int n_threads) {
  using namespace orc;
  if (scalar_type == 0)
    return gjkWorkCounters<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                                  (const float*)poses2, n, gjk_tol, gjk_max_iter, want_distance, out, n_threads);
  return gjkWorkCounters<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                                 (const double*)poses2, n, gjk_tol, gjk_max_iter, want_distance, out, n_threads);
}

int fclport_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

}  // extern "C"
