// fclb_bvh_build.h -- HOST builder for the flattened BVHModel<OBBRSS<S>> tree.
//
// SURVEY.md 8(a17) keeps tree construction on the host; this is the host-side
// mirror of BVHModel::beginModel / addSubModel / endModel so that a user of
// include/fcl_b200/fcl.h (and bench.py) can go from a triangle soup to a device
// tree without mind-fcl.  The OBB half of every node is bit-identical to what
// the reference's builder produces on the same vertices (tests/test_bvh_build.py
// compares against the tree exported from oracle/_ref), because every
// expression below keeps the reference's operand order and literal types:
//   tree loop + split            geometry/bvh/BVH_model-inl.h:402-570
//   OBBRSS fitter (OBB half)     geometry/bvh/detail/BV_fitter-inl.h:324-345
//   covariance / extent          math/geometry-inl.h:148-180,713-780
//   Jacobi eigen solver          math/geometry-inl.h:237-336 (eigen_old)
//   axis choice                  math/geometry-inl.h:340-366 (axisFromEigen)
//   mean split rule              geometry/bvh/detail/BV_splitter-inl.h:361-372,411-415,449-453,470-500
// The RSS half of OBBRSS is never read by collide (math/bv/OBBRSS-inl.h:130-135)
// and is not built.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#ifdef __CUDACC__
#define FCLB_HD __host__ __device__
#else
#define FCLB_HD
#endif

namespace fclb {
namespace hostbuild {

template <typename S>
struct P3 {
  S v[3];
  S operator[](int i) const { return v[i]; }
  S& operator[](int i) { return v[i]; }
};

// Symmetric 3x3 Jacobi sweep solver.  Output: d = eigenvalues, vec[r][c] =
// component r of eigenvector c.  Returns false when 50 sweeps did not
// converge (the reference then leaves its outputs unset).
// The literal types (0.2, 100.0, 0.5, 1.0 are double; 1 is int) decide where a
// float instantiation computes in double -- they are kept on purpose.
template <typename S>
FCLB_HD bool jacobi3(const S m[3][3], S d[3], S vec[3][3]) {
  S R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[i][j] = m[i][j];
      vec[i][j] = (i == j) ? S(1) : S(0);
    }
  S b[3], z[3];
  for (int k = 0; k < 3; k++) {
    b[k] = d[k] = R[k][k];
    z[k] = 0;
  }
  const int n = 3;
  for (int sweep = 0; sweep < 50; ++sweep) {
    S sm = 0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) sm += std::abs(R[p][q]);
    if (sm == 0.0) return true;
    S tresh;
    if (sweep < 3)
      tresh = 0.2 * sm / (n * n);
    else
      tresh = 0.0;
    for (int p = 0; p < n; ++p) {
      for (int q = p + 1; q < n; ++q) {
        S g = 100.0 * std::abs(R[p][q]);
        if (sweep > 3 && std::abs(d[p]) + g == std::abs(d[p]) && std::abs(d[q]) + g == std::abs(d[q])) {
          R[p][q] = 0.0;
        } else if (std::abs(R[p][q]) > tresh) {
          S h = d[q] - d[p];
          S t;
          if (std::abs(h) + g == std::abs(h)) {
            t = (R[p][q]) / h;
          } else {
            const S theta = 0.5 * h / (R[p][q]);
            t = 1.0 / (std::abs(theta) + std::sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          const S c = 1.0 / std::sqrt(1 + t * t);
          const S s = t * c;
          const S tau = s / (1.0 + c);
          h = t * R[p][q];
          z[p] -= h;
          z[q] += h;
          d[p] -= h;
          d[q] += h;
          R[p][q] = 0.0;
          auto rot = [&](S& x, S& y) {
            const S gx = x, hy = y;
            x = gx - s * (hy + gx * tau);
            y = hy + s * (gx - hy * tau);
          };
          for (int j = 0; j < p; ++j) rot(R[j][p], R[j][q]);
          for (int j = p + 1; j < q; ++j) rot(R[p][j], R[j][q]);
          for (int j = q + 1; j < n; ++j) rot(R[p][j], R[q][j]);
          for (int j = 0; j < n; ++j) rot(vec[j][p], vec[j][q]);
        }
      }
    }
    for (int k = 0; k < n; ++k) {
      b[k] += z[k];
      d[k] = b[k];
      z[k] = 0.0;
    }
  }
  return false;
}

// axisFromEigen (Matrix3 overload): columns 0/1 = eigenvectors of the largest /
// middle eigenvalue, column 2 = their cross product.  axis is row-major 3x3.
template <typename S>
FCLB_HD void axesFromEigen(const S vec[3][3], const S d[3], S axis[9]) {
  int mn, md, mx;
  if (d[0] > d[1]) {
    mx = 0;
    mn = 1;
  } else {
    mn = 0;
    mx = 1;
  }
  if (d[2] < d[mn]) {
    md = mn;
    mn = 2;
  } else if (d[2] > d[mx]) {
    md = mx;
    mx = 2;
  } else {
    md = 2;
  }
  (void)mn;
  S a0[3], a1[3];
  for (int r = 0; r < 3; r++) {
    a0[r] = vec[r][mx];
    a1[r] = vec[r][md];
  }
  const S a2[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
  for (int r = 0; r < 3; r++) {
    axis[3 * r + 0] = a0[r];
    axis[3 * r + 1] = a1[r];
    axis[3 * r + 2] = a2[r];
  }
}

template <typename S>
struct TreeOut {
  std::vector<S> obb;                // 15 S per node
  std::vector<int32_t> first_child;  // per node
  std::vector<S> tri;                // 9 S per triangle
};

template <typename S>
void buildObbTree(const double* verts_d, int n_verts, const int32_t* tris, int n_tris, TreeOut<S>& out) {
  std::vector<P3<S>> ps(n_verts);
  for (int i = 0; i < n_verts; i++)
    for (int k = 0; k < 3; k++) ps[i][k] = S(verts_d[3 * size_t(i) + k]);
  out.tri.resize(size_t(9) * n_tris);
  for (int t = 0; t < n_tris; t++)
    for (int v = 0; v < 3; v++)
      for (int k = 0; k < 3; k++) out.tri[size_t(9) * t + 3 * v + k] = ps[tris[3 * size_t(t) + v]][k];
  std::vector<unsigned> prim(n_tris);
  for (int i = 0; i < n_tris; i++) prim[i] = unsigned(i);
  const size_t n_nodes_max = size_t(2) * n_tris - 1;
  out.obb.assign(15 * n_nodes_max, S(0));
  out.first_child.assign(n_nodes_max, 0);
  size_t n_nodes = 1;

  struct Task {
    int node, first, count;
  };
  std::vector<Task> stack;
  stack.push_back({0, 0, n_tris});
  while (!stack.empty()) {
    const Task task = stack.back();
    stack.pop_back();
    unsigned* idx = prim.data() + task.first;
    // ---- fit: covariance of the triangle corners
    S S1[3] = {0, 0, 0};
    S c00 = 0, c11 = 0, c22 = 0, c01 = 0, c02 = 0, c12 = 0;
    int n_points = 0;
    for (int i = 0; i < task.count; ++i) {
      const int32_t* t = tris + 3 * size_t(idx[i]);
      const P3<S>&p1 = ps[t[0]], &p2 = ps[t[1]], &p3 = ps[t[2]];
      for (int k = 0; k < 3; k++) S1[k] += (p1[k] + p2[k]) + p3[k];
      c00 += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      c11 += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      c22 += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      c01 += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      c02 += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      c12 += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
      n_points += 3;
    }
    S M[3][3];
    M[0][0] = c00 - S1[0] * S1[0] / n_points;
    M[1][1] = c11 - S1[1] * S1[1] / n_points;
    M[2][2] = c22 - S1[2] * S1[2] / n_points;
    M[0][1] = c01 - S1[0] * S1[1] / n_points;
    M[1][2] = c12 - S1[1] * S1[2] / n_points;
    M[0][2] = c02 - S1[0] * S1[2] / n_points;
    M[1][0] = M[0][1];
    M[2][0] = M[0][2];
    M[2][1] = M[1][2];
    S d[3] = {0, 0, 0}, vec[3][3];
    if (!jacobi3<S>(M, d, vec)) {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
    }
    S* node = out.obb.data() + 15 * size_t(task.node);
    axesFromEigen<S>(vec, d, node);
    const S* ax = node;
    // ---- extent and centre along those axes
    const S big = std::numeric_limits<S>::max();
    S mn[3] = {big, big, big}, mxv[3] = {-big, -big, -big};
    for (int i = 0; i < task.count; ++i) {
      const int32_t* t = tris + 3 * size_t(idx[i]);
      for (int j = 0; j < 3; j++) {
        const P3<S>& p = ps[t[j]];
        for (int k = 0; k < 3; k++) {
          const S proj = (ax[0 + k] * p[0] + ax[3 + k] * p[1]) + ax[6 + k] * p[2];
          if (proj > mxv[k]) mxv[k] = proj;
          if (proj < mn[k]) mn[k] = proj;
        }
      }
    }
    S o[3];
    for (int k = 0; k < 3; k++) o[k] = (mxv[k] + mn[k]) / 2;
    for (int r = 0; r < 3; r++) node[9 + r] = (ax[3 * r] * o[0] + ax[3 * r + 1] * o[1]) + ax[3 * r + 2] * o[2];
    for (int k = 0; k < 3; k++) node[12 + k] = (mxv[k] - mn[k]) / 2;

    if (task.count == 1) {
      out.first_child[task.node] = -(int(idx[0]) + 1);
      continue;
    }
    // ---- mean split along the first axis
    const S sv[3] = {ax[0], ax[3], ax[6]};
    S c[3] = {0, 0, 0};
    for (int i = 0; i < task.count; ++i) {
      const int32_t* t = tris + 3 * size_t(idx[i]);
      const P3<S>&p1 = ps[t[0]], &p2 = ps[t[1]], &p3 = ps[t[2]];
      for (int k = 0; k < 3; k++) c[k] += (p1[k] + p2[k] + p3[k]) / 3;
    }
    const S split_value = (c[0] * sv[0] + c[1] * sv[1] + c[2] * sv[2]) / task.count;
    int c1 = 0;
    for (int i = 0; i < task.count; ++i) {
      const int32_t* t = tris + 3 * size_t(idx[i]);
      const P3<S>&p1 = ps[t[0]], &p2 = ps[t[1]], &p3 = ps[t[2]];
      S p[3];
      for (int k = 0; k < 3; k++) p[k] = ((p1[k] + p2[k]) + p3[k]) / S(3.0);
      const bool right = ((sv[0] * p[0] + sv[1] * p[1]) + sv[2] * p[2]) > split_value;
      if (!right) {
        const unsigned tmp = idx[i];
        idx[i] = idx[c1];
        idx[c1] = tmp;
        c1++;
      }
    }
    if (c1 == 0 || c1 == task.count) c1 = task.count / 2;
    const int left = int(n_nodes), rightn = int(n_nodes) + 1;
    n_nodes += 2;
    out.first_child[task.node] = left;
    stack.push_back({rightn, task.first + c1, task.count - c1});
    stack.push_back({left, task.first, c1});
  }
  out.obb.resize(15 * n_nodes);
  out.first_child.resize(n_nodes);
}

}  // namespace hostbuild
}  // namespace fclb
