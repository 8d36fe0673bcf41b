// fclb_bvh.cuh -- device-side BVH node / triangle records and the OBB overlap
// predicate shared by the mesh-mesh (fclb_bvh.cu) and mesh-shape
// (fclb_bvh_shape.cu) traversal kernels.
//   overlap(R0,T0,b1,b2) + obbDisjoint      math/bv/OBB-inl.h:305-436
//   BVNodeBase child / primitive encoding  geometry/bvh/BV_node_base.h:50-82
#pragma once
#include <map>
#include <vector>

#include "fclb_engine.h"
#include "fclb_math.cuh"

namespace fclb {

struct BvhDev {
  void* nodes = nullptr;  // 16 S per node: axis[9] row-major, To[3], extent[3], first_child bits
  void* tris = nullptr;   // 12 S per triangle (3 x {x,y,z,pad})
  int n_nodes = 0, n_tris = 0;
  int scalar_type = 0;
  // host copy in the upload layout (fclb_bvh_export)
  std::vector<unsigned char> h_obb, h_tri;
  std::vector<int32_t> h_child;
  // refit (fclb_bvh_refit_*): BVNodeBase::first_primitive / num_primitives of every node and primitive_indices_, derived
  // from the child links (the leaves of a subtree, left to right, are its slice of primitive_indices_)
  int2* d_range = nullptr;  // [n_nodes] (first, count)
  int* d_prim = nullptr;    // [n_tris]
  int *d_parent = nullptr, *d_leaf_node = nullptr, *d_arrived = nullptr;  // bottom-up refit: parent links, leaf of every triangle, arrival counters
  int* d_dfs_rank = nullptr;  // [n_tris] position of the triangle's leaf in a right-child-first depth-first walk (CCD contact order)
};

std::map<fclb_handle, BvhDev*>& bvhTable();  // fclb_bvh.cu

template <typename S>
struct NodeD {
  M3<S> axis;
  V3<S> To, extent;
  int first_child;
};

FCLB_DI NodeD<float> loadNode(const float* __restrict__ base, int i) {
  const float4* p = reinterpret_cast<const float4*>(base) + 4 * size_t(i);
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  NodeD<float> n;
  n.axis.m[0] = a.x; n.axis.m[1] = a.y; n.axis.m[2] = a.z; n.axis.m[3] = a.w;
  n.axis.m[4] = b.x; n.axis.m[5] = b.y; n.axis.m[6] = b.z; n.axis.m[7] = b.w;
  n.axis.m[8] = c.x;
  n.To = mk<float>(c.y, c.z, c.w);
  n.extent = mk<float>(d.x, d.y, d.z);
  n.first_child = __float_as_int(d.w);
  return n;
}
FCLB_DI NodeD<double> loadNode(const double* __restrict__ base, int i) {
  const double2* p = reinterpret_cast<const double2*>(base) + 8 * size_t(i);
  double2 v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = __ldg(p + k);
  NodeD<double> n;
  n.axis.m[0] = v[0].x; n.axis.m[1] = v[0].y; n.axis.m[2] = v[1].x; n.axis.m[3] = v[1].y;
  n.axis.m[4] = v[2].x; n.axis.m[5] = v[2].y; n.axis.m[6] = v[3].x; n.axis.m[7] = v[3].y;
  n.axis.m[8] = v[4].x;
  n.To = mk<double>(v[4].y, v[5].x, v[5].y);
  n.extent = mk<double>(v[6].x, v[6].y, v[7].x);
  n.first_child = int(__double_as_longlong(v[7].y));
  return n;
}
FCLB_DI void loadTri(const float* __restrict__ base, int t, V3<float> p[3]) {
  const float4* q = reinterpret_cast<const float4*>(base) + 3 * size_t(t);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float4 v = __ldg(q + k);
    p[k] = mk<float>(v.x, v.y, v.z);
  }
}
FCLB_DI void loadTri(const double* __restrict__ base, int t, V3<double> p[3]) {
  const double2* q = reinterpret_cast<const double2*>(base) + 6 * size_t(t);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double2 a = __ldg(q + 2 * k), b = __ldg(q + 2 * k + 1);
    p[k] = mk<double>(a.x, a.y, b.x);
  }
}

// obbDisjoint (math/bv/OBB-inl.h:319-436): 15-axis SAT, reps added to |B|
template <typename S>
FCLB_DI bool obbDisjoint(const M3<S>& B, const V3<S>& T, const V3<S>& a, const V3<S>& b) {
  S t, s;
  const S reps = S(1e-6);
  M3<S> Bf;
#pragma unroll
  for (int i = 0; i < 9; i++) Bf.m[i] = fabs_(B.m[i]) + reps;
  t = (T.x < 0) ? -T.x : T.x;
  if (t > (a.x + dot(row(Bf, 0), b))) return true;
  s = dot(col(B, 0), T);
  t = (s < 0) ? -s : s;
  if (t > (b.x + dot(col(Bf, 0), a))) return true;
  t = (T.y < 0) ? -T.y : T.y;
  if (t > (a.y + dot(row(Bf, 1), b))) return true;
  t = (T.z < 0) ? -T.z : T.z;
  if (t > (a.z + dot(row(Bf, 2), b))) return true;
  s = dot(col(B, 1), T);
  t = (s < 0) ? -s : s;
  if (t > (b.y + dot(col(Bf, 1), a))) return true;
  s = dot(col(B, 2), T);
  t = (s < 0) ? -s : s;
  if (t > (b.z + dot(col(Bf, 2), a))) return true;
#define FCLB_OBB_EDGE(SEXPR, RAD) \
  s = (SEXPR);                    \
  t = (s < 0) ? -s : s;           \
  if (t > (RAD)) return true;
  FCLB_OBB_EDGE(T.z * B(1, 0) - T.y * B(2, 0), a.y * Bf(2, 0) + a.z * Bf(1, 0) + b.y * Bf(0, 2) + b.z * Bf(0, 1))
  FCLB_OBB_EDGE(T.z * B(1, 1) - T.y * B(2, 1), a.y * Bf(2, 1) + a.z * Bf(1, 1) + b.x * Bf(0, 2) + b.z * Bf(0, 0))
  FCLB_OBB_EDGE(T.z * B(1, 2) - T.y * B(2, 2), a.y * Bf(2, 2) + a.z * Bf(1, 2) + b.x * Bf(0, 1) + b.y * Bf(0, 0))
  FCLB_OBB_EDGE(T.x * B(2, 0) - T.z * B(0, 0), a.x * Bf(2, 0) + a.z * Bf(0, 0) + b.y * Bf(1, 2) + b.z * Bf(1, 1))
  FCLB_OBB_EDGE(T.x * B(2, 1) - T.z * B(0, 1), a.x * Bf(2, 1) + a.z * Bf(0, 1) + b.x * Bf(1, 2) + b.z * Bf(1, 0))
  FCLB_OBB_EDGE(T.x * B(2, 2) - T.z * B(0, 2), a.x * Bf(2, 2) + a.z * Bf(0, 2) + b.x * Bf(1, 1) + b.y * Bf(1, 0))
  FCLB_OBB_EDGE(T.y * B(0, 0) - T.x * B(1, 0), a.x * Bf(1, 0) + a.y * Bf(0, 0) + b.y * Bf(2, 2) + b.z * Bf(2, 1))
  FCLB_OBB_EDGE(T.y * B(0, 1) - T.x * B(1, 1), a.x * Bf(1, 1) + a.y * Bf(0, 1) + b.x * Bf(2, 2) + b.z * Bf(2, 0))
  FCLB_OBB_EDGE(T.y * B(0, 2) - T.x * B(1, 2), a.x * Bf(1, 2) + a.y * Bf(0, 2) + b.x * Bf(2, 1) + b.y * Bf(2, 0))
#undef FCLB_OBB_EDGE
  return false;
}

// overlap(R0, T0, b1, b2) (math/bv/OBB-inl.h:305-316)
template <typename S>
FCLB_DI bool obbOverlap(const M3<S>& R0, const V3<S>& T0, const NodeD<S>& b1, const NodeD<S>& b2) {
  const M3<S> R0b2 = mulMM(R0, b2.axis);
  const M3<S> R = mulMtM(b1.axis, R0b2);
  const V3<S> Ttemp = (mulMV(R0, b2.To) + T0) - b1.To;
  const V3<S> T = mulMtV(b1.axis, Ttemp);
  return !obbDisjoint(R, T, b1.extent, b2.extent);
}

}  // namespace fclb
