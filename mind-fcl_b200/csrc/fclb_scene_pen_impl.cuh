// fclb_scene_pen_impl.cuh -- MPR penetration for the contacts of a scene-vs-shape query.
//
// collisionPenetrationMPR (narrowphase/collision_penetration-inl.h:189-252) runs the boolean collide and,
// for every reported contact, rebuilds the two leaf geometries (penetrationDistanceGetContactGJK, :34-95:
// mesh -> the contact's triangle in the mesh pose; heightmap / octree -> a Box of the contact's o1_bv at
// tf_geom translated by R * center; shape -> itself) and calls computePenetrationMPR (:107-186).
// Pass 1 (the traversal kernels) stored the leaf ids / boxes of each query's first max_keep contacts;
// this pass runs one thread per stored contact.
#pragma once
#include "fclb_internal.h"
#include "fclb_mpr_pen.cuh"

namespace fclb {

template <typename S>
__global__ void __launch_bounds__(kBlock) scenePenetrationKernel(ScenePenArgs a) {
  const size_t total = a.n * size_t(a.max_keep);
  const V3<S> dir_world = mk<S>(S(a.dir[0]), S(a.dir[1]), S(a.dir[2]));
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = i / a.max_keep;
    const uint32_t k = uint32_t(i % a.max_keep);
    S* out = static_cast<S*>(a.out_contacts) + i * 7;
    if (k >= a.counts[q]) {
#pragma unroll
      for (int j = 0; j < 7; j++) out[j] = S(0);
      continue;
    }
    const Pose<S> tf_scene = loadPose(static_cast<const S*>(a.poses_scene), q);
    const Pose<S> tf_shape = loadPose(static_cast<const S*>(a.poses_shape), q);
    MinkDiff<S, ST_DYNAMIC, ST_DYNAMIC> md;
    md.s1 = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), a.shape_ids[q]);
    md.s0.cvx = nullptr;
    Pose<S> tf1 = tf_scene;
    if (a.leaf_is_triangle) {
      const S* t = static_cast<const S*>(a.tris) + size_t(12) * size_t(a.b1[i]);
      md.s0.type = ST_TRIANGLE;
      md.s0.p0 = md.s0.p1 = md.s0.p2 = S(0);
#pragma unroll
      for (int v = 0; v < 3; v++) md.s0.tri[v] = mk<S>(t[4 * v], t[4 * v + 1], t[4 * v + 2]);
    } else {
      const S* b = static_cast<const S*>(a.box) + i * 6;
      const V3<S> mn = mk<S>(b[0], b[1], b[2]), mx = mk<S>(b[3], b[4], b[5]);
      const V3<S> side = mx - mn;
      const V3<S> center = (mn + mx) * S(0.5);
      md.s0.type = ST_BOX;
      md.s0.p0 = side.x;
      md.s0.p1 = side.y;
      md.s0.p2 = side.z;
      tf1.t = tf_scene.t + mulMV(tf_scene.R, center);  // tf.translation() += tf.linear() * center
    }
    md.setPoses(tf1, tf_shape);
    V3<S> pos, normal;
    S depth;
    computePenetrationMpr<S>(md, tf1, dir_world, a.incremental != 0, 128, S(a.tol), pos, normal, depth);
    out[0] = normal.x; out[1] = normal.y; out[2] = normal.z;
    out[3] = pos.x; out[4] = pos.y; out[5] = pos.z;
    out[6] = depth;
  }
}

template <typename S>
cudaError_t launchScenePenetration(const ScenePenArgs& a, cudaStream_t st) {
  const size_t total = a.n * size_t(a.max_keep);
  if (total == 0) return cudaSuccess;
  size_t grid = (total + kBlock - 1) / kBlock;
  if (grid > 148 * 16) grid = 148 * 16;
  scenePenetrationKernel<S><<<int(grid), kBlock, 0, st>>>(a);
  return cudaGetLastError();
}

// The same for the contacts of a scene-vs-scene query (fclb_scene_pair_impl.cuh): side 1 is always a pixel / voxel
// box (Contact::o1_bv at tf1), side 2 a box (o2_bv at tf2) or the mesh triangle b2 in the mesh pose.
template <typename S>
__global__ void __launch_bounds__(kBlock) scenePairPenetrationKernel(ScenePairPenArgs a) {
  const size_t total = a.n * size_t(a.max_keep);
  const V3<S> dir_world = mk<S>(S(a.dir[0]), S(a.dir[1]), S(a.dir[2]));
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = i / a.max_keep;
    const uint32_t k = uint32_t(i % a.max_keep);
    S* out = static_cast<S*>(a.out_contacts) + i * 7;
    if (k >= a.counts[q]) {
#pragma unroll
      for (int j = 0; j < 7; j++) out[j] = S(0);
      continue;
    }
    const Pose<S> tf_a = loadPose(static_cast<const S*>(a.poses1), q);
    const Pose<S> tf_b = loadPose(static_cast<const S*>(a.poses2), q);
    MinkDiff<S, ST_DYNAMIC, ST_DYNAMIC> md;
    md.s0.cvx = nullptr;
    md.s1.cvx = nullptr;
    auto boxGeom = [](const S* b, const Pose<S>& tf, ShapeInst<S>& g, Pose<S>& tf_g) {
      const V3<S> mn = mk<S>(b[0], b[1], b[2]), mx = mk<S>(b[3], b[4], b[5]);
      const V3<S> side = mx - mn;
      const V3<S> center = (mn + mx) * S(0.5);
      g.type = ST_BOX;
      g.p0 = side.x;
      g.p1 = side.y;
      g.p2 = side.z;
      tf_g = tf;
      tf_g.t = tf.t + mulMV(tf.R, center);  // tf.translation() += tf.linear() * center
    };
    Pose<S> tf1, tf2 = tf_b;
    boxGeom(static_cast<const S*>(a.box1) + i * 6, tf_a, md.s0, tf1);
    if (a.leaf2_is_triangle) {
      const S* t = static_cast<const S*>(a.tris) + size_t(12) * size_t(a.b2[i]);
      md.s1.type = ST_TRIANGLE;
      md.s1.p0 = md.s1.p1 = md.s1.p2 = S(0);
#pragma unroll
      for (int v = 0; v < 3; v++) md.s1.tri[v] = mk<S>(t[4 * v], t[4 * v + 1], t[4 * v + 2]);
    } else {
      boxGeom(static_cast<const S*>(a.box2) + i * 6, tf_b, md.s1, tf2);
    }
    md.setPoses(tf1, tf2);
    V3<S> pos, normal;
    S depth;
    computePenetrationMpr<S>(md, tf1, dir_world, a.incremental != 0, 128, S(a.tol), pos, normal, depth);
    out[0] = normal.x; out[1] = normal.y; out[2] = normal.z;
    out[3] = pos.x; out[4] = pos.y; out[5] = pos.z;
    out[6] = depth;
  }
}

template <typename S>
cudaError_t launchScenePairPenetration(const ScenePairPenArgs& a, cudaStream_t st) {
  const size_t total = a.n * size_t(a.max_keep);
  if (total == 0) return cudaSuccess;
  size_t grid = (total + kBlock - 1) / kBlock;
  if (grid > 148 * 16) grid = 148 * 16;
  scenePairPenetrationKernel<S><<<int(grid), kBlock, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace fclb
