// oracle/ref_harness_scene.cpp -- TEST INFRASTRUCTURE (the "reference" oracle,
// shape-vs-scene part).  Compiled into oracle/_ref/libfclref.so from the
// UNMODIFIED reference headers + oracle/eigen_shim.
//
//   * batched fcl::collide(BVHModel<OBBRSS>, tf1, Shape, tf2)
//     (-> OrientedNodeBVHSolver::MeshShapeIntersect, traversal/collision/bvh_solver-inl.h:8)
#include <climits>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#include "fcl/fcl.h"

namespace fclref {
std::shared_ptr<fcl::ShapeBase<float>> makeShapeF(const void* rec);
std::shared_ptr<fcl::ShapeBase<double>> makeShapeD(const void* rec);
const fcl::BVHModel<fcl::OBBRSS<float>>* meshF(int id);
const fcl::BVHModel<fcl::OBBRSS<double>>* meshD(int id);
}  // namespace fclref

namespace {

struct ShapeRec {
  uint32_t type;
  uint32_t geom;
  double p[3];
};
struct RequestRec {
  uint32_t max_contacts;
  uint32_t penetration_mode;
  double dir[3];
  double binary_tol, distance_tol;
  uint32_t gjk_max_iter, epa_max_faces, epa_max_iter;
  uint32_t flags;
};

template <typename S>
struct Sel;
template <>
struct Sel<float> {
  static std::shared_ptr<fcl::ShapeBase<float>> shape(const ShapeRec* r) { return fclref::makeShapeF(r); }
  static const fcl::BVHModel<fcl::OBBRSS<float>>* mesh(int id) { return fclref::meshF(id); }
};
template <>
struct Sel<double> {
  static std::shared_ptr<fcl::ShapeBase<double>> shape(const ShapeRec* r) { return fclref::makeShapeD(r); }
  static const fcl::BVHModel<fcl::OBBRSS<double>>* mesh(int id) { return fclref::meshD(id); }
};

template <typename S>
fcl::Transform3<S> loadPose(const S* p) {
  fcl::Transform3<S> tf;
  tf.setIdentity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) tf.linear()(i, j) = p[3 * i + j];
  for (int i = 0; i < 3; i++) tf.translation()[i] = p[9 + i];
  return tf;
}

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

template <typename S>
fcl::CollisionRequest<S> makeRequest(const RequestRec* rq) {
  fcl::CollisionRequest<S> req(rq->max_contacts);
  if (rq->penetration_mode == 1)
    req.useDefaultPenetration();
  else
    req.disablePenetration();
  return req;
}

template <typename S>
void meshShapeBatch(int mesh_id, const ShapeRec* shapes, uint32_t n_shapes, const uint32_t* shape_ids, const S* poses_mesh,
                    const S* poses_shape, size_t n, const RequestRec* rq, uint32_t* counts, int32_t* first_tri,
                    int threads) {
  const auto* mesh = Sel<S>::mesh(mesh_id);
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> objs;
  for (uint32_t i = 0; i < n_shapes; i++) objs.push_back(Sel<S>::shape(shapes + i));
  parallelFor(n, threads, [&](size_t b, size_t e) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(rq);
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(mesh, loadPose<S>(poses_mesh + 12 * q), objs[shape_ids[q]].get(),
                                       loadPose<S>(poses_shape + 12 * q), req, res);
      counts[q] = uint32_t(c);
      if (first_tri) first_tri[q] = c ? int32_t(res.getContact(0).b1) : -1;
    }
  });
}

// ---- heightmaps -----------------------------------------------------------------
struct HmRec {
  std::shared_ptr<fcl::HeightMapCollisionGeometry<float>> f;
  std::shared_ptr<fcl::HeightMapCollisionGeometry<double>> d;
};
std::vector<HmRec>& heightmaps() {
  static std::vector<HmRec> t;
  return t;
}
template <typename S>
const fcl::HeightMapCollisionGeometry<S>* getHm(int id);
template <>
const fcl::HeightMapCollisionGeometry<float>* getHm<float>(int id) {
  return heightmaps().at(id).f.get();
}
template <>
const fcl::HeightMapCollisionGeometry<double>* getHm<double>(int id) {
  return heightmaps().at(id).d.get();
}

template <typename S>
std::shared_ptr<fcl::HeightMapCollisionGeometry<S>> buildHm(const double* pts, size_t n, double res, int half_shape) {
  auto map = std::make_shared<fcl::heightmap::LayeredHeightMap<S>>(S(res), uint16_t(half_shape));
  map->updateHeightsByPointGenerationFunctor(
      [&](int i, S& x, S& y, S& z) {
        x = S(pts[3 * size_t(i)]);
        y = S(pts[3 * size_t(i) + 1]);
        z = S(pts[3 * size_t(i) + 2]);
      },
      int(n));
  return std::make_shared<fcl::HeightMapCollisionGeometry<S>>(map);
}

template <typename S>
void hmShapeBatch(int hm_id, const ShapeRec* shapes, uint32_t n_shapes, const uint32_t* shape_ids, const S* poses_hm,
                  const S* poses_shape, size_t n, const RequestRec* rq, uint32_t* counts, int32_t* first_pixel,
                  int threads) {
  const auto* hm = getHm<S>(hm_id);
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> objs;
  for (uint32_t i = 0; i < n_shapes; i++) objs.push_back(Sel<S>::shape(shapes + i));
  parallelFor(n, threads, [&](size_t b, size_t e) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(rq);
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(hm, loadPose<S>(poses_hm + 12 * q), objs[shape_ids[q]].get(),
                                       loadPose<S>(poses_shape + 12 * q), req, res);
      counts[q] = uint32_t(c);
      if (first_pixel) first_pixel[q] = c ? int32_t(res.getContact(0).b1) : -1;
    }
  });
}

}  // namespace

extern "C" {

int fclref_heightmap_create(const double* points, size_t n_points, double resolution, int half_shape) {
  HmRec r;
  r.f = buildHm<float>(points, n_points, resolution, half_shape);
  r.d = buildHm<double>(points, n_points, resolution, half_shape);
  heightmaps().push_back(r);
  return int(heightmaps().size()) - 1;
}
/* bottom layer heights (mm), index = y * full_x + x; returns height_upper_bound_mm */
int fclref_heightmap_export(int id, int scalar_type, uint16_t* heights) {
  auto dump = [&](const auto* g) {
    const auto& bottom = g->raw_heightmap()->bottom();
    for (uint16_t y = 0; y < bottom.full_shape_y(); y++)
      for (uint16_t x = 0; x < bottom.full_shape_x(); x++)
        heights[size_t(y) * bottom.full_shape_x() + x] = bottom.pixelHeight(fcl::heightmap::Pixel(x, y));
    return int(g->raw_heightmap()->height_upper_bound_mm());
  };
  return scalar_type == 0 ? dump(getHm<float>(id)) : dump(getHm<double>(id));
}
int fclref_heightmap_shape_collide_batch(int scalar_type, int hm_id, const void* shapes, uint32_t n_shapes,
                                         const uint32_t* shape_ids, const void* poses_hm, const void* poses_shape,
                                         size_t n, const void* request, uint32_t* counts, int32_t* first_pixel,
                                         int threads) {
  if (scalar_type == 0)
    hmShapeBatch<float>(hm_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses_hm,
                        (const float*)poses_shape, n, (const RequestRec*)request, counts, first_pixel, threads);
  else
    hmShapeBatch<double>(hm_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses_hm,
                         (const double*)poses_shape, n, (const RequestRec*)request, counts, first_pixel, threads);
  return 0;
}

int fclref_mesh_shape_collide_batch(int scalar_type, int mesh_id, const void* shapes, uint32_t n_shapes,
                                    const uint32_t* shape_ids, const void* poses_mesh, const void* poses_shape, size_t n,
                                    const void* request, uint32_t* counts, int32_t* first_tri, int threads) {
  if (scalar_type == 0)
    meshShapeBatch<float>(mesh_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses_mesh,
                          (const float*)poses_shape, n, (const RequestRec*)request, counts, first_tri, threads);
  else
    meshShapeBatch<double>(mesh_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses_mesh,
                           (const double*)poses_shape, n, (const RequestRec*)request, counts, first_tri, threads);
  return 0;
}

}  // extern "C"
