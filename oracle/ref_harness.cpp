// oracle/ref_harness.cpp  --  TEST INFRASTRUCTURE (the "reference" oracle).
//
// A thin extern "C" batch driver around the UNMODIFIED mind-fcl headers under
// /root/reference/include, compiled against oracle/eigen_shim (Eigen itself is
// absent from this image; see oracle/README.md).  Output: oracle/_ref/libfclref.so.
// Nothing in the product (mind-fcl_b200/, include/) links or loads this file;
// only tests/, __graft_entry__.smoke() and bench.py's CPU legs may.
//
// Every entry point loops the reference's own public call over a batch of
// queries, optionally across std::threads (the reference has no thread pool:
// "read-only narrowphase" = re-entrant fcl::collide on const geometry,
// reference README.md:13, collision_interface-inl.h:13-21).
//
// Conventions shared with include/fclb200.h (the product's C ABI):
//   scalar_type 0 = float, 1 = double
//   pose  = 12 S : R row-major (9) then t (3)
//   shape = { u32 type; u32 geom; double p[3] }   (FCLB_* type codes)
//   pair  = { u32 shape1; u32 shape2 }            (indices into the shape table)
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "fcl/fcl.h"

namespace {

struct ShapeRec {
  uint32_t type;
  uint32_t geom;
  double p[3];
};
struct PairRec {
  uint32_t s1, s2;
};
struct RequestRec {
  uint32_t max_contacts;
  uint32_t penetration_mode;  // 0 disabled, 1 default GJK/EPA, 2 directed, 3 incremental-minimum
  double dir[3];
  double binary_tol, distance_tol;
  uint32_t gjk_max_iter, epa_max_faces, epa_max_iter;
  uint32_t flags;
};

enum { T_BOX = 0, T_SPHERE = 1, T_ELLIPSOID = 2, T_CAPSULE = 3, T_CONE = 4, T_CYLINDER = 5, T_CONVEX = 6 };

struct ConvexRec {
  std::vector<double> verts;  // 3 per vertex
  std::vector<int> faces;     // n, v0..v(n-1), ...
  int num_faces;
};
std::vector<ConvexRec>& convexTable() {
  static std::vector<ConvexRec> t;
  return t;
}

template <typename S>
fcl::Transform3<S> loadPose(const S* p) {
  fcl::Transform3<S> tf;
  tf.setIdentity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) tf.linear()(i, j) = p[3 * i + j];
  for (int i = 0; i < 3; i++) tf.translation()[i] = p[9 + i];
  return tf;
}

template <typename S>
std::shared_ptr<fcl::ShapeBase<S>> makeShape(const ShapeRec& r) {
  using namespace fcl;
  switch (r.type) {
    case T_BOX:
      return std::make_shared<Box<S>>(S(r.p[0]), S(r.p[1]), S(r.p[2]));
    case T_SPHERE:
      return std::make_shared<Sphere<S>>(S(r.p[0]));
    case T_ELLIPSOID:
      return std::make_shared<Ellipsoid<S>>(S(r.p[0]), S(r.p[1]), S(r.p[2]));
    case T_CAPSULE:
      return std::make_shared<Capsule<S>>(S(r.p[0]), S(r.p[1]));
    case T_CONE:
      return std::make_shared<Cone<S>>(S(r.p[0]), S(r.p[1]));
    case T_CYLINDER:
      return std::make_shared<Cylinder<S>>(S(r.p[0]), S(r.p[1]));
    case T_CONVEX: {
      const ConvexRec& c = convexTable().at(r.geom);
      auto v = std::make_shared<std::vector<Vector3<S>>>();
      for (size_t i = 0; i + 2 < c.verts.size(); i += 3)
        v->emplace_back(S(c.verts[i]), S(c.verts[i + 1]), S(c.verts[i + 2]));
      auto f = std::make_shared<std::vector<int>>(c.faces);
      return std::make_shared<Convex<S>>(v, c.num_faces, f, false);
    }
    default:
      return nullptr;
  }
}

template <typename F>
void parallelFor(size_t n, int n_threads, F&& f) {
  if (n_threads <= 1 || n < 2) {
    f(size_t(0), n, 0);
    return;
  }
  std::vector<std::thread> ts;
  const size_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const size_t b = std::min(n, chunk * t), e = std::min(n, chunk * (t + 1));
    if (b >= e) break;
    ts.emplace_back([=, &f] { f(b, e, t); });
  }
  for (auto& t : ts) t.join();
}

// Double dispatch on the concrete shape type: GJKSolver::shapeDistance is a
// template over (Shape1, Shape2) with closed-form specialisations
// (gjk_solver-inl.h:902-988), so the static types matter.
template <typename S, typename Fn>
bool withShape(const fcl::ShapeBase<S>* s, Fn&& fn) {
  using namespace fcl;
  switch (s->getNodeType()) {
    case GEOM_BOX:
      return fn(*static_cast<const Box<S>*>(s));
    case GEOM_SPHERE:
      return fn(*static_cast<const Sphere<S>*>(s));
    case GEOM_ELLIPSOID:
      return fn(*static_cast<const Ellipsoid<S>*>(s));
    case GEOM_CAPSULE:
      return fn(*static_cast<const Capsule<S>*>(s));
    case GEOM_CONE:
      return fn(*static_cast<const Cone<S>*>(s));
    case GEOM_CYLINDER:
      return fn(*static_cast<const Cylinder<S>*>(s));
    case GEOM_CONVEX:
      return fn(*static_cast<const Convex<S>*>(s));
    default:
      return false;
  }
}

template <typename S>
int distanceBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const S* poses1, const S* poses2,
                  size_t n, double gjk_tol, uint32_t gjk_max_iter, S* dist, S* p1, S* p2, uint8_t* ok,
                  int n_threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) tab.push_back(makeShape<S>(shapes[i]));
  parallelFor(n, n_threads, [&](size_t b, size_t e, int) {
    fcl::detail::GJKSolver<S> solver;  // defaults: gjk_solver-inl.h:1121-1130
    if (gjk_tol > 0) solver.gjk_tolerance = S(gjk_tol);
    if (gjk_max_iter > 0) solver.gjk_max_iterations = gjk_max_iter;
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      const fcl::ShapeBase<S>* s1 = tab[pairs[q].s1].get();
      const fcl::ShapeBase<S>* s2 = tab[pairs[q].s2].get();
      S d = S(0);
      fcl::Vector3<S> a = fcl::Vector3<S>::Zero(), c = fcl::Vector3<S>::Zero();
      const bool r = withShape<S>(s1, [&](const auto& sh1) {
        return withShape<S>(s2, [&](const auto& sh2) { return solver.shapeDistance(sh1, tf1, sh2, tf2, &d, &a, &c); });
      });
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

template <typename S>
int signedDistanceBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const S* poses1, const S* poses2,
                        size_t n, S* dist, S* p1, S* p2, uint8_t* ok, int n_threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) tab.push_back(makeShape<S>(shapes[i]));
  parallelFor(n, n_threads, [&](size_t b, size_t e, int) {
    fcl::detail::GJKSolver<S> solver;
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      const fcl::ShapeBase<S>* s1 = tab[pairs[q].s1].get();
      const fcl::ShapeBase<S>* s2 = tab[pairs[q].s2].get();
      S d = S(0);
      fcl::Vector3<S> a = fcl::Vector3<S>::Zero(), c = fcl::Vector3<S>::Zero();
      const bool r = withShape<S>(s1, [&](const auto& sh1) {
        return withShape<S>(s2, [&](const auto& sh2) { return solver.shapeSignedDistance(sh1, tf1, sh2, tf2, &d, &a, &c); });
      });
      dist[q] = d;
      ok[q] = r ? 1 : 0;
      for (int k = 0; k < 3; k++) {
        p1[3 * q + k] = a[k];
        p2[3 * q + k] = c[k];
      }
    }
  });
  return 0;
}

template <typename S>
fcl::CollisionRequest<S> makeRequest(const RequestRec& r) {
  fcl::CollisionRequest<S> req(r.max_contacts);
  switch (r.penetration_mode) {
    case 0:
      req.disablePenetration();
      break;
    case 1:
      req.useDefaultPenetration();
      break;
    case 2:
      req.useDirectedPenetration(fcl::Vector3<S>(S(r.dir[0]), S(r.dir[1]), S(r.dir[2])));
      break;
    case 3:
      req.useIncrementalMinimumDistancePenetration(fcl::Vector3<S>(S(r.dir[0]), S(r.dir[1]), S(r.dir[2])));
      break;
  }
  if (r.binary_tol > 0) req.setBinaryCollisionTolerance(S(r.binary_tol));
  if (r.distance_tol > 0) req.setPenetrationDistanceTolerance(S(r.distance_tol));
  return req;
}

// out_contacts: max_keep records of 9 S each per query: {b1, b2, normal[3], pos[3], depth}
template <typename S>
int collideBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const S* poses1, const S* poses2,
                 size_t n, const RequestRec* rq, uint32_t max_keep, S* out_contacts, uint32_t* out_counts,
                 int n_threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) tab.push_back(makeShape<S>(shapes[i]));
  parallelFor(n, n_threads, [&](size_t b, size_t e, int) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(*rq);
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      fcl::CollisionResult<S> res;
      const size_t cnt = fcl::collide<S>(tab[pairs[q].s1].get(), tf1, tab[pairs[q].s2].get(), tf2, req, res);
      out_counts[q] = uint32_t(cnt);
      if (out_contacts) {
        for (uint32_t c = 0; c < max_keep; c++) {
          S* o = out_contacts + (size_t(q) * max_keep + c) * 9;
          if (c < cnt) {
            const auto& ct = res.getContact(c);
            o[0] = S(ct.b1);
            o[1] = S(ct.b2);
            for (int k = 0; k < 3; k++) o[2 + k] = ct.normal[k];
            for (int k = 0; k < 3; k++) o[5 + k] = ct.pos[k];
            o[8] = ct.penetration_depth;
          } else {
            for (int k = 0; k < 9; k++) o[k] = S(0);
          }
        }
      }
    }
  });
  return 0;
}

// Direct cvx_collide path, as the reference's own tests drive it
// (test/cvx_collide/test_epa2_with_gjk2.cpp:76-162): GJK(max_iter, tol) then,
// on Intersect, EPA(max_faces, max_iter, tol) on the GJK simplex.
// mode bit0: run MPR::Intersect first and report its status in out_mpr.
// out_status: GJK_Status as int;  out_epa: EPA_Status as int (or -1 when EPA not run)
// out_geom: 7 S per query {depth, p0[3], p1[3]} in shape-0 frame
// out_iters: 2 u32 per query {#support calls in GJK(+MPR), #support calls in EPA}
template <typename S>
int gjkEpaBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const S* poses1, const S* poses2,
                size_t n, const RequestRec* rq, int mode, int32_t* out_status, int32_t* out_epa, int32_t* out_mpr,
                S* out_geom, uint32_t* out_iters, int n_threads) {
  using namespace fcl;
  using namespace fcl::detail;
  std::vector<std::shared_ptr<ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) tab.push_back(makeShape<S>(shapes[i]));
  parallelFor(n, n_threads, [&](size_t b, size_t e, int) {
    const S gjk_tol = S(rq->binary_tol > 0 ? rq->binary_tol : 1e-6);
    const S epa_tol = S(rq->distance_tol > 0 ? rq->distance_tol : 1e-6);
    const size_t gjk_it = rq->gjk_max_iter ? rq->gjk_max_iter : 128;
    const size_t epa_faces = rq->epa_max_faces ? rq->epa_max_faces : 256;
    const size_t epa_it = rq->epa_max_iter ? rq->epa_max_iter : 255;
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      uint32_t n_support = 0;
      MinkowskiDiff<S> shape;
      shape.shapes[0] = constructGJKGeometry(tab[pairs[q].s1].get());
      shape.shapes[1] = constructGJKGeometry(tab[pairs[q].s2].get());
      // each supportVertex() makes two support_function calls
      shape.support_function = [&n_support](const GJKGeometryData<S>& g, const Vector3<S>& d) {
        n_support++;
        return computeSupport<S>(g, d);
      };
      shape.interior_function = computeInterior<S>;
      shape.toshape1.noalias() = tf2.linear().transpose() * tf1.linear();
      shape.toshape0 = tf1.inverse(Eigen::Isometry) * tf2;

      if ((mode & 1) && out_mpr) {
        MPR<S> mpr(gjk_it, gjk_tol);
        out_mpr[q] = int32_t(mpr.Intersect(shape));
      }
      GJK2<S> gjk(gjk_it, gjk_tol);
      GJKSimplex<S> simplex;
      const Vector3<S> guess(1, 0, 0);
      const auto st = gjk.Evaluate(shape, simplex, -guess);
      out_status[q] = int32_t(st);
      const uint32_t n_gjk = n_support;
      S depth = 0;
      Vector3<S> p0 = Vector3<S>::Zero(), p1 = Vector3<S>::Zero();
      int32_t est = -1;
      if (st == GJK_Status::Intersect && !(mode & 2)) {
        EPA2<S> epa(epa_faces, epa_it, epa_tol);
        est = int32_t(epa.Evaluate(simplex, shape, &depth, &p0, &p1));
      }
      if (out_epa) out_epa[q] = est;
      if (out_geom) {
        S* o = out_geom + 7 * q;
        o[0] = depth;
        for (int k = 0; k < 3; k++) o[1 + k] = p0[k];
        for (int k = 0; k < 3; k++) o[4 + k] = p1[k];
      }
      if (out_iters) {
        out_iters[2 * q] = n_gjk;
        out_iters[2 * q + 1] = n_support - n_gjk;
      }
    }
  });
  return 0;
}

}  // namespace

extern "C" {

int fclref_register_convex(const double* verts, int n_verts, const int* faces, int faces_len, int num_faces) {
  ConvexRec c;
  c.verts.assign(verts, verts + 3 * size_t(n_verts));
  c.faces.assign(faces, faces + faces_len);
  c.num_faces = num_faces;
  convexTable().push_back(std::move(c));
  return int(convexTable().size()) - 1;
}

int fclref_distance_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs, const void* poses1,
                          const void* poses2, size_t n, double gjk_tol, uint32_t gjk_max_iter, void* dist, void* p1,
                          void* p2, uint8_t* ok, int n_threads) {
  if (scalar_type == 0)
    return distanceBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                                (const float*)poses2, n, gjk_tol, gjk_max_iter, (float*)dist, (float*)p1,
                                (float*)p2, ok, n_threads);
  return distanceBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                               (const double*)poses2, n, gjk_tol, gjk_max_iter, (double*)dist, (double*)p1,
                               (double*)p2, ok, n_threads);
}

int fclref_collide_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs, const void* poses1,
                         const void* poses2, size_t n, const void* request, uint32_t max_keep, void* out_contacts,
                         uint32_t* out_counts, int n_threads) {
  if (scalar_type == 0)
    return collideBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                               (const float*)poses2, n, (const RequestRec*)request, max_keep, (float*)out_contacts,
                               out_counts, n_threads);
  return collideBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                              (const double*)poses2, n, (const RequestRec*)request, max_keep, (double*)out_contacts,
                              out_counts, n_threads);
}

int fclref_gjk_epa_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs, const void* poses1,
                         const void* poses2, size_t n, const void* request, int mode, int32_t* out_status,
                         int32_t* out_epa, int32_t* out_mpr, void* out_geom, uint32_t* out_iters, int n_threads) {
  if (scalar_type == 0)
    return gjkEpaBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                              (const float*)poses2, n, (const RequestRec*)request, mode, out_status, out_epa,
                              out_mpr, (float*)out_geom, out_iters, n_threads);
  return gjkEpaBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                             (const double*)poses2, n, (const RequestRec*)request, mode, out_status, out_epa,
                             out_mpr, (double*)out_geom, out_iters, n_threads);
}

int fclref_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

}  // extern "C"

extern "C" int fclref_signed_distance_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs,
                                            const void* poses1, const void* poses2, size_t n, void* dist, void* p1, void* p2,
                                            uint8_t* ok, int threads) {
  if (scalar_type == 0)
    return signedDistanceBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                                      (const float*)poses2, n, (float*)dist, (float*)p1, (float*)p2, ok, threads);
  return signedDistanceBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                                     (const double*)poses2, n, (double*)dist, (double*)p1, (double*)p2, ok, threads);
}

// fcl::translational_ccd(shape, tf1, displacement, shape, tf2, request, result) per query
// (narrowphase/continuous_collision-inl.h:21-36).  disp: 4 S per query = unit axis in shape 1's frame, scalar.
// out_hit: 1 when a contact is reported; out_toc: ContinuousCollisionContact::toc (lower, upper) or (-1, -1).
template <typename S>
int translationalCcdBatch(const ShapeRec* shapes, int n_shapes, const PairRec* pairs, const S* poses1, const S* poses2,
                          const S* disp, size_t n, int request_type, double zero_tol, double gjk_tol, int max_iter,
                          uint8_t* out_hit, S* out_toc, int n_threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) {
    tab.push_back(makeShape<S>(shapes[i]));
    tab.back()->computeLocalAABB();  // kBoxApproximate reads aabb_local; the reference's own CCD tests prepare it the same way
  }
  parallelFor(n, n_threads, [&](size_t b, size_t e, int) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = 1;
    if (zero_tol > 0) req.zero_movement_tolerance = S(zero_tol);
    if (gjk_tol > 0) req.gjk_tolerance = S(gjk_tol);
    if (max_iter > 0) req.max_gjk_iterations = max_iter;
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      fcl::translational_ccd<S>(tab[pairs[q].s1].get(), tf1, d, tab[pairs[q].s2].get(), tf2, req, res);
      const bool hit = res.num_contacts() > 0;
      out_hit[q] = hit ? 1 : 0;
      if (out_toc) {
        out_toc[2 * q] = hit ? res.raw_contacts()[0].toc.lower_bound : S(-1);
        out_toc[2 * q + 1] = hit ? res.raw_contacts()[0].toc.upper_bound : S(-1);
      }
    }
  });
  return 0;
}
extern "C" int fclref_translational_ccd_batch(int scalar_type, const void* shapes, int n_shapes, const void* pairs,
                                              const void* poses1, const void* poses2, const void* disp, size_t n,
                                              int request_type, double zero_tol, double gjk_tol, int max_iter,
                                              uint8_t* out_hit, void* out_toc, int threads) {
  if (scalar_type == 0)
    return translationalCcdBatch<float>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const float*)poses1,
                                        (const float*)poses2, (const float*)disp, n, request_type, zero_tol, gjk_tol, max_iter,
                                        out_hit, (float*)out_toc, threads);
  return translationalCcdBatch<double>((const ShapeRec*)shapes, n_shapes, (const PairRec*)pairs, (const double*)poses1,
                                       (const double*)poses2, (const double*)disp, n, request_type, zero_tol, gjk_tol, max_iter,
                                       out_hit, (double*)out_toc, threads);
}

// shape factory for the other harness translation units (ref_harness_scene.cpp)
namespace fclref {
std::shared_ptr<fcl::ShapeBase<float>> makeShapeF(const void* rec) { return makeShape<float>(*static_cast<const ShapeRec*>(rec)); }
std::shared_ptr<fcl::ShapeBase<double>> makeShapeD(const void* rec) { return makeShape<double>(*static_cast<const ShapeRec*>(rec)); }
}  // namespace fclref
