// oracle/ref_harness_mesh.cpp -- TEST INFRASTRUCTURE (the "reference" oracle,
// mesh part).  Compiled with ref_harness.cpp into oracle/_ref/libfclref.so from
// the UNMODIFIED reference headers + oracle/eigen_shim.
//
//   * BVHModel<OBBRSS<S>> construction by the reference's own builder
//     (geometry/bvh/BVH_model-inl.h:402-570) and export of the flattened OBB
//     nodes / triangles -- the same data a mind-fcl integration would hand to
//     fclb_bvh_upload;
//   * batched fcl::collide(BVH, BVH) (-> OrientedNodeBVHSolver::MeshIntersect,
//     narrowphase/detail/traversal/collision/bvh_solver-inl.h:75);
//   * visit counters (BV-pair tests, leaf-pair tests) from an instrumented
//     replica of that loop which calls the reference's own overlap() and
//     Intersect::intersect_Triangle -- the counts SURVEY.md 8(d) uses for the
//     algorithmic bytes of config C3.
#include <climits>
#include <cstdint>
#include <memory>
#include <stack>
#include <thread>
#include <vector>

#include "fcl/fcl.h"

namespace {

struct RequestRec {
  uint32_t max_contacts;
  uint32_t penetration_mode;
  double dir[3];
  double binary_tol, distance_tol;
  uint32_t gjk_max_iter, epa_max_faces, epa_max_iter;
  uint32_t flags;
};

template <typename S>
using Model = fcl::BVHModel<fcl::OBBRSS<S>>;

struct MeshRec {
  std::shared_ptr<Model<float>> f;
  std::shared_ptr<Model<double>> d;
};
std::vector<MeshRec>& meshes() {
  static std::vector<MeshRec> m;
  return m;
}
template <typename S>
Model<S>* get(int id);
template <>
Model<float>* get<float>(int id) {
  return meshes().at(id).f.get();
}
template <>
Model<double>* get<double>(int id) {
  return meshes().at(id).d.get();
}

template <typename S>
std::shared_ptr<Model<S>> build(const double* verts, int n_verts, const int* tris, int n_tris) {
  std::vector<fcl::Vector3<S>> pts;
  std::vector<fcl::MeshSimplex> simp;
  for (int i = 0; i < n_verts; i++) pts.emplace_back(S(verts[3 * i]), S(verts[3 * i + 1]), S(verts[3 * i + 2]));
  for (int i = 0; i < n_tris; i++) simp.emplace_back(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
  auto m = std::make_shared<Model<S>>();
  m->beginModel();
  m->addSubModel(pts, simp);
  m->endModel();
  return m;
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
void exportModel(int id, S* obb, int32_t* child, S* tri) {
  const Model<S>* m = get<S>(id);
  for (int i = 0; i < m->getNumBVs(); i++) {
    const auto& node = m->getBV(i);
    const fcl::OBB<S>& b = node.bv.obb;
    S* o = obb + 15 * size_t(i);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) o[3 * r + c] = b.axis(r, c);
    for (int k = 0; k < 3; k++) o[9 + k] = b.To[k];
    for (int k = 0; k < 3; k++) o[12 + k] = b.extent[k];
    child[i] = node.first_child;
  }
  for (int t = 0; t < m->num_simplex(); t++) {
    const fcl::Simplex<S> s = m->getSimplex(t);
    for (int v = 0; v < 3; v++)
      for (int k = 0; k < 3; k++) tri[9 * size_t(t) + 3 * v + k] = s[v][k];
  }
}

template <typename S>
void collideBatch(int id1, int id2, const S* poses1, const S* poses2, size_t n, const RequestRec* rq,
                  uint32_t* counts, int32_t* first_pair, int threads) {
  const Model<S>* m1 = get<S>(id1);
  const Model<S>* m2 = get<S>(id2);
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::CollisionRequest<S> req(rq->max_contacts);
    if (rq->penetration_mode == 1)
      req.useDefaultPenetration();
    else
      req.disablePenetration();
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(m1, loadPose<S>(poses1 + 12 * q), m2, loadPose<S>(poses2 + 12 * q), req, res);
      counts[q] = uint32_t(c);
      if (first_pair) {
        first_pair[2 * q] = c ? int32_t(res.getContact(0).b1) : -1;
        first_pair[2 * q + 1] = c ? int32_t(res.getContact(0).b2) : -1;
      }
    }
  });
}

// every contact of fcl::collide(BVH, BVH) with contact generation: ids (b1, b2) + {normal, pos, depth}
template <typename S>
void collideContactsBatch(int id1, int id2, const S* poses1, const S* poses2, size_t n, const RequestRec* rq,
                          uint32_t max_keep, uint32_t* counts, int32_t* ids, S* contacts, int threads) {
  const Model<S>* m1 = get<S>(id1);
  const Model<S>* m2 = get<S>(id2);
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::CollisionRequest<S> req(rq->max_contacts);
    const fcl::Vector3<S> dir(S(rq->dir[0]), S(rq->dir[1]), S(rq->dir[2]));
    if (rq->penetration_mode == 2)  // collisionPenetrationMPR (collision_penetration-inl.h:189-252)
      req.useDirectedPenetration(dir);
    else if (rq->penetration_mode == 3)
      req.useIncrementalMinimumDistancePenetration(dir);
    else
      req.useDefaultPenetration();
    if (rq->distance_tol > 0) req.setPenetrationDistanceTolerance(S(rq->distance_tol));
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(m1, loadPose<S>(poses1 + 12 * q), m2, loadPose<S>(poses2 + 12 * q), req, res);
      counts[q] = uint32_t(c);
      for (uint32_t k = 0; k < max_keep; k++) {
        int32_t* id = ids + (q * max_keep + k) * 2;
        S* o = contacts + (q * max_keep + k) * 7;
        if (k < c) {
          const auto& ct = res.getContact(k);
          id[0] = int32_t(ct.b1);
          id[1] = int32_t(ct.b2);
          for (int j = 0; j < 3; j++) {
            o[j] = ct.normal[j];
            o[3 + j] = ct.pos[j];
          }
          o[6] = ct.penetration_depth;
        } else {
          id[0] = id[1] = -1;
          for (int j = 0; j < 7; j++) o[j] = S(0);
        }
      }
    }
  });
}

// Instrumented replica of MeshIntersect's loop (bvh_solver-inl.h:90-160) in
// all-contacts, no-penetration mode: counts BV-pair tests and leaf-pair tests.
template <typename S>
void visitCounts(int id1, int id2, const S* poses1, const S* poses2, size_t n, uint64_t* n_bv, uint64_t* n_leaf,
                 uint32_t* n_hit, int threads) {
  using namespace fcl;
  const Model<S>* m1 = get<S>(id1);
  const Model<S>* m2 = get<S>(id2);
  parallelFor(n, threads, [&](size_t b, size_t e) {
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q), tf2 = loadPose<S>(poses2 + 12 * q);
      Matrix3<S> R;
      Vector3<S> t;
      relativeTransform(tf1.linear(), tf1.translation(), tf2.linear(), tf2.translation(), R, t);
      std::stack<std::pair<int, int>> st;
      st.emplace(0, 0);
      uint64_t bv = 0, leaf = 0;
      uint32_t hit = 0;
      while (!st.empty()) {
        const auto task = st.top();
        st.pop();
        const auto& a = m1->getBV(task.first);
        const auto& c = m2->getBV(task.second);
        bv++;
        if (!overlap(R, t, a.bv, c.bv)) continue;
        const bool l1 = a.isLeaf(), l2 = c.isLeaf();
        if (l1 && l2) {
          leaf++;
          const Simplex<S> s1 = m1->getSimplex(a.primitiveId());
          const Simplex<S> s2 = m2->getSimplex(c.primitiveId());
          if (detail::Intersect<S>::intersect_Triangle(s1[0], s1[1], s1[2], s2[0], s2[1], s2[2], R, t)) hit++;
        } else if (l2 || (!l1 && a.bv.size() > c.bv.size())) {
          st.push(std::make_pair(a.leftChild(), task.second));
          st.push(std::make_pair(a.rightChild(), task.second));
        } else {
          st.push(std::make_pair(task.first, c.leftChild()));
          st.push(std::make_pair(task.first, c.rightChild()));
        }
      }
      if (n_bv) n_bv[q] = bv;
      if (n_leaf) n_leaf[q] = leaf;
      if (n_hit) n_hit[q] = hit;
    }
  });
}

}  // namespace

extern "C" {

int fclref_bvh_create(const double* verts, int n_verts, const int* tris, int n_tris) {
  MeshRec r;
  r.f = build<float>(verts, n_verts, tris, n_tris);
  r.d = build<double>(verts, n_verts, tris, n_tris);
  meshes().push_back(r);
  return int(meshes().size()) - 1;
}
// BVHModel::beginReplaceModel / replaceSubModel / endReplaceModel(refit = true, bottomup) on BOTH scalar instances:
// the vertices are replaced (same count, same triangles), the hierarchy keeps its topology
// (geometry/bvh/BVH_model-inl.h:318-375, refitTree :568-637)
int fclref_bvh_refit(int id, const double* verts, int n_verts, int bottomup) {
  auto run = [&](auto* model) {
    using S = typename std::remove_pointer<decltype(model)>::type::S;
    std::vector<fcl::Vector3<S>> pts;
    for (int i = 0; i < n_verts; i++) pts.emplace_back(S(verts[3 * i]), S(verts[3 * i + 1]), S(verts[3 * i + 2]));
    int rc = model->beginReplaceModel();
    if (rc == 0) rc = model->replaceSubModel(pts);
    if (rc == 0) rc = model->endReplaceModel(true, bottomup != 0);
    return rc;
  };
  const int r1 = run(get<float>(id));
  const int r2 = run(get<double>(id));
  return r1 ? r1 : r2;
}
int fclref_bvh_num_nodes(int id, int scalar_type) {
  return scalar_type == 0 ? get<float>(id)->getNumBVs() : get<double>(id)->getNumBVs();
}
int fclref_bvh_num_tris(int id) { return get<double>(id)->num_simplex(); }
int fclref_bvh_export(int id, int scalar_type, void* obb, int32_t* child, void* tri) {
  if (scalar_type == 0)
    exportModel<float>(id, (float*)obb, child, (float*)tri);
  else
    exportModel<double>(id, (double*)obb, child, (double*)tri);
  return 0;
}
int fclref_bvh_collide_batch(int scalar_type, int id1, int id2, const void* poses1, const void* poses2, size_t n,
                             const void* request, uint32_t* counts, int32_t* first_pair, int threads) {
  if (scalar_type == 0)
    collideBatch<float>(id1, id2, (const float*)poses1, (const float*)poses2, n, (const RequestRec*)request, counts,
                        first_pair, threads);
  else
    collideBatch<double>(id1, id2, (const double*)poses1, (const double*)poses2, n, (const RequestRec*)request, counts,
                         first_pair, threads);
  return 0;
}
int fclref_bvh_collide_contacts_batch(int scalar_type, int id1, int id2, const void* poses1, const void* poses2, size_t n,
                                      const void* request, uint32_t max_keep, uint32_t* counts, int32_t* ids, void* contacts,
                                      int threads) {
  if (scalar_type == 0)
    collideContactsBatch<float>(id1, id2, (const float*)poses1, (const float*)poses2, n, (const RequestRec*)request, max_keep,
                                counts, ids, (float*)contacts, threads);
  else
    collideContactsBatch<double>(id1, id2, (const double*)poses1, (const double*)poses2, n, (const RequestRec*)request,
                                 max_keep, counts, ids, (double*)contacts, threads);
  return 0;
}
int fclref_bvh_visit_counts(int scalar_type, int id1, int id2, const void* poses1, const void* poses2, size_t n,
                            uint64_t* n_bv, uint64_t* n_leaf, uint32_t* n_hit, int threads) {
  if (scalar_type == 0)
    visitCounts<float>(id1, id2, (const float*)poses1, (const float*)poses2, n, n_bv, n_leaf, n_hit, threads);
  else
    visitCounts<double>(id1, id2, (const double*)poses1, (const double*)poses2, n, n_bv, n_leaf, n_hit, threads);
  return 0;
}

}  // extern "C"


// ---- translational continuous collision, shape vs mesh -----------------------------------------
// fcl::translational_ccd(shape, tf_shape, displacement, BVHModel<OBB<S>>, tf_mesh, request, result)
// (narrowphase/continuous_collision-inl.h; matrix entry ShapeBVH_TranslationalCollideImpl<Shape, OBB<S>>,
// detail/ccd/translational_collision_func_matrix-inl.h:469-477 -> bvh_ccd_solver-inl.h RunSweptBV).  The CCD matrix
// has OBB and AABB trees only, so the mesh is built as BVHModel<OBB<S>> (same builder, same OBBs as the OBB half of
// the OBBRSS tree: fclref_bvh_obb_export lets the tests check that).
namespace fclref {
std::shared_ptr<fcl::ShapeBase<float>> makeShapeF(const void* rec);
std::shared_ptr<fcl::ShapeBase<double>> makeShapeD(const void* rec);
}  // namespace fclref
namespace {
struct ShapeRecM {  // = ShapeRec of ref_harness.cpp
  int32_t type;
  uint32_t convex;
  double p[3];
};
template <typename S>
using ObbModel = fcl::BVHModel<fcl::OBB<S>>;
struct ObbMeshRec {
  std::shared_ptr<ObbModel<float>> f;
  std::shared_ptr<ObbModel<double>> d;
};
std::vector<ObbMeshRec>& obbMeshes() {
  static std::vector<ObbMeshRec> m;
  return m;
}
template <typename S>
std::shared_ptr<ObbModel<S>> buildObb(const double* verts, int n_verts, const int* tris, int n_tris) {
  std::vector<fcl::Vector3<S>> pts;
  std::vector<fcl::MeshSimplex> simp;
  for (int i = 0; i < n_verts; i++) pts.emplace_back(S(verts[3 * i]), S(verts[3 * i + 1]), S(verts[3 * i + 2]));
  for (int i = 0; i < n_tris; i++) simp.emplace_back(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
  auto m = std::make_shared<ObbModel<S>>();
  m->beginModel();
  m->addSubModel(pts, simp);
  m->endModel();
  m->computeLocalAABB();
  return m;
}
template <typename S>
struct ObbOf;
template <>
struct ObbOf<float> {
  static ObbModel<float>* get(int id) { return obbMeshes().at(id).f.get(); }
  static std::shared_ptr<fcl::ShapeBase<float>> shape(const ShapeRecM* r) { return fclref::makeShapeF(r); }
};
template <>
struct ObbOf<double> {
  static ObbModel<double>* get(int id) { return obbMeshes().at(id).d.get(); }
  static std::shared_ptr<fcl::ShapeBase<double>> shape(const ShapeRecM* r) { return fclref::makeShapeD(r); }
};

template <typename S>
void ccdMeshBatch(int id, const ShapeRecM* shapes, int n_shapes, const uint32_t* shape_ids, const S* poses_shape,
                  const S* poses_mesh, const S* disp, size_t n, int request_type, uint32_t max_contacts, double zero_tol,
                  int mesh_moves, uint32_t keep, uint32_t* counts, int64_t* prim, S* toc, int threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (int i = 0; i < n_shapes; i++) {
    tab.push_back(ObbOf<S>::shape(shapes + i));
    tab.back()->computeLocalAABB();
  }
  const ObbModel<S>* mesh = ObbOf<S>::get(id);
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = max_contacts;
    if (zero_tol > 0) req.zero_movement_tolerance = S(zero_tol);
    for (size_t q = b; q < e; q++) {
      const auto tf_s = loadPose<S>(poses_shape + 12 * q);
      const auto tf_m = loadPose<S>(poses_mesh + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      if (mesh_moves)  // the displacement is the MESH's, in the mesh frame (matrix entry [BV_OBB][GEOM_x])
        fcl::translational_ccd<S>(mesh, tf_m, d, tab[shape_ids[q]].get(), tf_s, req, res);
      else
        fcl::translational_ccd<S>(tab[shape_ids[q]].get(), tf_s, d, mesh, tf_m, req, res);
      counts[q] = uint32_t(res.num_contacts());
      for (uint32_t k = 0; k < keep && k < res.num_contacts(); k++) {
        const auto& c = res.raw_contacts()[k];
        prim[size_t(q) * keep + k] = c.b2;
        toc[(size_t(q) * keep + k) * 2] = c.toc.lower_bound;
        toc[(size_t(q) * keep + k) * 2 + 1] = c.toc.upper_bound;
      }
    }
  });
}
}  // namespace

extern "C" int fclref_bvh_obb_create(const double* verts, int n_verts, const int* tris, int n_tris) {
  ObbMeshRec r;
  r.f = buildObb<float>(verts, n_verts, tris, n_tris);
  r.d = buildObb<double>(verts, n_verts, tris, n_tris);
  obbMeshes().push_back(r);
  return int(obbMeshes().size()) - 1;
}
// 15 S per node (axis row-major, To, extent) + first_child, as fclref_bvh_export does for the OBBRSS model
extern "C" int fclref_bvh_obb_export(int id, int scalar_type, void* obb, int32_t* child) {
  auto run = [&](auto* m, auto* out) {
    for (int i = 0; i < m->getNumBVs(); i++) {
      const auto& nd = m->getBV(i);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) out[15 * i + 3 * r + c] = nd.bv.axis(r, c);
      for (int k = 0; k < 3; k++) out[15 * i + 9 + k] = nd.bv.To[k];
      for (int k = 0; k < 3; k++) out[15 * i + 12 + k] = nd.bv.extent[k];
      child[i] = nd.first_child;
    }
  };
  if (scalar_type == 0)
    run(ObbOf<float>::get(id), (float*)obb);
  else
    run(ObbOf<double>::get(id), (double*)obb);
  return 0;
}
extern "C" int fclref_translational_ccd_mesh_batch(int scalar_type, int id, const void* shapes, int n_shapes,
                                                   const uint32_t* shape_ids, const void* poses_shape, const void* poses_mesh,
                                                   const void* disp, size_t n, int request_type, uint32_t max_contacts,
                                                   double zero_tol, int mesh_moves, uint32_t keep, uint32_t* counts,
                                                   int64_t* prim, void* toc, int threads) {
  if (scalar_type == 0)
    ccdMeshBatch<float>(id, (const ShapeRecM*)shapes, n_shapes, shape_ids, (const float*)poses_shape, (const float*)poses_mesh,
                        (const float*)disp, n, request_type, max_contacts, zero_tol, mesh_moves, keep, counts, prim, (float*)toc,
                        threads);
  else
    ccdMeshBatch<double>(id, (const ShapeRecM*)shapes, n_shapes, shape_ids, (const double*)poses_shape,
                         (const double*)poses_mesh, (const double*)disp, n, request_type, max_contacts, zero_tol, mesh_moves, keep,
                         counts, prim, (double*)toc, threads);
  return 0;
}

// fcl::translational_ccd(BVHModel<OBB>, tf1, displacement, BVHModel<OBB>, tf2, request, result)
// (TranslationalDisplacementBVH_PairSolverImpl<S, OBB<S>>, bvh_ccd_solver-inl.h:425-551): contacts (b1, b2, toc) in order
namespace {
template <typename S>
void ccdMeshPairBatch(int id1, int id2, const S* poses1, const S* poses2, const S* disp, size_t n, int request_type,
                      uint32_t max_contacts, double zero_tol, uint32_t keep, uint32_t* counts, int64_t* prim, S* toc, int threads) {
  const ObbModel<S>* m1 = ObbOf<S>::get(id1);
  const ObbModel<S>* m2 = ObbOf<S>::get(id2);
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = max_contacts;
    if (zero_tol > 0) req.zero_movement_tolerance = S(zero_tol);
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      fcl::translational_ccd<S>(m1, tf1, d, m2, tf2, req, res);
      counts[q] = uint32_t(res.num_contacts());
      for (uint32_t k = 0; k < keep && k < res.num_contacts(); k++) {
        const auto& c = res.raw_contacts()[k];
        prim[(size_t(q) * keep + k) * 2] = c.b1;
        prim[(size_t(q) * keep + k) * 2 + 1] = c.b2;
        toc[(size_t(q) * keep + k) * 2] = c.toc.lower_bound;
        toc[(size_t(q) * keep + k) * 2 + 1] = c.toc.upper_bound;
      }
    }
  });
}
}  // namespace
extern "C" int fclref_translational_ccd_mesh_pair_batch(int scalar_type, int id1, int id2, const void* poses1, const void* poses2,
                                                        const void* disp, size_t n, int request_type, uint32_t max_contacts,
                                                        double zero_tol, uint32_t keep, uint32_t* counts, int64_t* prim, void* toc,
                                                        int threads) {
  if (scalar_type == 0)
    ccdMeshPairBatch<float>(id1, id2, (const float*)poses1, (const float*)poses2, (const float*)disp, n, request_type, max_contacts,
                            zero_tol, keep, counts, prim, (float*)toc, threads);
  else
    ccdMeshPairBatch<double>(id1, id2, (const double*)poses1, (const double*)poses2, (const double*)disp, n, request_type,
                             max_contacts, zero_tol, keep, counts, prim, (double*)toc, threads);
  return 0;
}

// mesh registry access for the other harness translation units (ref_harness_scene.cpp)
namespace fclref {
const fcl::BVHModel<fcl::OBB<float>>* obbMeshF(int id) { return ObbOf<float>::get(id); }
const fcl::BVHModel<fcl::OBB<double>>* obbMeshD(int id) { return ObbOf<double>::get(id); }
const fcl::BVHModel<fcl::OBBRSS<float>>* meshF(int id) { return get<float>(id); }
const fcl::BVHModel<fcl::OBBRSS<double>>* meshD(int id) { return get<double>(id); }
}  // namespace fclref
