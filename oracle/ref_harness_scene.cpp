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
  const fcl::Vector3<S> dir(S(rq->dir[0]), S(rq->dir[1]), S(rq->dir[2]));
  switch (rq->penetration_mode) {
    case 1:
      req.useDefaultPenetration();
      break;
    case 2:
      req.useDirectedPenetration(dir);
      break;
    case 3:
      req.useIncrementalMinimumDistancePenetration(dir);
      break;
    default:
      req.disablePenetration();
      break;
  }
  if (rq->binary_tol > 0) req.setBinaryCollisionTolerance(S(rq->binary_tol));
  if (rq->distance_tol > 0) req.setPenetrationDistanceTolerance(S(rq->distance_tol));
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

// ---- octrees ---------------------------------------------------------------------------
struct OctRec {
  std::shared_ptr<const fcl::Octree2CollisionGeometry<float>> f;
  std::shared_ptr<const fcl::Octree2CollisionGeometry<double>> d;
};
std::vector<OctRec>& octrees() {
  static std::vector<OctRec> t;
  return t;
}
template <typename S>
const fcl::Octree2CollisionGeometry<S>* getOct(int id);
template <>
const fcl::Octree2CollisionGeometry<float>* getOct<float>(int id) {
  return octrees().at(id).f.get();
}
template <>
const fcl::Octree2CollisionGeometry<double>* getOct<double>(int id) {
  return octrees().at(id).d.get();
}
template <typename S>
std::shared_ptr<fcl::Octree2CollisionGeometry<S>> buildOct(const double* pts, size_t n, double res, int half_shape) {
  auto tree = std::make_shared<fcl::octree2::Octree<S>>(S(res), uint16_t(half_shape));
  tree->rebuildTree(
      [&](int i, S& x, S& y, S& z) {
        x = S(pts[3 * size_t(i)]);
        y = S(pts[3 * size_t(i) + 1]);
        z = S(pts[3 * size_t(i) + 2]);
      },
      int(n));
  return std::make_shared<fcl::Octree2CollisionGeometry<S>>(tree);
}
template <typename S>
void octShapeBatch(int oct_id, const ShapeRec* shapes, uint32_t n_shapes, const uint32_t* shape_ids, const S* poses_oct,
                   const S* poses_shape, size_t n, const RequestRec* rq, uint32_t* counts, int64_t* first_node, int threads) {
  const auto* oct = getOct<S>(oct_id);
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> objs;
  for (uint32_t i = 0; i < n_shapes; i++) objs.push_back(Sel<S>::shape(shapes + i));
  parallelFor(n, threads, [&](size_t b, size_t e) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(rq);
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(oct, loadPose<S>(poses_oct + 12 * q), objs[shape_ids[q]].get(),
                                       loadPose<S>(poses_shape + 12 * q), req, res);
      counts[q] = uint32_t(c);
      if (first_node) first_node[q] = c ? int64_t(res.getContact(0).b1) : -1;
    }
  });
}

// every contact of a scene-vs-shape query (any request mode): b1 + {normal, pos, depth}
template <typename S>
void sceneContactsBatch(const fcl::CollisionGeometry<S>* scene, const ShapeRec* shapes, uint32_t n_shapes,
                        const uint32_t* shape_ids, const S* poses_scene, const S* poses_shape, size_t n,
                        const RequestRec* rq, uint32_t max_keep, uint32_t* counts, int64_t* b1, S* contacts, int threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> objs;
  for (uint32_t i = 0; i < n_shapes; i++) objs.push_back(Sel<S>::shape(shapes + i));
  parallelFor(n, threads, [&](size_t b, size_t e) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(rq);
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(scene, loadPose<S>(poses_scene + 12 * q), objs[shape_ids[q]].get(),
                                       loadPose<S>(poses_shape + 12 * q), req, res);
      counts[q] = uint32_t(c);
      for (uint32_t k = 0; k < max_keep; k++) {
        int64_t* id = b1 + q * max_keep + k;
        S* o = contacts + (q * max_keep + k) * 7;
        if (k < c) {
          const auto& ct = res.getContact(k);
          *id = int64_t(ct.b1);
          for (int j = 0; j < 3; j++) {
            o[j] = ct.normal[j];
            o[3 + j] = ct.pos[j];
          }
          o[6] = ct.penetration_depth;
        } else {
          *id = -1;
          for (int j = 0; j < 7; j++) o[j] = S(0);
        }
      }
    }
  });
}

// every contact of a scene-vs-scene query: counts + (b1, b2) of the first max_keep contacts
template <typename S>
void scenePairBatch(const fcl::CollisionGeometry<S>* g1, const fcl::CollisionGeometry<S>* g2, const S* poses1,
                    const S* poses2, size_t n, const RequestRec* rq, uint32_t max_keep, uint32_t* counts, int64_t* b1,
                    int64_t* b2, S* contacts, int threads) {
  parallelFor(n, threads, [&](size_t b, size_t e) {
    const fcl::CollisionRequest<S> req = makeRequest<S>(rq);
    for (size_t q = b; q < e; q++) {
      fcl::CollisionResult<S> res;
      const size_t c = fcl::collide<S>(g1, loadPose<S>(poses1 + 12 * q), g2, loadPose<S>(poses2 + 12 * q), req, res);
      counts[q] = uint32_t(c);
      for (uint32_t k = 0; k < max_keep; k++) {
        b1[q * max_keep + k] = k < c ? int64_t(res.getContact(k).b1) : -1;
        b2[q * max_keep + k] = k < c ? int64_t(res.getContact(k).b2) : -1;
        if (contacts) {
          S* o = contacts + (q * max_keep + k) * 7;
          for (int j = 0; j < 7; j++) o[j] = S(0);
          if (k < c) {
            const auto& ct = res.getContact(k);
            for (int j = 0; j < 3; j++) {
              o[j] = ct.normal[j];
              o[3 + j] = ct.pos[j];
            }
            o[6] = ct.penetration_depth;
          }
        }
      }
    }
  });
}
template <typename S>
const fcl::CollisionGeometry<S>* sceneGeom(int kind, int id) {
  return kind == 0   ? (const fcl::CollisionGeometry<S>*)Sel<S>::mesh(id)
         : kind == 1 ? (const fcl::CollisionGeometry<S>*)getHm<S>(id)
                     : (const fcl::CollisionGeometry<S>*)getOct<S>(id);
}

// ---- broadphase -----------------------------------------------------------------------
template <typename S>
struct TreeRec {
  fcl::BroadphaseAABB_Tree<S> tree;
};
struct BpRec {
  std::shared_ptr<TreeRec<float>> f;
  std::shared_ptr<TreeRec<double>> d;
};
std::vector<BpRec>& trees() {
  static std::vector<BpRec> t;
  return t;
}
template <typename S>
fcl::BroadphaseAABB_Tree<S>& getTree(int id);
template <>
fcl::BroadphaseAABB_Tree<float>& getTree<float>(int id) {
  return trees().at(id).f->tree;
}
template <>
fcl::BroadphaseAABB_Tree<double>& getTree<double>(int id) {
  return trees().at(id).d->tree;
}

template <typename S>
std::vector<fcl::BroadphaseObjectInfo<S>> loadObjects(const S* boxes, const uint64_t* ids, size_t n) {
  std::vector<fcl::BroadphaseObjectInfo<S>> objs(n);
  for (size_t i = 0; i < n; i++) {
    objs[i].bv.min_ = fcl::Vector3<S>(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]);
    objs[i].bv.max_ = fcl::Vector3<S>(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]);
    objs[i].user_id = ids[i];
  }
  return objs;
}

struct PairSink {
  uint64_t* out;
  size_t cap;
  size_t n;
};
bool sinkPair(std::uint64_t a, std::uint64_t b, void* data) {
  PairSink* s = static_cast<PairSink*>(data);
  if (s->out && s->n < s->cap) {
    s->out[2 * s->n] = a;
    s->out[2 * s->n + 1] = b;
  }
  s->n++;
  return false;
}

template <typename S>
void computeAabbs(const ShapeRec* shapes, uint32_t n_shapes, const uint32_t* shape_ids, const S* poses, size_t n, S* out) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> objs;
  for (uint32_t i = 0; i < n_shapes; i++) {
    objs.push_back(Sel<S>::shape(shapes + i));
    objs.back()->computeLocalAABB();
  }
  for (size_t q = 0; q < n; q++) {
    fcl::CollisionObject<S> obj(objs[shape_ids[q]], loadPose<S>(poses + 12 * q));
    obj.computeAABB();
    const fcl::AABB<S>& b = obj.getAABB();
    for (int k = 0; k < 3; k++) {
      out[6 * q + k] = b.min_[k];
      out[6 * q + 3 + k] = b.max_[k];
    }
  }
}

}  // namespace

extern "C" {

int fclref_compute_aabb_batch(int scalar_type, const void* shapes, uint32_t n_shapes, const uint32_t* shape_ids,
                              const void* poses, size_t n, void* out) {
  if (scalar_type == 0)
    computeAabbs<float>((const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses, n, (float*)out);
  else
    computeAabbs<double>((const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses, n, (double*)out);
  return 0;
}
/* BinaryAABB_Tree::Rebuild over n (AABB, user_id) objects, both scalar types share the id */
int fclref_broadphase_create(int scalar_type, const void* boxes, const uint64_t* ids, size_t n) {
  BpRec r;
  if (scalar_type == 0) {
    r.f = std::make_shared<TreeRec<float>>();
    auto objs = loadObjects<float>((const float*)boxes, ids, n);
    r.f->tree.Rebuild(objs.data(), uint32_t(n));
  } else {
    r.d = std::make_shared<TreeRec<double>>();
    auto objs = loadObjects<double>((const double*)boxes, ids, n);
    r.d->tree.Rebuild(objs.data(), uint32_t(n));
  }
  trees().push_back(r);
  return int(trees().size()) - 1;
}
size_t fclref_broadphase_self_pairs(int scalar_type, int tree, uint64_t* out, size_t cap) {
  PairSink s{out, cap, 0};
  if (scalar_type == 0)
    getTree<float>(tree).SelfCollision(sinkPair, &s);
  else
    getTree<double>(tree).SelfCollision(sinkPair, &s);
  return s.n;
}
size_t fclref_broadphase_tree_pairs(int scalar_type, int tree_a, int tree_b, uint64_t* out, size_t cap) {
  PairSink s{out, cap, 0};
  if (scalar_type == 0)
    getTree<float>(tree_a).TreeCollision(getTree<float>(tree_b), sinkPair, &s);
  else
    getTree<double>(tree_a).TreeCollision(getTree<double>(tree_b), sinkPair, &s);
  return s.n;
}
size_t fclref_broadphase_query_pairs(int scalar_type, int tree, const void* boxes, const uint64_t* ids, size_t n,
                                     uint64_t* out, size_t cap) {
  PairSink s{out, cap, 0};
  if (scalar_type == 0) {
    auto objs = loadObjects<float>((const float*)boxes, ids, n);
    for (auto& o : objs) getTree<float>(tree).SingleObjectCollision(o, sinkPair, &s);
  } else {
    auto objs = loadObjects<double>((const double*)boxes, ids, n);
    for (auto& o : objs) getTree<double>(tree).SingleObjectCollision(o, sinkPair, &s);
  }
  return s.n;
}
int fclref_broadphase_update(int scalar_type, int tree, const uint64_t* ids, const void* boxes, size_t n) {
  int ok = 1;
  if (scalar_type == 0) {
    auto objs = loadObjects<float>((const float*)boxes, ids, n);
    for (auto& o : objs) ok &= int(getTree<float>(tree).UpdateObjectAABB(o.user_id, o.bv));
  } else {
    auto objs = loadObjects<double>((const double*)boxes, ids, n);
    for (auto& o : objs) ok &= int(getTree<double>(tree).UpdateObjectAABB(o.user_id, o.bv));
  }
  return ok;
}
/* the whole C5 scene step on the CPU: computeAABB, Rebuild, SelfCollision with fcl::collide (boolean) on
 * every candidate; returns the number of colliding pairs, *n_candidates = candidate pairs */
size_t fclref_scene_self_collide(int scalar_type, const void* shapes, uint32_t n_shapes, const uint32_t* shape_ids,
                                 const void* poses, size_t n, size_t* n_candidates) {
  auto run = [&](auto tag) -> size_t {
    using S = decltype(tag);
    const S* P = (const S*)poses;
    std::vector<std::shared_ptr<fcl::ShapeBase<S>>> geoms;
    for (uint32_t i = 0; i < n_shapes; i++) {
      geoms.push_back(Sel<S>::shape((const ShapeRec*)shapes + i));
      geoms.back()->computeLocalAABB();
    }
    std::vector<fcl::CollisionObject<S>> objs;
    objs.reserve(n);
    std::vector<fcl::BroadphaseObjectInfo<S>> infos(n);
    for (size_t q = 0; q < n; q++) {
      objs.emplace_back(geoms[shape_ids[q]], loadPose<S>(P + 12 * q));
      objs.back().computeAABB();
      infos[q].bv = objs.back().getAABB();
      infos[q].user_id = q;
    }
    fcl::BroadphaseAABB_Tree<S> tree;
    tree.Rebuild(infos.data(), uint32_t(n));
    size_t cand = 0, hits = 0;
    fcl::CollisionRequest<S> req(1);
    req.disablePenetration();
    tree.SelfCollision(
        [&](std::uint64_t a, std::uint64_t b, void*) -> bool {
          cand++;
          fcl::CollisionResult<S> res;
          if (fcl::collide<S>(&objs[a], &objs[b], req, res) > 0) hits++;
          return false;
        },
        nullptr);
    if (n_candidates) *n_candidates = cand;
    return hits;
  };
  return scalar_type == 0 ? run(float(0)) : run(double(0));
}

/* kind: 0 mesh (BVHModel<OBBRSS>), 1 heightmap, 2 octree */
int fclref_scene_shape_contacts_batch(int scalar_type, int kind, int scene_id, const void* shapes, uint32_t n_shapes,
                                      const uint32_t* shape_ids, const void* poses_scene, const void* poses_shape, size_t n,
                                      const void* request, uint32_t max_keep, uint32_t* counts, int64_t* b1, void* contacts,
                                      int threads) {
  if (scalar_type == 0) {
    const fcl::CollisionGeometry<float>* g = kind == 0   ? (const fcl::CollisionGeometry<float>*)Sel<float>::mesh(scene_id)
                                             : kind == 1 ? (const fcl::CollisionGeometry<float>*)getHm<float>(scene_id)
                                                         : (const fcl::CollisionGeometry<float>*)getOct<float>(scene_id);
    sceneContactsBatch<float>(g, (const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses_scene,
                              (const float*)poses_shape, n, (const RequestRec*)request, max_keep, counts, b1, (float*)contacts,
                              threads);
  } else {
    const fcl::CollisionGeometry<double>* g = kind == 0   ? (const fcl::CollisionGeometry<double>*)Sel<double>::mesh(scene_id)
                                              : kind == 1 ? (const fcl::CollisionGeometry<double>*)getHm<double>(scene_id)
                                                          : (const fcl::CollisionGeometry<double>*)getOct<double>(scene_id);
    sceneContactsBatch<double>(g, (const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses_scene,
                               (const double*)poses_shape, n, (const RequestRec*)request, max_keep, counts, b1,
                               (double*)contacts, threads);
  }
  return 0;
}

/* fcl::collide between two scene geometries; kind: 0 mesh (BVHModel<OBBRSS>), 1 heightmap, 2 octree */
int fclref_scene_pair_collide_batch(int scalar_type, int kind1, int id1, int kind2, int id2, const void* poses1,
                                    const void* poses2, size_t n, const void* request, uint32_t max_keep, uint32_t* counts,
                                    int64_t* b1, int64_t* b2, void* contacts_or_null, int threads) {
  if (scalar_type == 0)
    scenePairBatch<float>(sceneGeom<float>(kind1, id1), sceneGeom<float>(kind2, id2), (const float*)poses1,
                          (const float*)poses2, n, (const RequestRec*)request, max_keep, counts, b1, b2,
                          (float*)contacts_or_null, threads);
  else
    scenePairBatch<double>(sceneGeom<double>(kind1, id1), sceneGeom<double>(kind2, id2), (const double*)poses1,
                           (const double*)poses2, n, (const RequestRec*)request, max_keep, counts, b1, b2,
                           (double*)contacts_or_null, threads);
  return 0;
}

int fclref_octree_create(const double* points, size_t n_points, double resolution, int half_shape) {
  OctRec r;
  r.f = buildOct<float>(points, n_points, resolution, half_shape);
  r.d = buildOct<double>(points, n_points, resolution, half_shape);
  octrees().push_back(r);
  return int(octrees().size()) - 1;
}
/* Octree2CollisionGeometry::pruneBy(obb, rebuild_octree = false) (octree_collision_geometry-inl.h, pruneOctreeByOBB,
 * octree_prune-inl.h:10-103): a new geometry id sharing the tree, with prune info.  obb: axis[9] row-major, To[3], extent[3] */
int fclref_octree_prune(int id, const double* obb) {
  auto make = [&](auto tag) {
    using S = decltype(tag);
    fcl::OBB<S> bv;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) bv.axis(i, j) = S(obb[3 * i + j]);
    for (int k = 0; k < 3; k++) {
      bv.To[k] = S(obb[9 + k]);
      bv.extent[k] = S(obb[12 + k]);
    }
    return bv;
  };
  OctRec r;
  r.f = octrees().at(id).f->pruneBy(make(float(0)), false);
  r.d = octrees().at(id).d->pruneBy(make(double(0)), false);
  octrees().push_back(r);
  return int(octrees().size()) - 1;
}
/* Octree2CollisionGeometry::pruneBy(obb, rebuild_octree = true): the pruned tree consolidated into a renumbered
 * octree (Octree::rebuildAccordingToPruneInfo, octree_construction-inl.h:247-369); a new geometry id */
int fclref_octree_prune_rebuild(int id, const double* obb) {
  auto make = [&](auto tag) {
    using S = decltype(tag);
    fcl::OBB<S> bv;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) bv.axis(i, j) = S(obb[3 * i + j]);
    for (int k = 0; k < 3; k++) {
      bv.To[k] = S(obb[9 + k]);
      bv.extent[k] = S(obb[12 + k]);
    }
    return bv;
  };
  OctRec r;
  r.f = octrees().at(id).f->pruneBy(make(float(0)), true);
  r.d = octrees().at(id).d->pruneBy(make(double(0)), true);
  octrees().push_back(r);
  return int(octrees().size()) - 1;
}
/* prune_internal_nodes as bytes; returns 0 when the geometry carries no prune info */
int fclref_octree_export_pruned(int id, int scalar_type, uint8_t* pruned) {
  auto dump = [&](const auto* g) {
    const std::vector<bool>* p = g->prune_internal_nodes();
    if (!p) return 0;
    for (size_t i = 0; i < p->size(); i++) pruned[i] = (*p)[i] ? 1 : 0;
    return 1;
  };
  return scalar_type == 0 ? dump(getOct<float>(id)) : dump(getOct<double>(id));
}

/* sizes[0..2] = n_inner, n_leaf, n_layers */
int fclref_octree_sizes(int id, int scalar_type, uint32_t* sizes) {
  auto get = [&](const auto* g) {
    sizes[0] = uint32_t(g->inner_nodes().size());
    sizes[1] = uint32_t(g->leaf_nodes().size());
    sizes[2] = uint32_t(g->raw_octree()->n_layers());
    return 0;
  };
  return scalar_type == 0 ? get(getOct<float>(id)) : get(getOct<double>(id));
}
int fclref_octree_export(int id, int scalar_type, uint32_t* inner_children, uint8_t* inner_full, uint8_t* leaf_bits,
                         double* root_aabb) {
  auto dump = [&](const auto* g) {
    const auto& inner = g->inner_nodes();
    const auto& full = g->inner_nodes_fully_occupied();
    const auto& leaf = g->leaf_nodes();
    for (size_t i = 0; i < inner.size(); i++) {
      for (int c = 0; c < 8; c++) inner_children[8 * i + c] = inner[i].children[c];
      inner_full[i] = full[i] ? 1 : 0;
    }
    for (size_t i = 0; i < leaf.size(); i++) {
      uint8_t bits = 0;
      for (uint8_t c = 0; c < 8; c++)
        if (leaf[i].child_occupied.test_i(c)) bits |= uint8_t(1u << c);
      leaf_bits[i] = bits;
    }
    const auto& bv = g->octree_root_bv();
    for (int k = 0; k < 3; k++) {
      root_aabb[k] = double(bv.min_[k]);
      root_aabb[3 + k] = double(bv.max_[k]);
    }
    return 0;
  };
  return scalar_type == 0 ? dump(getOct<float>(id)) : dump(getOct<double>(id));
}
int fclref_octree_shape_collide_batch(int scalar_type, int oct_id, const void* shapes, uint32_t n_shapes,
                                      const uint32_t* shape_ids, const void* poses_oct, const void* poses_shape, size_t n,
                                      const void* request, uint32_t* counts, int64_t* first_node, int threads) {
  if (scalar_type == 0)
    octShapeBatch<float>(oct_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses_oct,
                         (const float*)poses_shape, n, (const RequestRec*)request, counts, first_node, threads);
  else
    octShapeBatch<double>(oct_id, (const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses_oct,
                          (const double*)poses_shape, n, (const RequestRec*)request, counts, first_node, threads);
  return 0;
}

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

// ---- translational continuous collision, shape vs heightmap / octree ----------------------------
// fcl::translational_ccd(shape, tf_shape, displacement, scene, tf_scene, ...) and the scene-first entry
// (TranslationalDisplacementHeightMapSolver::RunShapeHeightMap / RunHeightMapShape, heightmap_ccd_solver-inl.h:8-166;
// TranslationalDisplacementOctreeSolver::RunShapeOctree / RunOctreeShape, octree2_ccd_solver-inl.h).
// kind: 1 heightmap, 2 octree.  Per contact: b2 (encodePixel / encodeOctree2Node), toc, o2_bv (6 S).
namespace {
template <typename S>
void ccdSceneBatch(const fcl::CollisionGeometry<S>* scene, const ShapeRec* shapes, uint32_t n_shapes, const uint32_t* shape_ids,
                   const S* poses_shape, const S* poses_scene, const S* disp, size_t n, int request_type, uint32_t max_contacts,
                   int scene_moves, uint32_t keep, uint32_t* counts, int64_t* code, S* toc, S* box, int threads) {
  std::vector<std::shared_ptr<fcl::ShapeBase<S>>> tab;
  for (uint32_t i = 0; i < n_shapes; i++) {
    tab.push_back(Sel<S>::shape(shapes + i));
    tab.back()->computeLocalAABB();
  }
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = max_contacts;
    for (size_t q = b; q < e; q++) {
      const auto tf_s = loadPose<S>(poses_shape + 12 * q);
      const auto tf_g = loadPose<S>(poses_scene + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      if (scene_moves)
        fcl::translational_ccd<S>(scene, tf_g, d, tab[shape_ids[q]].get(), tf_s, req, res);
      else
        fcl::translational_ccd<S>(tab[shape_ids[q]].get(), tf_s, d, scene, tf_g, req, res);
      counts[q] = uint32_t(res.num_contacts());
      for (uint32_t k = 0; k < keep && k < res.num_contacts(); k++) {
        const auto& c = res.raw_contacts()[k];
        const size_t o = size_t(q) * keep + k;
        // the scene geometry is o2 for the shape-first entry; record whichever side carries the id
        const bool scene_is_o2 = c.o2 == scene;
        code[o] = scene_is_o2 ? c.b2 : c.b1;
        const auto& bv = scene_is_o2 ? c.o2_bv : c.o1_bv;
        toc[2 * o] = c.toc.lower_bound;
        toc[2 * o + 1] = c.toc.upper_bound;
        for (int j = 0; j < 3; j++) {
          box[6 * o + j] = bv.min_[j];
          box[6 * o + 3 + j] = bv.max_[j];
        }
      }
    }
  });
}
}  // namespace
extern "C" int fclref_translational_ccd_scene_batch(int scalar_type, int kind, int scene_id, const void* shapes, uint32_t n_shapes,
                                                    const uint32_t* shape_ids, const void* poses_shape, const void* poses_scene,
                                                    const void* disp, size_t n, int request_type, uint32_t max_contacts,
                                                    int scene_moves, uint32_t keep, uint32_t* counts, int64_t* code, void* toc,
                                                    void* box, int threads) {
  if (scalar_type == 0) {
    const fcl::CollisionGeometry<float>* g = kind == 1 ? (const fcl::CollisionGeometry<float>*)getHm<float>(scene_id)
                                                        : (const fcl::CollisionGeometry<float>*)getOct<float>(scene_id);
    ccdSceneBatch<float>(g, (const ShapeRec*)shapes, n_shapes, shape_ids, (const float*)poses_shape, (const float*)poses_scene,
                         (const float*)disp, n, request_type, max_contacts, scene_moves, keep, counts, code, (float*)toc,
                         (float*)box, threads);
  } else {
    const fcl::CollisionGeometry<double>* g = kind == 1 ? (const fcl::CollisionGeometry<double>*)getHm<double>(scene_id)
                                                         : (const fcl::CollisionGeometry<double>*)getOct<double>(scene_id);
    ccdSceneBatch<double>(g, (const ShapeRec*)shapes, n_shapes, shape_ids, (const double*)poses_shape, (const double*)poses_scene,
                          (const double*)disp, n, request_type, max_contacts, scene_moves, keep, counts, code, (double*)toc,
                          (double*)box, threads);
  }
  return 0;
}

// ---- translational continuous collision, heightmap / octree vs mesh (BVHModel<OBB>) ----------------
// fcl::translational_ccd(scene, tf_scene, displacement, mesh, tf_mesh, ...) and the mesh-first entry
// (TranslationalDisplacementHeightMapSolver::RunHeightMapObbBVH / RunObbBVH_HeightMap, heightmap_ccd_solver-inl.h:168-366;
// TranslationalDisplacementOctreeSolver::RunOctreeObbBVH / RunObbBVH_Octree, octree2_ccd_solver-inl.h:225-470).
// Per contact: (b1 = pixel / node code, b2 = triangle id), toc, o1_bv.
namespace fclref {
const fcl::BVHModel<fcl::OBB<float>>* obbMeshF(int id);
const fcl::BVHModel<fcl::OBB<double>>* obbMeshD(int id);
}  // namespace fclref
namespace {
template <typename S>
void ccdSceneMeshBatch(const fcl::CollisionGeometry<S>* scene, const fcl::CollisionGeometry<S>* mesh, const S* poses_scene,
                       const S* poses_mesh, const S* disp, size_t n, int request_type, uint32_t max_contacts, int mesh_moves,
                       uint32_t keep, uint32_t* counts, int64_t* ids, S* toc, S* box, int threads) {
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = max_contacts;
    for (size_t q = b; q < e; q++) {
      const auto tf_g = loadPose<S>(poses_scene + 12 * q);
      const auto tf_m = loadPose<S>(poses_mesh + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      if (mesh_moves)
        fcl::translational_ccd<S>(mesh, tf_m, d, scene, tf_g, req, res);
      else
        fcl::translational_ccd<S>(scene, tf_g, d, mesh, tf_m, req, res);
      counts[q] = uint32_t(res.num_contacts());
      for (uint32_t k = 0; k < keep && k < res.num_contacts(); k++) {
        const auto& c = res.raw_contacts()[k];
        const size_t o = size_t(q) * keep + k;
        const bool scene_is_o1 = c.o1 == scene;
        ids[2 * o] = scene_is_o1 ? c.b1 : c.b2;
        ids[2 * o + 1] = scene_is_o1 ? c.b2 : c.b1;
        const auto& bv = scene_is_o1 ? c.o1_bv : c.o2_bv;
        toc[2 * o] = c.toc.lower_bound;
        toc[2 * o + 1] = c.toc.upper_bound;
        for (int j = 0; j < 3; j++) {
          box[6 * o + j] = bv.min_[j];
          box[6 * o + 3 + j] = bv.max_[j];
        }
      }
    }
  });
}
}  // namespace
extern "C" int fclref_translational_ccd_scene_mesh_batch(int scalar_type, int kind, int scene_id, int obb_mesh_id,
                                                         const void* poses_scene, const void* poses_mesh, const void* disp, size_t n,
                                                         int request_type, uint32_t max_contacts, int mesh_moves, uint32_t keep,
                                                         uint32_t* counts, int64_t* ids, void* toc, void* box, int threads) {
  if (scalar_type == 0) {
    const fcl::CollisionGeometry<float>* g = kind == 1 ? (const fcl::CollisionGeometry<float>*)getHm<float>(scene_id)
                                                        : (const fcl::CollisionGeometry<float>*)getOct<float>(scene_id);
    ccdSceneMeshBatch<float>(g, fclref::obbMeshF(obb_mesh_id), (const float*)poses_scene, (const float*)poses_mesh, (const float*)disp,
                             n, request_type, max_contacts, mesh_moves, keep, counts, ids, (float*)toc, (float*)box, threads);
  } else {
    const fcl::CollisionGeometry<double>* g = kind == 1 ? (const fcl::CollisionGeometry<double>*)getHm<double>(scene_id)
                                                         : (const fcl::CollisionGeometry<double>*)getOct<double>(scene_id);
    ccdSceneMeshBatch<double>(g, fclref::obbMeshD(obb_mesh_id), (const double*)poses_scene, (const double*)poses_mesh,
                              (const double*)disp, n, request_type, max_contacts, mesh_moves, keep, counts, ids, (double*)toc,
                              (double*)box, threads);
  }
  return 0;
}

// ---- translational continuous collision, heightmap / octree vs heightmap / octree ---------------------------------
// fcl::translational_ccd(scene1, tf1, displacement, scene2, tf2, ...): RunHeightMapPair, RunHeightMapOctree /
// RunOctreeHeightMap (heightmap_ccd_solver-inl.h:369-779), RunOctreePair (octree2_ccd_solver-inl.h:467-922).
// Per contact, in the CALLER's argument order: (b on geometry 1, b on geometry 2), toc, (bv on geometry 1, bv on geometry 2).
namespace {
template <typename S>
void ccdScenePairBatch(int kind1, int id1, int kind2, int id2, const S* poses1, const S* poses2, const S* disp, size_t n,
                       int request_type, uint32_t max_contacts, uint32_t keep, uint32_t* counts, int64_t* ids, S* toc, S* box,
                       int threads) {
  const fcl::CollisionGeometry<S>* g1 = sceneGeom<S>(kind1, id1);
  const fcl::CollisionGeometry<S>* g2 = sceneGeom<S>(kind2, id2);
  // the reference names the heightmap o1 when it is handed (octree, heightmap)
  const bool swapped = kind1 == 2 && kind2 == 1;
  parallelFor(n, threads, [&](size_t b, size_t e) {
    fcl::ContinuousCollisionRequest<S> req;
    req.request_type = static_cast<fcl::TimeOfCollisionRequestType>(request_type);
    req.num_max_contacts = max_contacts;
    for (size_t q = b; q < e; q++) {
      const auto tf1 = loadPose<S>(poses1 + 12 * q);
      const auto tf2 = loadPose<S>(poses2 + 12 * q);
      fcl::TranslationalDisplacement<S> d;
      d.unit_axis_in_shape1 = fcl::Vector3<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      d.scalar_displacement = disp[4 * q + 3];
      fcl::ContinuousCollisionResult<S> res;
      fcl::translational_ccd<S>(g1, tf1, d, g2, tf2, req, res);
      counts[q] = uint32_t(res.num_contacts());
      for (uint32_t k = 0; k < keep && k < res.num_contacts(); k++) {
        const auto& c = res.raw_contacts()[k];
        const size_t o = size_t(q) * keep + k;
        ids[2 * o] = swapped ? c.b2 : c.b1;
        ids[2 * o + 1] = swapped ? c.b1 : c.b2;
        const auto& bv1 = swapped ? c.o2_bv : c.o1_bv;
        const auto& bv2 = swapped ? c.o1_bv : c.o2_bv;
        toc[2 * o] = c.toc.lower_bound;
        toc[2 * o + 1] = c.toc.upper_bound;
        for (int j = 0; j < 3; j++) {
          box[12 * o + j] = bv1.min_[j];
          box[12 * o + 3 + j] = bv1.max_[j];
          box[12 * o + 6 + j] = bv2.min_[j];
          box[12 * o + 9 + j] = bv2.max_[j];
        }
      }
    }
  });
}
}  // namespace
extern "C" int fclref_translational_ccd_scene_pair_batch(int scalar_type, int kind1, int id1, int kind2, int id2, const void* poses1,
                                                         const void* poses2, const void* disp, size_t n, int request_type,
                                                         uint32_t max_contacts, uint32_t keep, uint32_t* counts, int64_t* ids,
                                                         void* toc, void* box, int threads) {
  if (scalar_type == 0)
    ccdScenePairBatch<float>(kind1, id1, kind2, id2, (const float*)poses1, (const float*)poses2, (const float*)disp, n, request_type,
                             max_contacts, keep, counts, ids, (float*)toc, (float*)box, threads);
  else
    ccdScenePairBatch<double>(kind1, id1, kind2, id2, (const double*)poses1, (const double*)poses2, (const double*)disp, n,
                              request_type, max_contacts, keep, counts, ids, (double*)toc, (double*)box, threads);
  return 0;
}
