// Host-side C++ API check, written the way the reference's own tests read
// (test/test_fcl_geometric_shapes.cpp shapeDistance_*, test_fcl_collision_penetration.cpp):
// known answers through fcl::collide / fcl::distance of include/fcl_b200/fcl.h.
// Built and run by tests/test_host_api_gpu.py on the GPU box.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "fcl_b200/fcl.h"

#define EXPECT_TRUE(c)                                                  \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);       \
      failures++;                                                       \
    }                                                                   \
  } while (0)

static int failures = 0;

template <typename S>
void run() {
  using namespace fcl;
  Transform3<S> I = Transform3<S>::Identity();
  auto at = [](S x, S y, S z) {
    Transform3<S> t;
    t.translation() = Vector3<S>(x, y, z);
    return t;
  };
  // shapeDistance_boxsphere (test_fcl_geometric_shapes.cpp): sphere r=20 vs box 5^3
  {
    Sphere<S> s1(20);
    Box<S> s2(5, 5, 5);
    DistanceRequest<S> req;
    DistanceResult<S> res;
    distance<S>(&s1, I, &s2, I, req, res);
    EXPECT_TRUE(!res.separated && res.min_distance < 0);
    distance<S>(&s1, I, &s2, at(S(22.6), 0, 0), req, res);
    EXPECT_TRUE(res.separated && std::fabs(res.min_distance - S(0.1)) < S(0.001));
    distance<S>(&s1, I, &s2, at(40, 0, 0), req, res);
    EXPECT_TRUE(res.separated && std::fabs(res.min_distance - S(17.5)) < S(0.001));
  }
  // signed distance: two unit spheres 1.5 apart penetrate by 0.5
  {
    Sphere<S> a(1), b(1);
    DistanceRequest<S> req;
    req.enable_signed_distance = true;
    DistanceResult<S> res;
    distance<S>(&a, I, &b, at(S(1.5), 0, 0), req, res);
    EXPECT_TRUE(res.separated && std::fabs(res.min_distance + S(0.5)) < S(2e-3));
    distance<S>(&a, I, &b, at(S(2.5), 0, 0), req, res);
    EXPECT_TRUE(res.separated && std::fabs(res.min_distance - S(0.5)) < S(2e-3));
  }
  // shapeDistance_cylindercylinder: GJK distance path
  {
    Cylinder<S> s1(5, 10), s2(5, 10);
    DistanceRequest<S> req;
    DistanceResult<S> res;
    distance<S>(&s1, I, &s2, at(S(10.1), 0, 0), req, res);
    EXPECT_TRUE(res.separated && std::fabs(res.min_distance - S(0.1)) < S(0.001));
  }
  // collide with penetration: "move shape2 by 1.05 * depth * normal => no collision"
  // (the criterion of test/test_fcl_collision_penetration.cpp:14-82)
  {
    Capsule<S> s1(S(0.3), S(0.8));
    Box<S> s2(S(0.8), S(0.6), S(0.4));
    Transform3<S> tf2 = at(S(0.35), S(0.1), S(0.05));
    CollisionRequest<S> req(1);
    req.useDefaultPenetration();
    CollisionResult<S> res;
    const std::size_t n = collide<S>(&s1, I, &s2, tf2, req, res);
    EXPECT_TRUE(n == 1);
    if (n == 1) {
      const Contact<S>& c = res.getContact(0);
      EXPECT_TRUE(c.penetration_depth > 0);
      Transform3<S> moved = tf2;
      for (int k = 0; k < 3; k++) moved.translation()[k] += S(1.05) * c.penetration_depth * c.normal[k];
      CollisionRequest<S> breq(1);
      CollisionResult<S> bres;
      EXPECT_TRUE(collide<S>(&s1, I, &s2, moved, breq, bres) == 0);
    }
  }
  // box-box contacts (boxBox2): up to 4 contacts, all with the same normal
  {
    Box<S> s1(2, 1, S(0.5)), s2(1, 1, 1);
    CollisionRequest<S> req(4);
    req.useDefaultPenetration();
    CollisionResult<S> res;
    const std::size_t n = collide<S>(&s1, I, &s2, at(0, 0, S(0.7)), req, res);
    EXPECT_TRUE(n >= 1 && n <= 4);
    for (std::size_t i = 0; i < n; i++) EXPECT_TRUE(std::fabs(std::fabs(res.getContact(i).normal[2]) - 1) < S(1e-5));
  }
  // max_contacts == 0 => warning + 0 (collision-inl.h:79-84)
  {
    Sphere<S> a(1), b(1);
    CollisionRequest<S> req(0);
    CollisionResult<S> res;
    EXPECT_TRUE(collide<S>(&a, I, &b, I, req, res) == 0);
  }
}

// scene geometry through the mirrored classes: a two-triangle floor mesh (beginModel / addSubModel / endModel),
// a heightmap rasterised from points, the MPR penetration request modes and the broadphase tree
template <typename S>
void runScene() {
  using namespace fcl;
  Transform3<S> I = Transform3<S>::Identity();
  auto at = [](S x, S y, S z) {
    Transform3<S> t;
    t.translation() = Vector3<S>(x, y, z);
    return t;
  };
  {
    BVHModel<OBBRSS<S>> floor;
    floor.beginModel();
    floor.addSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)},
                      {{0, 1, 2}, {0, 2, 3}});
    floor.endModel();
    EXPECT_TRUE(floor.getNumBVs() == 3);
    Sphere<S> ball(S(0.25));
    CollisionRequest<S> req(10);
    CollisionResult<S> hit, miss;
    EXPECT_TRUE(collide<S>(&floor, I, &ball, at(S(0.5), S(-0.5), S(0.2)), req, hit) == 1);   // touches one triangle
    EXPECT_TRUE(collide<S>(&floor, I, &ball, at(S(0.5), S(-0.5), S(0.3)), req, miss) == 0);  // just above the floor
    CollisionResult<S> both;
    EXPECT_TRUE(collide<S>(&floor, I, &ball, at(0, 0, S(0.1)), req, both) == 2);  // on the shared diagonal
  }
  {
    auto map = std::make_shared<heightmap::LayeredHeightMap<S>>(S(0.1), uint16_t(8));  // 16 x 16 pixels, +-0.8 m
    map->updateHeightsByPointGenerationFunctor(
        [](int i, S& x, S& y, S& z) {
          x = S(0.05) + S(0.1) * S(i % 4);
          y = S(0.05);
          z = S(0.5);
        },
        4);  // four 0.5 m columns in a row
    HeightMapCollisionGeometry<S> hm(map);
    Box<S> box(S(0.06), S(0.06), S(0.06));
    CollisionRequest<S> req(100);
    CollisionResult<S> on, over, off;
    EXPECT_TRUE(collide<S>(&hm, I, &box, at(S(0.05), S(0.05), S(0.3)), req, on) == 1);
    EXPECT_TRUE(on.numContacts() == 1 && on.getContact(0).b1 == ((8 << 16) | 8));  // encodePixel(x=8, y=8)
    EXPECT_TRUE(collide<S>(&hm, I, &box, at(S(0.05), S(0.05), S(0.6)), req, over) == 0);
    EXPECT_TRUE(collide<S>(&hm, I, &box, at(S(-0.3), S(0.05), S(0.3)), req, off) == 0);
    // heightmap vs mesh (HeightMapBVHCollide / BVHHeightMapCollide): a 2 x 2 m sheet cutting the four columns at
    // half height; shifted so that all four lie on the sheet's second triangle
    BVHModel<OBBRSS<S>> sheet;
    sheet.beginModel();
    sheet.addSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)},
                      {{0, 1, 2}, {0, 2, 3}});
    sheet.endModel();
    CollisionResult<S> cut, swapped, above;
    EXPECT_TRUE(collide<S>(&hm, I, &sheet, at(0, S(-0.5), S(0.25)), req, cut) == 4);
    for (std::size_t c = 0; c < cut.numContacts(); c++) {
      const auto& ct = cut.getContact(c);
      EXPECT_TRUE(ct.b2 == 1 && (ct.b1 >> 16) >= 8 && (ct.b1 >> 16) <= 11 && (ct.b1 & 0xffff) == 8);
    }
    EXPECT_TRUE(collide<S>(&sheet, at(0, S(-0.5), S(0.25)), &hm, I, req, swapped) == 4);
    EXPECT_TRUE(swapped.numContacts() == 4 && swapped.getContact(0).o1 == &hm);
    EXPECT_TRUE(collide<S>(&hm, I, &sheet, at(0, S(-0.5), S(0.6)), req, above) == 0);
  }
  {
    // octree2::Octree built from points (rebuildTree) and wrapped as the reference wraps it: four voxels in a row
    auto tree = std::make_shared<octree2::Octree<S>>(S(0.1), std::uint16_t(8));  // 16^3 voxels, +-0.8 m
    EXPECT_TRUE(tree->n_layers() == 5 && tree->n_inner_nodes() == 1 && tree->n_leaf_nodes() == 0);
    tree->rebuildTree(
        [](int i, S& x, S& y, S& z) {
          x = S(0.05) + S(0.1) * S(i % 4);
          y = S(0.05);
          z = S(0.05);
        },
        8);  // every voxel named twice
    EXPECT_TRUE(tree->n_inner_nodes() == 3 && tree->n_leaf_nodes() == 2);  // root + two levels above two 2x2x2 cells
    EXPECT_TRUE(tree->leaf_bits()[0] == 3 && tree->leaf_bits()[1] == 3);
    EXPECT_TRUE(tree->isPointOccupied(Vector3<S>(S(0.17), S(0.02), S(0.09))) && tree->isPointOccupied(Vector3<S>(S(0.35), S(0.05), S(0.05))));
    EXPECT_TRUE(!tree->isPointOccupied(Vector3<S>(S(0.45), S(0.05), S(0.05))) && !tree->isPointOccupied(Vector3<S>(S(0.05), S(0.15), S(0.05))) &&
                !tree->isPointOccupied(Vector3<S>(S(-0.05), S(0.05), S(0.05))) && !tree->isPointOccupied(Vector3<S>(S(0.9), 0, 0)));
    Octree2CollisionGeometry<S> oct(tree);
    Box<S> small(S(0.06), S(0.06), S(0.06)), bar(S(0.5), S(0.06), S(0.06));
    CollisionRequest<S> req(100);
    CollisionResult<S> one, none, four;
    EXPECT_TRUE(collide<S>(&oct, I, &small, at(S(0.05), S(0.05), S(0.05)), req, one) == 1);
    EXPECT_TRUE(collide<S>(&oct, I, &small, at(S(0.05), S(0.05), S(0.5)), req, none) == 0);
    EXPECT_TRUE(collide<S>(&oct, I, &bar, at(S(0.2), S(0.05), S(0.05)), req, four) == 4);
    // pruneBy: cut the two voxels with x < 0.2 out of the tree
    OBB<S> cut;
    cut.To = Vector3<S>(S(0.1), S(0.05), S(0.05));
    cut.extent = Vector3<S>(S(0.09), S(0.2), S(0.2));
    auto rest = oct.pruneBy(cut, false);
    EXPECT_TRUE(rest->prune_internal_nodes() != nullptr && rest->leaf_bits()[0] == 0 && rest->leaf_bits()[1] == 3);
    CollisionResult<S> gone, two;
    EXPECT_TRUE(collide<S>(rest.get(), I, &small, at(S(0.05), S(0.05), S(0.05)), req, gone) == 0);
    EXPECT_TRUE(collide<S>(rest.get(), I, &bar, at(S(0.2), S(0.05), S(0.05)), req, two) == 2);
    auto rebuilt = oct.pruneBy(cut, true);  // consolidated: same voxels, no prune info
    CollisionResult<S> gone2, two2;
    EXPECT_TRUE(rebuilt->prune_internal_nodes() == nullptr && rebuilt->leaf_bits().size() == 2);
    EXPECT_TRUE(collide<S>(rebuilt.get(), I, &small, at(S(0.05), S(0.05), S(0.05)), req, gone2) == 0);
    EXPECT_TRUE(collide<S>(rebuilt.get(), I, &bar, at(S(0.2), S(0.05), S(0.05)), req, two2) == 2);
  }
  {
    // directed penetration: two unit spheres 1.5 apart along x, escape direction +x => depth 0.5
    Sphere<S> a(1), b(1);
    CollisionRequest<S> req(1);
    req.useDirectedPenetration(Vector3<S>(1, 0, 0));
    CollisionResult<S> res;
    EXPECT_TRUE(collide<S>(&a, I, &b, at(S(1.5), 0, 0), req, res) == 1);
    if (res.numContacts() == 1) EXPECT_TRUE(std::fabs(res.getContact(0).penetration_depth - S(0.5)) < S(1e-3));
    CollisionRequest<S> inc(1);
    inc.useIncrementalMinimumDistancePenetration(Vector3<S>(0, S(0.6), S(0.8)));
    CollisionResult<S> res2;
    EXPECT_TRUE(collide<S>(&a, I, &b, at(S(1.5), 0, 0), inc, res2) == 1);
    if (res2.numContacts() == 1) {
      EXPECT_TRUE(res2.getContact(0).penetration_depth > 0 && res2.getContact(0).penetration_depth < S(1.3));
    }
  }
  {
    std::vector<BroadphaseObjectInfo<S>> objs(3);
    objs[0].bv = {0, 0, 0, 1, 1, 1};
    objs[0].user_id = 10;
    objs[1].bv = {S(0.5), S(0.5), S(0.5), 2, 2, 2};
    objs[1].user_id = 11;
    objs[2].bv = {5, 5, 5, 6, 6, 6};
    objs[2].user_id = 12;
    BroadphaseAABB_Tree<S> tree;
    tree.Rebuild(objs.data(), 3);
    int n_pairs = 0;
    std::uint64_t sum = 0;
    tree.SelfCollision(
        [&](std::uint64_t i, std::uint64_t j, void*) {
          n_pairs++;
          sum += i + j;
          return false;
        },
        nullptr);
    EXPECT_TRUE(n_pairs == 1 && sum == 21);
    EXPECT_TRUE(tree.UpdateObjectAABB(12, {S(1.5), S(1.5), S(1.5), 6, 6, 6}));
    n_pairs = 0;
    tree.SelfCollision([&](std::uint64_t, std::uint64_t, void*) { return ++n_pairs, false; }, nullptr);
    EXPECT_TRUE(n_pairs == 2);  // (10,11) and (11,12)
    Box<S> unit(1, 1, 1);
    CollisionObject<S> obj(std::make_shared<Box<S>>(1, 1, 1), at(3, 0, 0));
    const auto box = obj.computeAABB();
    EXPECT_TRUE(std::fabs(box[0] - S(2.5)) < S(1e-6) && std::fabs(box[3] - S(3.5)) < S(1e-6));
  }
}

// Batched dispatch of scene queries (include/fcl_b200/fcl.h collideBatch): ONE C-ABI call per (geometry, geometry)
// group, the reference's argument swaps, contact records for penetration requests, the host-side contact functor.
template <typename S>
void runBatch() {
  using namespace fcl;
  Transform3<S> I = Transform3<S>::Identity();
  auto at = [](S x, S y, S z) {
    Transform3<S> t;
    t.translation() = Vector3<S>(x, y, z);
    return t;
  };
  BVHModel<OBBRSS<S>> floor;
  floor.beginModel();
  floor.addSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)}, {{0, 1, 2}, {0, 2, 3}});
  floor.endModel();
  Sphere<S> ball(S(0.25));
  Box<S> brick(S(0.3), S(0.2), S(0.1));
  // 10k mesh-shape queries, alternating argument order and shape: exactly one batched call
  std::vector<CollisionQuery<S>> qs;
  for (int i = 0; i < 10000; i++) {
    const S x = S(-0.9) + S(1.8) * S(i % 100) / S(99), z = S(-0.1) + S(0.5) * S(i / 100) / S(99);
    const CollisionGeometry<S>* shape = (i % 3) ? static_cast<const CollisionGeometry<S>*>(&ball) : &brick;
    if (i & 1)
      qs.push_back({shape, at(x, S(0.3), z), &floor, I});  // (Shape, BVH): collide() swaps (collision-inl.h:91-100)
    else
      qs.push_back({&floor, I, shape, at(x, S(0.3), z)});
  }
  CollisionRequest<S> all(100);
  std::vector<CollisionResult<S>> res;
  const std::size_t calls0 = detail::abiBatchCalls();
  collideBatch(qs, all, res);
  EXPECT_TRUE(detail::abiBatchCalls() - calls0 == 1);
  std::size_t hits = 0;
  for (std::size_t i = 0; i < qs.size(); i++) {
    hits += res[i].numContacts();
    for (const auto& c : res[i].getContacts()) EXPECT_TRUE(c.o1 == &floor && (c.b1 == 0 || c.b1 == 1));
    // a query and its argument-swapped twin (same pose, one row later in the grid is a different pose: compare with single calls)
  }
  EXPECT_TRUE(hits > 1000);
  for (int i : {0, 1, 2, 3, 4999, 5000, 9999}) {  // the batch equals the single calls, whichever argument order
    CollisionResult<S> one;
    const auto& Q = qs[std::size_t(i)];
    const bool shape_first = Q.o1->isShape();
    const std::size_t n1 = shape_first ? collide<S>(Q.o2, Q.tf2, Q.o1, Q.tf1, all, one) : collide<S>(Q.o1, Q.tf1, Q.o2, Q.tf2, all, one);
    EXPECT_TRUE(n1 == res[std::size_t(i)].numContacts());
  }
  // penetration request on the same batch: contact records are filled, still one call
  CollisionRequest<S> pen(100);
  pen.useDefaultPenetration();
  std::vector<CollisionResult<S>> pres;
  const std::size_t calls1 = detail::abiBatchCalls();
  collideBatch(qs, pen, pres);
  EXPECT_TRUE(detail::abiBatchCalls() - calls1 == 1);
  std::size_t with_depth = 0;
  for (const auto& r : pres)
    for (const auto& c : r.getContacts()) {
      const S nn = c.normal[0] * c.normal[0] + c.normal[1] * c.normal[1] + c.normal[2] * c.normal[2];
      EXPECT_TRUE(std::fabs(nn - 1) < S(1e-4));
      with_depth++;
    }
  EXPECT_TRUE(with_depth > 1000);
  // heightmap / octree with the shape first (ShapeHeightMapCollide, ShapeOcTree2Collide): o1 is still the scene geometry
  auto map = std::make_shared<heightmap::LayeredHeightMap<S>>(S(0.1), uint16_t(8));
  map->updateHeightsByPointGenerationFunctor([](int i, S& x, S& y, S& z) { x = S(0.05) + S(0.1) * S(i % 4); y = S(0.05); z = S(0.5); }, 4);
  HeightMapCollisionGeometry<S> hm(map);
  auto tree = std::make_shared<octree2::Octree<S>>(S(0.1), std::uint16_t(8));
  tree->rebuildTree([](int i, S& x, S& y, S& z) { x = S(0.05) + S(0.1) * S(i % 4); y = S(0.05); z = S(0.05); }, 4);
  Octree2CollisionGeometry<S> oct(tree);
  Box<S> bar(S(0.5), S(0.06), S(0.06));
  CollisionResult<S> a, b, c, d;
  EXPECT_TRUE(collide<S>(&hm, I, &bar, at(S(0.2), S(0.05), S(0.3)), all, a) == 4);
  EXPECT_TRUE(collide<S>(&bar, at(S(0.2), S(0.05), S(0.3)), &hm, I, all, b) == 4 && b.getContact(0).o1 == &hm && b.getContact(0).o2 == &bar);
  EXPECT_TRUE(collide<S>(&oct, I, &bar, at(S(0.2), S(0.05), S(0.05)), all, c) == 4);
  EXPECT_TRUE(collide<S>(&bar, at(S(0.2), S(0.05), S(0.05)), &oct, I, all, d) == 4 && d.getContact(0).o1 == &oct);
  // DefaultGJK_EPA against a heightmap: box-box leaves make up to four contacts each (boxBox2), none is refused
  CollisionResult<S> e;
  EXPECT_TRUE(collide<S>(&hm, I, &bar, at(S(0.2), S(0.05), S(0.3)), pen, e) >= 4);
  for (const auto& ct : e.getContacts()) EXPECT_TRUE(ct.penetration_depth >= 0 && (ct.b1 & 0xffff) == 8);
  // translational continuous collision: a unit sphere sweeping 3 m along +x meets a box 2 m ahead; 0.5 m does not reach it
  {
    Sphere<S> mover(S(0.5));
    Box<S> wall(S(0.2), 2, 2);
    TranslationalDisplacement<S> far_sweep, short_sweep;
    far_sweep.unit_axis_in_shape1 = short_sweep.unit_axis_in_shape1 = Vector3<S>(1, 0, 0);
    far_sweep.scalar_displacement = 3;
    short_sweep.scalar_displacement = S(0.5);
    ContinuousCollisionRequest<S> creq;
    creq.request_type = TimeOfCollisionRequestType::kBoxApproximate;
    ContinuousCollisionResult<S> hit_r, miss_r;
    translational_ccd<S>(&mover, I, far_sweep, &wall, at(2, 0, 0), creq, hit_r);
    translational_ccd<S>(&mover, I, short_sweep, &wall, at(2, 0, 0), creq, miss_r);
    EXPECT_TRUE(hit_r.num_contacts() == 1 && miss_r.num_contacts() == 0);
    if (hit_r.num_contacts() == 1) {  // first touch after 1.4 m of 3 m, last after 2.6 m
      const auto& toc = hit_r.raw_contacts()[0].toc;
      EXPECT_TRUE(std::fabs(toc.lower_bound - S(1.4 / 3.0)) < S(1e-4) && std::fabs(toc.upper_bound - S(2.6 / 3.0)) < S(1e-4));
    }
  }
  // the same against a mesh, both argument orders: a sphere dropping 2 m onto the two-triangle floor touches both triangles
  // (contacts in the reference's order: right child first), a sideways sweep above it touches none
  {
    Sphere<S> mover(S(0.25));
    TranslationalDisplacement<S> down, sideways, up;
    down.unit_axis_in_shape1 = Vector3<S>(0, 0, -1);
    up.unit_axis_in_shape1 = Vector3<S>(0, 0, 1);
    sideways.unit_axis_in_shape1 = Vector3<S>(1, 0, 0);
    down.scalar_displacement = up.scalar_displacement = 2;
    sideways.scalar_displacement = 1;
    ContinuousCollisionRequest<S> creq;
    creq.num_max_contacts = 8;
    ContinuousCollisionResult<S> hit_r, miss_r, mesh_up;
    translational_ccd<S>(&mover, at(0, 0, 1), down, &floor, I, creq, hit_r);
    translational_ccd<S>(&mover, at(0, 0, 1), sideways, &floor, I, creq, miss_r);
    translational_ccd<S>(&floor, I, up, &mover, at(0, 0, 1), creq, mesh_up);  // the floor rises instead
    EXPECT_TRUE(hit_r.num_contacts() == 2 && miss_r.num_contacts() == 0 && mesh_up.num_contacts() == 2);
    if (hit_r.num_contacts() == 2 && mesh_up.num_contacts() == 2) {
      const auto& c0 = hit_r.raw_contacts()[0];
      EXPECT_TRUE(c0.o1 == &mover && c0.o2 == &floor && c0.b2 >= 0 && c0.b2 < 2 && c0.b2 != hit_r.raw_contacts()[1].b2);
      EXPECT_TRUE(c0.toc.lower_bound > S(0.3) && c0.toc.lower_bound < S(0.4));  // first touch after 0.75 m of 2 m
      EXPECT_TRUE(mesh_up.raw_contacts()[0].o1 == &mover && mesh_up.raw_contacts()[0].b2 == c0.b2);
    }
  }
  // mesh vs mesh: a second floor 1 m above the first, moving 2 m down, meets it (4 triangle pairs at most); moving up it does not
  {
    BVHModel<OBBRSS<S>> lid;
    lid.beginModel();
    lid.addSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)}, {{0, 1, 2}, {0, 2, 3}});
    lid.endModel();
    TranslationalDisplacement<S> down, up;
    down.unit_axis_in_shape1 = Vector3<S>(0, 0, -1);
    up.unit_axis_in_shape1 = Vector3<S>(0, 0, 1);
    down.scalar_displacement = up.scalar_displacement = 2;
    ContinuousCollisionRequest<S> creq;
    creq.num_max_contacts = 16;
    ContinuousCollisionResult<S> hit_r, miss_r;
    translational_ccd<S>(&lid, at(0, 0, 1), down, &floor, I, creq, hit_r);
    translational_ccd<S>(&lid, at(0, 0, 1), up, &floor, I, creq, miss_r);
    EXPECT_TRUE(hit_r.num_contacts() >= 2 && hit_r.num_contacts() <= 4 && miss_r.num_contacts() == 0);
    for (const auto& ct : hit_r.raw_contacts()) EXPECT_TRUE(ct.o1 == &lid && ct.o2 == &floor && ct.b1 >= 0 && ct.b1 < 2 && ct.b2 >= 0 && ct.b2 < 2);
  }
  // shape vs heightmap / octree: the bar dropping 1 m onto the four 0.5 m columns touches all four; the contact names the
  // pixel and carries its box; an upward sweep touches nothing; the same with the octree's four voxels
  {
    TranslationalDisplacement<S> down, up;
    down.unit_axis_in_shape1 = Vector3<S>(0, 0, -1);
    up.unit_axis_in_shape1 = Vector3<S>(0, 0, 1);
    down.scalar_displacement = up.scalar_displacement = 1;
    ContinuousCollisionRequest<S> creq;
    creq.num_max_contacts = 16;
    creq.request_type = TimeOfCollisionRequestType::kBoxApproximate;
    ContinuousCollisionResult<S> h_hit, h_miss, o_hit, hm_moves;
    translational_ccd<S>(&bar, at(S(0.2), S(0.05), S(1.0)), down, &hm, I, creq, h_hit);
    translational_ccd<S>(&bar, at(S(0.2), S(0.05), S(1.0)), up, &hm, I, creq, h_miss);
    translational_ccd<S>(&bar, at(S(0.2), S(0.05), S(0.6)), down, &oct, I, creq, o_hit);
    translational_ccd<S>(&hm, I, up, &bar, at(S(0.2), S(0.05), S(1.0)), creq, hm_moves);  // the map rises instead
    EXPECT_TRUE(h_hit.num_contacts() == 4 && h_miss.num_contacts() == 0 && o_hit.num_contacts() == 4 && hm_moves.num_contacts() == 4);
    for (const auto& ct : h_hit.raw_contacts()) {
      EXPECT_TRUE(ct.o1 == &bar && ct.o2 == &hm && (ct.b2 & 0xffff) == 8);
      EXPECT_TRUE(std::fabs(ct.o2_bv.max_[2] - S(0.5)) < S(1e-3) && ct.o2_bv.min_[2] == 0);
      EXPECT_TRUE(std::fabs(ct.toc.lower_bound - S(0.47)) < S(1e-3));  // bar bottom at 0.97 reaches the column tops at 0.5
    }
  }
  // heightmap / octree vs mesh and vs each other: the four 0.5 m columns rise 1 m into the floor mesh held 1 m above them
  // (o1 = the map, b1 = pixel, o1_bv its box, b2 = triangle -- in either argument order); into the octree's four voxels held
  // 1 m above (for (octree, heightmap) the reference names the map o1 as well); two copies of the octree sweeping through
  // each other
  {
    TranslationalDisplacement<S> down, up;
    down.unit_axis_in_shape1 = Vector3<S>(0, 0, -1);
    up.unit_axis_in_shape1 = Vector3<S>(0, 0, 1);
    down.scalar_displacement = up.scalar_displacement = 1;
    ContinuousCollisionRequest<S> creq;
    creq.num_max_contacts = 64;
    ContinuousCollisionResult<S> hm_mesh, hm_mesh_miss, mesh_hm, hm_oct, oct_hm, oct_oct, oct_oct_miss;
    translational_ccd<S>(&hm, I, up, &floor, at(0, 0, 1), creq, hm_mesh);
    translational_ccd<S>(&hm, I, down, &floor, at(0, 0, 1), creq, hm_mesh_miss);
    translational_ccd<S>(&floor, at(0, 0, 1), down, &hm, I, creq, mesh_hm);
    EXPECT_TRUE(hm_mesh.num_contacts() >= 4 && hm_mesh.num_contacts() <= 8 && hm_mesh_miss.num_contacts() == 0);
    EXPECT_TRUE(mesh_hm.num_contacts() == hm_mesh.num_contacts());
    for (const auto* r : {&hm_mesh, &mesh_hm})
      for (const auto& ct : r->raw_contacts()) {
        EXPECT_TRUE(ct.o1 == &hm && ct.o2 == &floor && (ct.b1 & 0xffff) == 8 && ct.b2 >= 0 && ct.b2 < 2);
        EXPECT_TRUE(std::fabs(ct.o1_bv.max_[2] - S(0.5)) < S(1e-3) && ct.o1_bv.min_[2] == 0);
      }
    translational_ccd<S>(&hm, I, up, &oct, at(0, 0, 1), creq, hm_oct);
    translational_ccd<S>(&oct, at(0, 0, 1), down, &hm, I, creq, oct_hm);
    EXPECT_TRUE(hm_oct.num_contacts() >= 4 && hm_oct.num_contacts() <= 16 && oct_hm.num_contacts() == hm_oct.num_contacts());
    for (const auto* r : {&hm_oct, &oct_hm})
      for (const auto& ct : r->raw_contacts()) {
        EXPECT_TRUE(ct.o1 == &hm && ct.o2 == &oct && (ct.b1 & 0xffff) == 8);
        EXPECT_TRUE(std::fabs(ct.o1_bv.max_[2] - S(0.5)) < S(1e-3) && std::fabs(ct.o2_bv.max_[2] - ct.o2_bv.min_[2] - S(0.1)) < S(1e-3));
        EXPECT_TRUE(std::fabs(ct.toc.lower_bound - S(0.5)) < S(1e-3));  // column tops at 0.5 reach the voxel bottoms at 1.0
      }
    translational_ccd<S>(&oct, I, up, &oct, at(0, 0, S(0.5)), creq, oct_oct);
    translational_ccd<S>(&oct, I, down, &oct, at(0, 0, S(0.5)), creq, oct_oct_miss);
    EXPECT_TRUE(oct_oct.num_contacts() >= 4 && oct_oct.num_contacts() <= 16 && oct_oct_miss.num_contacts() == 0);
    for (const auto& ct : oct_oct.raw_contacts()) EXPECT_TRUE(ct.o1 == &oct && ct.o2 == &oct && ct.b1 >= 0 && ct.b2 >= 0);
  }
  // replace protocol: the floor drops by 0.5 m (default arguments: refit bottom-up); a ball that touched it no longer does
  {
    BVHModel<OBBRSS<S>> sheet;
    sheet.beginModel();
    sheet.addSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)}, {{0, 1, 2}, {0, 2, 3}});
    sheet.endModel();
    CollisionRequest<S> one(1);
    CollisionResult<S> before, after, after_td;
    EXPECT_TRUE(collide<S>(&sheet, I, &ball, at(S(0.5), S(-0.5), S(0.2)), one, before) == 1);
    EXPECT_TRUE(sheet.beginReplaceModel() == 0);
    sheet.replaceSubModel({Vector3<S>(-1, -1, S(-0.5)), Vector3<S>(1, -1, S(-0.5)), Vector3<S>(1, 1, S(-0.5)), Vector3<S>(-1, 1, S(-0.5))});
    EXPECT_TRUE(sheet.endReplaceModel() == 0);
    EXPECT_TRUE(collide<S>(&sheet, I, &ball, at(S(0.5), S(-0.5), S(0.2)), one, after) == 0);
    EXPECT_TRUE(collide<S>(&sheet, I, &ball, at(S(0.5), S(-0.5), S(-0.3)), one, after) == 1);
    sheet.beginUpdateModel();
    sheet.updateSubModel({Vector3<S>(-1, -1, 0), Vector3<S>(1, -1, 0), Vector3<S>(1, 1, 0), Vector3<S>(-1, 1, 0)});
    EXPECT_TRUE(sheet.endUpdateModel(true, false) == 0);  // top-down refit
    EXPECT_TRUE(collide<S>(&sheet, I, &ball, at(S(0.5), S(-0.5), S(0.2)), one, after_td) == 1);
  }
  // UserContactProcessFunctor on the host: keep only contacts on triangle 1, stop after two
  std::vector<CollisionQuery<S>> one_q{{&floor, I, &ball, at(0, 0, S(0.1))}};
  std::vector<CollisionResult<S>> fr;
  int seen = 0;
  collideBatch<S>(one_q, all, [&](const Contact<S>& ct, bool& keep, bool& stop) { seen++; keep = ct.b1 == 1; stop = false; }, fr);
  EXPECT_TRUE(seen == 2 && fr[0].numContacts() == 1 && fr[0].getContact(0).b1 == 1);
}

// Contacts against the reference: the Python test writes a mesh, a box, poses and the contacts fcl::collide of the
// reference (oracle/_ref) reports with useDefaultPenetration(); every contact must come back bit for bit.
static void runReferenceFile(const char* path) {
  using namespace fcl;
  using S = double;
  std::FILE* f = std::fopen(path, "rb");
  if (!f) {
    std::printf("FAILED: cannot open %s\n", path);
    failures++;
    return;
  }
  int32_t hdr[4];
  EXPECT_TRUE(std::fread(hdr, 4, 4, f) == 4);
  const int nv = hdr[0], nt = hdr[1], nq = hdr[2], keep = hdr[3];
  std::vector<double> verts(3 * std::size_t(nv)), side(3), pm(12 * std::size_t(nq)), ps(12 * std::size_t(nq)),
      contacts(7 * std::size_t(nq) * keep);
  std::vector<int32_t> tris(3 * std::size_t(nt));
  std::vector<uint32_t> counts(nq);
  std::vector<int64_t> b1(std::size_t(nq) * keep);
  bool ok = std::fread(verts.data(), 8, verts.size(), f) == verts.size() && std::fread(tris.data(), 4, tris.size(), f) == tris.size() &&
            std::fread(side.data(), 8, 3, f) == 3 && std::fread(pm.data(), 8, pm.size(), f) == pm.size() &&
            std::fread(ps.data(), 8, ps.size(), f) == ps.size() && std::fread(counts.data(), 4, counts.size(), f) == counts.size() &&
            std::fread(b1.data(), 8, b1.size(), f) == b1.size() && std::fread(contacts.data(), 8, contacts.size(), f) == contacts.size();
  std::fclose(f);
  EXPECT_TRUE(ok);
  if (!ok) return;
  BVHModel<OBBRSS<S>> mesh;
  std::vector<Vector3<S>> pts;
  std::vector<std::array<int, 3>> tri;
  for (int i = 0; i < nv; i++) pts.emplace_back(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
  for (int i = 0; i < nt; i++) tri.push_back({tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]});
  mesh.beginModel();
  mesh.addSubModel(pts, tri);
  mesh.endModel();
  Box<S> box(side[0], side[1], side[2]);
  auto pose = [](const double* p) {
    Transform3<S> t;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) t.linear()(i, j) = p[3 * i + j];
    for (int i = 0; i < 3; i++) t.translation()[i] = p[9 + i];
    return t;
  };
  std::vector<CollisionQuery<S>> qs;
  for (int q = 0; q < nq; q++) {
    if (q & 1)
      qs.push_back({&box, pose(&ps[12 * q]), &mesh, pose(&pm[12 * q])});
    else
      qs.push_back({&mesh, pose(&pm[12 * q]), &box, pose(&ps[12 * q])});
  }
  CollisionRequest<S> req{std::size_t(keep)};
  req.useDefaultPenetration();
  std::vector<CollisionResult<S>> res;
  collideBatch(qs, req, res);
  std::size_t compared = 0, bad = 0;
  for (int q = 0; q < nq; q++) {
    EXPECT_TRUE(res[q].numContacts() == counts[q]);
    if (res[q].numContacts() != counts[q]) continue;
    for (const auto& c : res[q].getContacts()) {
      bool found = false;
      for (uint32_t k = 0; k < counts[q] && !found; k++) {
        const double* r = &contacts[(std::size_t(q) * keep + k) * 7];
        found = b1[std::size_t(q) * keep + k] == int64_t(c.b1) && r[0] == c.normal[0] && r[1] == c.normal[1] && r[2] == c.normal[2] &&
                r[3] == c.pos[0] && r[4] == c.pos[1] && r[5] == c.pos[2] && r[6] == c.penetration_depth;
      }
      compared++;
      if (!found) bad++;
    }
  }
  std::printf("reference file: %d queries, %zu contacts compared, %zu not bit-identical\n", nq, compared, bad);
  EXPECT_TRUE(compared > 0 && bad == 0);
}

int main(int argc, char** argv) {
  runBatch<float>();
  runBatch<double>();
  if (argc > 1) runReferenceFile(argv[1]);
  run<float>();
  run<double>();
  runScene<float>();
  runScene<double>();
  std::printf(failures ? "HOST API: %d FAILURES\n" : "HOST API: ALL OK\n", failures);
  return failures ? 1 : 0;
}
