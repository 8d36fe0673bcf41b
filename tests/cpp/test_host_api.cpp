// Host-side C++ API check, written the way the reference's own tests read
// (test/test_fcl_geometric_shapes.cpp shapeDistance_*, test_fcl_collision_penetration.cpp):
// known answers through fcl::collide / fcl::distance of include/fcl_b200/fcl.h.
// Built and run by tests/test_host_api_gpu.py on the GPU box.
#include <cmath>
#include <cstdio>

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

int main() {
  run<float>();
  run<double>();
  std::printf(failures ? "HOST API: %d FAILURES\n" : "HOST API: ALL OK\n", failures);
  return failures ? 1 : 0;
}
