// Host-only part of the C++ mirror (no GPU): octree2::Octree<S>::rebuildTree / isPointOccupied over fclb_octree_build_host.
#include "fcl_b200/fcl.h"
using namespace fcl;
template <typename S> int run() {
  auto tree = std::make_shared<octree2::Octree<S>>(S(0.1), std::uint16_t(8));
  tree->rebuildTree([](int i, S& x, S& y, S& z) { x = S(0.05) + S(0.1) * S(i % 4); y = S(0.05); z = S(0.05); }, 8);
  int bad = 0;
  bad += !tree->isPointOccupied(Vector3<S>(S(0.17), S(0.02), S(0.09)));
  bad += !tree->isPointOccupied(Vector3<S>(S(0.35), S(0.05), S(0.05)));
  bad += tree->isPointOccupied(Vector3<S>(S(0.45), S(0.05), S(0.05)));
  bad += tree->isPointOccupied(Vector3<S>(S(0.05), S(0.15), S(0.05)));
  bad += tree->isPointOccupied(Vector3<S>(S(-0.05), S(0.05), S(0.05)));
  bad += tree->isPointOccupied(Vector3<S>(S(0.9), 0, 0));
  // full grid: root fully occupied
  auto full = std::make_shared<octree2::Octree<S>>(S(1), std::uint16_t(2));
  full->rebuildTree([](int i, S& x, S& y, S& z) { x = S(i % 4) - S(1.5); y = S((i / 4) % 4) - S(1.5); z = S(i / 16) - S(1.5); }, 64);
  bad += !(full->inner_nodes_fully_occupied()[0] == 1 && full->isPointOccupied(Vector3<S>(S(1.9), S(-1.9), S(0.1))));
  return bad;
}
int main() { int b = run<float>() + run<double>(); std::printf("bad=%d\n", b); return b; }
