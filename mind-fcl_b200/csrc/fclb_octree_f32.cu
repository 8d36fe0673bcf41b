// explicit instantiation of the octree-shape traversal for S = float
#include "fclb_octree_impl.cuh"
namespace fclb {
template cudaError_t launchOctreeShape<float>(int, const OctreeArgs&, int, cudaStream_t);
}
