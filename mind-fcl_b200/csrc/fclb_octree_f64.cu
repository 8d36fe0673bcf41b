// explicit instantiation of the octree-shape traversal for S = double
#include "fclb_octree_impl.cuh"
namespace fclb {
template cudaError_t launchOctreeShape<double>(int, const OctreeArgs&, int, cudaStream_t);
}
