// explicit instantiation of the mesh-shape traversal for S = float
#include "fclb_bvh_shape_impl.cuh"
namespace fclb {
template cudaError_t launchBvhShape<float>(int, const BvhShapeArgs&, int, cudaStream_t);
}
