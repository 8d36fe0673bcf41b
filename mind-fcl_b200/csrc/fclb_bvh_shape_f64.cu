// explicit instantiation of the mesh-shape traversal for S = double
#include "fclb_bvh_shape_impl.cuh"
namespace fclb {
template cudaError_t launchBvhShape<double>(int, const BvhShapeArgs&, int, cudaStream_t);
}
