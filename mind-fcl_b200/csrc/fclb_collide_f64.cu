// explicit instantiation of the batched collide path for S = double
#include "fclb_collide_launch.cuh"
namespace fclb {
template cudaError_t launchCollide<double>(const BatchView&, const CollideLaunchArgs&, cudaStream_t, int*);
template cudaError_t launchMprPenetration<double>(const BatchView&, const CollideLaunchArgs&, cudaStream_t);
}
