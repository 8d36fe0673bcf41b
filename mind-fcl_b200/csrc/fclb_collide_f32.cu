// explicit instantiation of the batched collide path for S = float
#include "fclb_collide_launch.cuh"
namespace fclb {
template cudaError_t launchCollide<float>(const BatchView&, const CollideLaunchArgs&, cudaStream_t, int*);
template cudaError_t launchMprPenetration<float>(const BatchView&, const CollideLaunchArgs&, cudaStream_t);
}
