// explicit instantiation of the EPA stage for S = float
#include "fclb_epa_launch.cuh"
namespace fclb {
template cudaError_t launchEpa<float>(const BatchView&, const CollideLaunchArgs&, cudaStream_t);
}
