// explicit instantiation of the EPA stage for S = double
#include "fclb_epa_launch.cuh"
namespace fclb {
template cudaError_t launchEpa<double>(const BatchView&, const CollideLaunchArgs&, cudaStream_t);
}
