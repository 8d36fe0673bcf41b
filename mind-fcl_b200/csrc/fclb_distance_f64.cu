// explicit instantiation of the batched distance path for S = double
#include "fclb_distance_impl.cuh"
namespace fclb {
template cudaError_t launchDistance<double>(const BatchView&, const SolverParams&, const DistanceOut&, cudaStream_t, int*);
}
