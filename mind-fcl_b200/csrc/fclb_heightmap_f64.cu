// explicit instantiation of the heightmap-shape scan for S = double
#include "fclb_heightmap_impl.cuh"
namespace fclb {
template cudaError_t launchHeightmapShape<double>(int, const HeightmapArgs&, int, cudaStream_t);
}
