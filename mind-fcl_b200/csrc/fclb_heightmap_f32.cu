// explicit instantiation of the heightmap-shape scan for S = float
#include "fclb_heightmap_impl.cuh"
namespace fclb {
template cudaError_t launchHeightmapShape<float>(int, const HeightmapArgs&, int, cudaStream_t);
}
