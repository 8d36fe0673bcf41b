// explicit instantiation of the scene-pair traversal for S = double
#include "fclb_scene_pair_impl.cuh"
namespace fclb {
template cudaError_t launchScenePair<double>(const ScenePairArgs&, int, cudaStream_t);
}
