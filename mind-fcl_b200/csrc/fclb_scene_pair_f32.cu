// explicit instantiation of the scene-pair traversal for S = float
#include "fclb_scene_pair_impl.cuh"
namespace fclb {
template cudaError_t launchScenePair<float>(const ScenePairArgs&, int, cudaStream_t);
}
