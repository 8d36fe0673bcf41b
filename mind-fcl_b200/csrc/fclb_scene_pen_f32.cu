// explicit instantiation of the scene-contact MPR penetration pass for S = float
#include "fclb_scene_pen_impl.cuh"
namespace fclb {
template cudaError_t launchScenePenetration<float>(const ScenePenArgs&, cudaStream_t);
template cudaError_t launchScenePairPenetration<float>(const ScenePairPenArgs&, cudaStream_t);
}
