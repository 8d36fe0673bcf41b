// explicit instantiation of the scene-contact MPR penetration pass for S = double
#include "fclb_scene_pen_impl.cuh"
namespace fclb {
template cudaError_t launchScenePenetration<double>(const ScenePenArgs&, cudaStream_t);
template cudaError_t launchScenePairPenetration<double>(const ScenePairPenArgs&, cudaStream_t);
}
