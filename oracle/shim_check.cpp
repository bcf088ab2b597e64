// Compile-only check (TEST INFRASTRUCTURE): instantiates the ITMLib shim classes against the
// unmodified reference headers, proving the drop-in compiles behind the real virtual interfaces.
#include "ITMEngines_B200.h"
#include "ITMViewBuilder_B200.h"

using namespace ITMLib::Engine;

template class ITMLib::Engine::ITMSceneReconstructionEngine_B200<ITMVoxel, ITMVoxelIndex>;
template class ITMLib::Engine::ITMVisualisationEngine_B200<ITMVoxel, ITMVoxelIndex>;
template class ITMLib::Engine::ITMSwappingEngine_B200<ITMVoxel, ITMVoxelIndex>;
template class ITMLib::Engine::ITMMeshingEngine_B200<ITMVoxel, ITMVoxelIndex>;

// what ITMDenseMapper's / ITMMainEngine's new `case ITMLibSettings::DEVICE_B200:` would do
// (Engine/ITMDenseMapper.cpp:16-36, Engine/ITMMainEngine.cpp:24-54)
void *make_engines(ITMScene<ITMVoxel, ITMVoxelIndex> *scene, const ITMLibSettings *settings, Vector2i imgSize) {
  auto h = std::make_shared<B200EngineHandle>(0, settings->sdfLocalBlockNum, imgSize);
  ITMSceneReconstructionEngine<ITMVoxel, ITMVoxelIndex> *reco = new ITMSceneReconstructionEngine_B200<ITMVoxel, ITMVoxelIndex>(h);
  IITMVisualisationEngine *vis = new ITMVisualisationEngine_B200<ITMVoxel, ITMVoxelIndex>(scene, settings, h);
  ITMSwappingEngine<ITMVoxel, ITMVoxelIndex> *swap = new ITMSwappingEngine_B200<ITMVoxel, ITMVoxelIndex>(h);
  ITMViewBuilder *vb = new ITMViewBuilder_B200(nullptr, h);   // Engine/ITMMainEngine.cpp:24-54 picks the view builder
  (void)vis; (void)swap; (void)vb;
  return reco;
}
