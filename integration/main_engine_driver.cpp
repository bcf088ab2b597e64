// main_engine_driver.cpp — TEST / INTEGRATION EVIDENCE. Drives the reference's own top-level object, ITMMainEngine
// (Engine/ITMMainEngine.h), built from the PATCHED ITMLib (integration/itmlib_b200.patch): the constructor's
// `case DEVICE_CUDA` consults settings->engineBackend and, for BACKEND_B200, builds the B200 shim engines behind the
// abstract interfaces; everything downstream is the reference's unmodified host code. DynSLAM never calls
// ITMMainEngine::ProcessFrame (it runs TRACKER_EXTERNAL, whose TrackCamera throws): its InfiniTamDriver SUBCLASSES
// ITMMainEngine and calls the protected members in this order per frame (DS/InfiniTamDriver.h:132-158, :201-206,
// DS/InfiniTamDriver.cpp:211-224): viewBuilder->UpdateView, SetPose, denseMapper->ProcessFrame, trackingController->Prepare,
// denseMapper->Decay. The subclass below does exactly that.
// Built by integration/build_patched.sh into oracle/_ref/libitmpatched.so.
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "ITMLib/Engine/ITMMainEngine.h"

using namespace ITMLib::Engine;
using namespace ITMLib::Objects;

// what DS/InfiniTamDriver does with the protected members of ITMMainEngine
class DriverEngine : public ITMMainEngine {
 public:
  DriverEngine(const ITMLibSettings *settings, const ITMRGBDCalib *calib, Vector2i size) : ITMMainEngine(settings, calib, size, size) {}
  void UpdateView(ITMUChar4Image *rgb, ITMShortImage *raw) {
    this->viewBuilder->UpdateView(&view, rgb, raw, settings->useBilateralFilter, settings->modelSensorNoise);
  }
  void SetPoseM(const Matrix4f &M) { this->trackingState->pose_d->SetM(M); }
  void Integrate() { this->denseMapper->ProcessFrame(this->view, this->trackingState, this->scene, this->renderState_live); }
  void PrepareNextStep() {
    ITMRenderState_VH *rs = (ITMRenderState_VH *)this->renderState_live;
    if (rs->noVisibleBlocks > 0) this->trackingController->Prepare(this->trackingState, this->view, this->renderState_live);
  }
  void Decay(int maxWeight, int minAge) { this->denseMapper->Decay(scene, renderState_live, maxWeight, minAge, false); }
  int NoVisibleBlocks() { return ((ITMRenderState_VH *)this->renderState_live)->noVisibleBlocks; }
};

struct Driver {
  ITMLibSettings *settings;
  ITMRGBDCalib calib;
  DriverEngine *engine;
  ITMUChar4Image *rgb;
  ITMShortImage *raw;
  ITMUChar4Image *out;
  Vector2i size;
  char err[256];
};

extern "C" {

// backend: 0 = BACKEND_REFERENCE (the bundled CUDA engines), 1 = BACKEND_B200
Driver *med_create(int backend, int w, int h, float fx, float fy, float cx, float cy, float voxelSize, float mu, long numBlocks) {
  Driver *D = new Driver();
  D->err[0] = 0;
  D->size = Vector2i(w, h);
  D->settings = new ITMLibSettings();
  D->settings->engineBackend = backend ? ITMLibSettings::BACKEND_B200 : ITMLibSettings::BACKEND_REFERENCE;
  D->settings->sceneParams.voxelSize = voxelSize;
  D->settings->sceneParams.mu = mu;
  D->settings->sdfLocalBlockNum = numBlocks;
  D->settings->createMeshingEngine = false;
  D->calib.intrinsics_d.SetFrom(fx, fy, cx, cy, (float)w, (float)h);
  D->calib.intrinsics_rgb.SetFrom(fx, fy, cx, cy, (float)w, (float)h);
  D->calib.disparityCalib.type = ITMDisparityCalib::TRAFO_AFFINE;       // raw depth in millimetres (DynSLAM's KITTI input)
  D->calib.disparityCalib.params = Vector2f(1.0f / 1000.0f, 0.0f);
  try {
    D->engine = new DriverEngine(D->settings, &D->calib, D->size);
  } catch (std::exception &e) {
    snprintf(D->err, sizeof(D->err), "%s", e.what());
    D->engine = NULL;
  }
  D->rgb = new ITMUChar4Image(D->size, true, true);
  D->raw = new ITMShortImage(D->size, true, true);
  D->out = new ITMUChar4Image(D->size, true, true);
  return D;
}

const char *med_error(Driver *D) { return D->err; }

// one frame, InfiniTamDriver's sequence, pose supplied externally (DynSLAM sets it from libviso2: DS/InfiniTamDriver.h SetPose);
// returns 0, 2 on a std::runtime_error (VBA / excess list exhausted)
int med_process_frame(Driver *D, const short *rawDepth, const unsigned char *rgba, const float *M_d) {
  if (!D->engine) return 1;
  const size_t n = (size_t)D->size.x * D->size.y;
  memcpy(D->raw->GetData(MEMORYDEVICE_CPU), rawDepth, n * sizeof(short));
  memcpy(D->rgb->GetData(MEMORYDEVICE_CPU), rgba, n * 4);
  Matrix4f M; for (int i = 0; i < 16; ++i) M.m[i] = M_d[i];
  try {
    D->engine->UpdateView(D->rgb, D->raw);
    D->engine->SetPoseM(M);
    D->engine->Integrate();
    D->engine->PrepareNextStep();
    D->engine->Decay(1, 3);
  } catch (std::runtime_error &e) {
    snprintf(D->err, sizeof(D->err), "%s", e.what());
    return 2;
  }
  return 0;
}

// ITMMainEngine::GetImage(InfiniTAM_IMAGE_SCENERAYCAST): the live raycast image, copied to the host by the reference's code
void med_get_raycast_image(Driver *D, unsigned char *rgbaOut) {
  D->engine->GetImage(D->out, NULL, ITMMainEngine::InfiniTAM_IMAGE_SCENERAYCAST);
  memcpy(rgbaOut, D->out->GetData(MEMORYDEVICE_CPU), (size_t)D->size.x * D->size.y * 4);
}

void med_counters(Driver *D, int *lastFreeBlockId, int *noVisibleBlocks, int *allocatedEntries) {
  ITMScene<ITMVoxel, ITMVoxelIndex> *scene = D->engine->GetScene();
  *lastFreeBlockId = scene->localVBA.lastFreeBlockId;
  *noVisibleBlocks = D->engine->NoVisibleBlocks();
  // count allocated entries on the host
  const int n = ITMVoxelBlockHash::noTotalEntries;
  ITMHashEntry *h = new ITMHashEntry[n];
  ORcudaSafeCall(cudaMemcpy(h, scene->index.GetEntries(), sizeof(ITMHashEntry) * (size_t)n, cudaMemcpyDeviceToHost));
  int a = 0;
  for (int i = 0; i < n; ++i) a += (h[i].ptr >= 0);
  delete[] h;
  *allocatedEntries = a;
}

void med_destroy(Driver *D) {
  delete D->engine; delete D->rgb; delete D->raw; delete D->out; delete D->settings;
  delete D;
}

}  // extern "C"
