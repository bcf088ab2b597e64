// main_engine_driver.cpp — TEST / INTEGRATION EVIDENCE. Drives the reference's own top-level object, ITMMainEngine
// (Engine/ITMMainEngine.h), built from the PATCHED ITMLib (integration/itmlib_b200.patch): the constructor's
// `case DEVICE_CUDA` consults settings->engineBackend and, for BACKEND_B200, builds the B200 shim engines behind the
// abstract interfaces; everything downstream — ITMMainEngine::ProcessFrame -> view builder -> ITMDenseMapper::ProcessFrame ->
// ITMTrackingController::Prepare, GetImage — is the reference's unmodified host code. This is the path
// DS/InfiniTamDriver (a subclass of ITMMainEngine) takes in DynSLAM.
// Built by integration/build_patched.sh into oracle/_ref/libitmpatched.so.
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "ITMLib/Engine/ITMMainEngine.h"

using namespace ITMLib::Engine;
using namespace ITMLib::Objects;

struct Driver {
  ITMLibSettings *settings;
  ITMRGBDCalib calib;
  ITMMainEngine *engine;
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
    D->engine = new ITMMainEngine(D->settings, &D->calib, D->size, D->size);
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

// one frame through ITMMainEngine::ProcessFrame with the externally supplied pose (DynSLAM runs TRACKER_EXTERNAL and sets the
// pose from libviso2: DS/InfiniTamDriver.h SetPose); returns 0, 2 on a std::runtime_error (VBA / excess list exhausted)
int med_process_frame(Driver *D, const short *rawDepth, const unsigned char *rgba, const float *M_d) {
  if (!D->engine) return 1;
  const size_t n = (size_t)D->size.x * D->size.y;
  memcpy(D->raw->GetData(MEMORYDEVICE_CPU), rawDepth, n * sizeof(short));
  memcpy(D->rgb->GetData(MEMORYDEVICE_CPU), rgba, n * 4);
  Matrix4f M; for (int i = 0; i < 16; ++i) M.m[i] = M_d[i];
  D->engine->GetTrackingState()->pose_d->SetM(M);
  try {
    D->engine->ProcessFrame(D->rgb, D->raw);
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
  *noVisibleBlocks = -1;
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
