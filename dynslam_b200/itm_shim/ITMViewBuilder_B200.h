// ITMViewBuilder_B200.h — C++ host side of the view builder: an ITMViewBuilder
// subclass (Engine/ITMViewBuilder.h:19-58) that forwards to libb200fusion. Header-only, compiled against the
// unmodified reference headers. It takes the place of ITMViewBuilder_CUDA where ITMMainEngine picks the view
// builder (Engine/ITMMainEngine.cpp:24-54):
//     case ITMLibSettings::DEVICE_B200: viewBuilder = new ITMViewBuilder_B200(calib, handle); break;
// ITMView / ITMViewIMU objects are allocated exactly as ITMViewBuilder_CUDA allocates them (:37-49, :99-106);
// the floatImage scratch of the base class is not needed (the five filter passes run inside one kernel).
#pragma once

#include "ITMEngines_B200.h"

#include "ITMLib/Engine/ITMViewBuilder.h"

namespace ITMLib {
namespace Engine {

class ITMViewBuilder_B200 : public ITMViewBuilder {
  std::shared_ptr<B200EngineHandle> h;

  b200_view_calib Calib(const ITMView *view, bool useBilateralFilter, bool modelSensorNoise) const {
    b200_view_calib c{};
    const ITMRGBDCalib *cal = view->calib;
    c.trafoType = cal->disparityCalib.type == ITMDisparityCalib::TRAFO_KINECT ? 0 : 1;
    c.params[0] = cal->disparityCalib.params.x; c.params[1] = cal->disparityCalib.params.y;
    c.fx_depth = cal->intrinsics_d.projectionParamsSimple.fx;
    const Vector4f all = cal->intrinsics_d.projectionParamsSimple.all;
    c.intrinsics_d[0] = all.x; c.intrinsics_d[1] = all.y; c.intrinsics_d[2] = all.z; c.intrinsics_d[3] = all.w;
    c.useBilateralFilter = useBilateralFilter ? 1 : 0;
    c.modelSensorNoise = modelSensorNoise ? 1 : 0;
    return c;
  }

 public:
  ITMViewBuilder_B200(const ITMRGBDCalib *calib, std::shared_ptr<B200EngineHandle> handle) : ITMViewBuilder(calib), h(std::move(handle)) {}
  ~ITMViewBuilder_B200() {}

  // ITMViewBuilder_CUDA.cu:120-135
  void ConvertDisparityToDepth(ITMFloatImage *depth_out, const ITMShortImage *disp_in, const ITMIntrinsics *depthIntrinsics,
                               Vector2f disparityCalibParams) override {
    h->check(b200_convert_disparity_to_depth(h->e, depth_out->GetData(MEMORYDEVICE_CUDA), disp_in->GetData(MEMORYDEVICE_CUDA),
                                             disp_in->noDims.x, disp_in->noDims.y, disparityCalibParams.x, disparityCalibParams.y,
                                             depthIntrinsics->projectionParamsSimple.fx));
  }

  // ITMViewBuilder_CUDA.cu:137-148
  void ConvertDepthAffineToFloat(ITMFloatImage *depth_out, const ITMShortImage *depth_in, Vector2f depthCalibParams) override {
    h->check(b200_convert_depth_affine_to_float(h->e, depth_out->GetData(MEMORYDEVICE_CUDA), depth_in->GetData(MEMORYDEVICE_CUDA),
                                                depth_in->noDims.x, depth_in->noDims.y, depthCalibParams.x, depthCalibParams.y));
  }

  // ITMViewBuilder_CUDA.cu:150-161 (one pass)
  void DepthFiltering(ITMFloatImage *image_out, const ITMFloatImage *image_in) override {
    h->check(b200_depth_filtering(h->e, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA),
                                  image_in->noDims.x, image_in->noDims.y));
  }

  // ITMViewBuilder_CUDA.cu:163-178
  void ComputeNormalAndWeights(ITMFloat4Image *normal_out, ITMFloatImage *sigmaZ_out, const ITMFloatImage *depth_in,
                               Vector4f intrinsic) override {
    const float intr[4] = {intrinsic.x, intrinsic.y, intrinsic.z, intrinsic.w};
    h->check(b200_compute_normal_and_weights(h->e, (b200_vec4f *)normal_out->GetData(MEMORYDEVICE_CUDA),
                                             sigmaZ_out->GetData(MEMORYDEVICE_CUDA), depth_in->GetData(MEMORYDEVICE_CUDA),
                                             depth_in->noDims.x, depth_in->noDims.y, intr));
  }

  // ITMViewBuilder_CUDA.cu:33-84
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMShortImage *rawDepthImage, bool useBilateralFilter,
                  bool modelSensorNoise = false) override {
    if (*view_ptr == NULL) {
      *view_ptr = new ITMView(calib, rgbImage->noDims, rawDepthImage->noDims, true);
      if (this->shortImage != NULL) delete this->shortImage;
      this->shortImage = new ITMShortImage(rawDepthImage->noDims, true, true);
      if (modelSensorNoise) {
        (*view_ptr)->depthNormal = new ITMFloat4Image(rawDepthImage->noDims, true, true);
        (*view_ptr)->depthUncertainty = new ITMFloatImage(rawDepthImage->noDims, true, true);
      }
    }
    ITMView *view = *view_ptr;
    view->rgb->SetFrom(rgbImage, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CUDA);
    this->shortImage->SetFrom(rawDepthImage, ORUtils::MemoryBlock<short>::CPU_TO_CUDA);
    b200_view_calib c = Calib(view, useBilateralFilter, modelSensorNoise);
    // enqueue only: the reference's kernels are asynchronous too, and every later engine call of this volume runs on
    // the same stream
    h->check(b200_update_view_async(h->e, this->shortImage->GetData(MEMORYDEVICE_CUDA), rawDepthImage->noDims.x, rawDepthImage->noDims.y,
                                    &c, view->depth->GetData(MEMORYDEVICE_CUDA),
                                    modelSensorNoise ? (b200_vec4f *)view->depthNormal->GetData(MEMORYDEVICE_CUDA) : nullptr,
                                    modelSensorNoise ? view->depthUncertainty->GetData(MEMORYDEVICE_CUDA) : nullptr));
  }

  // ITMViewBuilder_CUDA.cu:86-95
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMFloatImage *depthImage) override {
    if (*view_ptr == NULL) *view_ptr = new ITMView(calib, rgbImage->noDims, depthImage->noDims, true);
    ITMView *view = *view_ptr;
    view->rgb->UpdateDeviceFromHost();
    view->depth->UpdateDeviceFromHost();
  }

  // ITMViewBuilder_CUDA.cu:97-113
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMShortImage *depthImage, bool useBilateralFilter,
                  ITMIMUMeasurement *imuMeasurement) override {
    if (*view_ptr == NULL) {
      *view_ptr = new ITMViewIMU(calib, rgbImage->noDims, depthImage->noDims, true);
      if (this->shortImage != NULL) delete this->shortImage;
      this->shortImage = new ITMShortImage(depthImage->noDims, true, true);
    }
    ITMViewIMU *imuView = (ITMViewIMU *)(*view_ptr);
    imuView->imu->SetFrom(imuMeasurement);
    this->UpdateView(view_ptr, rgbImage, depthImage, useBilateralFilter);
  }
};

}  // namespace Engine
}  // namespace ITMLib
