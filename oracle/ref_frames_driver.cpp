// ref_frames_driver.cpp — TEST INFRASTRUCTURE ONLY. Pins oracle/frames_oracle.c to the reference's own code.
//
// ProcessSilhouette_CPU / RemoveSilhouette_CPU / CompositeDepth / CompositeColor are free functions inside
// DS/InstRecLib/InstanceReconstructor.cpp, a translation unit that needs OpenCV, Eigen and Pangolin (none installed). Their
// bodies only need: ORUtils images and vectors (real headers, /root/reference), instreclib::utils::Mask + BoundingBox (real
// headers, /root/reference/src/DynSLAM/InstRecLib/Utils), a byte matrix with at<T>() behind Mask (oracle/stubs/opencv2, ours),
// and two Eigen fixed-size integer vectors (index access only; the two 6-line structs below, ours).
// oracle/build_ref.sh cuts the four functions out of the reference file AT BUILD TIME into oracle/_ref/instrec_extract.inc
// (git-ignored; nothing of the reference is stored in this repo) and this driver #includes that file unmodified.
#include <algorithm>
#include <cassert>
#include <cstring>
#include <limits>
#include <sys/types.h>

#include "ITMLib/Utils/ITMLibDefines.h"   // Vector4u, ITMFloatImage, ITMUChar4Image (reference)
#include "Mask.h"                          // reference InstRecLib/Utils/Mask.h -> BoundingBox.h, <opencv2/opencv.hpp> (stub)

#include "../include/b200fusion.h"

namespace Eigen {
struct Vector2i { int v[2]; int operator[](int i) const { return v[i]; } };
struct Vector4i { int v[4]; int operator()(int i) const { return v[i]; } };
}   // namespace Eigen

using namespace std;
using namespace instreclib::utils;

#include "_ref/instrec_extract.inc"

static Mask make_mask(const b200_mask *m, const unsigned char *data) {
  const int w = m->x1 - m->x0 + 1, h = m->y1 - m->y0 + 1;
  cv::Mat1b *mat = new cv::Mat1b(h > 0 ? h : 0, w > 0 ? w : 0);
  if (w > 0 && h > 0) memcpy(mat->bytes.data(), data, (size_t)w * h);
  return Mask(BoundingBox(m->x0, m->y0, m->x1, m->y1), mat);   // the Mask owns (and deletes) the matrix; never copied here
}

static_assert(sizeof(Vector4u) == sizeof(b200_vec4u), "vec4u layout");

extern "C" {

// ProcessSilhouette_CPU<float> (InstanceReconstructor.cpp:59-135); the delete mask is unused by the reference body
void ref_process_silhouette(b200_vec4u *src_rgb, float *src_depth, b200_vec4u *dst_rgb, float *dst_depth, int w, int h,
                            const b200_mask *copy_mask) {
  Mask cm = make_mask(copy_mask, copy_mask->d_data);
  Eigen::Vector2i dims = {{w, h}};
  ProcessSilhouette_CPU<float>(reinterpret_cast<Vector4u *>(src_rgb), src_depth, reinterpret_cast<Vector4u *>(dst_rgb), dst_depth, dims,
                               cm, cm);
}

// RemoveSilhouette_CPU<float> (:137-170)
void ref_remove_silhouette(b200_vec4u *rgb, float *depth, int w, int h, const b200_mask *mask) {
  Mask m = make_mask(mask, mask->d_data);
  Eigen::Vector2i dims = {{w, h}};
  RemoveSilhouette_CPU<float>(reinterpret_cast<Vector4u *>(rgb), depth, dims, m);
}

// CompositeDepth (:850-869) on host images of w x h
void ref_composite_depth(float *target, const float *source, int w, int h) {
  ITMFloatImage t(Vector2i(w, h), MEMORYDEVICE_CPU), s(Vector2i(w, h), MEMORYDEVICE_CPU);
  memcpy(t.GetData(MEMORYDEVICE_CPU), target, sizeof(float) * w * h);
  memcpy(s.GetData(MEMORYDEVICE_CPU), source, sizeof(float) * w * h);
  CompositeDepth(&t, &s);
  memcpy(target, t.GetData(MEMORYDEVICE_CPU), sizeof(float) * w * h);
}

// CompositeColor (:873-905)
void ref_composite_color(b200_vec4u *t_color, float *t_depth, const b200_vec4u *s_color, const float *s_depth, int w, int h,
                         const int32_t tint[4], float tint_strength) {
  ITMUChar4Image tc(Vector2i(w, h), MEMORYDEVICE_CPU), sc(Vector2i(w, h), MEMORYDEVICE_CPU);
  ITMFloatImage td(Vector2i(w, h), MEMORYDEVICE_CPU), sd(Vector2i(w, h), MEMORYDEVICE_CPU);
  memcpy(tc.GetData(MEMORYDEVICE_CPU), t_color, 4 * (size_t)w * h);
  memcpy(sc.GetData(MEMORYDEVICE_CPU), s_color, 4 * (size_t)w * h);
  memcpy(td.GetData(MEMORYDEVICE_CPU), t_depth, sizeof(float) * w * h);
  memcpy(sd.GetData(MEMORYDEVICE_CPU), s_depth, sizeof(float) * w * h);
  Eigen::Vector4i tn = {{tint[0], tint[1], tint[2], tint[3]}};
  CompositeColor(&tc, &td, &sc, &sd, tn, tint_strength);
  memcpy(t_color, tc.GetData(MEMORYDEVICE_CPU), 4 * (size_t)w * h);
  memcpy(t_depth, td.GetData(MEMORYDEVICE_CPU), sizeof(float) * w * h);
}

}   // extern "C"
