/* view_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported or called by dynslam_b200/).
 *
 * Serial C restatement of the reference's view builder, CUDA-build semantics:
 *   ITMViewBuilder_CUDA::UpdateView                Engine/DeviceSpecific/CUDA/ITMViewBuilder_CUDA.cu:33-84
 *   convertDisparityToDepth / convertDepthAffineToFloat / filterDepth / computeNormalAndWeight
 *                                                  Engine/DeviceAgnostic/ITMViewBuilder.h:7-114
 * Pinned bit for bit against those reference functions compiled from /root/reference
 * (oracle/ref_driver.cpp: ref_view_*; tests/test_oracle_vs_ref.py). exp/sqrt/acos are the host libm's
 * single-precision functions, exactly what the reference's host build calls; the CUDA path uses the CUDA
 * math library and is therefore compared within a tolerance (tests/test_gpu_view.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200fusion.h"

#define MEAN_SIGMA_L 1.2232f                                /* DA/ITMViewBuilder.h:30 */
#define ITM_PI ((float)3.1415926535897932384626433832795)   /* ORUtils/MathUtils.h:26 */

/* DA/ITMViewBuilder.h:7-20 */
void oracle_convert_disparity_to_depth(float *d_out, const int16_t *d_in, int w, int h, float p0, float p1, float fx_depth) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int locId = x + y * w;
    short disparity = d_in[locId];
    float disparity_tmp = p0 - (float)(disparity);
    float depth;
    if (disparity_tmp == 0) depth = 0.0;
    else depth = 8.0f * p1 * fx_depth / disparity_tmp;
    d_out[locId] = (depth > 0) ? depth : -1.0f;
  }
}

/* DA/ITMViewBuilder.h:22-28 */
void oracle_convert_depth_affine_to_float(float *d_out, const int16_t *d_in, int w, int h, float p0, float p1) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int locId = x + y * w;
    short depth_in = d_in[locId];
    d_out[locId] = ((depth_in <= 0) || (depth_in > 32000)) ? -1.0f : (float)depth_in * p0 + p1;
  }
}

/* DA/ITMViewBuilder.h:31-56 */
static void filter_depth(float *out, const float *in, int x, int y, int w) {
  float z, tmpz, dz, final_depth = 0.0f, wgt, w_sum = 0.0f;
  z = in[x + y * w];
  if (z < 0.0f) { out[x + y * w] = -1.0f; return; }
  float sigma_z = 1.0f / (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * 0.25f);
  for (int i = -2; i <= 2; i++) for (int j = -2; j <= 2; j++) {
    tmpz = in[(x + j) + (y + i) * w];
    if (tmpz < 0.0f) continue;
    dz = (tmpz - z); dz *= dz;
    wgt = expf(-0.5f * ((float)(abs(i) + abs(j)) * MEAN_SIGMA_L * MEAN_SIGMA_L + dz * sigma_z * sigma_z));
    w_sum += wgt;
    final_depth += wgt * tmpz;
  }
  final_depth /= w_sum;
  out[x + y * w] = final_depth;
}

/* One DepthFiltering pass with the CUDA build's border rule: pixels with x<2, x>=w-2, y<2, y>=h-2 of the
 * TARGET are not written (ITMViewBuilder_CUDA.cu:196-209; the CPU twin clears the target first). */
void oracle_depth_filtering(float *out, const float *in, int w, int h) {
  for (int y = 2; y < h - 2; y++) for (int x = 2; x < w - 2; x++) filter_depth(out, in, x, y, w);
}

/* ITMViewBuilder_CUDA.cu:211-227 + DA/ITMViewBuilder.h:59-114, in-image threads only (the reference's
 * out-of-image threads alias the next row through idx = x + y*w and race; not reproduced). */
void oracle_compute_normal_and_weights(b200_vec4f *normal_out, float *sigmaZ_out, const float *depth_in, int w, int h,
                                       const float intr[4]) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int idx = x + y * w;
    if (x < 2 || x > w - 2 || y < 2 || y > h - 2) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; continue; }
    float z = depth_in[idx];
    if (z < 0.0f) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; continue; }
    float xp1_z = depth_in[(x + 1) + y * w], yp1_z = depth_in[x + (y + 1) * w];
    float xm1_z = depth_in[(x - 1) + y * w], ym1_z = depth_in[x + (y - 1) * w];
    if (xp1_z <= 0 || yp1_z <= 0 || xm1_z <= 0 || ym1_z <= 0) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; continue; }
    float xp1_x = xp1_z * ((x + 1.0f) - intr[2]) * intr[0], xp1_y = xp1_z * (y - intr[3]) * intr[1];
    float xm1_x = xm1_z * ((x - 1.0f) - intr[2]) * intr[0], xm1_y = xm1_z * (y - intr[3]) * intr[1];
    float yp1_x = yp1_z * (x - intr[2]) * intr[0], yp1_y = yp1_z * ((y + 1.0f) - intr[3]) * intr[1];
    float ym1_x = ym1_z * (x - intr[2]) * intr[0], ym1_y = ym1_z * ((y - 1.0f) - intr[3]) * intr[1];
    float dxx = xp1_x - xm1_x, dxy = xp1_y - xm1_y, dxz = xp1_z - xm1_z;
    float dyx = yp1_x - ym1_x, dyy = yp1_y - ym1_y, dyz = yp1_z - ym1_z;
    float nx = (dxy * dyz - dxz * dyy);
    float ny = (dxz * dyx - dxx * dyz);
    float nz = (dxx * dyy - dxy * dyx);
    if (nx == 0.0f && ny == 0 && nz == 0) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; continue; }
    float norm = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= norm; ny *= norm; nz *= norm;
    normal_out[idx].x = nx; normal_out[idx].y = ny; normal_out[idx].z = nz; normal_out[idx].w = 1.0f;
    float theta = acosf(nz);
    float theta_diff = theta / (ITM_PI * 0.5f - theta);
    sigmaZ_out[idx] = (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * theta_diff * theta_diff);
  }
}

/* ITMViewBuilder_CUDA::UpdateView (:33-84) after its two H2D copies. floatImage is the builder's persistent
 * scratch image: zero-initialised when created (ORUtils/MemoryBlock.h:109-122) and owned by the caller here. */
void oracle_update_view(const int16_t *raw, int w, int h, const b200_view_calib *c, float *depth, float *floatImage,
                        b200_vec4f *depthNormal, float *depthUncertainty) {
  if (c->trafoType == 0) oracle_convert_disparity_to_depth(depth, raw, w, h, c->params[0], c->params[1], c->fx_depth);
  else oracle_convert_depth_affine_to_float(depth, raw, w, h, c->params[0], c->params[1]);
  if (c->useBilateralFilter) {
    oracle_depth_filtering(floatImage, depth, w, h);
    oracle_depth_filtering(depth, floatImage, w, h);
    oracle_depth_filtering(floatImage, depth, w, h);
    oracle_depth_filtering(depth, floatImage, w, h);
    oracle_depth_filtering(floatImage, depth, w, h);
    memcpy(depth, floatImage, (size_t)w * h * sizeof(float));
  }
  if (c->modelSensorNoise) oracle_compute_normal_and_weights(depthNormal, depthUncertainty, depth, w, h, c->intrinsics_d);
}
