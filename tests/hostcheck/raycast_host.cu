// Test infrastructure (CPU): the raycast kernel's own castRay chain (dynslam_b200/csrc/raycast_ray.cuh, __host__ __device__)
// compiled for the HOST and run over a whole image on host arrays, with k_raycast's pixel -> min/max-cell mapping
// (vis.cu, GenericRaycast, Vis_CUDA.cu:672-684). tests/test_raycast_host.py compares every ray with the CPU oracle's.
#include "../../dynslam_b200/csrc/raycast_ray.cuh"

extern "C" void hostcheck_raycast(const b200_voxel *voxels, const b200_hash_entry *table, int nb, int w, int h, const float *invM16,
                                  const float *proj, float voxelSize, float mu, const b200_vec2f *minmax, b200_vec4f *out) {
  Mat4 invM;
  for (int i = 0; i < 16; ++i) invM.m[i] = invM16[i];
  const float fx = proj[0], fy = proj[1], cxp = proj[2], cyp = proj[3];
  const float2 *mm = reinterpret_cast<const float2 *>(minmax);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int locId2 = (int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w;
      float4 o;
      cast_ray(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, mm[locId2]);
      out[x + y * w].x = o.x; out[x + y * w].y = o.y; out[x + y * w].z = o.z; out[x + y * w].w = o.w;
    }
}
