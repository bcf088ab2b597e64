// Test infrastructure (CPU): the raycast kernel's own castRay chain (dynslam_b200/csrc/raycast_ray.cuh, __host__ __device__)
// compiled for the HOST and run over a whole image on host arrays, with k_raycast's pixel -> min/max-cell mapping
// (vis.cu, GenericRaycast, Vis_CUDA.cu:672-684). tests/test_raycast_host.py compares every ray with the CPU oracle's.
#define RC_COUNT_WALKS
static long rc_walks = 0;      // chain walks of cast_ray_nbr (design statistics only)
#include "../../dynslam_b200/csrc/raycast_ray.cuh"

extern "C" long hostcheck_raycast_walks(int reset) { const long v = rc_walks; if (reset) rc_walks = 0; return v; }

extern "C" void hostcheck_raycast(const b200_voxel *voxels, const b200_hash_entry *table, int nb, int w, int h, const float *invM16,
                                  const float *proj, float voxelSize, float mu, const b200_vec2f *minmax, b200_vec4f *out, int variant) {
  Mat4 invM;
  for (int i = 0; i < 16; ++i) invM.m[i] = invM16[i];
  const float fx = proj[0], fy = proj[1], cxp = proj[2], cyp = proj[3];
  const float2 *mm = reinterpret_cast<const float2 *>(minmax);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int locId2 = (int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w;
      float4 o;
      if (variant == 1) { int nbr[8]; cast_ray_nbr(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, mm[locId2], nbr, 1); }
      else cast_ray(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, mm[locId2]);
      out[x + y * w].x = o.x; out[x + y * w].y = o.y; out[x + y * w].z = o.z; out[x + y * w].w = o.w;
    }
}

// walk statistics of the same march (scripts/raycast_stats.py): per ray, the number of steps through missing blocks and the
// number of interpolated samples. Mirrors cast_ray's loop; used for design decisions only, never for parity.
extern "C" void hostcheck_raycast_stats(const b200_voxel *voxels, const b200_hash_entry *table, int nb, int w, int h, const float *invM16,
                                        const float *proj, float voxelSize, float mu, const b200_vec2f *minmax, int *emptySteps, int *foundSteps) {
  Mat4 invM;
  for (int i = 0; i < 16; ++i) invM.m[i] = invM16[i];
  const float invfx = 1.0f / proj[0], invfy = 1.0f / proj[1], cxp = proj[2], cyp = proj[3], oneOverVoxelSize = 1.0f / voxelSize;
  const float2 *mmI = reinterpret_cast<const float2 *>(minmax);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float2 mm = mmI[(int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w];
      const float stepScale = mu * oneOverVoxelSize;
      float cz = mm.x;
      float cx = cz * (((float)x - cxp) * invfx), cy = cz * (((float)y - cyp) * invfy);
      float totalLength = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxelSize;
      Vec4 r = m4v4(invM, cx, cy, cz, 1.0f);
      const float sx = r.x * oneOverVoxelSize, sy = r.y * oneOverVoxelSize, sz = r.z * oneOverVoxelSize;
      cz = mm.y;
      cx = cz * (((float)x - cxp) * invfx); cy = cz * (((float)y - cyp) * invfy);
      const float totalLengthMax = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxelSize;
      r = m4v4(invM, cx, cy, cz, 1.0f);
      float dx = r.x * oneOverVoxelSize - sx, dy = r.y * oneOverVoxelSize - sy, dz = r.z * oneOverVoxelSize - sz;
      const float dn = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
      dx *= dn; dy *= dn; dz *= dn;
      float px = sx, py = sy, pz = sz;
      IdxCache cache; cache_init(cache);
      int ne = 0, nf = 0;
      while (totalLength < totalLengthMax) {
        const int base = block_base(table, nb, ((int)round_(px)) >> 3, ((int)round_(py)) >> 3, ((int)round_(pz)) >> 3, cache);
        float stepLength;
        if (base < 0) { stepLength = BS; ne++; }
        else {
          nf++;
          const float s = sdf_interp(voxels, table, nb, px, py, pz, cache);
          if (s <= 0.0f) break;
          stepLength = maxf_(s * stepScale, 1.0f);
        }
        px += stepLength * dx; py += stepLength * dy; pz += stepLength * dz;
        totalLength += stepLength;
      }
      emptySteps[x + y * w] = ne; foundSteps[x + y * w] = nf;
    }
}
