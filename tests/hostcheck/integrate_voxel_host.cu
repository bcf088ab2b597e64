// Test infrastructure (CPU): compiles the kernel's own per-voxel functions (dynslam_b200/csrc/integrate_voxel.cuh, which are
// __host__ __device__) for the HOST and checks, voxel by voxel, that the default kernel's fast path
//   pose products -> v3_stage_a -> depth fetch -> v3_stage_b -> { done | v3_colour | generic path }
// leaves exactly the bits that the generic per-voxel code (integrate_voxel: the expressions of
// DA/ITMSceneReconstructionEngine.h:14-171 with the `/` operator) leaves. No GPU is involved; MUFU.RCP is replaced by the
// IEEE reciprocal (see rcp_nr), the Newton step and the quotient sequence are the device's.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../dynslam_b200/csrc/integrate_voxel.cuh"

namespace {
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 1) {}
  uint32_t u32() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
  float uni(float a, float b) { return a + (b - a) * (float)(u32() & 0xffffff) / 16777216.0f; }
};

void rot_xyz(float rx, float ry, float rz, float R[9]) {   // row-major
  const float cx = cosf(rx), sx = sinf(rx), cy = cosf(ry), sy = sinf(ry), cz = cosf(rz), sz = sinf(rz);
  const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float T[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) T[i * 3 + j] += Ry[i * 3 + k] * Rx[k * 3 + j]; }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { R[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) R[i * 3 + j] += Rz[i * 3 + k] * T[k * 3 + j]; }
}
}   // namespace

// stats: [0] voxels, [1] fast path finished the voxel, [2] colour pass ran, [3] generic path taken, [4] voxels changed,
//        [5] behind the camera
extern "C" long long hostcheck_integrate(long long blocks, unsigned seed, float mu, float voxelSize, int depthWeighting, int maxW,
                                         int identityPose, long long *stats) {
  Rng rng(seed);
  const int w = 311, h = 94;
  FrameGeom g;
  memset(&g, 0, sizeof(g));
  // pose: world -> camera, M.m[col * 4 + row]
  float R[9];
  if (identityPose) rot_xyz(0, 0, 0, R);
  else rot_xyz(rng.uni(-0.2f, 0.2f), rng.uni(-3.1f, 3.1f), rng.uni(-0.1f, 0.1f), R);
  float C[3] = {rng.uni(-50, 50), rng.uni(-2, 2), rng.uni(-50, 50)};
  if (identityPose) C[0] = C[1] = C[2] = 0.0f;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) g.M_d.m[c * 4 + r] = R[r * 3 + c];
  for (int r = 0; r < 3; ++r) g.M_d.m[12 + r] = -(R[r * 3] * C[0] + R[r * 3 + 1] * C[1] + R[r * 3 + 2] * C[2]);
  g.M_d.m[15] = 1.0f;
  g.M_rgb = g.M_d;
  const float proj[4] = {707.0912f * 0.25f, 707.0912f * 0.25f, 609.74f * 0.25f, 185.58f * 0.25f};
  for (int i = 0; i < 4; ++i) g.proj_d[i] = g.proj_rgb[i] = proj[i];
  g.w = g.rgb_w = w; g.h = g.rgb_h = h;
  g.voxelSize = voxelSize; g.mu = mu; g.maxW = maxW; g.depthWeighting = depthWeighting; g.sameRgbCam = 1;
  { volatile float a = -1.0f, b = mu; g.negOneOverMu = a / b; }
  std::vector<float> depth((size_t)w * h);
  std::vector<b200_vec4u> rgb((size_t)w * h);
  for (size_t i = 0; i < depth.size(); ++i) {
    const uint32_t t = rng.u32() % 100;
    depth[i] = t < 6 ? 0.0f : (t < 8 ? -1.0f : floorf(rng.uni(500.0f, 20000.0f)) * 0.001f);
    const uint32_t c = rng.u32();
    rgb[i].x = (uint8_t)c; rgb[i].y = (uint8_t)(c >> 8); rgb[i].z = (uint8_t)(c >> 16); rgb[i].w = 255;
  }
  float div255[256], rcpW[272];
  for (int i = 0; i < 256; ++i) { volatile float a = (float)i, b = 255.0f; div255[i] = a / b; }
  for (int i = 0; i < 272; ++i) rcpW[i] = rcp_nr((float)i);
  V3K k;
  k.rcpMu = rcp_nr(g.mu); k.rcp255 = rcp_nr(255.0f); k.wm2 = (float)(g.w - 2); k.hm2 = (float)(g.h - 2);
  k.rejectColour = (!(fabsf(g.negOneOverMu) > 0.25f)) ? 2 : 0;
  const float m12x = g.M_d.m[12], m12y = g.M_d.m[13], m12z = g.M_d.m[14];
  const unsigned *rgbw = reinterpret_cast<const unsigned *>(rgb.data());
  long long bad = 0;
  for (long long b = 0; b < blocks; ++b) {
    // a block somewhere around the view frustum (including behind and beside the camera), in camera space, then to world
    const float pc[3] = {rng.uni(-12, 12), rng.uni(-4, 4), rng.uni(-3, 24)};
    int org[3];
    for (int r = 0; r < 3; ++r) {
      const float wv = C[r] + R[0 * 3 + r] * pc[0] + R[1 * 3 + r] * pc[1] + R[2 * 3 + r] * pc[2];
      org[r] = (int)floorf(wv / (8.0f * voxelSize)) * 8;
    }
    if (identityPose && (b % 4) == 0) org[b % 3] = 0;   // planes through the origin: exact zeros in the camera coordinates
    float4 prod[3][8];
    for (int axis = 0; axis < 3; ++axis) for (int i = 0; i < 8; ++i) {
      const float c = (float)(org[axis] + i) * g.voxelSize;
      prod[axis][i] = make_float4(g.M_d.m[axis * 4 + 0] * c, g.M_d.m[axis * 4 + 1] * c, g.M_d.m[axis * 4 + 2] * c, 0.0f);
    }
    for (int locId = 0; locId < 512; ++locId) {
      const int x = locId & 7, y = (locId >> 3) & 7, z = locId >> 6;
      // voxel content: fresh, typical and saturated
      const uint32_t t = rng.u32();
      unsigned lo, hi;
      if ((t & 7) == 0) { lo = 32767u; hi = 0; }
      else {
        const int sdf = (int)(rng.u32() % 65535) - 32767, wd = (t >> 3) % (maxW < 255 ? maxW + 1 : 256), wc = (t >> 12) % ((maxW & 0xff) + 1);
        const uint32_t c = rng.u32();
        lo = ((unsigned)sdf & 0xffffu) | ((unsigned)wd << 16) | ((c & 0xffu) << 24);
        hi = ((c >> 8) & 0xffffu) | ((unsigned)wc << 16);
      }
      unsigned glo = lo, ghi = hi;
      integrate_voxel(glo, ghi, locId, org[0], org[1], org[2], g, depth.data(), rgb.data(), div255);
      unsigned flo = lo, fhi = hi;
      const V3A a = v3_stage_a(prod[0][x], prod[1][y], prod[2][z], m12x, m12y, m12z, g, k);
      const float dm = depth[a.idx];
      int r = depthWeighting ? v3_stage_b<true>(flo, a, dm, g, k, rcpW) : v3_stage_b<false>(flo, a, dm, g, k, rcpW);
      stats[0]++;
      if (a.pcz <= 0.0f) stats[5]++;
      if (r == 2) { flo = lo; fhi = hi; integrate_voxel(flo, fhi, locId, org[0], org[1], org[2], g, depth.data(), rgb.data(), div255); stats[3]++; }
      else {
        stats[1]++;
        if (r == 1) { v3_colour(flo, fhi, a.ix, a.iy, g, k, rgbw, div255, rcpW); stats[2]++; }
      }
      if (flo != lo || fhi != hi) stats[4]++;
      if (flo != glo || fhi != ghi) bad++;
    }
  }
  return bad;
}


// The V4 pair path (v4_stage_a / v4_stage_b on two x-adjacent voxels, packed arithmetic evaluated element-wise on the host)
// against the generic per-voxel code. fast = 0: must be bit-identical (returns the number of differing voxels).
// fast = 1 (tolerance mode): returns the number of voxels whose weight, colour or colour weight differ or whose 16-bit TSDF
// code differs by more than 1; stats[6] counts the voxels whose TSDF code differs by exactly 1, stats[7] those where the
// colour gate flipped (colour fields differ while the weight agrees).
extern "C" long long hostcheck_integrate_v4(long long blocks, unsigned seed, float mu, float voxelSize, int depthWeighting, int maxW,
                                            int identityPose, int fast, long long *stats) {
  Rng rng(seed);
  const int w = 311, h = 94;
  FrameGeom g;
  memset(&g, 0, sizeof(g));
  float R[9];
  if (identityPose) rot_xyz(0, 0, 0, R);
  else rot_xyz(rng.uni(-0.2f, 0.2f), rng.uni(-3.1f, 3.1f), rng.uni(-0.1f, 0.1f), R);
  float C[3] = {rng.uni(-50, 50), rng.uni(-2, 2), rng.uni(-50, 50)};
  if (identityPose) C[0] = C[1] = C[2] = 0.0f;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) g.M_d.m[c * 4 + r] = R[r * 3 + c];
  for (int r = 0; r < 3; ++r) g.M_d.m[12 + r] = -(R[r * 3] * C[0] + R[r * 3 + 1] * C[1] + R[r * 3 + 2] * C[2]);
  g.M_d.m[15] = 1.0f;
  g.M_rgb = g.M_d;
  const float proj[4] = {707.0912f * 0.25f, 707.0912f * 0.25f, 609.74f * 0.25f, 185.58f * 0.25f};
  for (int i = 0; i < 4; ++i) g.proj_d[i] = g.proj_rgb[i] = proj[i];
  g.w = g.rgb_w = w; g.h = g.rgb_h = h;
  g.voxelSize = voxelSize; g.mu = mu; g.maxW = maxW; g.depthWeighting = depthWeighting; g.sameRgbCam = 1;
  { volatile float a = -1.0f, b = mu; g.negOneOverMu = a / b; }
  std::vector<float> depth((size_t)w * h);
  std::vector<b200_vec4u> rgb((size_t)w * h);
  for (size_t i = 0; i < depth.size(); ++i) {
    const uint32_t t = rng.u32() % 100;
    depth[i] = t < 6 ? 0.0f : (t < 8 ? -1.0f : floorf(rng.uni(500.0f, 20000.0f)) * 0.001f);
    const uint32_t c = rng.u32();
    rgb[i].x = (uint8_t)c; rgb[i].y = (uint8_t)(c >> 8); rgb[i].z = (uint8_t)(c >> 16); rgb[i].w = 255;
  }
  float div255[256], rcpW[272];
  for (int i = 0; i < 256; ++i) { volatile float a = (float)i, b = 255.0f; div255[i] = a / b; }
  for (int i = 0; i < 272; ++i) rcpW[i] = rcp_nr((float)i);
  V3K k;
  k.rcpMu = rcp_nr(g.mu); k.rcp255 = rcp_nr(255.0f); k.wm2 = (float)(g.w - 2); k.hm2 = (float)(g.h - 2);
  k.rejectColour = (!(fabsf(g.negOneOverMu) > 0.25f)) ? 2 : 0;
  const V4C cc = v4_constants(g, k);
  const unsigned *rgbw = reinterpret_cast<const unsigned *>(rgb.data());
  long long bad = 0;
  for (long long b = 0; b < blocks; ++b) {
    const float pc[3] = {rng.uni(-12, 12), rng.uni(-4, 4), rng.uni(-3, 24)};
    int org[3];
    for (int r = 0; r < 3; ++r) {
      const float wv = C[r] + R[0 * 3 + r] * pc[0] + R[1 * 3 + r] * pc[1] + R[2 * 3 + r] * pc[2];
      org[r] = (int)floorf(wv / (8.0f * voxelSize)) * 8;
    }
    if (identityPose && (b % 4) == 0) org[b % 3] = 0;
    float prod[3][8][3];
    for (int axis = 0; axis < 3; ++axis) for (int i = 0; i < 8; ++i) {
      const float c = (float)(org[axis] + i) * g.voxelSize;
      for (int comp = 0; comp < 3; ++comp) prod[axis][i][comp] = g.M_d.m[axis * 4 + comp] * c;
    }
    for (int z = 0; z < 8; ++z) for (int y = 0; y < 8; ++y) for (int xp = 0; xp < 4; ++xp) {
      unsigned lo[2], hi[2], glo[2], ghi[2];
      for (int e = 0; e < 2; ++e) {
        const uint32_t t = rng.u32();
        if ((t & 7) == 0) { lo[e] = 32767u; hi[e] = 0; }
        else {
          const int sdf = (int)(rng.u32() % 65535) - 32767, wd = (t >> 3) % (maxW < 255 ? maxW + 1 : 256), wc = (t >> 12) % ((maxW & 0xff) + 1);
          const uint32_t c = rng.u32();
          lo[e] = ((unsigned)sdf & 0xffffu) | ((unsigned)wd << 16) | ((c & 0xffu) << 24);
          hi[e] = ((c >> 8) & 0xffffu) | ((unsigned)wc << 16);
        }
        glo[e] = lo[e]; ghi[e] = hi[e];
        integrate_voxel(glo[e], ghi[e], 2 * xp + e + 8 * y + 64 * z, org[0], org[1], org[2], g, depth.data(), rgb.data(), div255);
      }
      float2 XY[3], Zd[3];
      for (int comp = 0; comp < 3; ++comp) {
        XY[comp] = f2add(make_float2(prod[0][2 * xp][comp], prod[0][2 * xp + 1][comp]), f2dup(prod[1][y][comp]));
        Zd[comp] = f2dup(prod[2][z][comp]);
      }
      V4A a;
      unsigned flo[2] = {lo[0], lo[1]}, fhi[2] = {hi[0], hi[1]};
      int r[2];
      if (fast) {
        a = v4_stage_a<true>(XY[0], XY[1], XY[2], Zd[0], Zd[1], Zd[2], g, k, cc);
        const float2 dm = make_float2(depth[a.idx0], depth[a.idx1]);
        if (depthWeighting) v4_stage_b<true, true>(flo[0], flo[1], a, dm, g, k, cc, rcpW, r[0], r[1]);
        else v4_stage_b<false, true>(flo[0], flo[1], a, dm, g, k, cc, rcpW, r[0], r[1]);
      } else {
        a = v4_stage_a<false>(XY[0], XY[1], XY[2], Zd[0], Zd[1], Zd[2], g, k, cc);
        const float2 dm = make_float2(depth[a.idx0], depth[a.idx1]);
        if (depthWeighting) v4_stage_b<true, false>(flo[0], flo[1], a, dm, g, k, cc, rcpW, r[0], r[1]);
        else v4_stage_b<false, false>(flo[0], flo[1], a, dm, g, k, cc, rcpW, r[0], r[1]);
      }
      for (int e = 0; e < 2; ++e) {
        stats[0]++;
        if (r[e] == 2) { flo[e] = lo[e]; fhi[e] = hi[e]; integrate_voxel(flo[e], fhi[e], 2 * xp + e + 8 * y + 64 * z, org[0], org[1], org[2], g, depth.data(), rgb.data(), div255); stats[3]++; }
        else {
          stats[1]++;
          if (r[e] == 1) { v3_colour(flo[e], fhi[e], e ? a.ix.y : a.ix.x, e ? a.iy.y : a.iy.x, g, k, rgbw, div255, rcpW); stats[2]++; }
        }
        if (flo[e] != lo[e] || fhi[e] != hi[e]) stats[4]++;
        if (!fast) { if (flo[e] != glo[e] || fhi[e] != ghi[e]) bad++; }
        else {
          const int ds = (int)(short)(flo[e] & 0xffff) - (int)(short)(glo[e] & 0xffff);
          const bool wsame = ((flo[e] >> 16) & 0xff) == ((glo[e] >> 16) & 0xff);
          const bool csame = (flo[e] >> 24) == (glo[e] >> 24) && fhi[e] == ghi[e];
          if (ds == 1 || ds == -1) stats[6]++;
          if (wsame && !csame) stats[7]++;
          if (!wsame || ds > 1 || ds < -1) bad++;
        }
      }
    }
  }
  return bad;
}
