// integrate_voxel.cuh — the per-voxel arithmetic of IntegrateIntoScene, shared by the kernels of integrate.cu.
//
// Everything here is `__host__ __device__`: the kernels use it on the device, and tests/hostcheck compiles the very same
// functions for the host to check, over millions of random voxels, that the fast two-stage path (v3_stage_a / v3_stage_b /
// v3_colour, plus the hand-off rules to the generic path) produces the same bits as the generic per-voxel code — a CPU test
// of the kernel's logic. (On the host rcp_nr starts from the IEEE reciprocal instead of MUFU.RCP; the division sequence
// itself is checked against `/` on the GPU by b200_selftest_divide.)
#pragma once
#include "engine.h"

#ifdef __CUDA_ARCH__
#define IV_LDG(p) __ldg(p)
#define IV_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define IV_MUL(a, b) __fmul_rn((a), (b))
#else
#include <math.h>
#define IV_LDG(p) (*(p))
#define IV_FMA(a, b, c) fmaf((a), (b), (c))
#define IV_MUL(a, b) ((a) * (b))
#endif

// ---- per-voxel update (bit-exact with the oracle) ---------------------------------------------
struct VoxelU { int sdf; int w_depth; int c0, c1, c2; int w_color; int pad; };

HD VoxelU unpack(unsigned lo, unsigned hi) {
  VoxelU v;
  v.sdf = (short)(lo & 0xffffu); v.w_depth = (lo >> 16) & 0xff; v.c0 = (lo >> 24) & 0xff;
  v.c1 = hi & 0xff; v.c2 = (hi >> 8) & 0xff; v.w_color = (hi >> 16) & 0xff; v.pad = (hi >> 24) & 0xff;
  return v;
}
HD void pack(const VoxelU &v, unsigned &lo, unsigned &hi) {
  lo = ((unsigned)v.sdf & 0xffffu) | ((unsigned)(v.w_depth & 0xff) << 16) | ((unsigned)(v.c0 & 0xff) << 24);
  hi = (unsigned)(v.c1 & 0xff) | ((unsigned)(v.c2 & 0xff) << 8) | ((unsigned)(v.w_color & 0xff) << 16) | ((unsigned)(v.pad & 0xff) << 24);
}

HD float bil(float a, float b, float c, float d, float dx, float dy) {
  return (a * (1.0f - dx) * (1.0f - dy) + b * dx * (1.0f - dy) + c * (1.0f - dx) * dy + d * dx * dy);
}

HD int to_uchar_round(float x) { return clampi_((int)round_(x), 0, 255); }

// ComputeUpdatedVoxelInfo<true,TVoxel>::compute — DA/ITMSceneReconstructionEngine.h:147-171, split in two so that
// the (expensive, minority) colour updates of a block can be compacted onto full warps.
// depth_part: computeUpdatedVoxelDepthInfo :14-88 + the colour gate :163-166. Returns true when the colour update
// must run; ix, iy hold the depth-camera projection when `projected`.
HD bool depth_part(VoxelU &v, float ptx, float pty, float ptz, const FrameGeom &g, const float *__restrict__ depth, float &ix,
                    float &iy, bool &projected) {
  float eta = -1.0f;
  float etaOverMu = 0.0f; bool haveQuot = false;   // eta / mu is needed twice (:63, :165); divide once
  ix = 0; iy = 0; projected = false;
  {
    Vec4 pc = m4v4(g.M_d, ptx, pty, ptz, 1.0f);
    bool done = false;
    if (pc.z <= 0) done = true;
    if (!done) {
      projected = true;
      ix = g.proj_d[0] * pc.x / pc.z + g.proj_d[2];
      iy = g.proj_d[1] * pc.y / pc.z + g.proj_d[3];
      if ((ix < 1) || (ix > g.w - 2) || (iy < 1) || (iy > g.h - 2)) done = true;
    }
    if (!done) {
      float dm = IV_LDG(depth + (int)(ix + 0.5f) + (int)(iy + 0.5f) * g.w);
      if (dm <= 0.0) done = true;
      else {
        eta = dm - pc.z;
        if (!(eta < -g.mu)) {
          float oldF = (float)(v.sdf) / 32767.0f;
          int oldW = v.w_depth;
          etaOverMu = eta / g.mu; haveQuot = true;
          float newF = minf_(1.0f, etaOverMu);
          int newW;
          if (g.depthWeighting) {
            newW = (int)(100.0 / dm);
            if (newW < 1) newW = 1;
            if (newW > 10) newW = 10;
          } else newW = 1;
          newF = oldW * oldF + newW * newF;
          newW = oldW + newW;
          newF /= newW;
          newW = mini_(newW, g.maxW);
          v.sdf = (short)((newF) * 32767.0f);
          v.w_depth = newW & 0xff;
        }
      }
    }
  }
  if (eta > g.mu) return false;
  if (!haveQuot) {
    // no update happened: either the voxel was rejected (eta == -1 exactly) or it lies more than mu behind the
    // surface (eta < -mu, so |eta / mu| >= 1 > 0.25 for any mu > 0). Both quotients are known without dividing.
    if (eta == -1.0f) etaOverMu = g.negOneOverMu;
    else if (g.mu > 0.0f && eta < -g.mu) return false;
    else etaOverMu = eta / g.mu;
  }
  return !(fabsf(etaOverMu) > 0.25f);
}

// colour_part: computeUpdatedVoxelColorInfo :91-128
HD void colour_part(VoxelU &v, float ptx, float pty, float ptz, float ix, float iy, bool projected, const FrameGeom &g,
                     const b200_vec4u *__restrict__ rgb, const float *__restrict__ div255) {
  const float oldW = (float)v.w_color;
  const float o0 = div255[v.c0], o1 = div255[v.c1], o2 = div255[v.c2];
  if (!(g.sameRgbCam && projected)) {   // same camera: the expressions below are the ones already evaluated
    Vec4 pc = m4v4(g.M_rgb, ptx, pty, ptz, 1.0f);
    ix = g.proj_rgb[0] * pc.x / pc.z + g.proj_rgb[2];
    iy = g.proj_rgb[1] * pc.y / pc.z + g.proj_rgb[3];
  }
  if ((ix < 1) || (ix > g.rgb_w - 2) || (iy < 1) || (iy > g.rgb_h - 2)) return;
  const int px = (int)floorf(ix), py = (int)floorf(iy);
  const float dx = ix - (float)px, dy = iy - (float)py;
  const unsigned *rgbw = reinterpret_cast<const unsigned *>(rgb);
  unsigned a = IV_LDG(rgbw + px + py * g.rgb_w), b = 0, c = 0, d = 0;
  if (dx != 0) b = IV_LDG(rgbw + (px + 1) + py * g.rgb_w);
  if (dy != 0) c = IV_LDG(rgbw + px + (py + 1) * g.rgb_w);
  if (dx != 0 && dy != 0) d = IV_LDG(rgbw + (px + 1) + (py + 1) * g.rgb_w);
  float m0 = bil((float)(a & 0xff), (float)(b & 0xff), (float)(c & 0xff), (float)(d & 0xff), dx, dy) / 255.0f;
  float m1 = bil((float)((a >> 8) & 0xff), (float)((b >> 8) & 0xff), (float)((c >> 8) & 0xff), (float)((d >> 8) & 0xff), dx, dy) / 255.0f;
  float m2 = bil((float)((a >> 16) & 0xff), (float)((b >> 16) & 0xff), (float)((c >> 16) & 0xff), (float)((d >> 16) & 0xff), dx, dy) / 255.0f;
  float newW = 5;
  float n0 = o0 * oldW + m0 * newW, n1 = o1 * oldW + m1 * newW, n2 = o2 * oldW + m2 * newW;
  newW = oldW + newW;
  n0 /= newW; n1 /= newW; n2 /= newW;
  const int maxWc = g.maxW & 0xff;   // maxW is passed as uchar (DA/...:93)
  newW = (newW < maxWc) ? newW : (float)maxWc;
  v.c0 = to_uchar_round(n0 * 255.0f); v.c1 = to_uchar_round(n1 * 255.0f); v.c2 = to_uchar_round(n2 * 255.0f);
  v.w_color = ((int)newW) & 0xff;
}

HD void update_voxel(VoxelU &v, float ptx, float pty, float ptz, const FrameGeom &g, const float *__restrict__ depth,
                      const b200_vec4u *__restrict__ rgb, const float *__restrict__ div255) {
  float ix, iy; bool projected;
  if (depth_part(v, ptx, pty, ptz, g, depth, ix, iy, projected)) colour_part(v, ptx, pty, ptz, ix, iy, projected, g, rgb, div255);
}

// processes voxel locId of the block at block coordinates (bx,by,bz); returns true if changed
HD bool integrate_voxel(unsigned &lo, unsigned &hi, int locId, int gx, int gy, int gz, const FrameGeom &g,
                         const float *__restrict__ depth, const b200_vec4u *__restrict__ rgb, const float *__restrict__ div255) {
  VoxelU v = unpack(lo, hi);
  if (g.stopMaxW) if (v.w_depth == g.maxW) return false;
  if (g.approx) if (v.w_depth != 0) return false;
  const int x = locId & 7, y = (locId >> 3) & 7, z = locId >> 6;
  update_voxel(v, (float)(gx + x) * g.voxelSize, (float)(gy + y) * g.voxelSize, (float)(gz + z) * g.voxelSize, g, depth, rgb, div255);
  unsigned nlo, nhi;
  pack(v, nlo, nhi);
  const bool changed = (nlo != lo) || (nhi != hi);
  lo = nlo; hi = nhi;
  return changed;
}


#define V3_SAFE_LO 9.094947017729282e-13f   // 2^-40
#define V3_SAFE_HI 1.099511627776e12f       // 2^40
#define V3_RCP_32767 3.0518509447574615e-05f   // RN(1 / 32767) = 0x1.0002p-15

HD float rcp_nr(float b) {
  float y0;
#ifdef __CUDA_ARCH__
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(b));
#else
  y0 = 1.0f / b;   // host stand-in for MUFU.RCP (tests/hostcheck)
#endif
  const float e = IV_FMA(-b, y0, 1.0f);
  return IV_FMA(y0, e, y0);
}
HD float div_nr(float a, float b, float y) {   // y = rcp_nr(b)
  const float q = IV_MUL(a, y);
  const float r = IV_FMA(-b, q, a);
  return IV_FMA(y, r, q);
}


struct V3K { float rcpMu, rcp255, wm2, hm2; int rejectColour; };


// computeUpdatedVoxelDepthInfo + colour gate, written without branches in two stages so that the two voxels of a lane
// interleave (both depth fetches are in flight together) and no lane pays for another lane's early exit:
//   stage A: camera point, projection, bounds, depth-image index (0 when the voxel will not sample the image);
//   stage B: the update, computed unconditionally and committed by a select.
// Result of stage B: 0 = done, 1 = colour update needed (a.ix, a.iy valid), 2 = take the generic path (voxel untouched).
struct V3A { float pcz, ix, iy; unsigned idx; bool ok, inb; };

HD V3A v3_stage_a(const float4 X, const float4 Y, const float4 Z, float m12x, float m12y, float m12z, const FrameGeom &g,
                   const V3K &k) {
  V3A a;
  const float pcx = X.x + Y.x + Z.x + m12x, pcy = X.y + Y.y + Z.y + m12y;
  a.pcz = X.z + Y.z + Z.z + m12z;
  const float ax = g.proj_d[0] * pcx, ay = g.proj_d[1] * pcy;
  const float aax = fabsf(ax), aay = fabsf(ay);
  const float lo3 = fminf(fminf(aax, aay), a.pcz), hi3 = fmaxf(fmaxf(aax, aay), a.pcz);
  a.ok = (lo3 >= V3_SAFE_LO) && (hi3 <= V3_SAFE_HI);     // false for pc.z <= 0, zero numerators, NaN z: generic path
  const float yz = rcp_nr(a.pcz);
  a.ix = div_nr(ax, a.pcz, yz) + g.proj_d[2];
  a.iy = div_nr(ay, a.pcz, yz) + g.proj_d[3];
  a.inb = !((a.ix < 1) | (a.ix > k.wm2) | (a.iy < 1) | (a.iy > k.hm2));   // (non-short-circuit: no branches)
  const int px = (int)(a.ix + 0.5f), py = (int)(a.iy + 0.5f);
  a.idx = (unsigned)(px + py * g.w) & (0u - (unsigned)(a.ok & a.inb));    // 0 when the voxel will not sample the image
  return a;
}

template <bool DW>
HD int v3_stage_b(unsigned &lo, const V3A &a, float dm, const FrameGeom &g, const V3K &k, const float *rcpW) {
  const bool rej = (dm <= 0.0f);
  const bool valid = a.ok && a.inb && !rej;
  const float eta = dm - a.pcz;
  const bool behind = eta < -g.mu;
  const float ae = fabsf(eta);
  const bool etaOK = ((ae >= V3_SAFE_LO) && (ae <= V3_SAFE_HI)) || (eta == 0.0f);
  const float eom = div_nr(eta, g.mu, k.rcpMu);
  float newF = minf_(1.0f, eom);
  const int oldW = (lo >> 16) & 0xff;
  const float oldF = div_nr((float)(short)(lo & 0xffffu), 32767.0f, V3_RCP_32767);
  int newW;
  if (DW) {
    newW = (int)(100.0 / dm);
    if (newW < 1) newW = 1;
    if (newW > 10) newW = 10;
  } else newW = 1;
  newF = oldW * oldF + newW * newF;
  newW = oldW + newW;
  newF = div_nr(newF, (float)newW, rcpW[newW]);
  newW = mini_(newW, g.maxW);
  const int ns = (short)((newF) * 32767.0f);
  const unsigned nlo = (lo & 0xff000000u) | ((unsigned)ns & 0xffffu) | ((unsigned)(newW & 0xff) << 16);
  const bool upd = valid && !behind && etaOK;
  lo = upd ? nlo : lo;
  // behind the camera (pc.z <= 0): computeUpdatedVoxelDepthInfo returns -1 before touching the voxel and, for mu < 4, the
  // colour gate rejects it too — nothing to do, no need for the generic path (blocks around the camera are full of these)
  const bool behindCamera = (a.pcz <= 0.0f) && (k.rejectColour == 0);
  const bool slow = (!a.ok && !behindCamera) || ((!a.inb || rej) && (k.rejectColour != 0)) || (valid && !behind && !etaOK);
  const bool col = upd && !(eta > g.mu) && !(fabsf(eom) > 0.25f);
  return slow ? 2 : (col ? 1 : 0);
}

// computeUpdatedVoxelColorInfo for the shared-camera case: (ix, iy) is the projection computed by v3_depth and has already
// passed the (identical) bounds test
HD bool v3_colour(unsigned &lo, unsigned &hi, float ix, float iy, const FrameGeom &g, const V3K &k, const unsigned *__restrict__ rgbw,
                   const float *div255, const float *rcpW) {
  const int c0 = lo >> 24, c1 = hi & 0xff, c2 = (hi >> 8) & 0xff, wc = (hi >> 16) & 0xff;
  const float oldW = (float)wc;
  const float o0 = div255[c0], o1 = div255[c1], o2 = div255[c2];
  const int px = (int)floorf(ix), py = (int)floorf(iy);
  const float dx = ix - (float)px, dy = iy - (float)py;
  const unsigned *p = rgbw + (unsigned)(px + py * g.rgb_w);
  unsigned a = IV_LDG(p), b = 0, c = 0, d = 0;
  if (dx != 0) b = IV_LDG(p + 1);
  if (dy != 0) c = IV_LDG(p + g.rgb_w);
  if (dx != 0 && dy != 0) d = IV_LDG(p + g.rgb_w + 1);
  const float m0 = div_nr(bil((float)(a & 0xff), (float)(b & 0xff), (float)(c & 0xff), (float)(d & 0xff), dx, dy), 255.0f, k.rcp255);
  const float m1 = div_nr(bil((float)((a >> 8) & 0xff), (float)((b >> 8) & 0xff), (float)((c >> 8) & 0xff), (float)((d >> 8) & 0xff), dx, dy), 255.0f, k.rcp255);
  const float m2 = div_nr(bil((float)((a >> 16) & 0xff), (float)((b >> 16) & 0xff), (float)((c >> 16) & 0xff), (float)((d >> 16) & 0xff), dx, dy), 255.0f, k.rcp255);
  float newW = 5;
  float n0 = o0 * oldW + m0 * newW, n1 = o1 * oldW + m1 * newW, n2 = o2 * oldW + m2 * newW;
  newW = oldW + newW;
  const float yw = rcpW[wc + 5];
  n0 = div_nr(n0, newW, yw); n1 = div_nr(n1, newW, yw); n2 = div_nr(n2, newW, yw);
  const int maxWc = g.maxW & 0xff;
  newW = (newW < maxWc) ? newW : (float)maxWc;
  const unsigned r0 = to_uchar_round(n0 * 255.0f), r1 = to_uchar_round(n1 * 255.0f), r2 = to_uchar_round(n2 * 255.0f);
  const unsigned nlo = (lo & 0x00ffffffu) | (r0 << 24);
  const unsigned nhi = (hi & 0xff000000u) | r1 | (r2 << 8) | ((unsigned)(((int)newW) & 0xff) << 16);
  const bool ch = (nlo != lo) || (nhi != hi);
  lo = nlo; hi = nhi;
  return ch;
}

