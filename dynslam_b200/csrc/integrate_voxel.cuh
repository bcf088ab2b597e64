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


// ================================================================================================================
// V4: the same two stages on PAIRS of x-adjacent voxels with Blackwell's packed binary32 instructions
// (add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2: one issue slot for two IEEE-rounded results; measured on B200 at 1.97
// warp-instructions per clock and SM against 2.81 for the scalar forms, i.e. 1.4x the arithmetic per issue slot). Every
// element goes through exactly the operations of the scalar path above, in the same order, so the bits are the same:
// tests/hostcheck runs v4 against the generic per-voxel code on the CPU like it does v3.
// FAST = true is the tolerance-mode build (north_star's bar, not the reference's bits): in the TSDF update (stage B) quotients
// become one multiplication by the tabulated reciprocal and sums are contracted into FMAs. The projection, the depth pixel it
// selects, weights and colours stay bit-exact; a TSDF value may land on the neighbouring 16-bit code (1 LSB = 3.05e-5 after
// SDF_valueToFloat) and the colour gate |eta / mu| <= 0.25 may flip on a tie — both counted by the tests.
// ================================================================================================================
#if defined(__CUDA_ARCH__)
#define F2_DEV 1
#else
#define F2_DEV 0
#endif
HD float2 f2add(float2 a, float2 b) {
#if F2_DEV
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
HD float2 f2mul(float2 a, float2 b) {
#if F2_DEV
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
HD float2 f2fma(float2 a, float2 b, float2 c) {
#if F2_DEV
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
HD float2 f2dup(float v) { return make_float2(v, v); }
// A product that must NOT be contracted with the addition that consumes it: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 even though both carry an explicit rounding mode and the build passes -fmad=false (seen in the SASS of the first V4
// build: 18 mul.rn.f32x2 in the PTX, 16 FMUL2 in the SASS, 1-LSB TSDF differences on a handful of voxels per frame). Two scalar
// products (FMUL, which ptxas leaves alone under -fmad=false) cost one instruction more per pair.
HD float2 f2mul_exact(float2 a, float2 b) {
#if F2_DEV
  return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
HD float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
HD float rcp_approx(float b) {
#ifdef __CUDA_ARCH__
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b)); return y;
#else
  return 1.0f / b;
#endif
}
HD float2 rcp_nr2(float2 nb, float2 y0) {                 // nb = -b, y0 = MUFU.RCP(b): one Newton step, as rcp_nr
  return f2fma(y0, f2fma(nb, y0, f2dup(1.0f)), y0);
}
HD float2 div_nr2(float2 a, float2 nb, float2 y) {        // nb = -b, y = rcp_nr(b): quotient + one residual correction, as div_nr
  const float2 q = f2mul(a, y);
  return f2fma(y, f2fma(nb, q, a), q);
}

struct V4C {            // constants of a launch, duplicated into both halves
  float2 fx, fy, cx, cy, tx, ty, tz, nmu, rcpMu, n32767, rcp32767, c32767, half;
};
HD V4C v4_constants(const FrameGeom &g, const V3K &k) {
  V4C c;
  c.fx = f2dup(g.proj_d[0]); c.fy = f2dup(g.proj_d[1]); c.cx = f2dup(g.proj_d[2]); c.cy = f2dup(g.proj_d[3]);
  c.tx = f2dup(g.M_d.m[12]); c.ty = f2dup(g.M_d.m[13]); c.tz = f2dup(g.M_d.m[14]);
  c.nmu = f2dup(-g.mu); c.rcpMu = f2dup(k.rcpMu); c.n32767 = f2dup(-32767.0f); c.rcp32767 = f2dup(V3_RCP_32767); c.c32767 = f2dup(32767.0f);
  c.half = f2dup(0.5f);
  return c;
}

struct V4A { float2 pcz, ix, iy; unsigned idx0, idx1; bool ok0, ok1, inb0, inb1; };

// XY* = X + Y of the pair (shared by the slabs a lane owns), Z* = the slab's products, both halves equal
template <bool FAST>
HD V4A v4_stage_a(float2 XYx, float2 XYy, float2 XYz, float2 Zx, float2 Zy, float2 Zz, const FrameGeom &g, const V3K &k, const V4C &c) {
  V4A a;
  const float2 pcx = f2add(f2add(XYx, Zx), c.tx), pcy = f2add(f2add(XYy, Zy), c.ty);
  a.pcz = f2add(f2add(XYz, Zz), c.tz);
  const float2 ax = f2mul(c.fx, pcx), ay = f2mul(c.fy, pcy);
  {
    const float lo0 = fminf(fminf(fabsf(ax.x), fabsf(ay.x)), a.pcz.x), hi0 = fmaxf(fmaxf(fabsf(ax.x), fabsf(ay.x)), a.pcz.x);
    const float lo1 = fminf(fminf(fabsf(ax.y), fabsf(ay.y)), a.pcz.y), hi1 = fmaxf(fmaxf(fabsf(ax.y), fabsf(ay.y)), a.pcz.y);
    a.ok0 = (lo0 >= V3_SAFE_LO) && (hi0 <= V3_SAFE_HI);
    a.ok1 = (lo1 >= V3_SAFE_LO) && (hi1 <= V3_SAFE_HI);
  }
  // The projection stays exact in tolerance mode too: it picks the depth pixel ((int)(u + 0.5f)), decides the bounds test and
  // is the colour sampling position — an approximate quotient would, on a tie, read a DIFFERENT pixel and change the voxel
  // by far more than an LSB ("Defining parity").
  const float2 y0 = make_float2(rcp_approx(a.pcz.x), rcp_approx(a.pcz.y));
  const float2 npcz = f2neg(a.pcz);
  const float2 yz = rcp_nr2(npcz, y0);
  a.ix = f2add(div_nr2(ax, npcz, yz), c.cx);
  a.iy = f2add(div_nr2(ay, npcz, yz), c.cy);
  a.inb0 = !((a.ix.x < 1) | (a.ix.x > k.wm2) | (a.iy.x < 1) | (a.iy.x > k.hm2));
  a.inb1 = !((a.ix.y < 1) | (a.ix.y > k.wm2) | (a.iy.y < 1) | (a.iy.y > k.hm2));
  const float2 rx = f2add(a.ix, c.half), ry = f2add(a.iy, c.half);
  a.idx0 = (unsigned)((int)rx.x + (int)ry.x * g.w) & (0u - (unsigned)(a.ok0 & a.inb0));
  a.idx1 = (unsigned)((int)rx.y + (int)ry.y * g.w) & (0u - (unsigned)(a.ok1 & a.inb1));
  return a;
}

// one element's bookkeeping of stage B (everything that is not binary32 arithmetic), shared by both halves
struct V4E { bool valid, behind, etaOK; int oldW, newW; };

template <bool DW, bool FAST>
HD void v4_stage_b(unsigned &lo0, unsigned &lo1, const V4A &a, float2 dm, const FrameGeom &g, const V3K &k, const V4C &c, const float *rcpW,
                    int &r0, int &r1) {
  const float2 eta = f2add(dm, f2neg(a.pcz));          // dm - pc.z
  const float2 eom = FAST ? f2mul(eta, c.rcpMu) : div_nr2(eta, c.nmu, c.rcpMu);
  const float2 newFm = make_float2(minf_(1.0f, eom.x), minf_(1.0f, eom.y));
  const int oldW0 = (lo0 >> 16) & 0xff, oldW1 = (lo1 >> 16) & 0xff;
  const float2 sd = make_float2((float)(short)(lo0 & 0xffffu), (float)(short)(lo1 & 0xffffu));
  const float2 oldF = FAST ? f2mul(sd, c.rcp32767) : div_nr2(sd, c.n32767, c.rcp32767);
  const float2 oldWf = make_float2((float)oldW0, (float)oldW1);
  int nw0 = 1, nw1 = 1;
  float2 num;
  if (DW) {
    nw0 = (int)(100.0 / dm.x); nw0 = nw0 < 1 ? 1 : (nw0 > 10 ? 10 : nw0);
    nw1 = (int)(100.0 / dm.y); nw1 = nw1 < 1 ? 1 : (nw1 > 10 ? 10 : nw1);
    const float2 nwf = make_float2((float)nw0, (float)nw1);
    num = FAST ? f2fma(oldWf, oldF, f2mul(nwf, newFm)) : f2add(f2mul_exact(oldWf, oldF), f2mul_exact(nwf, newFm));
  } else {
    num = FAST ? f2fma(oldWf, oldF, newFm) : f2add(f2mul_exact(oldWf, oldF), newFm);     // newW == 1: 1 * newF is newF, exactly
  }
  int W0 = oldW0 + nw0, W1 = oldW1 + nw1;
  const float2 Wf = make_float2((float)W0, (float)W1);
  const float2 yw = make_float2(rcpW[W0], rcpW[W1]);
  const float2 q = FAST ? f2mul(num, yw) : div_nr2(num, f2neg(Wf), yw);
  W0 = mini_(W0, g.maxW); W1 = mini_(W1, g.maxW);
  const float2 s = f2mul(q, c.c32767);
  const unsigned nlo0 = (lo0 & 0xff000000u) | ((unsigned)(int)(short)(s.x) & 0xffffu) | ((unsigned)(W0 & 0xff) << 16);
  const unsigned nlo1 = (lo1 & 0xff000000u) | ((unsigned)(int)(short)(s.y) & 0xffffu) | ((unsigned)(W1 & 0xff) << 16);
  {
    const bool rej = (dm.x <= 0.0f), valid = a.ok0 && a.inb0 && !rej, behind = eta.x < -g.mu;
    const float ae = fabsf(eta.x);
    const bool etaOK = ((ae >= V3_SAFE_LO) && (ae <= V3_SAFE_HI)) || (eta.x == 0.0f);
    const bool upd = valid && !behind && etaOK;
    lo0 = upd ? nlo0 : lo0;
    const bool behindCamera = (a.pcz.x <= 0.0f) && (k.rejectColour == 0);
    const bool slow = (!a.ok0 && !behindCamera) || ((!a.inb0 || rej) && (k.rejectColour != 0)) || (valid && !behind && !etaOK);
    const bool col = upd && !(eta.x > g.mu) && !(fabsf(eom.x) > 0.25f);
    r0 = slow ? 2 : (col ? 1 : 0);
  }
  {
    const bool rej = (dm.y <= 0.0f), valid = a.ok1 && a.inb1 && !rej, behind = eta.y < -g.mu;
    const float ae = fabsf(eta.y);
    const bool etaOK = ((ae >= V3_SAFE_LO) && (ae <= V3_SAFE_HI)) || (eta.y == 0.0f);
    const bool upd = valid && !behind && etaOK;
    lo1 = upd ? nlo1 : lo1;
    const bool behindCamera = (a.pcz.y <= 0.0f) && (k.rejectColour == 0);
    const bool slow = (!a.ok1 && !behindCamera) || ((!a.inb1 || rej) && (k.rejectColour != 0)) || (valid && !behind && !etaOK);
    const bool col = upd && !(eta.y > g.mu) && !(fabsf(eom.y) > 0.25f);
    r1 = slow ? 2 : (col ? 1 : 0);
  }
}
