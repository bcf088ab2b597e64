// raycast_ray.cuh — voxel access through the one-entry IndexCache, trilinear TSDF read and castRay, shared by the kernels of
// vis.cu (raycast, forward-projection fill-in). `__host__ __device__`, like integrate_voxel.cuh: tests/hostcheck compiles the
// same functions for the host and compares every ray with the CPU oracle's (tests/test_raycast_host.py).
#pragma once
#include "engine.h"

#ifdef __CUDA_ARCH__
#define RC_LDG(p) __ldg(p)
#else
#include <math.h>
#define RC_LDG(p) (*(p))
#endif

// load_entry (common.cuh) for host + device
HD Entry rc_load_entry(const b200_hash_entry *table, int idx) {
  const int *p = reinterpret_cast<const int *>(table) + (size_t)idx * 5;
  int w0 = RC_LDG(p), w1 = RC_LDG(p + 1);
  Entry e;
  e.x = (short)(w0 & 0xffff); e.y = (short)(w0 >> 16); e.z = (short)(w1 & 0xffff);
  e.offset = RC_LDG(p + 2); e.ptr = RC_LDG(p + 3);
  return e;
}

// ------------------------------------------------------------------------------------------------
// voxel access with the one-entry IndexCache (DA/ITMRepresentationAccess.h:176-220,
// Objects/ITMVoxelBlockHash.h:27-31)
// ------------------------------------------------------------------------------------------------
// The cache memoises the LAST block resolved, hit or miss (blockPtr = -1): the table is constant
// during a raycast, so remembering a miss returns exactly what the reference's chain walk would.
struct IdxCache { int bx, by, bz, blockPtr; };
HD void cache_init(IdxCache &c) { c.bx = c.by = c.bz = 0x7fffffff; c.blockPtr = -1; }

// resolves block (bx,by,bz): returns ptr*512 or -1
HD int block_base(const b200_hash_entry *__restrict__ table, int numBuckets, int bx, int by, int bz, IdxCache &c) {
  if (bx == c.bx && by == c.by && bz == c.bz) return c.blockPtr;
  int hashIdx = hash_index(bx, by, bz, numBuckets - 1);
  int res = -1;
  for (;;) {
    const Entry he = rc_load_entry(table, hashIdx);
    if (he.x == bx && he.y == by && he.z == bz && he.ptr >= 0) { res = he.ptr * BS3; break; }
    if (he.offset < 1) break;
    hashIdx = numBuckets + he.offset - 1;
  }
  c.bx = bx; c.by = by; c.bz = bz; c.blockPtr = res;
  return res;
}

// returns the voxel index in the VBA or -1 (pointToVoxelBlockPos: floor division by 8 == arithmetic shift)
HD int voxel_index(const b200_hash_entry *__restrict__ table, int numBuckets, int px, int py, int pz, IdxCache &c) {
  const int base = block_base(table, numBuckets, px >> 3, py >> 3, pz >> 3, c);
  return base < 0 ? -1 : base + (px & 7) + ((py & 7) << 3) + ((pz & 7) << 6);
}

HD float sdf_raw(const b200_voxel *__restrict__ voxels, int vi) {   // (float)voxel.sdf, missing -> TVoxel() = 32767
  return vi >= 0 ? (float)RC_LDG(reinterpret_cast<const short *>(voxels + vi)) : 32767.0f;
}

HD float rv_sdf(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, int x, int y, int z, IdxCache &c) {
  return sdf_raw(voxels, voxel_index(table, nb, x, y, z, c));
}

// uncached chain walk: ptr*512 or -1
HD int block_lookup(const b200_hash_entry *__restrict__ table, int numBuckets, int bx, int by, int bz) {
  int hashIdx = hash_index(bx, by, bz, numBuckets - 1);
  for (;;) {
    const Entry he = rc_load_entry(table, hashIdx);
    if (he.x == bx && he.y == by && he.z == bz && he.ptr >= 0) return he.ptr * BS3;
    if (he.offset < 1) return -1;
    hashIdx = numBuckets + he.offset - 1;
  }
}

HD float sdf_at(const b200_voxel *__restrict__ voxels, int base, int off) {   // (float)voxel.sdf or TVoxel() when the block is missing
  return base >= 0 ? (float)RC_LDG(reinterpret_cast<const short *>(voxels + base + off)) : 32767.0f;
}

// readFromSDF_float_interpolated — DA/ITMRepresentationAccess.h:252-278. The eight taps are the reference's, operand
// for operand; only the addressing differs. 67% of the positions have their 2x2x2 neighbourhood inside one block
// (one resolve, eight constant-offset 2-byte reads), 29% straddle exactly one block face (two resolves), the rest
// take the general per-tap walk. (A warp almost always contains straddling lanes, so the straddling paths must be
// cheap: the per-tap walk through a one-entry cache re-hashes on every tap.)
HD float sdf_interp(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, float px, float py, float pz,
                     IdxCache &c) {
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const float cx = px - fx, cy = py - fy, cz = pz - fz;
  const int x = (int)fx, y = (int)fy, z = (int)fz;
  float v000, v100, v010, v110, v001, v101, v011, v111;
  const int lx = x & 7, ly = y & 7, lz = z & 7;
  const int m = (lx == 7 ? 1 : 0) | (ly == 7 ? 2 : 0) | (lz == 7 ? 4 : 0);   // block faces crossed
  if ((m & (m - 1)) == 0) {
    // no face or exactly one face crossed
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    const int base0 = block_base(table, nb, bx, by, bz, c);
    const int local = lx + (ly << 3) + (lz << 6);
    int far = base0, farAdj = 0;                     // block of the taps beyond the crossed face, and their index correction
    if (m) {
      far = block_lookup(table, nb, bx + (m & 1), by + ((m >> 1) & 1), bz + (m >> 2));
      farAdj = (m & 1) ? -8 : ((m & 2) ? -64 : -512);   // (l + 1) & 7 == 0 on that axis: undo the carry of the tap offset
    }
#define TAP(bits, off) (((bits) & m) ? sdf_at(voxels, far, local + (off) + farAdj) : sdf_at(voxels, base0, local + (off)))
    v000 = sdf_at(voxels, base0, local);
    v100 = TAP(1, 1);
    v010 = TAP(2, 8);
    v110 = TAP(3, 9);
    v001 = TAP(4, 64);
    v101 = TAP(5, 65);
    v011 = TAP(6, 72);
    v111 = TAP(7, 73);
#undef TAP
  } else {
    v000 = rv_sdf(voxels, table, nb, x, y, z, c);         v100 = rv_sdf(voxels, table, nb, x + 1, y, z, c);
    v010 = rv_sdf(voxels, table, nb, x, y + 1, z, c);     v110 = rv_sdf(voxels, table, nb, x + 1, y + 1, z, c);
    v001 = rv_sdf(voxels, table, nb, x, y, z + 1, c);     v101 = rv_sdf(voxels, table, nb, x + 1, y, z + 1, c);
    v011 = rv_sdf(voxels, table, nb, x, y + 1, z + 1, c); v111 = rv_sdf(voxels, table, nb, x + 1, y + 1, z + 1, c);
  }
  float res1, res2;
  res1 = (1.0f - cx) * v000 + cx * v100;
  res1 = (1.0f - cy) * res1 + cy * ((1.0f - cx) * v010 + cx * v110);
  res2 = (1.0f - cx) * v001 + cx * v101;
  res2 = (1.0f - cy) * res2 + cy * ((1.0f - cx) * v011 + cx * v111);
  return ((1.0f - cz) * res1 + cz * res2) / 32767.0f;
}

// castRay — DA/ITMVisualisationEngine.h:93-179
__host__ __device__ inline bool cast_ray(float4 &out, int x, int y, const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table,
                         int nb, const Mat4 &invM, float invfx, float invfy, float cxp, float cyp, float oneOverVoxelSize, float mu,
                         float2 minmax) {
  const float stepScale = mu * oneOverVoxelSize * 1.0f;
  float cz = minmax.x;
  float cx = cz * (((float)x - cxp) * invfx), cy = cz * (((float)y - cyp) * invfy);
  float totalLength = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxelSize;
  Vec4 r = m4v4(invM, cx, cy, cz, 1.0f);
  const float sx = r.x * oneOverVoxelSize, sy = r.y * oneOverVoxelSize, sz = r.z * oneOverVoxelSize;
  cz = minmax.y;
  cx = cz * (((float)x - cxp) * invfx); cy = cz * (((float)y - cyp) * invfy);
  const float totalLengthMax = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxelSize;
  r = m4v4(invM, cx, cy, cz, 1.0f);
  float dx = r.x * oneOverVoxelSize - sx, dy = r.y * oneOverVoxelSize - sy, dz = r.z * oneOverVoxelSize - sz;
  const float direction_norm = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= direction_norm; dy *= direction_norm; dz *= direction_norm;
  float px = sx, py = sy, pz = sz;
  IdxCache cache; cache_init(cache);
  float sdfValue = 1.0f, stepLength;
  while (totalLength < totalLengthMax) {
    // readFromSDF_float_uninterpolated (:133): only its hash_found matters. When the block exists the value is
    // re-read interpolated unconditionally (the [-100, 20] window at :141-145 contains every sdf in [-1, 1]), so the
    // nearest voxel's 2-byte value is never consumed and is not loaded — one dependent memory latency less per step.
    const int base = block_base(table, nb, ((int)round_(px)) >> 3, ((int)round_(py)) >> 3, ((int)round_(pz)) >> 3, cache);
    if (base < 0) {
      sdfValue = 1.0f;   // TVoxel() = 32767 / 32767
      stepLength = BS;
      // Empty space is crossed in 8-voxel steps, each landing in a new block whose lookup is a dependent L2 read.
      // The next positions are known now: touch their bucket heads so that those lookups hit L1.
#pragma unroll
      for (int a = 1; a <= 2; ++a) {
        const float qx = px + (float)(a * BS) * dx, qy = py + (float)(a * BS) * dy, qz = pz + (float)(a * BS) * dz;   // hint only
        const int hidx = hash_index(((int)round_(qx)) >> 3, ((int)round_(qy)) >> 3, ((int)round_(qz)) >> 3, nb - 1);
#ifdef __CUDA_ARCH__
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const int *>(table) + (size_t)hidx * 5));
#else
        (void)hidx;
#endif
      }
    } else {
      sdfValue = sdf_interp(voxels, table, nb, px, py, pz, cache);
      if (sdfValue <= 0.0f) break;
      stepLength = maxf_(sdfValue * stepScale, 1.0f);
    }
    px += stepLength * dx; py += stepLength * dy; pz += stepLength * dz;
    totalLength += stepLength;
  }
  bool found;
  if (sdfValue <= 0.0f) {
    stepLength = sdfValue * stepScale;
    px += stepLength * dx; py += stepLength * dy; pz += stepLength * dz;
    sdfValue = sdf_interp(voxels, table, nb, px, py, pz, cache);
    stepLength = sdfValue * stepScale;
    px += stepLength * dx; py += stepLength * dy; pz += stepLength * dz;
    found = true;
  } else found = false;
  out = make_float4(px, py, pz, found ? 1.0f : 0.0f);
  return found;
}

