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

// ------------------------------------------------------------------------------------------------
// castRay with a neighbourhood cache (k_raycast). Same reads, same arithmetic, same results as cast_ray above; what changes
// is how the blocks are resolved. Measured on B200 (profiles/r02_raycast.md): cast_ray spends 465 warp-instructions per
// interpolated sample because a warp almost always holds lanes of all three tap paths (inside one block, across one face,
// across an edge or corner) and executes all of them, the last one re-hashing for each of its eight taps; and the block that
// hash_found asks for (the block of round(p)) is not the block the taps start in (the block of floor(p)), which thrashes a
// one-entry cache near every block face. Here every step works on the 2x2x2 BLOCK neighbourhood of floor(p)'s block:
//   * entry s (bit 0/1/2 = +1 block in x/y/z) is resolved at most once while floor(p) stays in the block (lazily, `valid` bits);
//   * the block of round(p) is entry sR of it (round(p) is floor(p) or floor(p)+1 on every axis), so hash_found costs no walk
//     of its own when the ray samples the block again;
//   * the eight taps read entry (t & m), m = the axes on which floor(p) sits on the block's last voxel: one straight-line
//     path for every lane.
// The entries live in shared memory (one 8-int column per thread: bank = lane, conflict-free for any index); the host build
// passes a local array.
// ------------------------------------------------------------------------------------------------
struct NbrCache { int kx, ky, kz; unsigned valid; };

HD int lowest_bit_(unsigned v) {
#ifdef __CUDA_ARCH__
  return __ffs((int)v) - 1;
#else
  return __builtin_ffs((int)v) - 1;
#endif
}

HD float s16_to_float_(int v) {   // exact: 1.5 * 2^23 + v stays in [2^23, 2^24) for |v| < 2^22 — two full-rate instructions instead of I2F
#ifdef __CUDA_ARCH__
  return __int_as_float(0x4B400000 + v) - 12582912.0f;
#else
  return (float)v;
#endif
}

// chain walk with three loads per entry (position words and ptr; `offset` only when the entry is not the block): ptr*512 or -1
HD int block_lookup3(const b200_hash_entry *__restrict__ table, int numBuckets, int bx, int by, int bz) {
#if defined(RC_COUNT_WALKS) && !defined(__CUDA_ARCH__)
  ++rc_walks;      // host build only (scripts/raycast_stats.py)
#endif
  int hashIdx = hash_index(bx, by, bz, numBuckets - 1);
  // entry positions are shorts: a block outside their range matches nothing (the reference compares short with int)
  const bool inRange = ((unsigned)(bx + 32768) | (unsigned)(by + 32768) | (unsigned)(bz + 32768)) < 65536u;
  const int key0 = (bx & 0xffff) | (int)((unsigned)by << 16), key1 = bz & 0xffff;
  for (;;) {
    const int *p = reinterpret_cast<const int *>(table) + (size_t)hashIdx * 5;
    const int w0 = RC_LDG(p), w1 = RC_LDG(p + 1), w3 = RC_LDG(p + 3);
    if (w0 == key0 && (w1 & 0xffff) == key1 && w3 >= 0 && inRange) return w3 * BS3;
    const int offset = RC_LDG(p + 2);
    if (offset < 1) return -1;
    hashIdx = numBuckets + offset - 1;
  }
}

#define NBR(s) nbr[(s) * nbrStride]

// New key: everything is resolved again on demand. (Carrying the entry of the block just entered / just left over a one-block
// move of the key saves 9.5 % of the walks — 3.26 M instead of 3.61 M on the KITTI frame of scripts/raycast_stats.py — for ~20
// instructions per key change: a wash, not built in.)
HD void nbr_rekey(NbrCache &c, int kx, int ky, int kz) {
  if (kx != c.kx || ky != c.ky || kz != c.kz) { c.kx = kx; c.ky = ky; c.kz = kz; c.valid = 0u; }
}

HD void nbr_ensure(unsigned need, NbrCache &c, int *nbr, int nbrStride, const b200_hash_entry *__restrict__ table, int nb) {
  unsigned todo = need & ~c.valid;
  while (todo) {
    const int s = lowest_bit_(todo);
    todo &= todo - 1;
    NBR(s) = block_lookup3(table, nb, c.kx + (s & 1), c.ky + ((s >> 1) & 1), c.kz + (s >> 2));
  }
  c.valid |= need;
}

// readFromSDF_float_interpolated (DA/ITMRepresentationAccess.h:252-278) on the neighbourhood cache
HD float sdf_interp_nbr(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, float px, float py, float pz,
                        NbrCache &c, int *nbr, int nbrStride) {
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const float cx = px - fx, cy = py - fy, cz = pz - fz;
  const int x = (int)fx, y = (int)fy, z = (int)fz;
  nbr_rekey(c, x >> 3, y >> 3, z >> 3);
  const int lx = x & 7, ly = y & 7, lz = z & 7;
  const int m = (lx == 7 ? 1 : 0) | (ly == 7 ? 2 : 0) | (lz == 7 ? 4 : 0);   // axes on which the +1 tap is in the next block
  // the blocks the taps touch: entries s with s a subset of m. As a bit set: (1 + 2a)(1 + 4b)(1 + 16c) for m = a + 2b + 4c —
  // the product's bits sit at the sums of the chosen exponents {1}, {2}, {4}, which are exactly the subsets
  nbr_ensure((1u + ((unsigned)m & 1u) * 2u) * (1u + ((unsigned)m & 2u) * 2u) * (1u + ((unsigned)m & 4u) * 4u), c, nbr, nbrStride, table, nb);
  const int ox0 = lx, ox1 = (lx + 1) & 7, oy0 = ly << 3, oy1 = ((ly + 1) & 7) << 3, oz0 = lz << 6, oz1 = ((lz + 1) & 7) << 6;
  const short *sv = reinterpret_cast<const short *>(voxels);
#define TAP(t, off) [&]() { const int b_ = NBR((t) & m); return s16_to_float_(b_ >= 0 ? (int)RC_LDG(sv + (size_t)(b_ + (off)) * 4) : 32767); }()
  const float v000 = TAP(0, ox0 + oy0 + oz0), v100 = TAP(1, ox1 + oy0 + oz0), v010 = TAP(2, ox0 + oy1 + oz0), v110 = TAP(3, ox1 + oy1 + oz0);
  const float v001 = TAP(4, ox0 + oy0 + oz1), v101 = TAP(5, ox1 + oy0 + oz1), v011 = TAP(6, ox0 + oy1 + oz1), v111 = TAP(7, ox1 + oy1 + oz1);
#undef TAP
  float res1, res2;
  res1 = (1.0f - cx) * v000 + cx * v100;
  res1 = (1.0f - cy) * res1 + cy * ((1.0f - cx) * v010 + cx * v110);
  res2 = (1.0f - cx) * v001 + cx * v101;
  res2 = (1.0f - cy) * res2 + cy * ((1.0f - cx) * v011 + cx * v111);
  return ((1.0f - cz) * res1 + cz * res2) / 32767.0f;
}

// castRay — DA/ITMVisualisationEngine.h:93-179 (see cast_ray for the notes on the loop)
__host__ __device__ inline bool cast_ray_nbr(float4 &out, int x, int y, const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table,
                                             int nb, const Mat4 &invM, float invfx, float invfy, float cxp, float cyp, float oneOverVoxelSize, float mu,
                                             float2 minmax, int *nbr, int nbrStride) {
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
  NbrCache c; c.kx = c.ky = c.kz = 0x7fffffff; c.valid = 0u;
  float sdfValue = 1.0f, stepLength;
  while (totalLength < totalLengthMax) {
    // hash_found of readFromSDF_float_uninterpolated(round(p)) (:133): entry sR of floor(p)'s neighbourhood
    const int kx = ((int)floorf(px)) >> 3, ky = ((int)floorf(py)) >> 3, kz = ((int)floorf(pz)) >> 3;
    nbr_rekey(c, kx, ky, kz);
    const int sR = ((((int)round_(px)) >> 3) - kx) | (((((int)round_(py)) >> 3) - ky) << 1) | (((((int)round_(pz)) >> 3) - kz) << 2);
    nbr_ensure(1u << sR, c, nbr, nbrStride, table, nb);
    if (NBR(sR) < 0) {
      sdfValue = 1.0f;   // TVoxel() = 32767 / 32767
      stepLength = BS;
#ifdef __CUDA_ARCH__
      {  // the next position is known now: touch its bucket head so that the next lookup hits L1 (hint only)
        const float qx = px + (float)BS * dx, qy = py + (float)BS * dy, qz = pz + (float)BS * dz;
        const int hidx = hash_index(((int)round_(qx)) >> 3, ((int)round_(qy)) >> 3, ((int)round_(qz)) >> 3, nb - 1);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const int *>(table) + (size_t)hidx * 5));
      }
#endif
    } else {
      sdfValue = sdf_interp_nbr(voxels, table, nb, px, py, pz, c, nbr, nbrStride);
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
    sdfValue = sdf_interp_nbr(voxels, table, nb, px, py, pz, c, nbr, nbrStride);
    stepLength = sdfValue * stepScale;
    px += stepLength * dx; py += stepLength * dy; pz += stepLength * dz;
    found = true;
  } else found = false;
  out = make_float4(px, py, pz, found ? 1.0f : 0.0f);
  return found;
}
#undef NBR
