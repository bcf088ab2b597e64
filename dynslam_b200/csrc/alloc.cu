// alloc.cu — ResetScene and AllocateSceneFromDepth for sm_100a.
//
// Replaces (reference, src/InfiniTAM/InfiniTAM/ITMLib/Engine/DeviceSpecific/CUDA/
// ITMSceneReconstructionEngine_CUDA.cu): memsetKernel/fillArrayKernel (:145-172), setToType3
// (:776-810), buildHashAllocAndVisibleType_device (:752-773 -> DeviceAgnostic/
// ITMSceneReconstructionEngine.h:176-313), allocateVoxelBlocksList_device (:812-905),
// buildVisibleList_device (:924-999) and the host-side visible-list snapshot (:302-317).
//
// Design (not a port): the reference resolves races with a 4 MiB lock array that is memset every
// frame, lets contended pixels skip ray steps, and hands out VBA slots with atomicSub, so its
// result differs run to run. Here every stage is deterministic and equal to the serial oracle:
//  * a request is a 64-bit atomicMax of (frame | pixel | step): the LAST pixel in raster order wins
//    the bucket, exactly as in the serial loop; the winner's block position is recomputed from
//    (pixel, step) when the request is served, so no 12 MB blockCoords array is written;
//  * requested entries are recorded in a 1-bit-per-entry bitmap; one prefix over the bitmap words
//    gives every request its rank in ascending entry order => slot = allocationList[lastFree - rank];
//  * the visible list is an ordered compaction (decoupled look-back scan over 16-entry/thread tiles
//    of the visibility bytes, 128-bit loads), not an atomicAdd of per-CTA group offsets;
//  * all counters stay on the device; nothing here synchronises with the host.
#include "engine.h"

// ------------------------------------------------------------------------------------------------
// ResetScene
// ------------------------------------------------------------------------------------------------
__global__ void k_reset(uint4 *voxels16, size_t nVox16, int *allocList, int numBlocks, int *hashWords, size_t nHashWords,
                        int *excessList, int excessSize, DevCounters *ctr) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  // two voxels {sdf 32767, everything else 0} per 16 bytes
  const uint4 v = make_uint4(0x00007fffu, 0u, 0x00007fffu, 0u);
  for (size_t i = tid; i < nVox16; i += nth) st_stream(voxels16 + i, v);
  for (size_t i = tid; i < (size_t)numBlocks; i += nth) allocList[i] = (int)i;
  // entry = {pos 0, pad 0, offset 0, ptr -2, allocatedTime 0}: word 3 of every 5 is -2
  for (size_t i = tid; i < nHashWords; i += nth) hashWords[i] = ((i % 5) == 3) ? -2 : 0;
  for (size_t i = tid; i < (size_t)excessSize; i += nth) excessList[i] = (int)i;
  if (tid == 0) { ctr->lastFreeBlockId = numBlocks - 1; ctr->lastFreeExcessListId = excessSize - 1; ctr->totalDecayed = 0; }
}

void launch_reset(b200_engine *e, const SceneRef &s) {
  size_t nVox16 = (size_t)s.numBlocks * BS3 / 2;
  k_reset<<<e->smCount * 8, 256, 0, e->stream>>>((uint4 *)s.voxels, nVox16, s.allocationList, s.numBlocks, (int *)s.hash,
                                                (size_t)s.noTotal * 5, s.excessList, s.excessSize, e->d_ctr);
  e->launches++;
}

// ------------------------------------------------------------------------------------------------
// prepare: setToType3 over the previous visible list + clear the request bitmaps
// ------------------------------------------------------------------------------------------------
__global__ void k_prepare(const b200_hash_entry *table, int numBuckets, const b200_vec3i *visiblePos, uint8_t *visType,
                          const DevCounters *ctr, unsigned *reqBits, unsigned *req2Bits, int noWords) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < noWords; i += nth) { reqBits[i] = 0u; req2Bits[i] = 0u; }
  const int n = ctr->noVisibleBlocks;
  for (int i = tid; i < n; i += nth) {
    b200_vec3i p = visiblePos[i];
    int idx = find_block<false>(table, numBuckets, p.x, p.y, p.z);
    if (idx >= 0) visType[idx] = 3;
  }
}

// ------------------------------------------------------------------------------------------------
// ray set-up shared by the marking kernel and by the request server (which replays one ray)
// ------------------------------------------------------------------------------------------------
struct Ray { float px, py, pz, dx, dy, dz; int noSteps; };

DEV bool make_ray(Ray &r, int x, int y, float d, const FrameGeom &g, float invfx, float invfy, float oneOverVoxelSize) {
  if (d <= 0 || (d - g.mu) < 0 || (d - g.mu) < g.vfmin || (d + g.mu) > g.vfmax) return false;
  float pz = d;
  float px = pz * (((float)x - g.proj_d[2]) * invfx);
  float py = pz * (((float)y - g.proj_d[3]) * invfy);
  float norm = sqrtf(px * px + py * py + pz * pz);
  Vec4 a = m4v4(g.invM_d, px * (1.0f - g.mu / norm), py * (1.0f - g.mu / norm), pz * (1.0f - g.mu / norm), 1.0f);
  float sx = a.x * oneOverVoxelSize, sy = a.y * oneOverVoxelSize, sz = a.z * oneOverVoxelSize;
  Vec4 b = m4v4(g.invM_d, px * (1.0f + g.mu / norm), py * (1.0f + g.mu / norm), pz * (1.0f + g.mu / norm), 1.0f);
  float ex = b.x * oneOverVoxelSize, ey = b.y * oneOverVoxelSize, ez = b.z * oneOverVoxelSize;
  float dx = ex - sx, dy = ey - sy, dz = ez - sz;
  norm = sqrtf(dx * dx + dy * dy + dz * dz);
  int noSteps = (int)ceilf(2.0f * norm);
  float den = (float)(noSteps - 1);
  r.px = sx; r.py = sy; r.pz = sz;
  r.dx = dx / den; r.dy = dy / den; r.dz = dz / den;
  r.noSteps = noSteps;
  return true;
}

#define KEY_PIXEL_BITS 24
#define KEY_STEP_BITS 16
DEV unsigned long long make_key(unsigned frameTag, unsigned pixel, unsigned step) {
  return ((unsigned long long)(frameTag & 0xffffffu) << (KEY_PIXEL_BITS + KEY_STEP_BITS)) |
         ((unsigned long long)pixel << KEY_STEP_BITS) | step;
}

// One warp covers an 8x4 pixel tile (neighbouring rays probe the same buckets, so the 20-byte
// entry loads of a warp collapse to a few L1/L2 transactions).
__global__ void __launch_bounds__(256)
k_mark(const float *__restrict__ depth, const b200_hash_entry *__restrict__ table, int numBuckets, uint8_t *visType,
       unsigned long long *reqKey, unsigned *reqBits, unsigned *req2Bits, FrameGeom g, unsigned frameTag) {
  const int tilesX = (g.w + 7) >> 3, tilesY = (g.h + 3) >> 2;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warpGlobal >= tilesX * tilesY) return;
  const int lane = threadIdx.x & 31;
  const int x = (warpGlobal % tilesX) * 8 + (lane & 7), y = (warpGlobal / tilesX) * 4 + (lane >> 3);
  if (x >= g.w || y >= g.h) return;
  const float invfx = 1.0f / g.proj_d[0], invfy = 1.0f / g.proj_d[1];
  const float oneOverVoxelSize = 1.0f / (g.voxelSize * BS);
  Ray r;
  if (!make_ray(r, x, y, __ldg(depth + x + y * g.w), g, invfx, invfy, oneOverVoxelSize)) return;
  const unsigned pixel = (unsigned)(x + y * g.w);
  float px = r.px, py = r.py, pz = r.pz;
  for (int i = 0; i < r.noSteps; i++) {
    int bx = (short)(int)floorf(px), by = (short)(int)floorf(py), bz = (short)(int)floorf(pz);
    int hashIdx = hash_index(bx, by, bz, numBuckets - 1);
    Entry he = load_entry(table, hashIdx);
    bool isFound = false;
    if (he.x == bx && he.y == by && he.z == bz && he.ptr >= -1) {
      visType[hashIdx] = (he.ptr == -1) ? 2 : 1;
      isFound = true;
    }
    if (!isFound) {
      bool isExcess = false;
      if (he.ptr >= -1) {
        while (he.offset >= 1) {
          hashIdx = numBuckets + he.offset - 1;
          he = load_entry(table, hashIdx);
          if (he.x == bx && he.y == by && he.z == bz && he.ptr >= -1) {
            visType[hashIdx] = (he.ptr == -1) ? 2 : 1;
            isFound = true;
            break;
          }
        }
        isExcess = true;
      }
      if (!isFound) {
        atomicMax(&reqKey[hashIdx], make_key(frameTag, pixel, (unsigned)i));
        const unsigned bit = 1u << (hashIdx & 31);
        if (!(reqBits[hashIdx >> 5] & bit)) atomicOr(&reqBits[hashIdx >> 5], bit);
        if (isExcess) { if (!(req2Bits[hashIdx >> 5] & bit)) atomicOr(&req2Bits[hashIdx >> 5], bit); }
        else visType[hashIdx] = 1;
      }
    }
    px += r.dx; py += r.dy; pz += r.dz;
  }
}

// ------------------------------------------------------------------------------------------------
// request ranking: exclusive prefix of popcounts over a bitmap (ordered, multi-CTA chained scan;
// 1024 words = 32768 entries per tile, 128-bit loads). WHICH 0: all requests, 1: excess requests.
// The second pass also commits the counters (every request decrements, served or not, :833, :857-858).
// ------------------------------------------------------------------------------------------------
#define BMP_TILE 1024
template <int WHICH>
__global__ void __launch_bounds__(256)
k_bitmap_prefix(const unsigned *__restrict__ bits, unsigned *prefix, int noWords, DevCounters *ctr, unsigned long long *scanDesc,
                unsigned gen) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const int noTiles = (noWords + BMP_TILE - 1) / BMP_TILE;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int first = tile * BMP_TILE + threadIdx.x * 4;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (first + 4 <= noWords) w = *reinterpret_cast<const uint4 *>(bits + first);
    else { if (first < noWords) w.x = bits[first]; if (first + 1 < noWords) w.y = bits[first + 1]; if (first + 2 < noWords) w.z = bits[first + 2]; }
    const unsigned c0 = __popc(w.x), c1 = __popc(w.y), c2 = __popc(w.z), c3 = __popc(w.w);
    unsigned total;
    const unsigned local = block_exclusive_scan(c0 + c1 + c2 + c3, sm, &total);
    if (threadIdx.x < 32) {
      const unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) {
        tileBase = ex;
        if (tile == noTiles - 1) {
          const int t = (int)(ex + total);
          if (WHICH == 0) { ctr->noRequests = t; }
          else {
            ctr->noRequestsExcess = t;
            ctr->allocBaseVba = ctr->lastFreeBlockId;
            ctr->allocBaseExl = ctr->lastFreeExcessListId;
            ctr->lastFreeBlockId -= ctr->noRequests;
            ctr->lastFreeExcessListId -= t;
          }
        }
      }
    }
    __syncthreads();
    const unsigned p = tileBase + local;
    if (first + 4 <= noWords) *reinterpret_cast<uint4 *>(prefix + first) = make_uint4(p, p + c0, p + c0 + c1, p + c0 + c1 + c2);
    else { if (first < noWords) prefix[first] = p; if (first + 1 < noWords) prefix[first + 1] = p + c0; if (first + 2 < noWords) prefix[first + 2] = p + c0 + c1; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// request server: one thread per bitmap word; ascending-entry-index slot assignment
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_request_apply(const float *__restrict__ depth, b200_hash_entry *table, int numBuckets, uint8_t *visType,
                const unsigned long long *reqKey, const unsigned *reqBits, const unsigned *req2Bits,
                const unsigned *reqPrefix, const unsigned *req2Prefix, int noWords, const int *allocList,
                const int *excessList, const DevCounters *ctr, FrameGeom g, int currentFrame) {
  const int wIdx = blockIdx.x * blockDim.x + threadIdx.x;
  if (wIdx >= noWords) return;
  unsigned bits = reqBits[wIdx];
  if (!bits) return;
  const unsigned bits2 = req2Bits[wIdx];
  const int baseVba = ctr->allocBaseVba, baseExl = ctr->allocBaseExl;
  unsigned rank = reqPrefix[wIdx], rank2 = req2Prefix[wIdx];
  const float invfx = 1.0f / g.proj_d[0], invfy = 1.0f / g.proj_d[1];
  const float oneOverVoxelSize = 1.0f / (g.voxelSize * BS);
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const int targetIdx = wIdx * 32 + b;
    const bool isExcess = (bits2 >> b) & 1u;
    const int vbaIdx = baseVba - (int)rank;
    const int exlIdx = baseExl - (int)rank2;
    rank++;
    if (isExcess) rank2++;
    if (vbaIdx < 0 || (isExcess && exlIdx < 0)) continue;   // exhausted: counters already went negative
    // replay the winning ray up to its step to recover the requested block position
    const unsigned long long key = reqKey[targetIdx];
    const unsigned step = (unsigned)(key & ((1u << KEY_STEP_BITS) - 1));
    const unsigned pixel = (unsigned)((key >> KEY_STEP_BITS) & ((1u << KEY_PIXEL_BITS) - 1));
    const int x = pixel % g.w, y = pixel / g.w;
    Ray r;
    make_ray(r, x, y, __ldg(depth + pixel), g, invfx, invfy, oneOverVoxelSize);
    float px = r.px, py = r.py, pz = r.pz;
    for (unsigned i = 0; i < step; ++i) { px += r.dx; py += r.dy; pz += r.dz; }
    const int bx = (short)(int)floorf(px), by = (short)(int)floorf(py), bz = (short)(int)floorf(pz);
    int *ew;
    if (!isExcess) {
      ew = reinterpret_cast<int *>(table) + (size_t)targetIdx * 5;
    } else {
      const int exlOffset = excessList[exlIdx];
      reinterpret_cast<int *>(table)[(size_t)targetIdx * 5 + 2] = exlOffset + 1;   // connect to child
      ew = reinterpret_cast<int *>(table) + (size_t)(numBuckets + exlOffset) * 5;
      visType[numBuckets + exlOffset] = 1;                                         // child visible
    }
    ew[0] = (int)(((unsigned)bx & 0xffffu) | ((unsigned)by << 16));
    ew[1] = (bz & 0xffff);
    ew[2] = 0;
    ew[3] = allocList[vbaIdx];
    ew[4] = currentFrame;
  }
}

// ------------------------------------------------------------------------------------------------
// visible list: frustum re-test of type-3 entries + ordered compaction of all types > 0
// ------------------------------------------------------------------------------------------------
DEV bool point_visible(const Mat4 &M, const float *proj, float x, float y, float z, int w, int h) {
  Vec4 b = m4v4(M, x, y, z, 1.0f);
  if (b.z < 1e-10f) return false;
  float u = proj[0] * b.x / b.z + proj[2];
  float v = proj[1] * b.y / b.z + proj[3];
  return (u >= 0 && u < w && v >= 0 && v < h);
}

// checkBlockVisibility<false> — DA/ITMSceneReconstructionEngine.h:338-397 (corner order and the
// incremental +=/-= updates are part of the arithmetic contract)
__device__ bool block_visible(int bx, int by, int bz, const Mat4 &M, const float *proj, float voxelSize, int w, int h) {
  const float factor = (float)BS * voxelSize;
  float x = (float)bx * factor, y = (float)by * factor, z = (float)bz * factor;
  if (point_visible(M, proj, x, y, z, w, h)) return true;
  z += factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  y += factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  x += factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  z -= factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  y -= factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  x -= factor; y += factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  x += factor; y -= factor; z += factor; if (point_visible(M, proj, x, y, z, w, h)) return true;
  return false;
}

// mode 0: reconstruction-engine list (types > 0, re-test type 3, writes the bytes back)
// mode 1: free-view list (every entry with ptr >= 0 that passes the frustum test; bytes untouched)
#define VIS_EPT 64                 // entries per thread (4 x 128-bit loads)
#define VIS_TILE (256 * VIS_EPT)
template <int MODE>
__global__ void __launch_bounds__(256, 4)
k_visible_list(const b200_hash_entry *__restrict__ table, int noTotal, uint8_t *visType, b200_vec3i *visiblePos, int capacity,
               DevCounters *ctr, unsigned long long *scanDesc, unsigned gen, Mat4 M, float p0, float p1, float p2, float p3,
               float voxelSize, int w, int h) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const float proj[4] = {p0, p1, p2, p3};
  const int noTiles = (noTotal + VIS_TILE - 1) / VIS_TILE;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int first = tile * VIS_TILE + threadIdx.x * VIS_EPT;
    unsigned long long mask = 0;   // bit k: entry first+k goes to the list
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < VIS_EPT / 16; ++q) {
        const int f16 = first + q * 16;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (f16 + 16 <= noTotal) raw = *reinterpret_cast<const uint4 *>(visType + f16);
        else for (int k = 0; k < 16; ++k) if (f16 + k < noTotal) reinterpret_cast<uint8_t *>(&raw)[k] = visType[f16 + k];
        if (raw.x | raw.y | raw.z | raw.w) {
          uint8_t *t = reinterpret_cast<uint8_t *>(&raw);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            uint8_t v = t[k];
            if (v == 3) {
              Entry en = load_entry(table, f16 + k);
              if (!block_visible(en.x, en.y, en.z, M, proj, voxelSize, w, h)) { v = 0; visType[f16 + k] = 0; }
            }
            if (v > 0) mask |= 1ull << (q * 16 + k);
          }
        }
      }
    } else {
      for (int k = 0; k < VIS_EPT; ++k) {
        const int idx = first + k;
        if (idx < noTotal) {
          Entry en = load_entry(table, idx);
          if (en.ptr >= 0 && block_visible(en.x, en.y, en.z, M, proj, voxelSize, w, h)) mask |= 1ull << k;
        }
      }
    }
    unsigned total;
    unsigned local = block_exclusive_scan(__popcll(mask), sm, &total);
    if (threadIdx.x < 32) {
      unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) {
        tileBase = ex;
        if (tile == noTiles - 1) ctr->noVisibleBlocks = (int)(ex + total);
      }
    }
    __syncthreads();
    unsigned o = tileBase + local;
    while (mask) {
      const int k = __ffsll((long long)mask) - 1;
      mask &= mask - 1;
      Entry en = load_entry(table, first + k);
      if ((int)o < capacity) { b200_vec3i p; p.x = en.x; p.y = en.y; p.z = en.z; visiblePos[o] = p; }
      o++;
    }
    __syncthreads();
  }
}

// snapshot of the visible list into the decay ring (Reco_CUDA.cu:302-317 without the per-frame
// cudaMalloc / blocking copies): one bump allocation on the device + a coalesced copy
__global__ void k_snapshot(const b200_vec3i *visiblePos, int capacity, DevCounters *ctr, b200_vec3i *ring, long long ringCap,
                           long long *snapStart, int *snapCount, int slot, int oldestSlot) {
  __shared__ long long start;
  int n = ctr->noVisibleBlocks;
  if (n > capacity) n = capacity;
  if (threadIdx.x == 0) {
    // single CTA: plain read-modify-write of the cursor
    start = ctr->ringHead;
    if (oldestSlot >= 0) {
      long long live = start + n - snapStart[oldestSlot];
      if (live > ringCap) ctr->errorFlags |= 1;
    } else if (n > ringCap) ctr->errorFlags |= 1;
    ctr->ringHead = start + n;
    snapStart[slot] = start;
    snapCount[slot] = n;
  }
  __syncthreads();
  const long long s0 = start;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ring[(s0 + i) % ringCap] = visiblePos[i];
}

void launch_allocate(b200_engine *e, const SceneRef &s, const FrameGeom &g, const float *depth, bool onlyVisible, int frameIdx,
                     int snapSlot) {
  cudaStream_t st = e->stream;
  const int noWords = e->noWords;
  const unsigned frameTag = (unsigned)(frameIdx + 1) & 0xffffffu;
  if (frameTag == 0) cudaMemsetAsync(e->d_reqKey, 0, sizeof(unsigned long long) * (size_t)s.noTotal, st);
  k_prepare<<<e->smCount * 4, 256, 0, st>>>(s.hash, s.numBuckets, s.visiblePos, s.visType, e->d_ctr, e->d_reqBits, e->d_req2Bits,
                                           noWords);
  const int tiles = ((g.w + 7) / 8) * ((g.h + 3) / 4);
  k_mark<<<(tiles + 7) / 8, 256, 0, st>>>(depth, s.hash, s.numBuckets, s.visType, e->d_reqKey, e->d_reqBits, e->d_req2Bits, g,
                                         frameTag);
  e->launches += 2;
  if (!onlyVisible) {
    const int bmpTiles = (noWords + BMP_TILE - 1) / BMP_TILE;
    const int bmpGrid = persistent_grid(e, 2, bmpTiles);
    k_bitmap_prefix<0><<<bmpGrid, 256, 0, st>>>(e->d_reqBits, e->d_reqPrefix, noWords, e->d_ctr, e->d_scanDesc, ++e->scanGen);
    k_bitmap_prefix<1><<<bmpGrid, 256, 0, st>>>(e->d_req2Bits, e->d_req2Prefix, noWords, e->d_ctr, e->d_scanDesc, ++e->scanGen);
    k_request_apply<<<(noWords + 255) / 256, 256, 0, st>>>(depth, s.hash, s.numBuckets, s.visType, e->d_reqKey, e->d_reqBits,
                                                           e->d_req2Bits, e->d_reqPrefix, e->d_req2Prefix, noWords,
                                                           s.allocationList, s.excessList, e->d_ctr, g, frameIdx);
    e->launches += 3;
  }
  const int noTiles = (s.noTotal + VIS_TILE - 1) / VIS_TILE;
  const int grid = persistent_grid(e, 4, noTiles);
  k_visible_list<0><<<grid, 256, 0, st>>>(s.hash, s.noTotal, s.visType, s.visiblePos, s.numBlocks, e->d_ctr, e->d_scanDesc,
                                         ++e->scanGen, g.M_d, g.proj_d[0], g.proj_d[1], g.proj_d[2], g.proj_d[3],
                                         g.voxelSize, g.w, g.h);
  const int oldest = e->qSize > 0 ? (e->qHead % SNAP_SLOTS) : -1;
  k_snapshot<<<1, 1024, 0, st>>>(s.visiblePos, s.numBlocks, e->d_ctr, e->d_ring, e->ringCap, e->d_snapStart, e->d_snapCount,
                                 snapSlot, oldest);
  e->launches += 2;
}

void launch_find_visible(b200_engine *e, const SceneRef &s, const Mat4 &M, const float proj[4], int w, int h, float voxelSize) {
  const int noTiles = (s.noTotal + VIS_TILE - 1) / VIS_TILE;
  const int grid = persistent_grid(e, 4, noTiles);
  k_visible_list<1><<<grid, 256, 0, e->stream>>>(s.hash, s.noTotal, s.visType, s.visiblePos, s.numBlocks, e->d_ctr,
                                                e->d_scanDesc, ++e->scanGen, M, proj[0], proj[1], proj[2], proj[3], voxelSize,
                                                w, h);
  e->launches++;
}
