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
// visibility test (DA/ITMSceneReconstructionEngine.h:315-397)
// ------------------------------------------------------------------------------------------------
DEV bool point_visible(const Mat4 &M, const float *proj, float x, float y, float z, int w, int h) {
  Vec4 b = m4v4(M, x, y, z, 1.0f);
  if (b.z < 1e-10f) return false;
  float u = proj[0] * b.x / b.z + proj[2];
  float v = proj[1] * b.y / b.z + proj[3];
  return (u >= 0 && u < w && v >= 0 && v < h);
}

// checkBlockVisibility<false>: corner order and the incremental +=/-= updates are part of the arithmetic contract
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

// Transient visibility codes, only alive between k_prepare and k_visible_list of ONE allocate call:
// the reference marks every previously visible block 3 (setToType3) and re-tests the survivors'
// frustum visibility inside the full-table sweep. Here the re-test is done eagerly, one previously
// visible block per thread, and its verdict is parked in the byte; the sweep then only decodes it.
#define VT_PREV_VISIBLE 5   // was 3, frustum test passed  -> ends as 3 unless re-observed (1)
#define VT_PREV_HIDDEN 4    // was 3, frustum test failed  -> ends as 0 unless re-observed (1)

// ------------------------------------------------------------------------------------------------
// prepare: setToType3 over the previous visible list (+ eager frustum re-test) + clear the bitmaps
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prepare(const b200_hash_entry *__restrict__ table, int numBuckets, const b200_vec3i *__restrict__ visiblePos, uint8_t *visType,
          DevCounters *ctr, unsigned *reqBits, unsigned *req2Bits, int noWords, Mat4 M, float p0, float p1, float p2, float p3,
          float voxelSize, int w, int h, int capacity, float2 *minmaxDead, int mw, int mh) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < noWords; i += nth) { reqBits[i] = 0u; req2Bits[i] = 0u; }
  if (minmaxDead) {
    if (tid == 0) ctr->noRenderingBlocks = 0;   // k_visible_list of this frame sums the rendering-tile counts into it   // fused frame: initialise the cells of the expected-depth image outside its live 1/8-res corner here
    const int liveX = (mw - 1) / B200_MINMAX_SUBSAMPLE, liveY = (mh - 1) / B200_MINMAX_SUBSAMPLE;
    const float2 v = make_float2(B200_FAR_AWAY, B200_VERY_CLOSE);
    for (int i = tid; i < mw * mh; i += nth) { const int y = i / mw, x = i - y * mw; if (x > liveX || y > liveY) minmaxDead[i] = v; }
  }
  const float proj[4] = {p0, p1, p2, p3};
  int n = ctr->noVisibleBlocks;
  if (n > capacity) n = capacity;
  for (int i = tid; i < n; i += nth) {
    const b200_vec3i p = visiblePos[i];
    const int idx = find_block<false>(table, numBuckets, p.x, p.y, p.z);
    if (idx >= 0) visType[idx] = block_visible(p.x, p.y, p.z, M, proj, voxelSize, w, h) ? VT_PREV_VISIBLE : VT_PREV_HIDDEN;
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
        // Neighbouring rays miss the same block at the same step: the lanes of the warp that request the same
        // entry elect the one with the largest key (lane order == raster order inside the 8x4 tile, the step
        // is warp-uniform) and only that lane issues the atomics.
        const unsigned peers = __match_any_sync(__activemask(), hashIdx);
        if (lane == 31 - __clz(peers)) {
          atomicMax(&reqKey[hashIdx], make_key(frameTag, pixel, (unsigned)i));
          const unsigned bit = 1u << (hashIdx & 31);
          if (!(reqBits[hashIdx >> 5] & bit)) atomicOr(&reqBits[hashIdx >> 5], bit);
          if (isExcess) { if (!(req2Bits[hashIdx >> 5] & bit)) atomicOr(&req2Bits[hashIdx >> 5], bit); }
          else visType[hashIdx] = 1;
        }
      }
    }
    px += r.dx; py += r.dy; pz += r.dz;
  }
}

// ------------------------------------------------------------------------------------------------
// request server: ONE kernel ranks every request in ascending entry order (two chained scans over
// the request bitmaps, 1024 words = 32768 entries per tile) and serves it:
//   vbaIdx = lastFree - rank(all requests),  exlIdx = lastFreeExcess - rank(excess requests)
// (every request decrements the counters, served or not — Reco_CUDA.cu:833, :857-858); the winning
// pixel's ray is replayed up to its step to recover the block position.
// ------------------------------------------------------------------------------------------------
#define BMP_TILE 1024
__global__ void __launch_bounds__(256)
k_serve_requests(const float *__restrict__ depth, b200_hash_entry *table, int numBuckets, uint8_t *visType,
                 const unsigned long long *__restrict__ reqKey, const unsigned *__restrict__ reqBits, const unsigned *__restrict__ req2Bits,
                 int noWords, const int *__restrict__ allocList, const int *__restrict__ excessList, DevCounters *ctr, FrameGeom g,
                 int currentFrame, unsigned long long *scanDesc, unsigned long long *scanDesc2, unsigned gen) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase, tileBase2;
  const int noTiles = (noWords + BMP_TILE - 1) / BMP_TILE;
  const int baseVba = ctr->lastFreeBlockId, baseExl = ctr->lastFreeExcessListId;   // only the last tile updates them, at its very end
  const float invfx = 1.0f / g.proj_d[0], invfy = 1.0f / g.proj_d[1];
  const float oneOverVoxelSize = 1.0f / (g.voxelSize * BS);
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int first = tile * BMP_TILE + threadIdx.x * 4;
    unsigned wv[4] = {0, 0, 0, 0}, wx[4] = {0, 0, 0, 0};
    if (first + 4 <= noWords) {
      const uint4 a = *reinterpret_cast<const uint4 *>(reqBits + first), b = *reinterpret_cast<const uint4 *>(req2Bits + first);
      wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wx[0] = b.x; wx[1] = b.y; wx[2] = b.z; wx[3] = b.w;
    } else for (int k = 0; k < 4; ++k) if (first + k < noWords) { wv[k] = reqBits[first + k]; wx[k] = req2Bits[first + k]; }
    const unsigned c = __popc(wv[0]) + __popc(wv[1]) + __popc(wv[2]) + __popc(wv[3]);
    const unsigned c2 = __popc(wx[0]) + __popc(wx[1]) + __popc(wx[2]) + __popc(wx[3]);
    unsigned total, total2;
    unsigned rank = block_exclusive_scan(c, sm, &total);
    unsigned rank2 = block_exclusive_scan(c2, sm, &total2);
    if (threadIdx.x < 32) {
      const unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      const unsigned ex2 = scan_lookback(scanDesc2, gen, tile, total2);
      if (threadIdx.x == 0) { tileBase = ex; tileBase2 = ex2; }
    }
    __syncthreads();
    rank += tileBase; rank2 += tileBase2;
    const bool last = (tile == noTiles - 1);
    const unsigned grand = tileBase + total, grand2 = tileBase2 + total2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned bits = wv[k];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int targetIdx = (first + k) * 32 + b;
        const bool isExcess = (wx[k] >> b) & 1u;
        const int vbaIdx = baseVba - (int)rank;
        const int exlIdx = baseExl - (int)rank2;
        rank++;
        if (isExcess) rank2++;
        if (vbaIdx < 0 || (isExcess && exlIdx < 0)) continue;   // exhausted: the counters still go negative
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
    __syncthreads();
    if (last && threadIdx.x == 0) {
      ctr->allocBaseVba = baseVba; ctr->allocBaseExl = baseExl;
      ctr->noRequests = (int)grand; ctr->noRequestsExcess = (int)grand2;
      ctr->lastFreeBlockId = baseVba - (int)grand;
      ctr->lastFreeExcessListId = baseExl - (int)grand2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// visible list (reconstruction engine): decode the visibility bytes and compact every entry with
// type > 0 in ascending entry order — one pass over the 1.5 MB byte array with 128-bit accesses,
// block scan + decoupled look-back for the global order, shared-memory compaction so that the
// scattered 20-byte entry reads and the list writes are spread over all threads. The same pass
// writes the snapshot for the decay queue into the ring (Reco_CUDA.cu:302-317 without the per-frame
// cudaMalloc / blocking copies) and the resolved VBA pointer of every item (saves IntegrateIntoScene
// and CreateExpectedDepths their hash lookups while the table is unchanged).
// ------------------------------------------------------------------------------------------------
#define VIS_EPT 32                 // entries per thread (2 x 128-bit loads)
#define VIS_TILE (256 * VIS_EPT)   // 8192 entries per tile
__global__ void __launch_bounds__(256, 4)
k_visible_list(const b200_hash_entry *__restrict__ table, int numBuckets, int noTotal, uint8_t *visType, b200_vec3i *visiblePos,
               int *visiblePtr, int capacity, DevCounters *ctr, unsigned long long *scanDesc, unsigned gen, Mat4 M, float p0, float p1,
               float p2, float p3, float voxelSize, int w, int h, b200_vec3i *ring, long long ringCap, long long *snapStart,
               int *snapCount, int slot, int oldestSlot, BlockRec *recs, int rw, int rh, unsigned maxRB) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  __shared__ int hits[VIS_TILE];
  unsigned myTiles = 0;   // rendering tiles of the blocks this thread projected (fused frame only)
  const float proj[4] = {p0, p1, p2, p3};
  const long long ringStart = ctr->ringHead;   // advanced by the last tile only, at its very end
  const int noTiles = (noTotal + VIS_TILE - 1) / VIS_TILE;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int first = tile * VIS_TILE + threadIdx.x * VIS_EPT;
    unsigned mask = 0;   // bit k: entry first+k goes to the list
#pragma unroll
    for (int q = 0; q < VIS_EPT / 16; ++q) {
      const int f16 = first + q * 16;
      if (f16 >= noTotal) break;
      uint4 raw = *reinterpret_cast<const uint4 *>(visType + f16);   // noTotal is a multiple of 32
      if (raw.x | raw.y | raw.z | raw.w) {
        uint8_t *t = reinterpret_cast<uint8_t *>(&raw);
        bool dirty = false;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          uint8_t v = t[k];
          if (v == VT_PREV_VISIBLE) { v = 3; dirty = true; }
          else if (v == VT_PREV_HIDDEN) { v = 0; dirty = true; }
          else if (v == 3) {   // a 3 this frame's k_prepare did not produce (decay moved it, Reco_CUDA.cu:1109): test it now
            Entry en = load_entry(table, f16 + k);
            if (!block_visible(en.x, en.y, en.z, M, proj, voxelSize, w, h)) { v = 0; dirty = true; }
          }
          t[k] = v;
          if (v > 0) mask |= 1u << (q * 16 + k);
        }
        if (dirty) *reinterpret_cast<uint4 *>(visType + f16) = raw;
      }
    }
    unsigned total;
    const unsigned local = block_exclusive_scan(__popc(mask), sm, &total);
    if (threadIdx.x < 32) {
      const unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) tileBase = ex;
    }
    unsigned o = local;
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      hits[o++] = first + k;
    }
    __syncthreads();
    const unsigned base = tileBase;
    for (unsigned t = threadIdx.x; t < total; t += blockDim.x) {
      const int idx = hits[t];
      const Entry en = load_entry(table, idx);
      const long long out = (long long)base + t;
      if (out < capacity) {
        b200_vec3i p; p.x = en.x; p.y = en.y; p.z = en.z;
        visiblePos[out] = p;
        ring[(ringStart + out) % ringCap] = p;
        int ptr = en.ptr;
        if (ptr < 0) { if (find_block<false>(table, numBuckets, en.x, en.y, en.z, &ptr) < 0) ptr = -1; }   // stale entry: what findBlock(pos) would hit
        visiblePtr[out] = ptr;
        if (recs) {
          // fused frame: CreateExpectedDepths renders from this very pose, so the block's 1/8-resolution box is produced here
          // (ProjectSingleBlock) and the expected-depth pass only rasterises. All blocks are assumed drawn; k_project_blocks
          // re-does the job with the ordered MAX_RENDERING_BLOCKS rule if the tile total exceeds the cap (never at KITTI sizes).
          BlockRec r; r.ulx = 1; r.uly = 1; r.lrx = 0; r.lry = 0; r.zmin = 0; r.zmax = 0;
          int ulx, uly, lrx, lry; float zmin, zmax;
          if (ptr >= 0 && project_single_block(en.x, en.y, en.z, M, proj, rw, rh, voxelSize, ulx, uly, lrx, lry, zmin, zmax)) {
            r.ulx = (short)ulx; r.uly = (short)uly; r.lrx = (short)lrx; r.lry = (short)lry; r.zmin = zmin; r.zmax = zmax;
            myTiles += rendering_tiles(ulx, uly, lrx, lry);
          }
          recs[out] = r;
        }
      }
    }
    __syncthreads();
    if (tile == noTiles - 1 && threadIdx.x == 0) {
      const int n = (int)(base + total);
      ctr->noVisibleBlocks = n;
      ctr->noIntegrated = 0;            // IntegrateIntoScene of this frame counts from zero (no separate memset)
      const int kept = n < capacity ? n : capacity;
      // (a snapshot that wraps onto older live ones simply overwrites them: decay.cu recognises an overwritten snapshot by
      // ringHead - snapStart > ringCap when its turn comes and sweeps nothing — the oldest snapshots are dropped, never an error)
      snapStart[slot] = ringStart;
      snapCount[slot] = kept;
      ctr->ringHead = ringStart + kept;
    }
  }
  if (recs) {
    for (int o = 16; o > 0; o >>= 1) myTiles += __shfl_xor_sync(0xffffffffu, myTiles, o);
    if ((threadIdx.x & 31) == 0 && myTiles) atomicAdd(&ctr->noRenderingBlocks, myTiles);
    // The CTA that finishes last knows the tile total. In the (pathological) case that it breaks MAX_RENDERING_BLOCKS it
    // re-applies the reference's ordered rule (Vis_CUDA.cu:609: a block is dropped when the running tile count would pass the
    // cap) on its own, serially over the list — so that no extra launch sits between this kernel and the expected-depth fill.
    __shared__ bool lastCta;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) lastCta = (atomicAdd(&ctr->visCtasDone, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!lastCta) return;
    if (threadIdx.x == 0) ctr->visCtasDone = 0;
    __threadfence();
    if (ctr->noRenderingBlocks <= maxRB) return;
    int n = ctr->noVisibleBlocks; if (n > capacity) n = capacity;
    unsigned running = 0;
    for (int base = 0; base < n; base += blockDim.x) {
      const int item = base + threadIdx.x;
      unsigned required = 0;
      BlockRec r; r.ulx = 1; r.uly = 1; r.lrx = 0; r.lry = 0; r.zmin = 0; r.zmax = 0;
      if (item < n) { r = recs[item]; if (r.ulx <= r.lrx) required = rendering_tiles(r.ulx, r.uly, r.lrx, r.lry); }
      unsigned total;
      const unsigned local = running + block_exclusive_scan(required, sm, &total);
      if (item < n && required > 0 && local + required > maxRB) {
        r.ulx = 1; r.uly = 1; r.lrx = 0; r.lry = 0; r.zmin = 0; r.zmax = 0;
        recs[item] = r;
      }
      running += total;
    }
  }
}

// free-view visible list (FindVisibleBlocks, Vis_CUDA.cu:151-180, :536-569): every entry with ptr >= 0
// that passes the frustum test, ascending entry order; the visibility bytes are not touched.
#define FV_EPT 16
#define FV_TILE (256 * FV_EPT)
__global__ void __launch_bounds__(256, 4)
k_freeview_list(const b200_hash_entry *__restrict__ table, int noTotal, b200_vec3i *visiblePos, int capacity, DevCounters *ctr,
                unsigned long long *scanDesc, unsigned gen, Mat4 M, float p0, float p1, float p2, float p3, float voxelSize, int w, int h) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const float proj[4] = {p0, p1, p2, p3};
  const int noTiles = (noTotal + FV_TILE - 1) / FV_TILE;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    // strided assignment inside the tile keeps the 20-byte entry reads of a warp adjacent
    unsigned mask = 0;
    for (int k = 0; k < FV_EPT; ++k) {
      const int idx = tile * FV_TILE + k * 256 + threadIdx.x;
      if (idx < noTotal) {
        Entry en = load_entry(table, idx);
        if (en.ptr >= 0 && block_visible(en.x, en.y, en.z, M, proj, voxelSize, w, h)) mask |= 1u << k;
      }
    }
    // order inside the tile must be ascending entry index: rank = sum over k' < k of count(k') + rank within row k
    unsigned total = 0;
    unsigned offs[FV_EPT];
#pragma unroll
    for (int k = 0; k < FV_EPT; ++k) {
      unsigned rowTotal;
      const unsigned r = block_exclusive_scan((mask >> k) & 1u, sm, &rowTotal);
      offs[k] = total + r;
      total += rowTotal;
    }
    if (threadIdx.x < 32) {
      const unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) { tileBase = ex; if (tile == noTiles - 1) ctr->noVisibleBlocks = (int)(ex + total); }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FV_EPT; ++k) if ((mask >> k) & 1u) {
      const int idx = tile * FV_TILE + k * 256 + threadIdx.x;
      const Entry en = load_entry(table, idx);
      const unsigned o = tileBase + offs[k];
      if ((int)o < capacity) { b200_vec3i p; p.x = en.x; p.y = en.y; p.z = en.z; visiblePos[o] = p; }
    }
    __syncthreads();
  }
}

void launch_allocate(b200_engine *e, const SceneRef &s, const FrameGeom &g, const float *depth, bool onlyVisible, int frameIdx,
                     int snapSlot, b200_vec2f *minmaxDead, int mw, int mh) {
  cudaStream_t st = e->stream;
  const int noWords = e->noWords;
  const unsigned frameTag = (unsigned)(frameIdx + 1) & 0xffffffu;
  if (frameTag == 0) cudaMemsetAsync(e->d_reqKey, 0, sizeof(unsigned long long) * (size_t)s.noTotal, st);
  trace_begin(e, st, "k_prepare");
  k_prepare<<<e->smCount * 4, 256, 0, st>>>(s.hash, s.numBuckets, s.visiblePos, s.visType, e->d_ctr, e->d_reqBits, e->d_req2Bits, noWords,
                                           g.M_d, g.proj_d[0], g.proj_d[1], g.proj_d[2], g.proj_d[3], g.voxelSize, g.w, g.h, s.numBlocks,
                                           (float2 *)minmaxDead, mw, mh);
  trace_end(e, st);
  const int tiles = ((g.w + 7) / 8) * ((g.h + 3) / 4);
  trace_begin(e, st, "k_mark");
  k_mark<<<(tiles + 7) / 8, 256, 0, st>>>(depth, s.hash, s.numBuckets, s.visType, e->d_reqKey, e->d_reqBits, e->d_req2Bits, g,
                                         frameTag);
  trace_end(e, st);
  e->launches += 2;
  if (!onlyVisible) {
    const int bmpTiles = (noWords + BMP_TILE - 1) / BMP_TILE;
    const unsigned gen = ++e->scanGen;
    trace_begin(e, st, "k_serve_requests");
    k_serve_requests<<<persistent_grid(e, 2, bmpTiles), 256, 0, st>>>(depth, s.hash, s.numBuckets, s.visType, e->d_reqKey, e->d_reqBits,
                                                                     e->d_req2Bits, noWords, s.allocationList, s.excessList, e->d_ctr,
                                                                     g, frameIdx, e->d_scanDesc, e->d_scanDesc + e->scanDescCap / 2, gen);
    trace_end(e, st);
    e->launches += 1;
  }
  const int noTiles = (s.noTotal + VIS_TILE - 1) / VIS_TILE;
  const int oldest = e->qSize > 0 ? (e->qHead % SNAP_SLOTS) : -1;
  trace_begin(e, st, "k_visible_list");
  k_visible_list<<<persistent_grid(e, 2, noTiles), 256, 0, st>>>(s.hash, s.numBuckets, s.noTotal, s.visType, s.visiblePos, e->d_visiblePtr,
                                                                s.numBlocks, e->d_ctr, e->d_scanDesc, ++e->scanGen, g.M_d, g.proj_d[0],
                                                                g.proj_d[1], g.proj_d[2], g.proj_d[3], g.voxelSize, g.w, g.h, e->d_ring,
                                                                e->ringCap, e->d_snapStart, e->d_snapCount, snapSlot, oldest,
                                                                minmaxDead ? (BlockRec *)e->d_blockRecs : nullptr, mw, mh,
                                                                (unsigned)e->maxRenderingBlocks);
  trace_end(e, st);
  e->launches += 1;
}

void launch_find_visible(b200_engine *e, const SceneRef &s, const Mat4 &M, const float proj[4], int w, int h, float voxelSize) {
  const int noTiles = (s.noTotal + FV_TILE - 1) / FV_TILE;
  k_freeview_list<<<persistent_grid(e, 4, noTiles), 256, 0, e->stream>>>(s.hash, s.noTotal, s.visiblePos, s.numBlocks, e->d_ctr,
                                                                        e->d_scanDesc, ++e->scanGen, M, proj[0], proj[1], proj[2], proj[3],
                                                                        voxelSize, w, h);
  e->launches++;
}
