// vis.cu — ITMVisualisationEngine for sm_100a: expected depths, hierarchical raycast, shading,
// ICP maps, forward render, point cloud.
//
// Replaces (reference ITMVisualisationEngine_CUDA.cu): memsetKernel + projectAndSplitBlocks_device +
// fillBlocks_device (:194-240, :572-637), genericRaycast_device (:672-684) with castRay
// (DeviceAgnostic/ITMVisualisationEngine.h:93-179), the five render*_device kernels (:735-886),
// renderICP_device (:716-724), forwardProject/findMissingPoints/genericRaycastMissingPoints/
// renderForward (:639-733) and renderPointCloud_device (:820-871).
//
// Differences in shape: no materialised RenderingBlock list and no blocking copy of its length —
// each visible block rasterises its own bounding box into the 1/8-resolution min/max image with
// native integer atomics on the (positive) float bit patterns, the MAX_RENDERING_BLOCKS rule is
// applied from a list-order prefix of the tile counts; rays are traced by 8x4-pixel warps so that
// a warp's rays walk the same voxel blocks; the voxel reads fetch the 2-byte sdf only.
#include "engine.h"
#include "raycast_ray.cuh"
#include <cstdlib>

// ------------------------------------------------------------------------------------------------
// expected depths
// ------------------------------------------------------------------------------------------------
// ProjectSingleBlock (DA/ITMVisualisationEngine.h:29-71) and the 16-byte block record live in common.cuh: the visible-list
// pass of the fused frame (alloc.cu) produces the records too.
// Stage 1: project every visible block (one thread each, small CTAs so that a few thousand blocks spread over
// many SMs), apply the MAX_RENDERING_BLOCKS rule from the list-order prefix of the tile counts (Vis_CUDA.cu:609)
// and emit a 16-byte record {bbox, z-range}. Cells of a bounding box that lie outside the live 1/8-resolution
// corner (the reference clamps boxes to the FULL-resolution bounds, DA/ITMVisualisationEngine.h:57-60) are
// rasterised here with global atomics; they are rare and never read by the raycast.
// Stage 2 (k_fill_minmax): the live corner is cut into 8x8-cell tiles, one CTA per tile; the CTA scans all
// records, rasterises the boxes that overlap its tile into shared memory with shared-memory atomics and
// writes the tile out with plain stores — no global atomics on the hot cells that hundreds of far blocks
// cover, and no separate initialisation of the live corner.

#define PRJ_THREADS 128   // small CTAs: they must fit beside the integrate kernel's CTAs (register file)
__global__ void __launch_bounds__(PRJ_THREADS)
k_project_blocks(const b200_hash_entry *__restrict__ table, int numBuckets, const b200_vec3i *__restrict__ visiblePos,
                 const int *__restrict__ visiblePtr, int capacity, DevCounters *ctr, Mat4 M, float p0, float p1, float p2, float p3, int w,
                 int h, float voxelSize, float2 *minmax, BlockRec *recs, unsigned long long *scanDesc, unsigned gen, int recsReady, unsigned maxRB) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const float intr[4] = {p0, p1, p2, p3};
  int n = ctr->noVisibleBlocks;
  if (n > capacity) n = capacity;
  const int noTiles = (n + PRJ_THREADS - 1) / PRJ_THREADS;
  const int liveX = (w - 1) / B200_MINMAX_SUBSAMPLE, liveY = (h - 1) / B200_MINMAX_SUBSAMPLE;   // last live cell
  const int lane = threadIdx.x & 31;
  // Fused frame (recsReady): k_serve_list (alloc.cu) already wrote the records, cap rule included, and rasterised the live corner; what is left is
  // the dead-cell part of the boxes, which nothing in the frame reads (a block next to the camera can cover 10^5 of them), so
  // it runs AFTER the fill the raycast waits for. Stand-alone CreateExpectedDepths (recsReady 0) does everything here.
  const bool fast = recsReady != 0;
  const bool doDead = true;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int item = tile * PRJ_THREADS + threadIdx.x;
    int ulx = 0, uly = 0, lrx = -1, lry = -1; float zmin = 0, zmax = 0;
    unsigned required = 0;
    bool draw;
    if (fast) {
      if (item < n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(recs) + item);
        ulx = (short)(q.x & 0xffff); uly = (short)(q.x >> 16); lrx = (short)(q.y & 0xffff); lry = (short)(q.y >> 16);
        zmin = __uint_as_float(q.z); zmax = __uint_as_float(q.w);
      }
      draw = ulx <= lrx;
    } else {
    if (item < n) {
      const b200_vec3i p = visiblePos[item];
      const bool allocated = visiblePtr ? (visiblePtr[item] >= 0) : (find_block<false>(table, numBuckets, p.x, p.y, p.z) >= 0);
      if (allocated) {
        if (project_single_block(p.x, p.y, p.z, M, intr, w, h, voxelSize, ulx, uly, lrx, lry, zmin, zmax)) {
          required = rendering_tiles(ulx, uly, lrx, lry);
        }
      }
    }
    unsigned total;
    const unsigned local = block_exclusive_scan(required, sm, &total);
    if (threadIdx.x < 32) {
      const unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) { tileBase = ex; if (tile == noTiles - 1) ctr->noRenderingBlocks = ex + total; }
    }
    __syncthreads();
    const unsigned out_offset = tileBase + local;
    draw = required > 0 && (out_offset + required <= maxRB);   // :609
    if (item < n) {
      BlockRec r;
      if (draw) { r.ulx = (short)ulx; r.uly = (short)uly; r.lrx = (short)lrx; r.lry = (short)lry; r.zmin = zmin; r.zmax = zmax; }
      else { r.ulx = 1; r.uly = 1; r.lrx = 0; r.lry = 0; r.zmin = 0; r.zmax = 0; }
      recs[item] = r;
    }
    }
    // the part of a box outside the live corner (dead cells): warp-cooperative, one box at a time
    unsigned todo = __ballot_sync(0xffffffffu, doDead && item < n && draw && (lrx > liveX || lry > liveY));
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int ax = __shfl_sync(0xffffffffu, ulx, src), ay = __shfl_sync(0xffffffffu, uly, src);
      const int bx = __shfl_sync(0xffffffffu, lrx, src), by = __shfl_sync(0xffffffffu, lry, src);
      const float zn = __shfl_sync(0xffffffffu, zmin, src), zx = __shfl_sync(0xffffffffu, zmax, src);
      const int bw = bx - ax + 1, cnt = bw * (by - ay + 1);
      for (int k = lane; k < cnt; k += 32) {
        const int yy = ay + k / bw, xx = ax + k % bw;
        if (xx > liveX || yy > liveY) { float2 *px = &minmax[xx + yy * w]; atomic_min_posf(&px->x, zn); atomic_max_posf(&px->y, zx); }
      }
    }
    __syncthreads();
  }
  if (!fast && noTiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctr->noRenderingBlocks = 0;
}

#define FILL_T 8          // tile edge in 1/8-resolution cells
#define FILL_THREADS 128
__global__ void __launch_bounds__(FILL_THREADS)
k_fill_minmax(const BlockRec *__restrict__ recs, const DevCounters *ctr, int capacity, int w, int h, float2 *minmax) {
  __shared__ int smin[FILL_T * FILL_T], smax[FILL_T * FILL_T];
  const int liveX = (w - 1) / B200_MINMAX_SUBSAMPLE, liveY = (h - 1) / B200_MINMAX_SUBSAMPLE;
  const int tx0 = blockIdx.x * FILL_T, ty0 = blockIdx.y * FILL_T;
  const int tx1 = min(tx0 + FILL_T - 1, liveX), ty1 = min(ty0 + FILL_T - 1, liveY);
  if (threadIdx.x < FILL_T * FILL_T) { smin[threadIdx.x] = __float_as_int(B200_FAR_AWAY); smax[threadIdx.x] = __float_as_int(B200_VERY_CLOSE); }
  __syncthreads();
  int n = ctr->noVisibleBlocks;
  if (n > capacity) n = capacity;
  const uint4 *r4 = reinterpret_cast<const uint4 *>(recs);
  for (int i0 = threadIdx.x; i0 < n; i0 += FILL_THREADS * 8) {
    uint4 q8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * FILL_THREADS; q8[u] = (i < n) ? __ldg(r4 + i) : make_uint4(1u | (1u << 16), 0u, 0u, 0u); }   // 8 loads in flight
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 q = q8[u];
      const int ulx = (short)(q.x & 0xffff), uly = (short)(q.x >> 16), lrx = (short)(q.y & 0xffff), lry = (short)(q.y >> 16);
      const int ax = max(ulx, tx0), bx = min(lrx, tx1), ay = max(uly, ty0), by = min(lry, ty1);
      if (ax > bx || ay > by) continue;
      const int zn = (int)q.z, zx = (int)q.w;   // positive floats: integer order == float order
      for (int yy = ay; yy <= by; ++yy)
        for (int xx = ax; xx <= bx; ++xx) {
          const int c = (yy - ty0) * FILL_T + (xx - tx0);
          atomicMin(&smin[c], zn);
          atomicMax(&smax[c], zx);
        }
    }
  }
  __syncthreads();
  if (threadIdx.x < FILL_T * FILL_T) {
    const int xx = tx0 + (threadIdx.x % FILL_T), yy = ty0 + (threadIdx.x / FILL_T);
    if (xx <= tx1 && yy <= ty1) minmax[xx + yy * w] = make_float2(__int_as_float(smin[threadIdx.x]), __int_as_float(smax[threadIdx.x]));
  }
}

// initialises every cell OUTSIDE the live corner (the live corner is written in full by k_fill_minmax)
__global__ void k_minmax_init_dead(float2 *minmax, int w, int h) {
  const int liveX = (w - 1) / B200_MINMAX_SUBSAMPLE, liveY = (h - 1) / B200_MINMAX_SUBSAMPLE;
  const float2 v = make_float2(B200_FAR_AWAY, B200_VERY_CLOSE);
  const int n = w * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    if (x > liveX || y > liveY) minmax[i] = v;
  }
}

void launch_expected_depths(b200_engine *e, const SceneRef &s, const Mat4 &M, const float proj[4], int w, int h, float voxelSize,
                            b200_vec2f *minmax, bool deadInitDone, bool recsReady) {
  if (!deadInitDone) { k_minmax_init_dead<<<e->smCount * 4, 256, 0, e->stream>>>((float2 *)minmax, w, h); e->launches++; }
  const int noTiles = (s.numBlocks + PRJ_THREADS - 1) / PRJ_THREADS;
  if (!recsReady) {   // otherwise the records come from k_serve_list of this frame; the dead cells follow in launch_expected_depths_dead
    trace_begin(e, e->stream, "k_project_blocks");
    k_project_blocks<<<persistent_grid(e, 2, noTiles), PRJ_THREADS, 0, e->stream>>>(s.hash, s.numBuckets, s.visiblePos, fresh_ptr_list(e, s),
                                                                                   s.numBlocks, e->d_ctr, M, proj[0], proj[1], proj[2], proj[3],
                                                                                   w, h, voxelSize, (float2 *)minmax, (BlockRec *)e->d_blockRecs,
                                                                                   e->d_scanDesc, ++e->scanGen, 0, (unsigned)e->maxRenderingBlocks);
    trace_end(e, e->stream);
    e->launches++;
  }
  dim3 grid(((w - 1) / B200_MINMAX_SUBSAMPLE) / FILL_T + 1, ((h - 1) / B200_MINMAX_SUBSAMPLE) / FILL_T + 1);
  trace_begin(e, e->stream, "k_fill_minmax");
  k_fill_minmax<<<grid, FILL_THREADS, 0, e->stream>>>((const BlockRec *)e->d_blockRecs, e->d_ctr, s.numBlocks, w, h, (float2 *)minmax);
  trace_end(e, e->stream);
  e->launches += 1;
}

// ---- stand-alone CreateExpectedDepths, fast form ------------------------------------------------------------------------
// One pass over the visible list: every listed block goes to an 8-lane group (a corner per lane, project_block_group), its box
// — clamped to the FULL-resolution bounds, as the reference does — is rasterised with atomic min / max into the image that
// k_minmax_init_all reset: by the group up to 64 cells, by the warp up to 512, by the whole CTA beyond (the boxes of blocks
// that leave the frustum on the right or at the bottom spill over up to 10^5 cells of the part of the image nothing reads).
// The MAX_RENDERING_BLOCKS rule is order dependent; this pass only COUNTS the rendering tiles, and when the total breaks the cap
// (never at KITTI sizes) the host runs the ordered two-kernel form above instead (engine.cu).
#define XD_BIG 64
__global__ void k_minmax_init_all(float2 *minmax, int n, DevCounters *ctr) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ctr->noRenderingBlocks = 0;
  const float2 v = make_float2(B200_FAR_AWAY, B200_VERY_CLOSE);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) minmax[i] = v;
}

__global__ void __launch_bounds__(256)
k_expected_depths_fast(const b200_hash_entry *__restrict__ table, int numBuckets, const b200_vec3i *__restrict__ visiblePos,
                       const int *__restrict__ visiblePtr, int capacity, DevCounters *ctr, Mat4 M, float p0, float p1, float p2, float p3, int w,
                       int h, float voxelSize, float2 *minmax) {
  __shared__ BlockRec bigRecs[XD_BIG];
  __shared__ int bigCount;
  const float intr[4] = {p0, p1, p2, p3};
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  if (threadIdx.x == 0) bigCount = 0;
  __syncthreads();
  int n = ctr->noVisibleBlocks;
  if (n > capacity) n = capacity;
  unsigned myTiles = 0;
  for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {
    const int item = i0 + grp;
    bool have = false;
    int ex = 0, ey = 0, ez = 0;
    if (item < n) {
      const b200_vec3i p = visiblePos[item];
      ex = p.x; ey = p.y; ez = p.z;
      have = visiblePtr ? (__ldg(visiblePtr + item) >= 0) : (find_block<false>(table, numBuckets, p.x, p.y, p.z) >= 0);
    }
    int ulx, uly, lrx, lry; float zl, zh;
    const bool draw = project_block_group(have, ex, ey, ez, M, intr, w, h, voxelSize, ulx, uly, lrx, lry, zl, zh);
    if (draw && sub == 0) myTiles += rendering_tiles(ulx, uly, lrx, lry);
    const int cells = draw ? (lrx - ulx + 1) * (lry - uly + 1) : 0;
    if (cells > 0 && cells <= 64) raster_box_group(minmax, w, ulx, uly, lrx, lry, zl, zh);
    unsigned todo = __ballot_sync(0xffffffffu, sub == 0 && cells > 64);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int ax = __shfl_sync(0xffffffffu, ulx, src), ay = __shfl_sync(0xffffffffu, uly, src);
      const int cx = __shfl_sync(0xffffffffu, lrx, src), cy = __shfl_sync(0xffffffffu, lry, src);
      const float zn = __shfl_sync(0xffffffffu, zl, src), zx = __shfl_sync(0xffffffffu, zh, src);
      if ((cx - ax + 1) * (cy - ay + 1) > 512) {
        int slotBig = -1;
        if (lane == 0) slotBig = atomicAdd(&bigCount, 1);
        slotBig = __shfl_sync(0xffffffffu, slotBig, 0);
        if (slotBig < XD_BIG) {
          if (lane == 0) { BlockRec b; b.ulx = (short)ax; b.uly = (short)ay; b.lrx = (short)cx; b.lry = (short)cy; b.zmin = zn; b.zmax = zx; bigRecs[slotBig] = b; }
          continue;
        }
      }
      raster_box_warp(minmax, w, ax, ay, cx, cy, zn, zx);
    }
  }
  for (int o = 16; o > 0; o >>= 1) myTiles += __shfl_xor_sync(0xffffffffu, myTiles, o);
  if (lane == 0 && myTiles) atomicAdd(&ctr->noRenderingBlocks, myTiles);
  __syncthreads();
  const int nb = bigCount < XD_BIG ? bigCount : XD_BIG;
  for (int b = 0; b < nb; ++b) {
    const BlockRec br = bigRecs[b];
    const int bw = br.lrx - br.ulx + 1, bh = br.lry - br.uly + 1;
    // rows of the box over the warps, columns over the lanes: no division per cell
    for (int yy = br.uly + (int)(threadIdx.x >> 5); yy < br.uly + bh; yy += (int)(blockDim.x >> 5))
      for (int xx = br.ulx + lane; xx < br.ulx + bw; xx += 32) {
        float2 *px = &minmax[xx + yy * w];
        atomic_min_posf(&px->x, br.zmin); atomic_max_posf(&px->y, br.zmax);
      }
  }
}

void launch_expected_depths_fast(b200_engine *e, const SceneRef &s, const Mat4 &M, const float proj[4], int w, int h, float voxelSize,
                                 b200_vec2f *minmax) {
  k_minmax_init_all<<<e->smCount * 4, 256, 0, e->stream>>>((float2 *)minmax, w * h, e->d_ctr);
  const int groupsOf32 = (s.numBlocks + 31) / 32;
  trace_begin(e, e->stream, "k_expected_depths_fast");
  k_expected_depths_fast<<<persistent_grid(e, 4, groupsOf32), 256, 0, e->stream>>>(s.hash, s.numBuckets, s.visiblePos, fresh_ptr_list(e, s), s.numBlocks,
                                                                                  e->d_ctr, M, proj[0], proj[1], proj[2], proj[3], w, h, voxelSize,
                                                                                  (float2 *)minmax);
  trace_end(e, e->stream);
  e->launches += 2;
}

// second half of the fused frame's expected-depth work: the dead cells (after launch_expected_depths(..., recsReady = true))
void launch_expected_depths_dead(b200_engine *e, const SceneRef &s, const Mat4 &M, const float proj[4], int w, int h, float voxelSize,
                                 b200_vec2f *minmax) {
  const int noTiles = (s.numBlocks + PRJ_THREADS - 1) / PRJ_THREADS;
  trace_begin(e, e->stream, "k_project_blocks/dead");
  k_project_blocks<<<persistent_grid(e, 2, noTiles), PRJ_THREADS, 0, e->stream>>>(s.hash, s.numBuckets, s.visiblePos, fresh_ptr_list(e, s),
                                                                                 s.numBlocks, e->d_ctr, M, proj[0], proj[1], proj[2], proj[3],
                                                                                 w, h, voxelSize, (float2 *)minmax, (BlockRec *)e->d_blockRecs,
                                                                                 e->d_scanDesc, ++e->scanGen, 2, (unsigned)e->maxRenderingBlocks);
  trace_end(e, e->stream);
  e->launches++;
}

// GenericRaycast — Vis_CUDA.cu:242-265, :672-684. One ray per thread, 8x4-pixel warps (a warp's rays walk the same
// voxel blocks). Measured on B200: the march is bound by the chain of dependent L2 reads per step, not by lane
// utilisation — a ray-refill variant (fewer, longer-lived warps) was 1.1-3x slower (profiles/), so the kernel keeps
// the maximum number of independent rays in flight and shortens the chain per step instead (see cast_ray).
// 64-thread CTAs: rays finish at very different times, small CTAs hand their SM slots back sooner.
#define RC_THREADS 64
template <bool useNbr>
__global__ void __launch_bounds__(RC_THREADS)
k_raycast(float4 *out, const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, int w, int h, Mat4 invM,
          float fx, float fy, float cxp, float cyp, float voxelSize, float mu, const float2 *__restrict__ minmax, int twLog2, int centreRow,
          int tilesX, int tilesY, unsigned tilesXMagic) {
  // warp tile: (1 << twLog2) x (32 >> twLog2) pixels
  const int tw = 1 << twLog2, th = 32 >> twLog2;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warpGlobal >= tilesX * tilesY) return;
  const int lane = threadIdx.x & 31;
  const int rowLinear = (int)__umulhi((unsigned)warpGlobal, tilesXMagic);      // warpGlobal / tilesX (magic = ceil(2^32 / tilesX))
  int tileRow = rowLinear;
  if (centreRow >= 0) {
    // dispatch order = tile rows sorted by distance from `centreRow` (c, c+1, c-1, c+2, ...): the rows around the horizon
    // hold the rays that skim the ground inside its truncation band for hundreds of steps; started first, they overlap
    // the rest of the image instead of forming the kernel's tail. Every pixel is still written exactly once.
    const int c = centreRow, up = tilesY - 1 - c, down = c, m = up < down ? up : down, k = tileRow;
    if (k <= 2 * m) { const int off = (k + 1) >> 1; tileRow = (k & 1) ? c + off : c - off; }
    else { const int rest = k - 2 * m; tileRow = (up > down) ? c + m + rest : c - m - rest; }
  }
  const int x = (warpGlobal - rowLinear * tilesX) * tw + (lane & (tw - 1)), y = tileRow * th + (lane >> twLog2);
  if (x >= w || y >= h) return;
  const int locId2 = (int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w;
  float4 o;
  if (useNbr) {
    __shared__ int nbr[8][RC_THREADS];     // per thread: the 2x2x2 block neighbourhood of the ray's current block (raycast_ray.cuh)
    cast_ray_nbr(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, __ldg(minmax + locId2), &nbr[0][threadIdx.x], RC_THREADS);
  } else {
    cast_ray(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, __ldg(minmax + locId2));
  }
  out[x + y * w] = o;
}

void launch_raycast(b200_engine *e, const SceneRef &s, const Mat4 &invM, const float proj[4], int w, int h, float voxelSize, float mu,
                    const b200_vec2f *minmax, b200_vec4f *out) {
  static int twLog2 = -1;
  if (twLog2 < 0) { const char *v = getenv("B200_RC_TILE"); twLog2 = v ? atoi(v) : 3; if (twLog2 < 0 || twLog2 > 5) twLog2 = 3; }   // 3: 8x4 (default)
  const int tw = 1 << twLog2, th = 32 >> twLog2;
  const int tiles = ((w + tw - 1) / tw) * ((h + th - 1) / th);
  const int warpsPerCta = RC_THREADS / 32;
  static int order = -1;
  if (order < 0) { const char *v = getenv("B200_RC_ORDER"); order = v ? atoi(v) : 1; }   // default: principal-point row first; 0 = plain row-major
  int centreRow = -1;
  if (order == 1) {   // principal-point row first (level camera: the horizon)
    const int tilesY = (h + th - 1) / th;
    centreRow = (int)(proj[3] / (float)th);
    if (centreRow < 0) centreRow = 0;
    if (centreRow > tilesY - 1) centreRow = tilesY - 1;
  }
  const int tilesXh = (w + tw - 1) / tw, tilesYh = (h + th - 1) / th;
  const unsigned magic = (unsigned)((0x100000000ull + (unsigned)tilesXh - 1) / (unsigned)tilesXh);
  static int impl = -1;
  if (impl < 0) { const char *v = getenv("B200_RC_IMPL"); impl = (v && v[0] == 'o') ? 0 : 1; }   // "old": per-lane one-entry cache (cross-check); default: neighbourhood cache
  trace_begin(e, e->stream, "k_raycast");
  if (impl)
    k_raycast<true><<<(tiles + warpsPerCta - 1) / warpsPerCta, RC_THREADS, 0, e->stream>>>((float4 *)out, s.voxels, s.hash, s.numBuckets, w, h, invM,
                                                                                          proj[0], proj[1], proj[2], proj[3], voxelSize, mu,
                                                                                          (const float2 *)minmax, twLog2, centreRow, tilesXh, tilesYh, magic);
  else
    k_raycast<false><<<(tiles + warpsPerCta - 1) / warpsPerCta, RC_THREADS, 0, e->stream>>>((float4 *)out, s.voxels, s.hash, s.numBuckets, w, h, invM,
                                                                                           proj[0], proj[1], proj[2], proj[3], voxelSize, mu,
                                                                                           (const float2 *)minmax, twLog2, centreRow, tilesXh, tilesYh, magic);
  trace_end(e, e->stream);
  e->launches++;
}

// ------------------------------------------------------------------------------------------------
// shading from the volume
// ------------------------------------------------------------------------------------------------
DEV float rvs(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, int x, int y, int z) {
  IdxCache c; cache_init(c);   // the 4-argument readVoxel builds a fresh cache per read (:222-228)
  return rv_sdf(voxels, table, nb, x, y, z, c);
}

// computeSingleNormalFromSDF — DA/ITMRepresentationAccess.h:336-449
__device__ void normal_from_sdf(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, float px, float py,
                                float pz, float &rx, float &ry, float &rz) {
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const float cx = px - fx, cy = py - fy, cz = pz - fz;
  const int X = (int)fx, Y = (int)fy, Z = (int)fz;
  const float nx = 1.0f - cx, ny = 1.0f - cy, nz = 1.0f - cz;
#define RV(a, b, c) rvs(voxels, table, nb, X + (a), Y + (b), Z + (c))
  const float f0 = RV(0, 0, 0), f1 = RV(1, 0, 0), f2 = RV(0, 1, 0), f3 = RV(1, 1, 0);
  const float b0 = RV(0, 0, 1), b1 = RV(1, 0, 1), b2 = RV(0, 1, 1), b3 = RV(1, 1, 1);
  float t0, t1, t2, t3, p1, p2, v1;
  p1 = f0 * ny * nz + f2 * cy * nz + b0 * ny * cz + b2 * cy * cz;
  t0 = RV(-1, 0, 0); t1 = RV(-1, 1, 0); t2 = RV(-1, 0, 1); t3 = RV(-1, 1, 1);
  p2 = t0 * ny * nz + t1 * cy * nz + t2 * ny * cz + t3 * cy * cz;
  v1 = p1 * cx + p2 * nx;
  p1 = f1 * ny * nz + f3 * cy * nz + b1 * ny * cz + b3 * cy * cz;
  t0 = RV(2, 0, 0); t1 = RV(2, 1, 0); t2 = RV(2, 0, 1); t3 = RV(2, 1, 1);
  p2 = t0 * ny * nz + t1 * cy * nz + t2 * ny * cz + t3 * cy * cz;
  rx = (p1 * nx + p2 * cx - v1) / 32767.0f;
  p1 = f0 * nx * nz + f1 * cx * nz + b0 * nx * cz + b1 * cx * cz;
  t0 = RV(0, -1, 0); t1 = RV(1, -1, 0); t2 = RV(0, -1, 1); t3 = RV(1, -1, 1);
  p2 = t0 * nx * nz + t1 * cx * nz + t2 * nx * cz + t3 * cx * cz;
  v1 = p1 * cy + p2 * ny;
  p1 = f2 * nx * nz + f3 * cx * nz + b2 * nx * cz + b3 * cx * cz;
  t0 = RV(0, 2, 0); t1 = RV(1, 2, 0); t2 = RV(0, 2, 1); t3 = RV(1, 2, 1);
  p2 = t0 * nx * nz + t1 * cx * nz + t2 * nx * cz + t3 * cx * cz;
  ry = (p1 * ny + p2 * cy - v1) / 32767.0f;
  p1 = f0 * nx * ny + f1 * cx * ny + f2 * nx * cy + f3 * cx * cy;
  t0 = RV(0, 0, -1); t1 = RV(1, 0, -1); t2 = RV(0, 1, -1); t3 = RV(1, 1, -1);
  p2 = t0 * nx * ny + t1 * cx * ny + t2 * nx * cy + t3 * cx * cy;
  v1 = p1 * cz + p2 * nz;
  p1 = b0 * nx * ny + b1 * cx * ny + b2 * nx * cy + b3 * cx * cy;
  t0 = RV(0, 0, 2); t1 = RV(1, 0, 2); t2 = RV(0, 1, 2); t3 = RV(1, 1, 2);
  p2 = t0 * nx * ny + t1 * cx * ny + t2 * nx * cy + t3 * cx * cy;
  rz = (p1 * nz + p2 * cz - v1) / 32767.0f;
#undef RV
}

// computeNormalAndAngle<TVoxel,TIndex> — DA/ITMVisualisationEngine.h:196-210
DEV void normal_and_angle_sdf(bool &found, float px, float py, float pz, const b200_voxel *__restrict__ voxels,
                              const b200_hash_entry *__restrict__ table, int nb, float lx, float ly, float lz, float &nx, float &ny,
                              float &nz, float &angle) {
  if (!found) return;
  normal_from_sdf(voxels, table, nb, px, py, pz, nx, ny, nz);
  const float normScale = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  nx *= normScale; ny *= normScale; nz *= normScale;
  angle = nx * lx + ny * ly + nz * lz;
  if (!(angle > 0.0)) found = false;
}

// readFromSDF_color4u_interpolated_noalpha — DA/ITMRepresentationAccess.h:280-318
__device__ void color_interp(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, float px, float py,
                             float pz, float &r0, float &r1, float &r2) {
  IdxCache c; cache_init(c);
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const float cx = px - fx, cy = py - fy, cz = pz - fz;
  const int X = (int)fx, Y = (int)fy, Z = (int)fz;
  r0 = r1 = r2 = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ox = k & 1, oy = (k >> 1) & 1, oz = k >> 2;
    const int vi = voxel_index(table, nb, X + ox, Y + oy, Z + oz, c);
    float cr = 0.0f, cg = 0.0f, cb = 0.0f;
    if (vi >= 0) {
      const unsigned lo = __ldg(reinterpret_cast<const unsigned *>(voxels + vi)), hi = __ldg(reinterpret_cast<const unsigned *>(voxels + vi) + 1);
      cr = (float)((lo >> 24) & 0xff); cg = (float)(hi & 0xff); cb = (float)((hi >> 8) & 0xff);
    }
    const float wgt = (ox ? cx : (1.0f - cx)) * (oy ? cy : (1.0f - cy)) * (oz ? cz : (1.0f - cz));
    r0 += wgt * cr; r1 += wgt * cg; r2 += wgt * cb;
  }
  r0 = r0 / 255.0f; r1 = r1 / 255.0f; r2 = r2 / 255.0f;
}

DEV uchar4 grey_px(float angle) {   // drawPixelGrey, DA/ITMVisualisationEngine.h:277-281
  const float outRes = (0.8f * angle + 0.2f) * 255.0f;
  const unsigned char g = (unsigned char)outRes;
  return make_uchar4(g, g, g, g);
}

DEV uchar4 draw_colour(const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, float px, float py, float pz) {
  float c0, c1, c2;
  color_interp(voxels, table, nb, px, py, pz, c0, c1, c2);
  return make_uchar4((unsigned char)(c0 * 255.0f), (unsigned char)(c1 * 255.0f), (unsigned char)(c2 * 255.0f), 255);
}

// RenderImage_common kernels — Vis_CUDA.cu:735-886. WeightRenderingParams(1.0,false,maxW,2) (:303-311).
__global__ void __launch_bounds__(256)
k_shade(const float4 *__restrict__ rays, const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, int w, int h,
        Mat4 M, float lx, float ly, float lz, float voxelSize, int maxW, uchar4 *outChar, float *outFloat, int type) {
  const int tilesX = (w + 7) >> 3, tilesY = (h + 3) >> 2;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warpGlobal >= tilesX * tilesY) return;
  const int lane = threadIdx.x & 31;
  const int x = (warpGlobal % tilesX) * 8 + (lane & 7), y = (warpGlobal / tilesX) * 4 + (lane >> 3);
  if (x >= w || y >= h) return;
  const int locId = x + y * w;
  const float4 pr = rays[locId];
  bool found = pr.w > 0;
  float nx = 0, ny = 0, nz = 0, angle = 0;
  if (type == B200_RENDER_DEPTH_MAP) {
    if (found) {   // drawPixelDepth :304-320
      Vec4 pc = m4v4(M, pr.x * voxelSize, pr.y * voxelSize, pr.z * voxelSize, 1.0f);
      outFloat[locId] = pc.z / pc.w;
    } else outFloat[locId] = 0.0f;
    return;
  }
  normal_and_angle_sdf(found, pr.x, pr.y, pr.z, voxels, table, nb, lx, ly, lz, nx, ny, nz, angle);
  switch (type) {
  case B200_RENDER_COLOUR_FROM_VOLUME:
    outChar[locId] = found ? draw_colour(voxels, table, nb, pr.x, pr.y, pr.z) : make_uchar4(0, 0, 0, 255);
    break;
  case B200_RENDER_COLOUR_FROM_NORMAL:
    if (found) {   // drawPixelNormal :283-288 leaves alpha untouched
      uchar4 o = outChar[locId];
      o.x = (unsigned char)((0.3f + (-nx + 1.0f) * 0.35f) * 255.0f);
      o.y = (unsigned char)((0.3f + (-ny + 1.0f) * 0.35f) * 255.0f);
      o.z = (unsigned char)((0.3f + (-nz + 1.0f) * 0.35f) * 255.0f);
      outChar[locId] = o;
    } else outChar[locId] = make_uchar4(0, 0, 0, 0);
    break;
  case B200_RENDER_COLOUR_FROM_DEPTH_WEIGHT:
    if (found) {   // drawPixelWeight :322-383
      uchar4 dest = draw_colour(voxels, table, nb, pr.x, pr.y, pr.z);
      IdxCache c; cache_init(c);
      const int ix = (int)pr.x, iy = (int)pr.y, iz = (int)pr.z;
      const int vi = voxel_index(table, nb, ix, iy, iz, c);
      const int wd = vi >= 0 ? (int)((__ldg(reinterpret_cast<const unsigned *>(voxels + vi)) >> 16) & 0xff) : 0;
      const unsigned char intensity = (unsigned char)(255.0f * (((float)wd) / maxW));
      const int blockIdx_ = find_block<false>(table, nb, floordiv8(ix), floordiv8(iy), floordiv8(iz));
      uchar4 ov;
      if (blockIdx_ < 0) ov = make_uchar4(255, 255, 255, 255);
      else if (wd <= 2) ov = make_uchar4(255, 0, 0, 255);
      else if (wd == maxW) ov = make_uchar4(0, 0, 255, 255);
      else ov = make_uchar4(intensity, intensity, intensity, 255);
      const float overlayWeight = 1.0f;
      const float a = (float)(1.0 - overlayWeight);
      dest.x = (unsigned char)(((float)dest.x * a) + ((float)ov.x * overlayWeight));
      dest.y = (unsigned char)(((float)dest.y * a) + ((float)ov.y * overlayWeight));
      dest.z = (unsigned char)(((float)dest.z * a) + ((float)ov.z * overlayWeight));
      dest.w = (unsigned char)(((float)dest.w * a) + ((float)ov.w * overlayWeight));
      outChar[locId] = dest;
    } else outChar[locId] = make_uchar4(0, 0, 0, 0);
    break;
  default:
    outChar[locId] = found ? grey_px(angle) : make_uchar4(0, 0, 0, 0);
    break;
  }
}

void launch_shade(b200_engine *e, const SceneRef &s, const Mat4 &M, const Mat4 &invM, int w, int h, float voxelSize, int maxW,
                  const b200_vec4f *rays, b200_vec4u *outChar, float *outFloat, int type) {
  const int tiles = ((w + 7) / 8) * ((h + 3) / 4);
  k_shade<<<(tiles + 7) / 8, 256, 0, e->stream>>>((const float4 *)rays, s.voxels, s.hash, s.numBuckets, w, h, M, -invM.m[8], -invM.m[9],
                                                 -invM.m[10], voxelSize, maxW, (uchar4 *)outChar, outFloat, type);
  e->launches++;
}

// ------------------------------------------------------------------------------------------------
// image-space normals: processPixelICP<true> / processPixelForwardRender<true>
// ------------------------------------------------------------------------------------------------
// computeNormalAndAngle<useSmoothing=true> — DA/ITMVisualisationEngine.h:212-275
DEV void normal_and_angle_img(bool &found, int x, int y, const float4 *__restrict__ pr, float lx, float ly, float lz, float voxelSize,
                              int w, int h, float &nx, float &ny, float &nz, float &angle) {
  if (!found) return;
  if (y <= 2 || y >= h - 3 || x <= 2 || x >= w - 3) { found = false; return; }
  float4 xp1 = pr[(x + 2) + y * w], yp1 = pr[x + (y + 2) * w], xm1 = pr[(x - 2) + y * w], ym1 = pr[x + (y - 2) * w];
  float dxx = 0, dxy = 0, dxz = 0, dyx = 0, dyy = 0, dyz = 0;
  bool doPlus1 = false;
  if (xp1.w <= 0 || yp1.w <= 0 || xm1.w <= 0 || ym1.w <= 0) doPlus1 = true;
  else {
    dxx = xp1.x - xm1.x; dxy = xp1.y - xm1.y; dxz = xp1.z - xm1.z;
    dyx = yp1.x - ym1.x; dyy = yp1.y - ym1.y; dyz = yp1.z - ym1.z;
    const float length_diff = maxf_(dxx * dxx + dxy * dxy + dxz * dxz, dyx * dyx + dyy * dyy + dyz * dyz);
    if (length_diff * voxelSize * voxelSize > (0.15f * 0.15f)) doPlus1 = true;
  }
  if (doPlus1) {
    xp1 = pr[(x + 1) + y * w]; yp1 = pr[x + (y + 1) * w]; xm1 = pr[(x - 1) + y * w]; ym1 = pr[x + (y - 1) * w];
    dxx = xp1.x - xm1.x; dxy = xp1.y - xm1.y; dxz = xp1.z - xm1.z;
    dyx = yp1.x - ym1.x; dyy = yp1.y - ym1.y; dyz = yp1.z - ym1.z;
    if (xp1.w <= 0 || yp1.w <= 0 || xm1.w <= 0 || ym1.w <= 0) { found = false; return; }
  }
  nx = -(dxy * dyz - dxz * dyy);
  ny = -(dxz * dyx - dxx * dyz);
  nz = -(dxx * dyy - dxy * dyx);
  const float normScale = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  nx *= normScale; ny *= normScale; nz *= normScale;
  angle = nx * lx + ny * ly + nz * lz;
  if (!(angle > 0.0)) found = false;
}

__global__ void __launch_bounds__(256)
k_icp(const float4 *__restrict__ rays, int w, int h, float voxelSize, float lx, float ly, float lz, uchar4 *outImg, float4 *points,
      float4 *normals) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int locId = x + y * w;
  const float4 p = rays[locId];
  bool found = p.w > 0.0f;
  float nx = 0, ny = 0, nz = 0, angle = 0;
  normal_and_angle_img(found, x, y, rays, lx, ly, lz, voxelSize, w, h, nx, ny, nz, angle);
  if (found) {
    outImg[locId] = grey_px(angle);
    if (points) { points[locId] = make_float4(p.x * voxelSize, p.y * voxelSize, p.z * voxelSize, 1.0f); normals[locId] = make_float4(nx, ny, nz, 0.0f); }
  } else {
    outImg[locId] = make_uchar4(0, 0, 0, 0);
    if (points) { points[locId] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); normals[locId] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); }
  }
}

void launch_icp(b200_engine *e, const Mat4 &invM, int w, int h, float voxelSize, const b200_vec4f *rays, b200_vec4u *outImg,
                b200_vec4f *points, b200_vec4f *normals) {
  dim3 grid((w + 31) / 32, (h + 7) / 8);
  trace_begin(e, e->stream, "k_icp");
  k_icp<<<grid, 256, 0, e->stream>>>((const float4 *)rays, w, h, voxelSize, -invM.m[8], -invM.m[9], -invM.m[10], (uchar4 *)outImg,
                                     (float4 *)points, (float4 *)normals);
  trace_end(e, e->stream);
  e->launches++;
}

// ------------------------------------------------------------------------------------------------
// forward render (approximate raycast path; dead under DynSLAM settings, kept for the interface)
// Canonical order (oracle): raster order; the forward splat keeps the LAST source pixel that lands
// on a target pixel -> atomicMax on the source index, then a gather.
// ------------------------------------------------------------------------------------------------
__global__ void k_fwd_splat(const float4 *__restrict__ rays, int w, int h, Mat4 M, float p0, float p1, float p2, float p3, float voxelSize,
                            int *winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const float4 px = rays[i];
  // forwardProjectPixel(pixel * voxelSize, ...) — DA/ITMVisualisationEngine.h:181-194
  Vec4 q = m4v4(M, px.x * voxelSize, px.y * voxelSize, px.z * voxelSize, 1.0f);
  const float ix = p0 * q.x / q.z + p2, iy = p1 * q.y / q.z + p3;
  if ((ix < 0) || (ix > w - 1) || (iy < 0) || (iy > h - 1)) return;
  atomicMax(&winner[(int)(ix + 0.5f) + (int)(iy + 0.5f) * w], i);
}

#define FWD_TILE 1024
__global__ void __launch_bounds__(256)
k_fwd_gather_missing(const float4 *__restrict__ rays, const int *__restrict__ winner, float4 *fwd, const float *__restrict__ depth,
                     const float2 *__restrict__ minmax, int w, int h, int *missing, DevCounters *ctr, unsigned long long *scanDesc,
                     unsigned gen) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const int n = w * h, noTiles = (n + FWD_TILE - 1) / FWD_TILE;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    unsigned mask = 0;
    const int first = tile * FWD_TILE + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int locId = first + k;
      if (locId >= n) continue;
      const int src = winner[locId];
      const float4 fp = src >= 0 ? rays[src] : make_float4(0, 0, 0, 0);
      fwd[locId] = fp;
      const int y = locId / w, x = locId - y * w;
      const int locId2 = (int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w;
      const float2 mm = minmax[locId2];
      const float d = depth[locId];
      if ((fp.w <= 0) && ((fp.x == 0 && fp.y == 0 && fp.z == 0) || (d > 0)) && (mm.x < mm.y)) mask |= 1u << k;
    }
    unsigned total;
    unsigned local = block_exclusive_scan(__popc(mask), sm, &total);
    if (threadIdx.x < 32) {
      unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) { tileBase = ex; if (tile == noTiles - 1) ctr->noFwdMissing = (int)(ex + total); }
    }
    __syncthreads();
    unsigned o = tileBase + local;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (mask & (1u << k)) missing[o++] = first + k;
    __syncthreads();
  }
}

__global__ void k_fwd_raycast_missing(float4 *fwd, const int *__restrict__ missing, const DevCounters *ctr, const b200_voxel *__restrict__ voxels,
                                      const b200_hash_entry *__restrict__ table, int nb, int w, Mat4 invM, float fx, float fy, float cxp,
                                      float cyp, float voxelSize, float mu, const float2 *__restrict__ minmax) {
  const int n = ctr->noFwdMissing;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int locId = missing[i];
    const int y = locId / w, x = locId - y * w;
    const int locId2 = (int)floorf((float)x / B200_MINMAX_SUBSAMPLE) + (int)floorf((float)y / B200_MINMAX_SUBSAMPLE) * w;
    float4 o;
    cast_ray(o, x, y, voxels, table, nb, invM, 1.0f / fx, 1.0f / fy, cxp, cyp, 1.0f / voxelSize, mu, minmax[locId2]);
    fwd[locId] = o;
  }
}

void launch_forward_render(b200_engine *e, const SceneRef &s, const FrameGeom &g, const float *depth, const b200_vec2f *minmax,
                           const b200_vec4f *rays, b200_vec4f *fwd, int *missing, b200_vec4u *outImg) {
  cudaStream_t st = e->stream;
  const int n = g.w * g.h;
  int *winner = reinterpret_cast<int *>(e->d_tileCounts);   // per target pixel: last source pixel landing on it
  cudaMemsetAsync(winner, 0xff, sizeof(int) * (size_t)n, st);   // -1 = none
  k_fwd_splat<<<(n + 255) / 256, 256, 0, st>>>((const float4 *)rays, g.w, g.h, g.M_d, g.proj_d[0], g.proj_d[1], g.proj_d[2], g.proj_d[3],
                                              g.voxelSize, winner);
  const int noTiles = (n + FWD_TILE - 1) / FWD_TILE;
  k_fwd_gather_missing<<<persistent_grid(e, 4, noTiles), 256, 0, st>>>((const float4 *)rays, winner, (float4 *)fwd, depth,
                                                                      (const float2 *)minmax, g.w, g.h, missing, e->d_ctr, e->d_scanDesc,
                                                                      ++e->scanGen);
  k_fwd_raycast_missing<<<e->smCount * 4, 128, 0, st>>>((float4 *)fwd, missing, e->d_ctr, s.voxels, s.hash, s.numBuckets, g.w, g.invM_d,
                                                       g.proj_d[0], g.proj_d[1], g.proj_d[2], g.proj_d[3], g.voxelSize, g.mu,
                                                       (const float2 *)minmax);
  dim3 grid((g.w + 31) / 32, (g.h + 7) / 8);
  trace_begin(e, st, "k_icp");
  k_icp<<<grid, 256, 0, st>>>((const float4 *)fwd, g.w, g.h, g.voxelSize, -g.invM_d.m[8], -g.invM_d.m[9], -g.invM_d.m[10], (uchar4 *)outImg,
                              nullptr, nullptr);
  trace_end(e, st);
  e->launches += 4;
}

// ------------------------------------------------------------------------------------------------
// point cloud for the colour tracker (dead under DynSLAM settings, kept for the interface);
// canonical order of the compacted cloud: raster order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_point_cloud(const float4 *__restrict__ rays, const b200_voxel *__restrict__ voxels, const b200_hash_entry *__restrict__ table, int nb, int w,
              int h, float lx, float ly, float lz, float voxelSize, int skipPoints, uchar4 *outImg, float4 *locations, float4 *colours,
              DevCounters *ctr, unsigned long long *scanDesc, unsigned gen) {
  __shared__ unsigned sm[33];
  __shared__ unsigned tileBase;
  const int n = w * h, noTiles = (n + 255) / 256;
  for (int tile = blockIdx.x; tile < noTiles; tile += gridDim.x) {
    const int locId = tile * 256 + threadIdx.x;
    bool found = false; float4 pr = make_float4(0, 0, 0, 0);
    if (locId < n) {
      const int y = locId / w, x = locId - y * w;
      pr = rays[locId];
      found = pr.w > 0;
      float nx = 0, ny = 0, nz = 0, angle = 0;
      normal_and_angle_sdf(found, pr.x, pr.y, pr.z, voxels, table, nb, lx, ly, lz, nx, ny, nz, angle);
      outImg[locId] = found ? grey_px(angle) : make_uchar4(0, 0, 0, 0);
      if (skipPoints && ((x % 2 == 0) || (y % 2 == 0))) found = false;
    }
    unsigned total;
    unsigned local = block_exclusive_scan(found ? 1u : 0u, sm, &total);
    if (threadIdx.x < 32) {
      unsigned ex = scan_lookback(scanDesc, gen, tile, total);
      if (threadIdx.x == 0) { tileBase = ex; if (tile == noTiles - 1) ctr->noTotalPoints = ex + total; }
    }
    __syncthreads();
    if (found) {
      const unsigned o = tileBase + local;
      float c0, c1, c2;
      color_interp(voxels, table, nb, pr.x, pr.y, pr.z, c0, c1, c2);
      float4 tmp = make_float4(c0, c1, c2, 1.0f);
      tmp.x /= tmp.w; tmp.y /= tmp.w; tmp.z /= tmp.w;
      colours[o] = tmp;
      locations[o] = make_float4(pr.x * voxelSize, pr.y * voxelSize, pr.z * voxelSize, 1.0f);
    }
    __syncthreads();
  }
}

void launch_point_cloud(b200_engine *e, const SceneRef &s, const Mat4 &invM, int w, int h, float voxelSize, int skipPoints,
                        const b200_vec4f *rays, b200_vec4u *outImg, b200_vec4f *locations, b200_vec4f *colours) {
  const int noTiles = (w * h + 255) / 256;
  k_point_cloud<<<persistent_grid(e, 2, noTiles), 256, 0, e->stream>>>((const float4 *)rays, s.voxels, s.hash, s.numBuckets, w, h,
                                                                      -invM.m[8], -invM.m[9], -invM.m[10], voxelSize, skipPoints,
                                                                      (uchar4 *)outImg, (float4 *)locations, (float4 *)colours, e->d_ctr,
                                                                      e->d_scanDesc, ++e->scanGen);
  e->launches++;
}
