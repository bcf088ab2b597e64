// view.cu — the view builder that feeds the fusion path.
//
// Replaces ITMViewBuilder_CUDA::UpdateView (Engine/DeviceSpecific/CUDA/ITMViewBuilder_CUDA.cu:33-84):
// raw short depth -> float metres (convertDepthAffineToFloat / convertDisparityToDepth,
// DeviceAgnostic/ITMViewBuilder.h:7-28), five passes of the 5x5 bilateral filter (filterDepth, :30-56,
// ping-ponging view->depth <-> floatImage, then a device-to-device copy back) and, for the weighted-ICP
// tracker, ComputeNormalAndWeights (:59-114).
//
// B200 design. The filter is bound by arithmetic (25 exp() per pixel and pass), not by memory: a first version that
// fused all five passes over a shared-memory tile with a 10-pixel halo measured 2x SLOWER (123 us) than five plain passes (62 us),
// because the halo recomputation (+42 %) costs more than the L2 round trips it saves (1.9 MB images, L2-resident).
// What is fused is everything that is not arithmetic: the raw->float conversion happens while pass 1 stages its tile,
// the final device-to-device copy disappears by letting view->depth play the role of floatImage (passes 1, 3, 5 write
// it) and an engine-owned scratch image the role of view->depth, so UpdateView is 5 launches instead of 7 and never
// materialises the converted image. Each pass stages a (32+4)x(16+4) tile in shared memory once, two pixels per thread.
//
// Border semantics are the CUDA reference's, not the CPU twin's (which clears the target of every pass,
// CPU/ITMViewBuilder_CPU.cpp:116-127): filterDepth_device leaves the two outermost rows/columns of its
// TARGET untouched (ITMViewBuilder_CUDA.cu:196-209), so
//   * floatImage's border is the zero MemoryBlock's constructor put there and is never written
//     (passes 1, 3, 5 write floatImage) -> the final depth image has a 2-pixel border of 0,
//   * view->depth's border still holds the converted raw values when passes 3 and 5 read it,
//   * those border values ARE read as taps by the neighbouring inner pixels (0 is not < 0).
//
// Arithmetic: conversions, masks and borders are bit-exact; the bilateral weights use MUFU.EX2 on pre-scaled arguments and
// the final quotient a fast division (see filter_depth_two) — like the reference's own CUDA build, which compiles this file
// with --use_fast_math — so the filtered depth is compared with the oracle (host libm) within a stated 2e-5 relative
// tolerance (tests/test_gpu_view.py), not bit for bit. acos() in ComputeNormalAndWeights is CUDA's acosf.
#include "engine.h"

namespace {

constexpr float MEAN_SIGMA_L = 1.2232f;   // DeviceAgnostic/ITMViewBuilder.h:30

// convertDepthAffineToFloat (DA/ITMViewBuilder.h:22-28)
__device__ __forceinline__ float convert_affine(short d, float p0, float p1) {
  return ((d <= 0) || (d > 32000)) ? -1.0f : (float)d * p0 + p1;
}

// convertDisparityToDepth (DA/ITMViewBuilder.h:7-20)
__device__ __forceinline__ float convert_disparity(short disparity, float p0, float p1, float fx) {
  const float disparity_tmp = p0 - (float)disparity;
  float depth;
  if (disparity_tmp == 0) depth = 0.0f;
  else depth = 8.0f * p1 * fx / disparity_tmp;
  return (depth > 0) ? depth : -1.0f;
}

// The 25 bilateral weights exp(-0.5 (a sigma_L^2 + dz^2 sigma_z^2)) are evaluated as 2^(k_a + dz^2 * kz) with the constants
// pre-scaled by log2(e): one FFMA and one MUFU.EX2 per tap (the reference's own CUDA build does the same through
// --use_fast_math, ITMLib/CMakeLists.txt:230). Against the exactly rounded filter (host libm) the weights are off by a few
// ulp, the filtered depth by <= 2e-5 relative (tests/test_gpu_view.py states and checks the bound); conversions, the
// invalid mask and the border semantics stay bit-exact. Round 1 used an expf-grade exponential (<= 2 ulp, 16 instructions per
// tap): 14.7 us per pass, slower than the reference's kernel; this form is 7 instructions per tap.
__device__ __forceinline__ float ex2_approx(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}

// filterDepth (DA/ITMViewBuilder.h:31-56) for TWO vertically adjacent pixels at c[0] and c[stride] (rows `stride` floats
// apart): the 6 x 5 taps the two windows cover are read once. An invalid tap (tmpz < 0) gets weight 0, as the reference's
// `continue`. Returns the filtered values (-1 where the centre is invalid).
__device__ __forceinline__ void filter_depth_two(const float *c, int stride, float &out0, float &out1) {
  constexpr float L2E = 1.44269504088896341f;
  constexpr float KA = -0.5f * MEAN_SIGMA_L * MEAN_SIGMA_L * L2E;   // per unit of |i| + |j|
  const float z0 = c[0], z1 = c[stride];
  // sigma_z as the reference computes it (IEEE division and sqrt); only the exponential is approximated
  const float sz0 = 1.0f / (0.0012f + 0.0019f * (z0 - 0.4f) * (z0 - 0.4f) + 0.0001f / sqrtf(fmaxf(z0, 1e-6f)) * 0.25f);
  const float sz1 = 1.0f / (0.0012f + 0.0019f * (z1 - 0.4f) * (z1 - 0.4f) + 0.0001f / sqrtf(fmaxf(z1, 1e-6f)) * 0.25f);
  const float kz0 = -0.5f * L2E * sz0 * sz0, kz1 = -0.5f * L2E * sz1 * sz1;
  float f0 = 0.0f, w0 = 0.0f, f1 = 0.0f, w1 = 0.0f;
#pragma unroll
  for (int r = -2; r <= 3; r++) {        // tap row relative to pixel 0; pixel 1 sees it as row r - 1
#pragma unroll
    for (int j = -2; j <= 2; j++) {
      const float t = c[r * stride + j];
      const bool ok = !(t < 0.0f);
      if (r <= 2) {
        const float d = t - z0;
        const int a = (r < 0 ? -r : r) + (j < 0 ? -j : j);
        float w = ex2_approx(__fmaf_rn(d * d, kz0, (float)a * KA));
        w = ok ? w : 0.0f;
        w0 += w; f0 = __fmaf_rn(w, t, f0);
      }
      if (r >= -1) {
        const float d = t - z1;
        const int a = (r - 1 < 0 ? 1 - r : r - 1) + (j < 0 ? -j : j);
        float w = ex2_approx(__fmaf_rn(d * d, kz1, (float)a * KA));
        w = ok ? w : 0.0f;
        w1 += w; f1 = __fmaf_rn(w, t, f1);
      }
    }
  }
  out0 = (z0 < 0.0f) ? -1.0f : __fdividef(f0, w0);
  out1 = (z1 < 0.0f) ? -1.0f : __fdividef(f1, w1);
}

__device__ __forceinline__ bool on_border(int x, int y, int w, int h) { return x < 2 || x >= w - 2 || y < 2 || y >= h - 2; }

// ---- stand-alone stages (interface completeness: ITMViewBuilder::ConvertDisparityToDepth,
// ConvertDepthAffineToFloat, DepthFiltering) ---------------------------------------------------------
__global__ void k_convert(const short *__restrict__ in, float *__restrict__ out, int n, int type, float p0, float p1, float fx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = type == 0 ? convert_disparity(in[i], p0, p1, fx) : convert_affine(in[i], p0, p1);
}

// One DepthFiltering pass. 32x16 output pixels per CTA of 256 threads (two vertically adjacent pixels per thread), the
// (32+4)x(16+4) input neighbourhood staged in shared memory (converted from the raw image on the fly when RAW). Border pixels
// of the target are left untouched (reference semantics) unless zeroBorder, which writes the 0 that floatImage's border
// holds in the reference; borderCopy, if given, receives the staged (converted) value of border pixels — the border
// view->depth keeps through all passes.
constexpr int FP_TW = 32, FP_TH = 16, FP_SW = FP_TW + 4, FP_SH = FP_TH + 4, FP_THREADS = 256;

template <bool RAW>
__global__ void __launch_bounds__(FP_THREADS) k_filter_pass(const short *__restrict__ raw, const float *__restrict__ in,
                                                             float *__restrict__ out, float *__restrict__ borderCopy, int w, int h,
                                                             int type, float p0, float p1, float fx, int zeroBorder) {
  __shared__ float tile[FP_SH * FP_SW];
  const int x0 = blockIdx.x * FP_TW - 2, y0 = blockIdx.y * FP_TH - 2;
  for (int c = threadIdx.x; c < FP_SW * FP_SH; c += FP_THREADS) {
    const int gx = x0 + c % FP_SW, gy = y0 + c / FP_SW;
    float v = 0.0f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      const int g = gx + gy * w;
      v = RAW ? (type == 0 ? convert_disparity(raw[g], p0, p1, fx) : convert_affine(raw[g], p0, p1)) : in[g];
    }
    tile[c] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = (threadIdx.x >> 5) * 2;
  const int x = x0 + 2 + tx, y = y0 + 2 + ty;
  if (x >= w || y >= h) return;
  const float *c = tile + (ty + 2) * FP_SW + tx + 2;
  float r0, r1;
  filter_depth_two(c, FP_SW, r0, r1);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int yy = y + k;
    if (yy >= h) break;
    if (on_border(x, yy, w, h)) {
      if (zeroBorder) out[x + yy * w] = 0.0f;
      if (borderCopy) borderCopy[x + yy * w] = c[k * FP_SW];
    } else {
      out[x + yy * w] = k ? r1 : r0;
    }
  }
}

// ---- ComputeNormalAndWeights (ITMViewBuilder_CUDA.cu:211-227, DA/ITMViewBuilder.h:59-114) -----------
// In-image threads only: the reference's kernel also lets its out-of-image threads (x >= w in the last
// column of 16x16 CTAs) write w=-1 through idx = x + y*w, which aliases pixels of the next row and races
// with their owners; that write is not reproduced (documented deviation).
__global__ void k_normal_weight(const float *__restrict__ depth_in, float4 *__restrict__ normal_out, float *__restrict__ sigmaZ_out,
                                int w, int h, float ix, float iy, float iz, float iw) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int idx = x + y * w;
  if (x < 2 || x > w - 2 || y < 2 || y > h - 2) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; return; }
  const float z = depth_in[idx];
  if (z < 0.0f) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; return; }
  const float zxp = depth_in[(x + 1) + y * w], zyp = depth_in[x + (y + 1) * w];
  const float zxm = depth_in[(x - 1) + y * w], zym = depth_in[x + (y - 1) * w];
  if (zxp <= 0 || zyp <= 0 || zxm <= 0 || zym <= 0) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; return; }
  // unprojected neighbours (the reference multiplies by intrinparam.x/.y, i.e. fx/fy, as written)
  const float xp1_x = zxp * ((x + 1.0f) - iz) * ix, xp1_y = zxp * (y - iw) * iy;
  const float xm1_x = zxm * ((x - 1.0f) - iz) * ix, xm1_y = zxm * (y - iw) * iy;
  const float yp1_x = zyp * (x - iz) * ix, yp1_y = zyp * ((y + 1.0f) - iw) * iy;
  const float ym1_x = zym * (x - iz) * ix, ym1_y = zym * ((y - 1.0f) - iw) * iy;
  const float dxx = xp1_x - xm1_x, dxy = xp1_y - xm1_y, dxz = zxp - zxm;
  const float dyx = yp1_x - ym1_x, dyy = yp1_y - ym1_y, dyz = zyp - zym;
  float nx = (dxy * dyz - dxz * dyy);
  float ny = (dxz * dyx - dxx * dyz);
  float nz = (dxx * dyy - dxy * dyx);
  if (nx == 0.0f && ny == 0 && nz == 0) { normal_out[idx].w = -1.0f; sigmaZ_out[idx] = -1; return; }
  const float norm = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  nx *= norm; ny *= norm; nz *= norm;
  normal_out[idx] = make_float4(nx, ny, nz, 1.0f);
  const float theta = acosf(nz);
  const float theta_diff = theta / (3.1415926535897932384626433832795f * 0.5f - theta);
  sigmaZ_out[idx] = (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * theta_diff * theta_diff);
}

}  // namespace

void launch_view_convert(b200_engine *e, const int16_t *raw, float *out, int w, int h, int type, float p0, float p1, float fx) {
  const int n = w * h;
  k_convert<<<(n + 255) / 256, 256, 0, e->stream>>>((const short *)raw, out, n, type, p0, p1, fx);
  e->launches++;
}

void launch_view_filter_pass(b200_engine *e, const float *in, float *out, int w, int h) {
  dim3 grid((w + FP_TW - 1) / FP_TW, (h + FP_TH - 1) / FP_TH);
  trace_begin(e, e->stream, "k_filter_pass<false>");
  k_filter_pass<false><<<grid, FP_THREADS, 0, e->stream>>>(nullptr, in, out, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 0);
  trace_end(e, e->stream);
  e->launches++;
}

// UpdateView's device part. `scratch` (w*h floats, engine-owned) plays view->depth, `out` plays floatImage:
//   pass 1: raw -> out (+ converted border -> scratch), 2: out -> scratch, 3: scratch -> out, 4: out -> scratch, 5: scratch -> out
void launch_update_view(b200_engine *e, const int16_t *raw, float *out, float *scratch, int w, int h, int type, float p0, float p1,
                        float fx, bool filter) {
  if (!filter) { launch_view_convert(e, raw, out, w, h, type, p0, p1, fx); return; }
  dim3 grid((w + FP_TW - 1) / FP_TW, (h + FP_TH - 1) / FP_TH);
  const int T = FP_THREADS;
  trace_begin(e, e->stream, "k_filter_pass<true>");
  k_filter_pass<true><<<grid, T, 0, e->stream>>>((const short *)raw, nullptr, out, scratch, w, h, type, p0, p1, fx, 1);
  trace_end(e, e->stream);
  trace_begin(e, e->stream, "k_filter_pass<false>");
  k_filter_pass<false><<<grid, T, 0, e->stream>>>(nullptr, out, scratch, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 0);
  trace_end(e, e->stream);
  trace_begin(e, e->stream, "k_filter_pass<false>");
  k_filter_pass<false><<<grid, T, 0, e->stream>>>(nullptr, scratch, out, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 1);
  trace_end(e, e->stream);
  trace_begin(e, e->stream, "k_filter_pass<false>");
  k_filter_pass<false><<<grid, T, 0, e->stream>>>(nullptr, out, scratch, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 0);
  trace_end(e, e->stream);
  trace_begin(e, e->stream, "k_filter_pass<false>");
  k_filter_pass<false><<<grid, T, 0, e->stream>>>(nullptr, scratch, out, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 1);
  trace_end(e, e->stream);
  e->launches += 5;
}

void launch_view_normals(b200_engine *e, const float *depth, b200_vec4f *normal, float *sigmaZ, int w, int h, const float intr[4]) {
  dim3 grid((w + 31) / 32, (h + 7) / 8);
  k_normal_weight<<<grid, 256, 0, e->stream>>>(depth, (float4 *)normal, sigmaZ, w, h, intr[0], intr[1], intr[2], intr[3]);
  e->launches++;
}
