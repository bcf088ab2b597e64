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
// materialises the converted image. Each pass stages a (32+4)x(8+4) tile in shared memory once.
//
// Border semantics are the CUDA reference's, not the CPU twin's (which clears the target of every pass,
// CPU/ITMViewBuilder_CPU.cpp:116-127): filterDepth_device leaves the two outermost rows/columns of its
// TARGET untouched (ITMViewBuilder_CUDA.cu:196-209), so
//   * floatImage's border is the zero MemoryBlock's constructor put there and is never written
//     (passes 1, 3, 5 write floatImage) -> the final depth image has a 2-pixel border of 0,
//   * view->depth's border still holds the converted raw values when passes 3 and 5 read it,
//   * those border values ARE read as taps by the neighbouring inner pixels (0 is not < 0).
//
// Arithmetic: the expressions keep the reference's operation order; the library is compiled without
// contraction, division and sqrt are IEEE. exp() is the CUDA math library's algorithm (<= 2 ulp), acos() its acosf, where the
// oracle uses the host libm, so this file is compared within a stated tolerance (tests/test_gpu_view.py),
// not bit for bit; the reference's own CUDA build uses --use_fast_math here.
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

// exp(x) for x <= 0, the algorithm of the CUDA math library's expf (magic-number rounding of x*log2(e), two-term
// Cody-Waite reduction in ln 2, ex2.approx, exponent splice; <= 2 ulp) without its overflow/underflow special cases:
// the argument is clamped at -86 instead. A bilateral weight below e^-86 = 4e-38 can never change the sums it is added
// to (the centre tap contributes weight 1 and depth >= 1e-3), so the filter output is the same bits either way.
__device__ __forceinline__ float exp_nonpositive(float x) {
  x = fmaxf(x, -86.0f);
  const float t = __fmaf_rn(x, 1.44269504088896341f, 12582912.0f);
  const float j = t - 12582912.0f;
  float r = __fmaf_rn(j, -0.693145751953125f, x);
  r = __fmaf_rn(j, -1.42860682030941723e-06f, r);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(r * 1.44269504088896341f));
  return __int_as_float(__float_as_int(e) + (__float_as_int(t) << 23));
}

// filterDepth (DA/ITMViewBuilder.h:31-56) for the pixel at *c; rows are `stride` floats apart. Branch-free form of the
// reference's `if (tmpz < 0) continue;`: an invalid tap gets weight +0, and x + 0 and x + (0 * tmpz) leave every partial
// sum bit-identical (the sums are >= 0; +0 + -0 = +0 in round-to-nearest).
__device__ __forceinline__ float filter_depth_at(const float *c, int stride) {
  const float z = c[0];
  if (z < 0.0f) return -1.0f;
  const float sigma_z = 1.0f / (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * 0.25f);
  float final_depth = 0.0f, w_sum = 0.0f;
#pragma unroll
  for (int i = -2; i <= 2; i++) {
#pragma unroll
    for (int j = -2; j <= 2; j++) {
      const float tmpz = c[i * stride + j];
      float dz = (tmpz - z); dz *= dz;
      const int a = (i < 0 ? -i : i) + (j < 0 ? -j : j);
      float w = exp_nonpositive(-0.5f * ((float)a * MEAN_SIGMA_L * MEAN_SIGMA_L + dz * sigma_z * sigma_z));
      w = tmpz < 0.0f ? 0.0f : w;
      w_sum += w;
      final_depth += w * tmpz;
    }
  }
  return final_depth / w_sum;
}

__device__ __forceinline__ bool on_border(int x, int y, int w, int h) { return x < 2 || x >= w - 2 || y < 2 || y >= h - 2; }

// ---- stand-alone stages (interface completeness: ITMViewBuilder::ConvertDisparityToDepth,
// ConvertDepthAffineToFloat, DepthFiltering) ---------------------------------------------------------
__global__ void k_convert(const short *__restrict__ in, float *__restrict__ out, int n, int type, float p0, float p1, float fx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = type == 0 ? convert_disparity(in[i], p0, p1, fx) : convert_affine(in[i], p0, p1);
}

// One DepthFiltering pass. 32x8 output pixels per CTA, the (32+4)x(8+4) input neighbourhood staged in shared memory
// (converted from the raw image on the fly when RAW). Border pixels of the target are left untouched (reference
// semantics) unless zeroBorder, which writes the 0 that floatImage's border holds in the reference; borderCopy, if
// given, receives the staged (converted) value of border pixels — the border view->depth keeps through all passes.
constexpr int FP_TW = 32, FP_TH = 8, FP_SW = FP_TW + 4, FP_SH = FP_TH + 4;

template <bool RAW>
__global__ void __launch_bounds__(FP_TW * FP_TH) k_filter_pass(const short *__restrict__ raw, const float *__restrict__ in,
                                                                  float *__restrict__ out, float *__restrict__ borderCopy, int w, int h,
                                                                  int type, float p0, float p1, float fx, int zeroBorder) {
  __shared__ float tile[FP_SH * FP_SW];
  const int x0 = blockIdx.x * FP_TW - 2, y0 = blockIdx.y * FP_TH - 2;
  for (int c = threadIdx.x; c < FP_SW * FP_SH; c += FP_TW * FP_TH) {
    const int gx = x0 + c % FP_SW, gy = y0 + c / FP_SW;
    float v = 0.0f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      const int g = gx + gy * w;
      v = RAW ? (type == 0 ? convert_disparity(raw[g], p0, p1, fx) : convert_affine(raw[g], p0, p1)) : in[g];
    }
    tile[c] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + 2 + tx, y = y0 + 2 + ty;
  if (x >= w || y >= h) return;
  const float *c = tile + (ty + 2) * FP_SW + tx + 2;
  if (on_border(x, y, w, h)) {
    if (zeroBorder) out[x + y * w] = 0.0f;
    if (borderCopy) borderCopy[x + y * w] = c[0];
    return;
  }
  out[x + y * w] = filter_depth_at(c, FP_SW);
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
  k_filter_pass<false><<<grid, FP_TW * FP_TH, 0, e->stream>>>(nullptr, in, out, nullptr, w, h, 1, 0.0f, 0.0f, 0.0f, 0);
  trace_end(e, e->stream);
  e->launches++;
}

// UpdateView's device part. `scratch` (w*h floats, engine-owned) plays view->depth, `out` plays floatImage:
//   pass 1: raw -> out (+ converted border -> scratch), 2: out -> scratch, 3: scratch -> out, 4: out -> scratch, 5: scratch -> out
void launch_update_view(b200_engine *e, const int16_t *raw, float *out, float *scratch, int w, int h, int type, float p0, float p1,
                        float fx, bool filter) {
  if (!filter) { launch_view_convert(e, raw, out, w, h, type, p0, p1, fx); return; }
  dim3 grid((w + FP_TW - 1) / FP_TW, (h + FP_TH - 1) / FP_TH);
  const int T = FP_TW * FP_TH;
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
