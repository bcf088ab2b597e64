// frames.cu — the two per-frame image stages either side of the volumes.
//
// (1) Instance frame splitting. The reference downloads the whole frame, cuts every detection out on the CPU
//     (ProcessSilhouette_CPU / RemoveSilhouette_CPU, DS/InstRecLib/InstanceReconstructor.cpp:59-170, driven by
//     InstanceReconstructor::ProcessSilhouette, :226-285) and uploads the main frame and one full frame per car
//     again (:180-181, :196-197, :262-263); its author asks for a CUDA version (:57, :66-72). Here the frame
//     stays on the device: ONE kernel visits every pixel once, walks the detections in the reference's order
//     (a later detection sees the pixels an earlier one blanked) and writes all instance frames and the main
//     frame. Only the bounding-box-sized masks cross PCIe.
// (2) Compositing of the per-volume renders (CompositeDepth :850-869, CompositeColor :873-905,
//     CompositeInstances :932-987): the consumer of the multi-GPU gather. One kernel dims the
//     background and z-composites every instance layer in order.
// Byte/integer work, compared bit for bit with oracle/frames_oracle.c.
#include "engine.h"

namespace {

__device__ __forceinline__ bool mask_hit(const b200_mask &m, int x, int y) {
  if (x < m.x0 || x > m.x1 || y < m.y0 || y > m.y1) return false;       // rows/cols of the box that fall outside the frame never match a pixel
  return m.d_data[(size_t)(y - m.y0) * (m.x1 - m.x0 + 1) + (x - m.x0)] == 1;
}

constexpr int SIL_MAX_OPS = 24;
struct SilOps { b200_silhouette_op op[SIL_MAX_OPS]; int n; };

__global__ void k_process_silhouettes(uchar4 *__restrict__ rgb, float *__restrict__ depth, int w, int h, const __grid_constant__ SilOps ops) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int i = x + y * w;
  uchar4 c = rgb[i];
  float d = depth[i];
  bool changed = false;
  for (int k = 0; k < ops.n; ++k) {
    const b200_silhouette_op &o = ops.op[k];
    if (o.action == 2) {
      // ProcessSilhouette_CPU: destination = white / depth 0 (the two memsets), the silhouette copied in
      uchar4 oc = make_uchar4(255, 255, 255, 255);
      float od = 0.0f;
      if (mask_hit(o.copy_mask, x, y)) { oc = c; od = d; }
      ((uchar4 *)o.d_dest_rgb)[i] = oc;
      o.d_dest_depth[i] = od;
    }
    if (o.action != 0 && mask_hit(o.delete_mask, x, y)) {   // RemoveSilhouette_CPU
      c = make_uchar4(0, 0, 0, 0); d = 0.0f; changed = true;
    }
  }
  if (changed) { rgb[i] = c; depth[i] = d; }
}

// CompositeDepth (InstanceReconstructor.cpp:850-869)
__global__ void k_composite_depth(float *__restrict__ t, const float *__restrict__ s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tv = t[i], sv = s[i];
  if (tv == 0) t[i] = sv;
  else if (sv != 0) t[i] = fminf(tv, sv);
}

__device__ __forceinline__ unsigned char boost(unsigned char s, double col_strength, int tint, float tint_strength) {
  // static_cast<uchar>(min(255.0, s * col_strength + tint * tint_strength)): the tint product is a float product
  const double v = (double)s * col_strength + (double)((float)tint * tint_strength);
  return (unsigned char)(v < 255.0 ? v : 255.0);
}

constexpr int CMP_MAX_LAYERS = 16;
struct Layers { b200_instance_layer l[CMP_MAX_LAYERS]; int n; };

// dim (optional) + CompositeColor for every layer in order (InstanceReconstructor.cpp:873-905, :944-985)
// (bgcol / bgdep: where the background is read from — the target itself, or the render it would otherwise be copied from first)
__global__ void k_composite_layers(uchar4 *tcol, float *tdep, const uchar4 *bgcol, const float *bgdep, int n, const __grid_constant__ Layers layers,
                                   int dim, float dim_factor, float tint_strength) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uchar4 c = bgcol[i];
  float d = bgdep[i];
  if (dim) {
    const double f = 1.0 - (double)dim_factor;
    c.x = (unsigned char)((double)c.x * f); c.y = (unsigned char)((double)c.y * f); c.z = (unsigned char)((double)c.z * f);
  }
  const double col_strength = 1.0 + (double)0.50f - (double)tint_strength;
  for (int k = 0; k < layers.n; ++k) {
    const b200_instance_layer &l = layers.l[k];
    const float sd = l.d_depth[i];
    if (sd != 0 && (d == 0 || d > sd)) {
      const uchar4 sc = ((const uchar4 *)l.d_color)[i];
      d = sd;
      c.x = boost(sc.x, col_strength, l.tint[0], tint_strength);
      c.y = boost(sc.y, col_strength, l.tint[1], tint_strength);
      c.z = boost(sc.z, col_strength, l.tint[2], tint_strength);
    }
  }
  tcol[i] = c;
  tdep[i] = d;
}

}  // namespace

void launch_process_silhouettes(b200_engine *e, b200_vec4u *rgb, float *depth, int w, int h, const b200_silhouette_op *ops, int n) {
  dim3 grid((w + 31) / 32, (h + 7) / 8);
  for (int base = 0; base < n; base += SIL_MAX_OPS) {       // chunks keep the reference's order: the frame is updated in place
    SilOps so;
    so.n = n - base < SIL_MAX_OPS ? n - base : SIL_MAX_OPS;
    for (int k = 0; k < so.n; ++k) so.op[k] = ops[base + k];
    trace_begin(e, e->stream, "k_process_silhouettes");
    k_process_silhouettes<<<grid, 256, 0, e->stream>>>((uchar4 *)rgb, depth, w, h, so);
    trace_end(e, e->stream);
    e->launches++;
  }
}

void launch_composite_depth(b200_engine *e, float *target, const float *source, int n) {
  k_composite_depth<<<(n + 255) / 256, 256, 0, e->stream>>>(target, source, n);
  e->launches++;
}

void launch_composite_layers(b200_engine *e, b200_vec4u *tcol, float *tdep, int n, const b200_instance_layer *layers, int nLayers,
                             bool dim, float dimFactor, float tintStrength, cudaStream_t other, const b200_vec4u *bgcol, const float *bgdep) {
  // other != 0: launched on that stream by the exchange's own host thread (comm.cu) — the engine's state is not touched
  const cudaStream_t st = other ? other : e->stream;
  int base = 0;
  do {
    Layers L;
    L.n = nLayers - base < CMP_MAX_LAYERS ? nLayers - base : CMP_MAX_LAYERS;
    for (int k = 0; k < L.n; ++k) L.l[k] = layers[base + k];
    if (!other) trace_begin(e, st, "k_composite_layers");
    // the first chunk reads the background from bgcol / bgdep when given (the exchange composites rank 0's render without
    // copying it into the destination first); later chunks continue in place
    const uchar4 *bc = (base == 0 && bgcol) ? (const uchar4 *)bgcol : (const uchar4 *)tcol;
    const float *bd = (base == 0 && bgdep) ? bgdep : tdep;
    k_composite_layers<<<(n + 255) / 256, 256, 0, st>>>((uchar4 *)tcol, tdep, bc, bd, n, L, (dim && base == 0) ? 1 : 0, dimFactor, tintStrength);
    if (!other) { trace_end(e, st); e->launches++; }
    base += CMP_MAX_LAYERS;
  } while (base < nLayers);
}
