// evalc.cu — the evaluation consumer of the float raycast depth for sm_100a (second half).
//
// Replaces the serial loop of Evaluation::EvaluateDepth (reference DS/Evaluation/Evaluation.cpp:241-304) with ProjectLidar
// (:214-238) and, per callback, EvaluationCallback::ProcessLidarPoint / ComputeAccuracy (DS/Evaluation/EvaluationCallback.cpp:15-103;
// static / dynamic split: SegmentedEvaluationCallback.cpp:8-41). The reference walks ~120 k LIDAR returns per frame and calls
// 14 virtual callbacks on each, after copying the rendered depth preview to the host; here the depth stays on the device, one
// thread takes one return through all callbacks, and only the counters travel.
//
// Arithmetic: the projection is the reference's double-precision chain (two matrix products and two perspective divisions; a row
// is the left-to-right sum of its four products — Eigen leaves the order open, oracle/eval_oracle.c states the same choice), the
// rest its float / double mixture expression by expression, so every count equals the oracle's. FP64 runs at full rate on
// sm_100a and a frame has 10^5 points: the kernel is launch-latency sized, not a roofline kernel.
// Counting: a CTA accumulates in shared memory (9 counters x 2 classes x callbacks) and adds its totals to global memory once.
#include "engine.h"

namespace {

struct EvalArgs {
  b200_eval_params p;
  b200_eval_callback cb[B200_EVAL_MAX_CALLBACKS];
  int nCallbacks, hasDynamic;
};

// counter layout per (class, callback): measurement_count, rendered.{missing, error, correct, missing_separate}, input.{...}
constexpr int EV_PER = 9;
constexpr int EV_SUMMARY = 4;   // valid, epi_errors, negative_disparities, skipped

DEV void mat_vec4(const double *m, int rows, const double *x, double *y) {
  for (int r = 0; r < rows; ++r)
    y[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[r], x[0]), __dmul_rn(m[rows + r], x[1])), __dmul_rn(m[2 * rows + r], x[2])), __dmul_rn(m[3 * rows + r], x[3]));
}

__global__ void __launch_bounds__(256)
k_evaluate_depth(const __grid_constant__ EvalArgs a, const float4 *__restrict__ lidar, int n, const float *__restrict__ rendered,
                 const short *__restrict__ inputMm, const uint8_t *__restrict__ association, unsigned long long *counters) {
  __shared__ unsigned sm[2 * B200_EVAL_MAX_CALLBACKS * EV_PER + EV_SUMMARY];
  const int nCounters = 2 * a.nCallbacks * EV_PER + EV_SUMMARY;
  for (int i = threadIdx.x; i < nCounters; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  unsigned *summary = sm + 2 * a.nCallbacks * EV_PER;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 pt = __ldg(lidar + i);
    // ProjectLidar (Evaluation.cpp:214-238)
    const double velo[4] = {(double)pt.x, (double)pt.y, (double)pt.z, 1.0};
    double cam[4];
    mat_vec4(a.p.velo_to_cam, 4, velo, cam);
    const double w = cam[3];
    for (int k = 0; k < 4; ++k) cam[k] = __ddiv_rn(cam[k], w);
    const double veloZ = cam[2];
    if (veloZ < (double)a.p.min_depth_m || veloZ > (double)a.p.max_depth_m) continue;
    double left[3], right[3];
    mat_vec4(a.p.proj_left, 3, cam, left);
    mat_vec4(a.p.proj_right, 3, cam, right);
    const double wl = left[2], wr = right[2];
    for (int k = 0; k < 3; ++k) { left[k] = __ddiv_rn(left[k], wl); right[k] = __ddiv_rn(right[k], wr); }
    // EvaluateDepth (:253-296)
    const int rowLeft = (int)round(left[1]), colLeft = (int)round(left[0]), rowRight = (int)round(right[1]);
    if (colLeft < 0 || colLeft >= a.p.frame_width || rowLeft < 0 || rowLeft >= a.p.frame_height) continue;
    if (rowLeft != rowRight) {
      const float fdelta = (float)__dsub_rn(left[1], right[1]);
      if ((double)fabsf(fdelta) > 1.2) atomicAdd(&summary[1], 1u);
    }
    const float lidarDisp = (float)__dsub_rn(left[0], right[0]);
    if (lidarDisp < 0.0f) { atomicAdd(&summary[2], 1u); continue; }      // the reference throws here; the host turns the count into an error
    atomicAdd(&summary[0], 1u);
    const int idx = rowLeft * a.p.frame_width + colLeft;
    const float renderedM = __ldg(rendered + idx);
    const float inputM = __fdiv_rn((float)(int)__ldg(inputMm + idx), 1000.0f);
    const float bf = __fmul_rn(a.p.baseline_m, a.p.left_focal_length_px);
    const float renderedDisp = __fdiv_rn(bf, renderedM), inputDisp = __fdiv_rn(bf, inputM);
    const int cls = association ? (int)__ldg(association + idx) : B200_EVAL_STATIC;
    if (!(cls == B200_EVAL_STATIC || (cls == B200_EVAL_DYNAMIC && a.hasDynamic))) { atomicAdd(&summary[3], 1u); continue; }
    // ComputeAccuracy (EvaluationCallback.cpp:48-103): what does not depend on the callback is computed once
    const float renDelta = fabsf(__fsub_rn(renderedDisp, lidarDisp)), inDelta = fabsf(__fsub_rn(inputDisp, lidarDisp));
    const bool missingInput = fabs((double)inputM) < 1e-5, missingRendered = fabs((double)renderedM) < 1e-5;
    const double fivePercent = __dmul_rn(0.05, (double)lidarDisp);
    const bool inOver5 = (double)inDelta > fivePercent, renOver5 = (double)renDelta > fivePercent;
    unsigned *base = sm + (cls == B200_EVAL_DYNAMIC ? a.nCallbacks * EV_PER : 0);
    for (int c = 0; c < a.nCallbacks; ++c) {
      unsigned *ctr = base + c * EV_PER;
      const float dm = a.cb[c].delta_max;
      const bool kitti = a.cb[c].kitti_style != 0;
      atomicAdd(&ctr[0], 1u);
      if (missingRendered) atomicAdd(&ctr[4], 1u);
      if (missingInput) atomicAdd(&ctr[8], 1u);
      if (a.cb[c].compare_on_intersection && (missingInput || missingRendered)) {
        atomicAdd(&ctr[1], 1u); atomicAdd(&ctr[5], 1u);
      } else {
        if (missingInput) atomicAdd(&ctr[5], 1u);
        else atomicAdd(&ctr[(kitti ? (inDelta > dm && inOver5) : (inDelta > dm)) ? 6 : 7], 1u);
        if (missingRendered) atomicAdd(&ctr[1], 1u);
        else atomicAdd(&ctr[(kitti ? (renDelta > dm && renOver5) : (renDelta > dm)) ? 2 : 3], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nCounters; i += blockDim.x)
    if (sm[i]) atomicAdd(&counters[i], (unsigned long long)sm[i]);
}

}  // namespace

int eval_counter_words(int nCallbacks) { return 2 * nCallbacks * EV_PER + EV_SUMMARY; }

void launch_evaluate_depth(b200_engine *e, const b200_eval_params *p, const b200_eval_callback *cb, int nCallbacks, int hasDynamic,
                           const float *lidar, int n, const float *rendered, const int16_t *inputMm, const uint8_t *association,
                           unsigned long long *counters) {
  EvalArgs a;
  a.p = *p;
  for (int c = 0; c < nCallbacks; ++c) a.cb[c] = cb[c];
  for (int c = nCallbacks; c < B200_EVAL_MAX_CALLBACKS; ++c) a.cb[c] = b200_eval_callback{0.0f, 0, 0};
  a.nCallbacks = nCallbacks; a.hasDynamic = hasDynamic;
  cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * (size_t)eval_counter_words(nCallbacks), e->stream);
  int ctas = (n + 255) / 256;
  if (ctas > e->smCount * 4) ctas = e->smCount * 4;
  if (ctas < 1) ctas = 1;
  trace_begin(e, e->stream, "k_evaluate_depth");
  k_evaluate_depth<<<ctas, 256, 0, e->stream>>>(a, reinterpret_cast<const float4 *>(lidar), n, rendered, (const short *)inputMm, association, counters);
  trace_end(e, e->stream);
  e->launches++;
}
