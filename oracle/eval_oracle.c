/* eval_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported or called by dynslam_b200/).
 *
 * Serial C restatement of the evaluation consumer of the float raycast depth (second half):
 *   Evaluation::ProjectLidar                          DS/Evaluation/Evaluation.cpp:214-238
 *   Evaluation::EvaluateDepth                         DS/Evaluation/Evaluation.cpp:241-304
 *   EvaluationCallback::ProcessLidarPoint / ComputeAccuracy   DS/Evaluation/EvaluationCallback.cpp:15-103
 *   the static / dynamic split of SegmentedEvaluationCallback::ProcessLidarPoint   DS/Evaluation/SegmentedEvaluationCallback.cpp:8-41
 *     (GetPointAssociation's verdict — which needs the tracker and the segmentation — arrives here as one byte per pixel)
 *
 * PINNED (tests/test_eval_oracle.py): compared count for count with ProjectLidar, EvaluateDepth, ProcessLidarPoint and
 * ComputeAccuracy compiled from the reference files themselves (oracle/build_ref.sh cuts the function bodies out at build
 * time; oracle/ref_eval_driver.cpp supplies class skeletons with the reference's member names, and oracle/stubs/Eigen/Eigen a
 * stand-in for the six Eigen operations they use). Eigen's own evaluation order of the two matrix products is unspecified;
 * here, as in the stand-in, a row is the left-to-right sum of its four products in double.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/b200fusion.h"

static void mat_vec(const double *m, int rows, const double *x, double *y) {   /* column-major rows x 4 */
  for (int r = 0; r < rows; ++r) y[r] = m[0 * rows + r] * x[0] + m[1 * rows + r] * x[1] + m[2 * rows + r] * x[2] + m[3 * rows + r] * x[3];
}

/* Evaluation.cpp:214-238 */
static int project_lidar(const b200_eval_params *p, const float *velodyne_reading, double *left, double *right) {
  double velo_point[4] = {(double)velodyne_reading[0], (double)velodyne_reading[1], (double)velodyne_reading[2], (double)velodyne_reading[3]};
  velo_point[3] = 1.0f;                       /* the reflectance is ignored */
  double cam_point[4];
  mat_vec(p->velo_to_cam, 4, velo_point, cam_point);
  const double w = cam_point[3];
  for (int i = 0; i < 4; ++i) cam_point[i] = cam_point[i] / w;
  const double velo_z = cam_point[2];
  if (velo_z < p->min_depth_m || velo_z > p->max_depth_m) return 0;
  mat_vec(p->proj_left, 3, cam_point, left);
  mat_vec(p->proj_right, 3, cam_point, right);
  const double wl = left[2], wr = right[2];
  for (int i = 0; i < 3; ++i) { left[i] = left[i] / wl; right[i] = right[i] / wr; }
  return 1;
}

/* EvaluationCallback.cpp:48-103 */
static void compute_accuracy(const b200_eval_callback *cb, float rendered_disp, float rendered_depth_m, float input_disp, float input_depth_m,
                             float lidar_disp, b200_eval_stats *input_stats, b200_eval_stats *rendered_stats) {
  const float ren_disp_delta = fabsf(rendered_disp - lidar_disp);
  const float input_disp_delta = fabsf(input_disp - lidar_disp);
  const int missing_input = (fabs((double)input_depth_m) < 1e-5);
  const int missing_rendered = (fabs((double)rendered_depth_m) < 1e-5);
  if (missing_input) input_stats->missing_separate++;
  if (missing_rendered) rendered_stats->missing_separate++;
  if (cb->compare_on_intersection && (missing_input || missing_rendered)) {
    input_stats->missing++;
    rendered_stats->missing++;
  } else {
    if (missing_input) input_stats->missing++;
    else {
      const int is_error = cb->kitti_style ? (input_disp_delta > cb->delta_max && ((double)input_disp_delta > 0.05 * (double)lidar_disp))
                                           : (input_disp_delta > cb->delta_max);
      if (is_error) input_stats->error++; else input_stats->correct++;
    }
    if (missing_rendered) rendered_stats->missing++;
    else {
      const int is_error = cb->kitti_style ? (ren_disp_delta > cb->delta_max && ((double)ren_disp_delta > 0.05 * (double)lidar_disp))
                                           : (ren_disp_delta > cb->delta_max);
      if (is_error) rendered_stats->error++; else rendered_stats->correct++;
    }
  }
}

/* Evaluation.cpp:241-304. Returns 0, or -1 for the reference's "Negative disparity in ground truth." exception (the counts
 * then hold what had been accumulated before the offending point, like the reference's callbacks at the throw). */
int oracle_evaluate_depth(const b200_eval_params *p, const float *lidar_points, int n, const float *rendered_depth, const int16_t *input_depth_mm,
                          const uint8_t *association, const b200_eval_callback *callbacks, int n_callbacks, b200_eval_result *out_static,
                          b200_eval_result *out_dynamic, b200_eval_summary *summary) {
  memset(out_static, 0, sizeof(*out_static) * (size_t)n_callbacks);
  if (out_dynamic) memset(out_dynamic, 0, sizeof(*out_dynamic) * (size_t)n_callbacks);
  memset(summary, 0, sizeof(*summary));
  for (int i = 0; i < n; ++i) {
    double velo_2d_left[3], velo_2d_right[3];
    if (!project_lidar(p, lidar_points + (size_t)i * 4, velo_2d_left, velo_2d_right)) continue;
    const int row_left = (int)round(velo_2d_left[1]);
    const int col_left = (int)round(velo_2d_left[0]);
    const int row_right = (int)round(velo_2d_right[1]);
    if (col_left < 0 || col_left >= p->frame_width || row_left < 0 || row_left >= p->frame_height) continue;
    if (row_left != row_right) {
      const float fdelta = (float)(velo_2d_left[1] - velo_2d_right[1]);
      if (fabsf(fdelta) > 1.2) summary->epi_errors++;          /* float compared with the double 1.2 */
    }
    const float lidar_disp = (float)(velo_2d_left[0] - velo_2d_right[0]);
    if (lidar_disp < 0.0f) { summary->negative_disparities++; return -1; }
    summary->valid_lidar_points++;
    const int idx_in_rendered = row_left * p->frame_width + col_left;
    const float rendered_depth_m = rendered_depth[idx_in_rendered];
    const float input_depth_m = input_depth_mm[idx_in_rendered] / 1000.0f;
    const float rendered_disp = p->baseline_m * p->left_focal_length_px / rendered_depth_m;
    const float input_disp = p->baseline_m * p->left_focal_length_px / input_depth_m;
    /* SegmentedEvaluationCallback.cpp:19-41: kDynamicReconstructed -> the dynamic evaluation, kStaticMap -> the static one,
       kNeither -> skipped. Without an association image every point is static (Evaluation::EvaluateFrame, :150-211). */
    const int a = association ? association[idx_in_rendered] : B200_EVAL_STATIC;
    b200_eval_result *dst = (a == B200_EVAL_STATIC) ? out_static : ((a == B200_EVAL_DYNAMIC && out_dynamic) ? out_dynamic : 0);
    if (!dst) { summary->skipped_lidar_points++; continue; }
    for (int c = 0; c < n_callbacks; ++c) {
      dst[c].measurement_count++;
      compute_accuracy(&callbacks[c], rendered_disp, rendered_depth_m, input_disp, input_depth_m, lidar_disp, &dst[c].input, &dst[c].rendered);
    }
  }
  return 0;
}
