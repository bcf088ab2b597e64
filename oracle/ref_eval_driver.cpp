// ref_eval_driver.cpp — TEST INFRASTRUCTURE ONLY. Pins oracle/eval_oracle.c to the reference's own code.
//
// Evaluation::ProjectLidar / EvaluateDepth (DS/Evaluation/Evaluation.cpp:214-304) and EvaluationCallback::ProcessLidarPoint /
// ComputeAccuracy (DS/Evaluation/EvaluationCallback.cpp:15-28, :48-103) live in translation units that need Eigen, OpenCV,
// Pangolin and the whole DynSlam class (none buildable here). Their bodies only need: six Eigen operations
// (oracle/stubs/Eigen/Eigen, ours), a 16-bit matrix with at<short>() (oracle/stubs/opencv2, ours), the reference's own
// ILidarEvalCallback.h (real header), and the data members of the two classes — declared below with the reference's names and
// types (Evaluation.h:143-151, EvaluationCallback.h:14-24, :57-62; the class definitions themselves cannot be included).
// oracle/build_ref.sh cuts the four function bodies out of the reference files AT BUILD TIME into oracle/_ref/eval_extract.inc
// (git-ignored; nothing of the reference is stored in this repo) and this driver #includes that file unmodified.
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

#include <Eigen/Eigen>
#include <opencv2/opencv.hpp>
#include "ILidarEvalCallback.h"   // reference DS/Evaluation/ILidarEvalCallback.h

#include "../include/b200fusion.h"

using namespace std;

namespace dynslam {
namespace eval {

struct Stats { long missing = 0; long error = 0; long correct = 0; long missing_separate = 0; };

class EvaluationCallback : public ILidarEvalCallback {
 public:
  const float delta_max;
  const bool compare_on_intersection;
  const bool kitti_style;
  EvaluationCallback(float d, bool c, bool k) : delta_max(d), compare_on_intersection(c), kitti_style(k) {}
  void ProcessLidarPoint(int idx, const Eigen::Vector3d &velo_2d_homo_px, float rendered_disp, float rendered_depth_m, float input_disp,
                         float input_depth_m, float lidar_disp, int frame_width, int frame_height) override;
  void ComputeAccuracy(float rendered_disp, float rendered_depth_m, float input_disp, float input_depth_m, float lidar_disp, Stats &input_stats,
                       Stats &rendered_stats);
  Stats input_stats_;
  Stats rendered_stats_;
  long measurement_count_ = 0;
};

class Evaluation {
 public:
  Eigen::Matrix4d velo_to_left_gray_cam_;
  Eigen::Matrix34d proj_left_color_;
  Eigen::Matrix34d proj_right_color_;
  float baseline_m_;
  int frame_width_;
  int frame_height_;
  float min_depth_m_;
  float max_depth_m_;
  float left_focal_length_px_;
  bool ProjectLidar(const Eigen::Vector4f &velodyne_reading, Eigen::Vector3d &out_velo_2d_left, Eigen::Vector3d &out_velo_2d_right) const;
  void EvaluateDepth(const Eigen::MatrixX4f &lidar_points, const float *const rendered_depth, const cv::Mat1s &input_depth_mm,
                     const std::vector<ILidarEvalCallback *> &callbacks) const;
};

#include "_ref/eval_extract.inc"

}  // namespace eval
}  // namespace dynslam

namespace {
// ours: SegmentedEvaluationCallback's dispatch (SegmentedEvaluationCallback.cpp:8-41) with GetPointAssociation's verdict read from a
// byte image at (round(px), round(py)) (SegmentedCallback.cpp:17-18) — the tracker / segmentation behind it are out of scope
struct SplitCallback : public ILidarEvalCallback {
  dynslam::eval::EvaluationCallback static_eval, dynamic_eval;
  const uint8_t *association; int w; bool has_dynamic; long skipped = 0;
  SplitCallback(const b200_eval_callback &c, const uint8_t *a, int w_, bool d)
      : static_eval(c.delta_max, c.compare_on_intersection != 0, c.kitti_style != 0), dynamic_eval(c.delta_max, c.compare_on_intersection != 0, c.kitti_style != 0),
        association(a), w(w_), has_dynamic(d) {}
  void ProcessLidarPoint(int idx, const Eigen::Vector3d &p, float rd, float rm, float id, float im, float ld, int fw, int fh) override {
    const int px = static_cast<int>(round(p(0))), py = static_cast<int>(round(p(1)));
    const int a = association ? association[py * w + px] : B200_EVAL_STATIC;
    if (a == B200_EVAL_DYNAMIC && has_dynamic) dynamic_eval.ProcessLidarPoint(idx, p, rd, rm, id, im, ld, fw, fh);
    else if (a == B200_EVAL_STATIC) static_eval.ProcessLidarPoint(idx, p, rd, rm, id, im, ld, fw, fh);
    else skipped++;
  }
};
void fill(b200_eval_result &r, const dynslam::eval::EvaluationCallback &c) {
  r.measurement_count = c.measurement_count_;
  r.rendered.missing = c.rendered_stats_.missing; r.rendered.error = c.rendered_stats_.error; r.rendered.correct = c.rendered_stats_.correct;
  r.rendered.missing_separate = c.rendered_stats_.missing_separate;
  r.input.missing = c.input_stats_.missing; r.input.error = c.input_stats_.error; r.input.correct = c.input_stats_.correct;
  r.input.missing_separate = c.input_stats_.missing_separate;
}
}  // namespace

extern "C" int ref_evaluate_depth(const b200_eval_params *p, const float *lidar_points, int n, const float *rendered_depth, const int16_t *input_depth_mm,
                                  const uint8_t *association, const b200_eval_callback *callbacks, int n_callbacks, b200_eval_result *out_static,
                                  b200_eval_result *out_dynamic, long *skipped) {
  dynslam::eval::Evaluation ev;
  memcpy(ev.velo_to_left_gray_cam_.m, p->velo_to_cam, sizeof(p->velo_to_cam));
  memcpy(ev.proj_left_color_.m, p->proj_left, sizeof(p->proj_left));
  memcpy(ev.proj_right_color_.m, p->proj_right, sizeof(p->proj_right));
  ev.baseline_m_ = p->baseline_m; ev.frame_width_ = p->frame_width; ev.frame_height_ = p->frame_height;
  ev.min_depth_m_ = p->min_depth_m; ev.max_depth_m_ = p->max_depth_m; ev.left_focal_length_px_ = p->left_focal_length_px;
  Eigen::MatrixX4f pts; pts.data = lidar_points; pts.n = n;
  cv::Mat1s depth(p->frame_height, p->frame_width);
  memcpy(depth.bytes.data(), input_depth_mm, (size_t)p->frame_width * p->frame_height * 2);
  std::vector<ILidarEvalCallback *> cbs;
  for (int c = 0; c < n_callbacks; ++c) cbs.push_back(new SplitCallback(callbacks[c], association, p->frame_width, out_dynamic != nullptr));
  int rc = 0;
  try { ev.EvaluateDepth(pts, rendered_depth, depth, cbs); } catch (const std::runtime_error &) { rc = -1; }
  for (int c = 0; c < n_callbacks; ++c) {
    SplitCallback *s = static_cast<SplitCallback *>(cbs[c]);
    fill(out_static[c], s->static_eval);
    if (out_dynamic) fill(out_dynamic[c], s->dynamic_eval);
    if (skipped) *skipped = s->skipped;
    delete s;
  }
  return rc;
}
