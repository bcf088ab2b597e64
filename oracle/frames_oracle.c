/* frames_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported or called by dynslam_b200/).
 *
 * Serial C restatement of the per-frame image stages either side of the volumes:
 *   ProcessSilhouette_CPU / RemoveSilhouette_CPU      DS/InstRecLib/InstanceReconstructor.cpp:59-170
 *   the per-track dispatch of ProcessSilhouette       DS/InstRecLib/InstanceReconstructor.cpp:226-285
 *   CompositeDepth / CompositeColor                    DS/InstRecLib/InstanceReconstructor.cpp:850-905
 *   background dimming + loop of CompositeInstances    DS/InstRecLib/InstanceReconstructor.cpp:932-987
 *
 * PINNED (tests/test_frames_oracle.py): process_silhouette, remove_silhouette, oracle_composite_depth and
 * oracle_composite_color are compared byte for byte with ProcessSilhouette_CPU<float>, RemoveSilhouette_CPU<float>,
 * CompositeDepth and CompositeColor compiled from the reference file itself (oracle/build_ref.sh cuts the four free functions
 * out of InstanceReconstructor.cpp at build time and compiles them against the reference's Mask / BoundingBox / ORUtils headers;
 * oracle/_ref/libinstrecref.so). NOT pinned, because they are member functions entangled with the tracker, Pangolin and the
 * renderers: the three-statement background dimming loop of CompositeInstances (:944-952) and the per-track dispatch
 * (:210-285); they are restated statement by statement and fixed by hand-computed known answers.
 */
#include <stdint.h>
#include <string.h>

#include "../include/b200fusion.h"

static int mask_at(const b200_mask *m, int row, int col) {           /* mask.GetData()->at<u_char>(row, col) */
  return m->d_data[(size_t)row * (m->x1 - m->x0 + 1) + col];
}

/* InstanceReconstructor.cpp:59-135 (DEPTH_T = float; the min-depth bookkeeping has no effect) */
static void process_silhouette(const b200_vec4u *source_rgb, const float *source_depth, b200_vec4u *dest_rgb, float *dest_depth,
                               int frame_width, int frame_height, const b200_mask *copy_mask) {
  int copy_box_width = copy_mask->x1 - copy_mask->x0 + 1, copy_box_height = copy_mask->y1 - copy_mask->y0 + 1;
  memset(dest_rgb, 255, (size_t)frame_width * frame_height * sizeof(*source_rgb));
  memset(dest_depth, 0, (size_t)frame_width * frame_height * sizeof(float));
  for (int row = 0; row < copy_box_height; ++row) {
    for (int col = 0; col < copy_box_width; ++col) {
      int copy_row = row + copy_mask->y0;
      int copy_col = col + copy_mask->x0;
      if (copy_row < 0 || copy_row >= frame_height || copy_col < 0 || copy_col >= frame_width) continue;
      int copy_idx = copy_row * frame_width + copy_col;
      if (mask_at(copy_mask, row, col) == 1) {
        dest_rgb[copy_idx] = source_rgb[copy_idx];
        dest_depth[copy_idx] = source_depth[copy_idx];
      } else {
        dest_rgb[copy_idx].x = 255; dest_rgb[copy_idx].y = 255; dest_rgb[copy_idx].z = 255;
      }
    }
  }
}

/* InstanceReconstructor.cpp:138-170 */
static void remove_silhouette(b200_vec4u *source_rgb, float *source_depth, int frame_width, int frame_height, const b200_mask *mask) {
  int box_width = mask->x1 - mask->x0 + 1, box_height = mask->y1 - mask->y0 + 1;
  for (int row = 0; row < box_height; ++row) {
    for (int col = 0; col < box_width; ++col) {
      int frame_row = row + mask->y0;
      int frame_col = col + mask->x0;
      if (frame_row < 0 || frame_row >= frame_height || frame_col < 0 || frame_col >= frame_width) continue;
      int frame_idx = frame_row * frame_width + frame_col;
      if (mask_at(mask, row, col) == 1) {
        source_rgb[frame_idx].x = 0; source_rgb[frame_idx].y = 0; source_rgb[frame_idx].z = 0; source_rgb[frame_idx].w = 0;
        source_depth[frame_idx] = 0.0f;
      }
    }
  }
}

/* The loop of InstanceReconstructor::UpdateTracks (:210-224) over ProcessSilhouette's three outcomes (:226-285).
 * All pointers are HOST pointers here. */
void oracle_process_silhouettes(b200_vec4u *rgb, float *depth, int w, int h, const b200_silhouette_op *ops, int n) {
  for (int k = 0; k < n; ++k) {
    const b200_silhouette_op *o = &ops[k];
    if (o->action == 2) process_silhouette(rgb, depth, o->d_dest_rgb, o->d_dest_depth, w, h, &o->copy_mask);
    if (o->action != 0) remove_silhouette(rgb, depth, w, h, &o->delete_mask);
  }
}

/* InstanceReconstructor.cpp:850-869 (its idx = i*noDims[1]+j walk visits every element exactly once) */
void oracle_composite_depth(float *t_data, const float *s_data, int n) {
  for (int idx = 0; idx < n; ++idx) {
    if (t_data[idx] == 0) t_data[idx] = s_data[idx];
    else if (s_data[idx] != 0) t_data[idx] = t_data[idx] < s_data[idx] ? t_data[idx] : s_data[idx];
  }
}

static unsigned char to_uchar_min255(double v) { return (unsigned char)(v < 255.0 ? v : 255.0); }   /* static_cast<uchar>(min(255.0, v)) */

/* InstanceReconstructor.cpp:873-905 */
void oracle_composite_color(b200_vec4u *t_color, float *t_depth, const b200_vec4u *s_color, const float *s_depth, int n,
                            const int32_t tint[4], float tint_strength) {
  const float kColorBoost = 0.50f;
  for (int idx = 0; idx < n; ++idx) {
    int instance_on_top = (s_depth[idx] != 0 && (t_depth[idx] == 0 || t_depth[idx] > s_depth[idx]));
    if (instance_on_top) {
      t_depth[idx] = s_depth[idx];
      double col_strength = 1.0 + kColorBoost - tint_strength;
      t_color[idx].x = to_uchar_min255(s_color[idx].x * col_strength + tint[0] * tint_strength);
      t_color[idx].y = to_uchar_min255(s_color[idx].y * col_strength + tint[1] * tint_strength);
      t_color[idx].z = to_uchar_min255(s_color[idx].z * col_strength + tint[2] * tint_strength);
    }
  }
}

/* InstanceReconstructor.cpp:932-987 without the renders: dim the background, then CompositeColor per layer */
void oracle_composite_instances(b200_vec4u *out_color, float *out_depth, int n, const b200_instance_layer *layers, int n_layers,
                                float dim_factor, float tint_strength) {
  if (dim_factor >= 0.0f) {
    for (int idx = 0; idx < n; ++idx) {
      out_color[idx].x = (unsigned char)(out_color[idx].x * (1.0 - dim_factor));
      out_color[idx].y = (unsigned char)(out_color[idx].y * (1.0 - dim_factor));
      out_color[idx].z = (unsigned char)(out_color[idx].z * (1.0 - dim_factor));
    }
  }
  for (int k = 0; k < n_layers; ++k)
    oracle_composite_color(out_color, out_depth, layers[k].d_color, layers[k].d_depth, n, layers[k].tint, tint_strength);
}
