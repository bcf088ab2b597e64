/*
 * b200fusion.h — C-ABI of the B200-native voxel-hashed TSDF fusion / raycast engine.
 *
 * This is the drop-in boundary behind DynSLAM's ITMLib engine interfaces. Every entry point
 * below replaces one virtual of the reference (paths relative to
 * src/InfiniTAM/InfiniTAM/ITMLib/ in AndreiBarsan/DynSLAM):
 *
 *   ITMSceneReconstructionEngine<TVoxel,ITMVoxelBlockHash>   Engine/ITMSceneReconstructionEngine.h:33-78
 *   IITMVisualisationEngine / ITMVisualisationEngine          Engine/ITMVisualisationEngine.h:19-127
 *   ITMSwappingEngine<TVoxel,ITMVoxelBlockHash>               Engine/ITMSwappingEngine.h:22-31
 *
 * Plain pointers and sizes only; all pointers named d_* are CUDA device pointers owned by the
 * caller (exactly the buffers ITMScene / ITMRenderState_VH / ITMView already own), the engine
 * handle owns only its scratch. Byte layouts are those of the reference:
 *   ITMHashEntry   20 B  Utils/ITMLibDefines.h:69-84
 *   ITMVoxel_s_rgb  8 B  Utils/ITMLibDefines.h:138-169 (one 8x8x8 block = 4096 B)
 *   Matrix4f       m[col*4+row]  ORUtils/Matrix.h:23-33
 *
 * Error convention (reference: ORcudaSafeCall exits, logic errors throw std::runtime_error,
 * Reco_CUDA.cu:348-357): every function returns a b200_status; b200_last_error() gives text.
 * The C++ shim (dynslam_b200/itm_shim) maps the codes back onto throw / exit.
 */
#ifndef B200FUSION_H
#define B200FUSION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_SDF_BLOCK_SIZE 8
#define B200_SDF_BLOCK_SIZE3 512
#define B200_MAX_RENDERING_BLOCKS (65536 * 4) /* DeviceAgnostic/ITMVisualisationEngine.h:25 */
#define B200_MINMAX_SUBSAMPLE 8               /* DeviceAgnostic/ITMVisualisationEngine.h:27 */
#define B200_FAR_AWAY 999999.9f
#define B200_VERY_CLOSE 0.05f
#define B200_TRANSFER_BLOCK_NUM 0x1000        /* Utils/ITMLibDefines.h:39 */

typedef enum {
  B200_OK = 0,
  B200_ERR_CUDA = 1,          /* reference: ORcudaSafeCall -> exit(-1) */
  B200_ERR_VBA_FULL = 2,      /* reference: throw runtime_error, Reco_CUDA.cu:348-351 */
  B200_ERR_EXCESS_FULL = 3,   /* reference: throw runtime_error, Reco_CUDA.cu:353-357 */
  B200_ERR_INVALID = 4,
  B200_ERR_DECAY_RING_FULL = 5, /* no longer returned: a full decay queue drops its oldest snapshots (see b200_engine_config) */
  B200_ERR_UNSUPPORTED = 6,
  B200_ERR_NEGATIVE_DISPARITY = 7  /* reference: throw runtime_error("Negative disparity in ground truth."), Evaluation.cpp:273-275 */
} b200_status;

/* ---- byte-exact mirrors of the reference PODs ---------------------------------------- */

typedef struct {            /* ITMHashEntry */
  int16_t pos[3];
  int16_t _pad;
  int32_t offset;           /* 1-based index into the excess part, <1 = end of chain */
  int32_t ptr;              /* >=0 VBA block id, -1 swapped out, <-1 free */
  int32_t allocatedTime;
} b200_hash_entry;

typedef struct {            /* ITMVoxel_s_rgb */
  int16_t sdf;              /* tsdf * 32767 */
  uint8_t w_depth;
  uint8_t clr[3];
  uint8_t w_color;
  uint8_t _pad;
} b200_voxel;

typedef struct { int32_t x, y, z; } b200_vec3i;   /* Vector3i, visible-list item */
typedef struct { float x, y; } b200_vec2f;        /* Vector2f, min/max image pixel */
typedef struct { float x, y, z, w; } b200_vec4f;  /* Vector4f */
typedef struct { uint8_t x, y, z, w; } b200_vec4u;/* Vector4u, RGBA8 */

/* ---- scene (ITMScene = ITMVoxelBlockHash + ITMLocalVBA + ITMSceneParams) --------------- */

typedef struct {
  b200_voxel *d_voxels;          /* localVBA.GetVoxelBlocks(): numBlocks*512 voxels */
  int32_t *d_allocationList;     /* localVBA.GetAllocationList(): numBlocks ints */
  b200_hash_entry *d_hash;       /* index.GetEntries(): numBuckets+excessSize entries */
  int32_t *d_excessList;         /* index.GetExcessAllocationList(): excessSize ints */
  uint8_t *d_swapStates;         /* globalCache->GetSwapStates(true) or NULL (swapping off) */
  int32_t numBlocks;             /* sdfLocalBlockNum */
  int32_t numBuckets;            /* SDF_BUCKET_NUM, power of two */
  int32_t excessSize;            /* SDF_EXCESS_LIST_SIZE */
  /* host-visible counters, valid on return of every synchronous call */
  int32_t lastFreeBlockId;       /* localVBA.lastFreeBlockId */
  int32_t lastFreeExcessListId;  /* index.lastFreeExcessListId */
  /* ITMSceneParams (Objects/ITMSceneParams.h:14-71) */
  float voxelSize, mu;
  int32_t maxW;
  float viewFrustum_min, viewFrustum_max;
  int32_t stopIntegratingAtMaxW;
  int32_t useSwapping;
} b200_scene;

/* ---- render state (ITMRenderState_VH + base) ------------------------------------------ */

typedef struct {
  b200_vec3i *d_visibleBlockPositions; /* GetVisibleBlockPositions(): numBlocks items */
  uint8_t *d_entriesVisibleType;       /* GetEntriesVisibleType(): one byte per hash entry */
  b200_vec2f *d_minmax;                /* renderingRangeImage, w*h, 1/8-res corner used */
  b200_vec4f *d_raycastResult;         /* raycastResult, w*h */
  b200_vec4f *d_forwardProjection;     /* forwardProjection, w*h (may be NULL) */
  int32_t *d_fwdProjMissingPoints;     /* fwdProjMissingPoints, w*h (may be NULL) */
  b200_vec4u *d_raycastImage;          /* raycastImage, w*h */
  int32_t img_w, img_h;
  int32_t noVisibleBlocks;             /* host-visible, in/out */
  int32_t noFwdProjMissingPoints;      /* host-visible, out */
} b200_render_state;

/* ---- per-frame inputs (ITMView + ITMTrackingState::pose_d + calib) ----------------------- */

typedef struct {
  const float *d_depth;          /* view->depth, metres, w*h */
  const b200_vec4u *d_rgb;       /* view->rgb, RGBA8 */
  int32_t depth_w, depth_h, rgb_w, rgb_h;
  float M_d[16];                 /* trackingState->pose_d->GetM() */
  float invM_d[16];              /* its inverse (Matrix4::inv, ORUtils/Matrix.h:162-224) */
  float M_rgb[16];               /* calib.trafo_rgb_to_depth.calib_inv * M_d */
  float proj_d[4];               /* intrinsics_d.projectionParamsSimple.all = fx,fy,cx,cy */
  float proj_rgb[4];
  int32_t depthWeighting;        /* WeightParams, Engine/ITMSceneReconstructionEngine.h:21-23 */
  int32_t requiresFullRendering; /* trackingState->requiresFullRendering (approx. integration off) */
} b200_view;

typedef struct {               /* pose + intrinsics pair used by the visualisation calls */
  float M[16];
  float invM[16];
  float proj[4];
} b200_camera;

typedef enum {                 /* IITMVisualisationEngine::RenderImageType, Vis.h:22-32 */
  B200_RENDER_SHADED_GREYSCALE = 0,
  B200_RENDER_COLOUR_FROM_VOLUME = 1,
  B200_RENDER_COLOUR_FROM_NORMAL = 2,
  B200_RENDER_COLOUR_FROM_DEPTH_WEIGHT = 3,
  B200_RENDER_DEPTH_MAP = 4
} b200_render_type;

/* ---- engine handle ---------------------------------------------------------------------- */

typedef struct b200_engine b200_engine;

typedef struct {
  int32_t device;              /* CUDA ordinal */
  int32_t numBlocks;           /* capacity the scratch is sized for */
  int32_t numBuckets;
  int32_t excessSize;
  int32_t img_w, img_h;
  int64_t decayRingItems;      /* capacity (visible-list items) of the decay snapshot ring; 0 = default 24*numBlocks.
                                  The reference keeps one copy of the visible list per frame in an unbounded std::queue
                                  (Reco_CUDA.cu:302-317); here the queue holds at most 4095 frames and decayRingItems list
                                  items. Beyond either bound the OLDEST snapshots are dropped (a Decay() that would have
                                  swept a dropped snapshot sweeps nothing; b200_frame_stats.droppedSnapshots counts them) —
                                  size it as min_decay_age x (visible blocks per frame) with headroom. Never an error. */
  void *stream;                /* optional caller cudaStream_t; NULL = engine-owned stream */
} b200_engine_config;

/* On failure nothing is left allocated, *out is NULL and b200_last_error(NULL) gives the reason (thread-local). */
b200_status b200_engine_create(const b200_engine_config *cfg, b200_engine **out);
void b200_engine_destroy(b200_engine *e);
const char *b200_last_error(const b200_engine *e);
void *b200_engine_stream(b200_engine *e);
int32_t b200_frame_index(const b200_engine *e);      /* frameIdx, Reco_CUDA.h:37-45 */

/* ---- ITMSceneReconstructionEngine ------------------------------------------------------- */

/* ResetScene — Reco_CUDA.cu:145-172 */
b200_status b200_reset_scene(b200_engine *e, b200_scene *scene);
/* AllocateSceneFromDepth — Reco_CUDA.cu:175-358 */
b200_status b200_allocate_from_depth(b200_engine *e, b200_scene *scene, b200_render_state *rs,
                                     const b200_view *view, int onlyUpdateVisibleList);
/* IntegrateIntoScene — Reco_CUDA.cu:361-427 */
b200_status b200_integrate(b200_engine *e, b200_scene *scene, b200_render_state *rs,
                           const b200_view *view);
/* Decay — Reco_CUDA.cu:509-560 */
b200_status b200_decay(b200_engine *e, b200_scene *scene, b200_render_state *rs, int maxWeight,
                       int minAge, int forceAllVoxels);
/* GetDecayedBlockCount — Reco_CUDA.cu:563-566 */
size_t b200_decayed_block_count(const b200_engine *e);

/* ---- ITMVisualisationEngine ---------------------------------------------------------------- */

/* FindVisibleBlocks — Vis_CUDA.cu:151-180 */
b200_status b200_find_visible_blocks(b200_engine *e, const b200_scene *scene,
                                     b200_render_state *rs, const b200_camera *cam);
/* CreateExpectedDepths — Vis_CUDA.cu:194-240 */
b200_status b200_expected_depths(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                                 const b200_camera *cam);
/* FindSurface / GenericRaycast — Vis_CUDA.cu:242-265, :484-489 */
b200_status b200_find_surface(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                              const b200_camera *cam);
/* RenderImage — Vis_CUDA.cu:267-343 */
b200_status b200_render_image(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                              const b200_camera *cam, b200_vec4u *d_outChar, float *d_outFloat,
                              int out_w, int out_h, b200_render_type type);
/* CreateICPMaps — Vis_CUDA.cu:372-390. d_points/d_normals = pointCloud->locations/colours */
b200_status b200_icp_maps(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                          const b200_view *view, b200_vec4f *d_points, b200_vec4f *d_normals);
/* ForwardRender — Vis_CUDA.cu:393-453 */
b200_status b200_forward_render(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                                const b200_view *view);
/* CreatePointCloud — Vis_CUDA.cu:346-369. *noTotalPoints = pointCloud->noTotalPoints */
b200_status b200_point_cloud(b200_engine *e, const b200_scene *scene, b200_render_state *rs,
                             const b200_view *view, const float *calib_rgb_to_depth,
                             int skipPoints, b200_vec4f *d_locations, b200_vec4f *d_colours,
                             uint32_t *noTotalPoints);

/* ---- ITMSwappingEngine ------------------------------------------------------------------- */

typedef struct {               /* the device half of ITMGlobalCache, Objects/ITMGlobalCache.h */
  b200_voxel *d_syncedVoxelBlocks;  /* SDF_TRANSFER_BLOCK_NUM blocks */
  uint8_t *d_hasSyncedData;         /* bool per transfer slot */
  int32_t *d_neededEntryIDs;        /* SDF_TRANSFER_BLOCK_NUM ints */
} b200_transfer_buffers;

/* IntegrateGlobalIntoLocal, device half — Swap_CUDA.cu:44-124.
   step 1: list entries to swap in; returns the count (<= SDF_TRANSFER_BLOCK_NUM) */
b200_status b200_swap_list_in(b200_engine *e, b200_scene *scene, b200_transfer_buffers *tb,
                              int *noNeeded);
/* step 2: merge the blocks the host has copied into tb into the local VBA */
b200_status b200_swap_integrate_in(b200_engine *e, b200_scene *scene, b200_transfer_buffers *tb,
                                   int noNeeded);
/* SaveToGlobalMemory, device half — Swap_CUDA.cu:126-216 */
b200_status b200_swap_out(b200_engine *e, b200_scene *scene, b200_render_state *rs,
                          b200_transfer_buffers *tb, int *noNeeded);

/* ---- fused fast path (no reference twin: removes the per-call blocking copies) */

typedef struct {
  int32_t doDecay, decayMaxWeight, decayMinAge;
  int32_t doRaycast;           /* CreateExpectedDepths + CreateICPMaps */
  /* Optional per-frame renders for compositing (the multi-volume configuration): shaded from the ray points of THIS frame's
     raycast (same pose), i.e. RenderImage's shading pass without a second raycast, before Decay touches the volume.
     NULL = not rendered. d_colourRender: RENDER_COLOUR_FROM_VOLUME (w*h RGBA8); d_depthRender: RENDER_DEPTH_MAP (w*h f32). */
  b200_vec4u *d_colourRender;
  float *d_depthRender;
} b200_frame_opts;

/* Enqueue allocate -> integrate -> expected depths -> ICP maps -> decay for one frame on the
   engine stream without any host synchronisation. Counters live on the device; the host
   fields of scene / rs are refreshed by b200_sync(). */
b200_status b200_process_frame_async(b200_engine *e, b200_scene *scene, b200_render_state *rs,
                                     const b200_view *view, b200_vec4f *d_points,
                                     b200_vec4f *d_normals, const b200_frame_opts *opts);
b200_status b200_sync(b200_engine *e, b200_scene *scene, b200_render_state *rs);

/* Same, from host buffers: depth (float metres) and rgb are copied H2D from pinned staging
   inside the call and the grey raycast image is copied back to h_outImage (may be NULL). */
b200_status b200_process_frame_host(b200_engine *e, b200_scene *scene, b200_render_state *rs,
                                    b200_view *view, const float *h_depth,
                                    const b200_vec4u *h_rgb, float *d_depth_stage,
                                    b200_vec4u *d_rgb_stage, b200_vec4f *d_points,
                                    b200_vec4f *d_normals, const b200_frame_opts *opts,
                                    b200_vec4u *h_outImage);

/* Pipelined variant for streams of host frames: two staging slots and a copy stream inside the engine, so that the
   H2D copy of frame f+1 and the D2H copy of frame f's image overlap the kernels of the neighbouring frames.
   b200_host_frame_submit(slot) enqueues {H2D depth+RGB from (pinned) host memory, the fused frame, D2H of the grey raycast
   image into h_outImage} without blocking; b200_host_frame_wait(slot) blocks until that slot's image has landed.
   A slot may be resubmitted only after it has been waited for. Host counters are refreshed by b200_sync(). */
b200_status b200_host_frame_submit(b200_engine *e, b200_scene *scene, b200_render_state *rs, b200_view *view,
                                   const float *h_depth, const b200_vec4u *h_rgb, b200_vec4f *d_points, b200_vec4f *d_normals,
                                   const b200_frame_opts *opts, b200_vec4u *h_outImage, int slot);
b200_status b200_host_frame_wait(b200_engine *e, int slot);

/* ---- view builder: the step immediately before fusion ------------------------
   Replaces ITMViewBuilder_CUDA (Engine/DeviceSpecific/CUDA/ITMViewBuilder_CUDA.cu). All images are
   device pointers, row-major, w*h elements. Border semantics are the CUDA reference's (the filter leaves
   the two outermost rows/columns of its TARGET untouched, :196-209). */

typedef struct {
  int32_t trafoType;           /* ITMDisparityCalib::TrafoType (Objects/ITMDisparityCalib.h:24-29): 0 TRAFO_KINECT, 1 TRAFO_AFFINE */
  float params[2];             /* disparityCalib.params */
  float fx_depth;              /* intrinsics_d.projectionParamsSimple.fx (TRAFO_KINECT only) */
  float intrinsics_d[4];       /* intrinsics_d.projectionParamsSimple.all, for ComputeNormalAndWeights */
  int32_t useBilateralFilter;  /* ITMLibSettings::useBilateralFilter (default true, Utils/ITMLibSettings.cpp:63) */
  int32_t modelSensorNoise;    /* ITMLibSettings::modelSensorNoise (TRACKER_WICP only, :74-75) */
} b200_view_calib;

/* ITMViewBuilder::ConvertDisparityToDepth (ITMViewBuilder_CUDA.cu:120-135, DA/ITMViewBuilder.h:7-20) */
b200_status b200_convert_disparity_to_depth(b200_engine *e, float *d_out, const int16_t *d_in, int w, int h,
                                            float p0, float p1, float fx_depth);
/* ITMViewBuilder::ConvertDepthAffineToFloat (ITMViewBuilder_CUDA.cu:137-148, DA/ITMViewBuilder.h:22-28) */
b200_status b200_convert_depth_affine_to_float(b200_engine *e, float *d_out, const int16_t *d_in, int w, int h,
                                               float p0, float p1);
/* ITMViewBuilder::DepthFiltering — ONE pass (ITMViewBuilder_CUDA.cu:150-161, :196-209, DA/ITMViewBuilder.h:31-56);
   d_out's 2-pixel border is left untouched */
b200_status b200_depth_filtering(b200_engine *e, float *d_out, const float *d_in, int w, int h);
/* ITMViewBuilder::ComputeNormalAndWeights (ITMViewBuilder_CUDA.cu:163-178, :211-227, DA/ITMViewBuilder.h:59-114) */
b200_status b200_compute_normal_and_weights(b200_engine *e, b200_vec4f *d_normal, float *d_sigmaZ, const float *d_depth,
                                            int w, int h, const float intrinsic[4]);
/* ITMViewBuilder::UpdateView(view, rgb, rawDepth, useBilateralFilter, modelSensorNoise) without its two H2D
   copies (ITMViewBuilder_CUDA.cu:33-84): conversion + five filter passes + copy back in one kernel.
   d_depthNormal / d_depthUncertainty are used only with modelSensorNoise. The _async form only enqueues
   on the engine stream (the fused frame that follows is ordered behind it). */
b200_status b200_update_view(b200_engine *e, const int16_t *d_rawDepth, int w, int h, const b200_view_calib *calib,
                             float *d_depth, b200_vec4f *d_depthNormal, float *d_depthUncertainty);
b200_status b200_update_view_async(b200_engine *e, const int16_t *d_rawDepth, int w, int h, const b200_view_calib *calib,
                                   float *d_depth, b200_vec4f *d_depthNormal, float *d_depthUncertainty);
/* b200_host_frame_submit for a RAW sensor frame: h_rawDepth (int16, e.g. millimetres) and h_rgb are copied H2D,
   UpdateView runs on the device into the slot's float depth image, then the fused frame. Halves the depth upload. */
b200_status b200_host_frame_submit_raw(b200_engine *e, b200_scene *scene, b200_render_state *rs, b200_view *view,
                                       const int16_t *h_rawDepth, const b200_vec4u *h_rgb, const b200_view_calib *calib,
                                       b200_vec4f *d_points, b200_vec4f *d_normals, const b200_frame_opts *opts,
                                       b200_vec4u *h_outImage, int slot);

/* ---- instance frame splitting ---------------------------------------------------
   Replaces the download / CPU masking / upload round trip of InstanceReconstructor::ProcessFrame
   (DS/InstRecLib/InstanceReconstructor.cpp:180-197) around ProcessSilhouette_CPU / RemoveSilhouette_CPU
   (:59-170). The frame (view->rgb, view->depth) and the instance frames stay on the device. */

typedef struct {
  int32_t x0, y0, x1, y1;      /* instreclib::utils::BoundingBox::r, inclusive (Utils/BoundingBox.h:35-37) */
  const uint8_t *d_data;       /* Mask::mask_data_ (cv::Mat1b, box-sized, row-major; 1 = inside) on the DEVICE */
} b200_mask;

typedef struct {
  int32_t action;              /* what InstanceReconstructor::ProcessSilhouette (:226-285) decides for the track:
                                  0 = keep in the main map (static / static class),
                                  1 = RemoveSilhouette_CPU(delete_mask) only (uncertain or unreconstructable),
                                  2 = ProcessSilhouette_CPU(copy_mask -> instance frame) then RemoveSilhouette_CPU(delete_mask) */
  b200_mask copy_mask, delete_mask;
  b200_vec4u *d_dest_rgb;      /* instance_view->rgb, full frame size (action 2) */
  float *d_dest_depth;         /* instance_view->depth */
} b200_silhouette_op;

/* Applies ops[0..n) in order to the frame, exactly like the reference's loop over the active tracks
   (InstanceReconstructor::UpdateTracks, :210-224): a later op sees the pixels an earlier one blanked. */
b200_status b200_process_silhouettes(b200_engine *e, b200_vec4u *d_rgb, float *d_depth, int w, int h,
                                     const b200_silhouette_op *ops, int n);
b200_status b200_process_silhouettes_async(b200_engine *e, b200_vec4u *d_rgb, float *d_depth, int w, int h,
                                           const b200_silhouette_op *ops, int n);

/* ---- compositing of per-volume renders (the consumer of the multi-GPU gather) ------ */

typedef struct {
  const b200_vec4u *d_color;   /* the instance's render (GetImage) */
  const float *d_depth;        /* its depth render (GetFloatImage kDepth); 0 = nothing */
  int32_t tint[4];             /* kMatplotlib2Palette[track id % size] (InstanceReconstructor.cpp:43-55) */
} b200_instance_layer;

/* CompositeDepth (InstanceReconstructor.cpp:850-869) */
b200_status b200_composite_depth(b200_engine *e, float *d_target, const float *d_source, int n);
/* CompositeColor (InstanceReconstructor.cpp:873-905) */
b200_status b200_composite_color(b200_engine *e, b200_vec4u *d_target_color, float *d_target_depth,
                                 const b200_vec4u *d_instance_color, const float *d_instance_depth, int n,
                                 const int32_t tint[4], float tint_strength);
/* CompositeInstances (InstanceReconstructor.cpp:932-987) without the renders: dims the background by
   dim_factor (< 0: no dimming) and composites layers[0..n_layers) in order, all in one pass over the image. */
b200_status b200_composite_instances(b200_engine *e, b200_vec4u *d_out_color, float *d_out_depth, int n,
                                     const b200_instance_layer *layers, int n_layers, float dim_factor,
                                     float tint_strength);

/* ---- evaluation consumer of the float raycast: Evaluation::EvaluateDepth ----------------------------
   DS/Evaluation/Evaluation.cpp:241-304 with ProjectLidar (:214-238) and EvaluationCallback::ProcessLidarPoint /
   ComputeAccuracy (DS/Evaluation/EvaluationCallback.cpp:15-103): every LIDAR return of a frame is projected into the left
   and right colour cameras (double precision), the rendered depth (the float raycast / composite preview, metres) and the
   input depth (int16 millimetres) under it become disparities, and each callback — one per delta_max, plus the KITTI-style
   one — counts the point as missing, erroneous or correct for the fused and for the input depth. One launch for all points
   and all callbacks; only the counters come back. */

typedef struct b200_eval_params {
  double velo_to_cam[16];      /* Evaluation::velo_to_left_gray_cam_ (Eigen::Matrix4d, column-major) */
  double proj_left[12];        /* proj_left_color_  (Eigen::Matrix<double, 3, 4>, column-major) */
  double proj_right[12];       /* proj_right_color_ */
  float baseline_m, left_focal_length_px, min_depth_m, max_depth_m;
  int32_t frame_width, frame_height;
} b200_eval_params;

typedef struct b200_eval_callback {   /* EvaluationCallback's constructor arguments (EvaluationCallback.h:14-24) */
  float delta_max;
  int32_t compare_on_intersection, kitti_style;
} b200_eval_callback;

typedef struct b200_eval_stats { int64_t missing, error, correct, missing_separate; } b200_eval_stats;   /* Evaluation.h:27-33 */
typedef struct b200_eval_result { int64_t measurement_count; b200_eval_stats rendered, input; } b200_eval_result;
typedef struct b200_eval_summary { int64_t valid_lidar_points, epi_errors, negative_disparities, skipped_lidar_points; } b200_eval_summary;

/* SegmentedCallback::GetPointAssociation's verdict for the pixel a point falls on (SegmentedCallback.h), one byte per pixel */
enum { B200_EVAL_STATIC = 0, B200_EVAL_DYNAMIC = 1, B200_EVAL_NEITHER = 2 };
#define B200_EVAL_MAX_CALLBACKS 16

/* d_lidar: n points of four floats (x, y, z, reflectance — the KITTI velodyne .bin record); d_rendered_depth, d_input_depth_mm:
   frame_width x frame_height; d_association: one byte per pixel or NULL (every point static: Evaluation::EvaluateFrame);
   callbacks / out_static / out_dynamic: host arrays of n_callbacks (out_dynamic may be NULL: dynamic points are skipped).
   Synchronous, like the reference. A negative ground-truth disparity returns B200_ERR_NEGATIVE_DISPARITY (the reference
   throws at the first one; the counts are then unspecified). */
b200_status b200_evaluate_depth(b200_engine *e, const b200_eval_params *params, const float *d_lidar, int n,
                                const float *d_rendered_depth, const int16_t *d_input_depth_mm, const uint8_t *d_association,
                                const b200_eval_callback *callbacks, int n_callbacks, b200_eval_result *out_static,
                                b200_eval_result *out_dynamic, b200_eval_summary *summary);

/* ---- meshing: ITMMeshingEngine<TVoxel, ITMVoxelBlockHash>::MeshScene --------------------------------
   Engine/DeviceSpecific/CUDA/ITMMeshingEngine_CUDA.cu:37-81 with DeviceAgnostic/ITMMeshingEngine.h. d_triangles is
   mesh->triangles->GetData(MEMORYDEVICE_CUDA) (ITMMesh::Triangle, Objects/ITMMesh.h:20-23: six Vector3f), noMaxTriangles
   mesh->noMaxTriangles; *noTotalTriangles receives mesh->noTotalTriangles. Triangles come out in the order of the reference's
   serial CPU engine (ascending hash entry, then z, y, x, then the case table), not in the CUDA engine's atomicAdd order. */

typedef struct { float p0[3], p1[3], p2[3], c0[3], c1[3], c2[3]; } b200_triangle;

b200_status b200_mesh_scene(b200_engine *e, const b200_scene *scene, b200_triangle *d_triangles, uint32_t noMaxTriangles,
                            uint32_t *noTotalTriangles);

/* ---- the exchange step of the multi-volume configuration -----------------------------------------------
   One volume per GPU, one process per GPU: the static map on rank 0, one ITMScene per car on the other ranks
   (DS/InstRecLib/InstanceReconstructor.cpp:363-389). Fusion, allocation and decay need no communication; per frame every
   rank hands the colour (RGBA8) and depth (f32) renders of its volume to rank 0, which z-composites them over its own render
   (CompositeInstances, InstanceReconstructor.cpp:932-987). NCCL send/recv over NVLink on the communicator's own stream,
   double-buffered, so the next frame's kernels never wait for this frame's rendez-vous. libnccl.so.2 is loaded at run time. */

#define B200_COMM_ID_BYTES 128
typedef struct b200_comm b200_comm;

/* rank 0 makes the rendez-vous id (ncclGetUniqueId); the host distributes it to the other ranks over any channel */
b200_status b200_comm_unique_id(char id[B200_COMM_ID_BYTES]);
b200_status b200_comm_create(int device, int nranks, int rank, const char id[B200_COMM_ID_BYTES], int img_w, int img_h,
                             b200_comm **out);
void b200_comm_destroy(b200_comm *c);
const char *b200_comm_last_error(const b200_comm *c);   /* c == NULL: why the last create / unique_id failed on this thread */
/* Per frame, every rank: d_color / d_depth (w*h each) are this volume's renders, produced by work already enqueued on e's
   stream. Ranks > 0 send them to rank 0. Rank 0 receives the other ranks' layers, copies its own render to d_out_color /
   d_out_depth and composites the layers over it in rank order (tints: 4 ints per layer, rank r at tints[4*(r-1)];
   dim_factor < 0: no dimming). Only enqueues. slot in {0, 1} double-buffers the hand-over. */
b200_status b200_gather_composite_submit(b200_comm *c, b200_engine *e, const b200_vec4u *d_color, const float *d_depth,
                                         b200_vec4u *d_out_color, float *d_out_depth, const int32_t *tints, float dim_factor,
                                         float tint_strength, int slot);
/* e's stream waits on the device until slot's exchange has drained (call before overwriting the buffers handed over with it) */
b200_status b200_gather_composite_release(b200_comm *c, b200_engine *e, int slot);
/* the host waits until slot's exchange (on rank 0: and the composite) has finished */
b200_status b200_gather_composite_wait(b200_comm *c, int slot);

#ifdef __cplusplus
}
#endif
#endif /* B200FUSION_H */
