// ref_driver.cpp — TEST INFRASTRUCTURE ONLY. Thin extern "C" driver around the REFERENCE's own
// per-element functions, compiled from the headers where they lie under /root/reference
// (never copied into this repo) into oracle/_ref/libitmref.so by oracle/build_ref.sh.
//
// Purpose: pin oracle/tsdf_oracle.c. The reference's _CPU hash engines are gutted
// (CPU/ITMSceneReconstructionEngine_CPU.cpp:73-115, :143-298 are inside /* */), so there is no
// runnable reference pipeline for allocate/integrate on the CPU; what IS live and host-compilable
// is every _CPU_AND_GPU_CODE_ function in Engine/DeviceAgnostic/. This driver loops over pixels /
// voxels / entries in the canonical serial order and calls those functions, so that a byte-for-byte
// comparison with the oracle checks the oracle's arithmetic against the reference's code.
// The sequencing loops here mirror the commented serial CPU code; only the loops are ours.
//
// Build flags: -DCOMPILE_WITHOUT_CUDA -D__device__=   (buildHashAllocAndVisibleTypePP is
// declared __device__-only, DeviceAgnostic/ITMSceneReconstructionEngine.h:176).
#include "ITMLib/Engine/ITMVisualisationEngine.h"
#include "ITMLib/Engine/DeviceAgnostic/ITMSceneReconstructionEngine.h"
#include "ITMLib/Engine/DeviceAgnostic/ITMVisualisationEngine.h"
#include "ITMLib/Engine/DeviceAgnostic/ITMSwappingEngine.h"
#include "ITMLib/Engine/DeviceAgnostic/ITMRepresentationAccess.h"
#include "ITMLib/Engine/DeviceAgnostic/ITMViewBuilder.h"
#include "ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.h"
#include "ITMLib/Objects/ITMScene.h"
#include "ITMLib/Objects/ITMMesh.h"

#include "../include/b200fusion.h"

#include <cstring>

using namespace ITMLib::Engine;

static_assert(sizeof(ITMHashEntry) == sizeof(b200_hash_entry), "hash entry layout");
static_assert(sizeof(ITMVoxel) == sizeof(b200_voxel), "voxel layout");
static_assert(sizeof(Vector3i) == sizeof(b200_vec3i), "vec3i layout");
static_assert(sizeof(Vector4f) == sizeof(b200_vec4f), "vec4f layout");
static_assert(sizeof(Vector2f) == sizeof(b200_vec2f), "vec2f layout");
static_assert(sizeof(Vector4u) == sizeof(b200_vec4u), "vec4u layout");
static_assert(sizeof(ITMMesh::Triangle) == sizeof(b200_triangle), "triangle layout");

static Matrix4f M4(const float *m) { Matrix4f r; for (int i = 0; i < 16; ++i) r.m[i] = m[i]; return r; }
static Vector4f V4(const float *v) { return Vector4f(v[0], v[1], v[2], v[3]); }

extern "C" {

int ref_table_sizes(int *numBuckets, int *excessSize) {
  *numBuckets = (int)SDF_BUCKET_NUM; *excessSize = (int)SDF_EXCESS_LIST_SIZE;
  return (int)sizeof(ITMHashEntry);
}

int ref_mat4_inv(const float *m, float *out) {
  Matrix4f M = M4(m), inv; bool ok = M.inv(inv);
  for (int i = 0; i < 16; ++i) out[i] = inv.m[i];
  return ok ? 1 : 0;
}

void ref_mat4_mul(const float *a, const float *b, float *out) {
  Matrix4f r = M4(a) * M4(b);
  for (int i = 0; i < 16; ++i) out[i] = r.m[i];
}

int ref_find_block(const b200_hash_entry *table, int x, int y, int z) {
  bool isFound = false;
  int idx = findBlock((const ITMHashEntry *)table, Vector3i(x, y, z), isFound);
  return (isFound && idx >= 0) ? idx : -1;
}

// Marking pass over the whole image in raster order — buildHashAllocAndVisibleTypePP per pixel.
void ref_mark_image(uint8_t *allocType, uint8_t *visType, int16_t *blockCoords, const b200_scene *s, const b200_view *v) {
  Vector4f invProj = V4(v->proj_d);
  invProj.x = 1.0f / invProj.x; invProj.y = 1.0f / invProj.y;
  float oneOverVoxelSize = 1.0f / (s->voxelSize * SDF_BLOCK_SIZE);
  Vector2i imgSize(v->depth_w, v->depth_h);
  Matrix4f invM = M4(v->invM_d);
  for (int locId = 0; locId < v->depth_w * v->depth_h; locId++) {
    int y = locId / v->depth_w, x = locId - y * v->depth_w;
    buildHashAllocAndVisibleTypePP(allocType, visType, x, y, (Vector4s *)blockCoords, v->d_depth, invM, invProj, s->mu,
                                   imgSize, oneOverVoxelSize, (const ITMHashEntry *)s->d_hash, s->viewFrustum_min,
                                   s->viewFrustum_max, (int *)0);
  }
}

int ref_block_visible(const int16_t *pos, const float *M, const float *proj, float voxelSize, int w, int h) {
  bool vis, visE;
  Vector3s p(pos[0], pos[1], pos[2]);
  checkBlockVisibility<false>(vis, visE, p, M4(M), V4(proj), voxelSize, Vector2i(w, h));
  return vis ? 1 : 0;
}

// Integration of one block through ComputeUpdatedVoxelInfo<true, ITMVoxel>::compute, voxel order
// z, y, x as in the commented CPU loop (CPU/ITMSceneReconstructionEngine_CPU.cpp:96-113).
void ref_integrate_block(b200_voxel *blk, const int16_t *pos, const b200_scene *s, const b200_view *v) {
  Matrix4f M_d = M4(v->M_d), M_rgb = M4(v->M_rgb);
  Vector4f pd = V4(v->proj_d), prgb = V4(v->proj_rgb);
  Vector2i ds(v->depth_w, v->depth_h), rs(v->rgb_w, v->rgb_h);
  WeightParams wp; wp.depthWeighting = v->depthWeighting != 0;
  ITMVoxel *vb = (ITMVoxel *)blk;
  Vector3i globalPos(pos[0], pos[1], pos[2]); globalPos *= SDF_BLOCK_SIZE;
  bool stopMaxW = s->stopIntegratingAtMaxW != 0, approx = !v->requiresFullRendering;
  for (int z = 0; z < SDF_BLOCK_SIZE; z++) for (int y = 0; y < SDF_BLOCK_SIZE; y++) for (int x = 0; x < SDF_BLOCK_SIZE; x++) {
    int locId = x + y * SDF_BLOCK_SIZE + z * SDF_BLOCK_SIZE * SDF_BLOCK_SIZE;
    if (stopMaxW) if (vb[locId].w_depth == s->maxW) continue;
    if (approx) if (vb[locId].w_depth != 0) continue;
    Vector4f pt_model;
    pt_model.x = (float)(globalPos.x + x) * s->voxelSize;
    pt_model.y = (float)(globalPos.y + y) * s->voxelSize;
    pt_model.z = (float)(globalPos.z + z) * s->voxelSize;
    pt_model.w = 1.0f;
    ComputeUpdatedVoxelInfo<true, ITMVoxel>::compute(vb[locId], pt_model, M_d, pd, M_rgb, prgb, s->mu, s->maxW,
                                                     v->d_depth, ds, (const Vector4u *)v->d_rgb, rs, wp);
  }
}

int ref_project_single_block(const int16_t *pos, const float *pose, const float *intr, int w, int h, float voxelSize,
                             int *ul, int *lr, float *zr) {
  Vector2i a, b; Vector2f z;
  Vector3s p(pos[0], pos[1], pos[2]);
  Matrix4f P = M4(pose); Vector4f I = V4(intr); Vector2i sz(w, h);
  bool ok = ProjectSingleBlock(p, P, I, sz, voxelSize, a, b, z);
  ul[0] = a.x; ul[1] = a.y; lr[0] = b.x; lr[1] = b.y; zr[0] = z.x; zr[1] = z.y;
  return ok ? 1 : 0;
}

// castRay over the whole image (CPU twin: CPU/ITMVisualisationEngine_CPU.cpp:158-192)
void ref_raycast(const b200_scene *s, b200_render_state *rs, const float *invM, const float *proj) {
  Vector4f invProj = V4(proj); invProj.x = 1.0f / invProj.x; invProj.y = 1.0f / invProj.y;
  float oneOverVoxelSize = 1.0f / s->voxelSize;
  Matrix4f iM = M4(invM);
  const Vector2f *minmax = (const Vector2f *)rs->d_minmax;
  for (int y = 0; y < rs->img_h; ++y) for (int x = 0; x < rs->img_w; ++x) {
    int locId = x + y * rs->img_w;
    int locId2 = (int)floor((float)x / minmaximg_subsample) + (int)floor((float)y / minmaximg_subsample) * rs->img_w;
    castRay<ITMVoxel, ITMVoxelIndex>(((Vector4f *)rs->d_raycastResult)[locId], x, y, (const ITMVoxel *)s->d_voxels,
                                     (const ITMHashEntry *)s->d_hash, iM, invProj, oneOverVoxelSize, s->mu, minmax[locId2]);
  }
}

// per-pixel shading — processPixel{Grey,Colour,Normal,ColourWeight,ColourDepth}
void ref_shade(const b200_scene *s, const b200_render_state *rs, const b200_camera *cam, b200_vec4u *outChar, float *outFloat, int type) {
  Matrix4f invM = M4(cam->invM), M = M4(cam->M);
  Vector3f light = -Vector3f(invM.getColumn(2));
  const ITMVoxel *vox = (const ITMVoxel *)s->d_voxels; const ITMHashEntry *idx = (const ITMHashEntry *)s->d_hash;
  WeightRenderingParams params(1.0, false, s->maxW, 2);
  for (int locId = 0; locId < rs->img_w * rs->img_h; ++locId) {
    Vector4f ptRay = ((const Vector4f *)rs->d_raycastResult)[locId];
    Vector4u &o = ((Vector4u *)outChar)[locId];
    switch (type) {
    case 1: processPixelColour<ITMVoxel, ITMVoxelIndex>(o, ptRay.toVector3(), ptRay.w > 0, vox, idx, light); break;
    case 2: processPixelNormal<ITMVoxel, ITMVoxelIndex>(o, ptRay.toVector3(), ptRay.w > 0, vox, idx, light); break;
    case 3: processPixelColourWeight<ITMVoxel, ITMVoxelIndex>(o, ptRay.toVector3(), ptRay.w > 0, vox, idx, light, params); break;
    case 4: processPixelColourDepth<ITMVoxel, ITMVoxelIndex>(outFloat[locId], ptRay.toVector3(), ptRay.w > 0, M, s->voxelSize); break;
    default: processPixelGrey<ITMVoxel, ITMVoxelIndex>(o, ptRay.toVector3(), ptRay.w > 0, vox, idx, light); break;
    }
  }
}

// processPixelICP<true> over the image (CPU twin: CPU/ITMVisualisationEngine_CPU.cpp:275-296)
void ref_icp(const b200_scene *s, b200_render_state *rs, const float *invM_d, b200_vec4f *points, b200_vec4f *normals) {
  Matrix4f invM = M4(invM_d);
  Vector3f light = -Vector3f(invM.getColumn(2));
  Vector2i imgSize(rs->img_w, rs->img_h);
  for (int y = 0; y < rs->img_h; ++y) for (int x = 0; x < rs->img_w; ++x)
    processPixelICP<true>((Vector4u *)rs->d_raycastImage, (Vector4f *)points, (Vector4f *)normals,
                          (const Vector4f *)rs->d_raycastResult, imgSize, x, y, s->voxelSize, light);
}

void ref_combine_block(const b200_voxel *src, b200_voxel *dst, int maxW) {
  for (int i = 0; i < SDF_BLOCK_SIZE3; ++i)
    CombineVoxelInformation<true, ITMVoxel>::compute(((const ITMVoxel *)src)[i], ((ITMVoxel *)dst)[i], maxW);
}

int ref_forward_project_pixel(const float *px, const float *M, const float *proj, int w, int h) {
  return forwardProjectPixel(V4(px), M4(M), V4(proj), Vector2i(w, h));
}

// ---- view builder (DeviceAgnostic/ITMViewBuilder.h); the loops are the CUDA build's index ranges ----
void ref_view_convert_disparity(float *out, const short *in, int w, int h, float p0, float p1, float fx) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) convertDisparityToDepth(out, x, y, in, Vector2f(p0, p1), fx, Vector2i(w, h));
}

void ref_view_convert_affine(float *out, const short *in, int w, int h, float p0, float p1) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) convertDepthAffineToFloat(out, x, y, in, Vector2i(w, h), Vector2f(p0, p1));
}

// filterDepth_device's guard (ITMViewBuilder_CUDA.cu:196-209): the target's 2-pixel border is not written
void ref_view_filter_pass(float *out, const float *in, int w, int h) {
  for (int y = 2; y < h - 2; y++) for (int x = 2; x < w - 2; x++) filterDepth(out, in, x, y, Vector2i(w, h));
}

// ComputeNormalAndWeight_device's guard (ITMViewBuilder_CUDA.cu:211-227), in-image threads
void ref_view_normal_weight(const float *depth, b200_vec4f *normal, float *sigmaZ, int w, int h, const float *intr) {
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int idx = x + y * w;
    if (x < 2 || x > w - 2 || y < 2 || y > h - 2) { ((Vector4f *)normal)[idx].w = -1.0f; sigmaZ[idx] = -1; }
    else computeNormalAndWeight(depth, (Vector4f *)normal, sigmaZ, x, y, Vector2i(w, h), V4(intr));
  }
}


// ---- meshing: the reference's own serial engine, ITMMeshingEngine_CPU<ITMVoxel, ITMVoxelBlockHash>::MeshScene
// (Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp:19-80), on a host ITMScene filled with the caller's table and voxels.
// The table must have the reference's compile-time size (SDF_BUCKET_NUM + SDF_EXCESS_LIST_SIZE entries).
unsigned ref_mesh_scene(const b200_hash_entry *hash, const b200_voxel *voxels, long numBlocks, float voxelSize, b200_triangle *out,
                        unsigned noMaxTriangles) {
  ITMSceneParams params(0.75f, 50, voxelSize, 0.1f, 300.0f, false);
  ITMScene<ITMVoxel, ITMVoxelIndex> scene(&params, false, MEMORYDEVICE_CPU, numBlocks);
  memcpy(scene.index.GetEntries(), hash, sizeof(ITMHashEntry) * (size_t)ITMVoxelBlockHash::noTotalEntries);
  memcpy(scene.localVBA.GetVoxelBlocks(), voxels, sizeof(ITMVoxel) * (size_t)numBlocks * SDF_BLOCK_SIZE3);
  ITMMesh mesh(MEMORYDEVICE_CPU, numBlocks);
  ITMMeshingEngine_CPU<ITMVoxel, ITMVoxelIndex> engine;
  engine.MeshScene(&mesh, &scene);
  const unsigned n = mesh.noTotalTriangles < noMaxTriangles ? mesh.noTotalTriangles : noMaxTriangles;
  memcpy(out, mesh.triangles->GetData(MEMORYDEVICE_CPU), sizeof(ITMMesh::Triangle) * (size_t)n);
  return mesh.noTotalTriangles;
}
int ref_table_entries(void) { return ITMVoxelBlockHash::noTotalEntries; }
} // extern "C"
