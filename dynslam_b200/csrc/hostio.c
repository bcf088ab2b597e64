/* hostio.c — the on-disk formats either side of the path, host code, built into libb200host.so.
 * NOT part of libb200fusion / the C-ABI: these are the files DynSLAM's own readers and writers handle (with OpenCV, pfmLib and
 * ITMMesh), restated without those dependencies for hosts that do not link them (the Python mirror, tools, tests):
 *   b200h_read_pfm            ReadFilePFM, src/pfmLib/ImageIOpfm.cpp:49-156 (DispNet disparity / depth maps, PrecomputedDepthProvider.cpp:27-31)
 *   b200h_read_depth_xml      the OpenCV FileStorage XML dump of a CV_16SC1 matrix named "depth-frame" (PrecomputedDepthProvider.cpp:32-42)
 *   b200h_clamp_max_depth_*   the "ensure the max depth" loop (PrecomputedDepthProvider.cpp:53-72)
 *   b200h_read_mask_txt       ReadMask, the numpy text dump of an instance mask (DS/InstRecLib/PrecomputedSegmentationProvider.cpp:24-71)
 *   b200h_write_obj           ITMMesh::WriteOBJ (ITMLib/Objects/ITMMesh.h:46-122) on the triangle array b200_mesh_scene fills
 * What they produce is exactly what the C-ABI consumes (int16 millimetre depth, byte masks of a bounding box) and what it fills
 * (ITMMesh::Triangle). tests/test_hostio.py pins each against the reference's own code where that compiles here.
 * Return: 0 on success, a negative code otherwise (-1 cannot open, -2 malformed, -3 size mismatch, -4 out of memory). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void b200h_free(void *p) { free(p); }

/* ---- PFM ("Pf" 1 band / "PF" 3 bands; rows stored bottom-up; the sign of the scale gives the byte order) ---- */
int b200h_read_pfm(const char *path, int *width, int *height, int *bands, float **data) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  char tag[8] = {0};
  int w = 0, h = 0;
  float scale = 0.0f;
  if (fscanf(f, "%7s %d %d %f", tag, &w, &h, &scale) != 4 || w <= 0 || h <= 0) { fclose(f); return -2; }
  int c = fgetc(f);                    /* a SINGLE newline after the scale (optionally <cr> first) */
  if (c == '\r') c = fgetc(f);
  if (c != '\n') { fclose(f); return -2; }
  const int nb = strcmp(tag, "Pf") == 0 ? 1 : (strcmp(tag, "PF") == 0 ? 3 : 0);
  if (!nb) { fclose(f); return -2; }
  const int one = 1;
  const int littleMachine = *(const unsigned char *)&one == 1, littleFile = scale < 0.0f, swap = littleMachine != littleFile;
  float *out = (float *)malloc(sizeof(float) * (size_t)w * h * nb);
  if (!out) { fclose(f); return -4; }
  for (int i = h - 1; i >= 0; --i) {
    float *row = out + (size_t)i * w * nb;
    if (fread(row, sizeof(float), (size_t)w * nb, f) != (size_t)w * nb) {
      /* the reference reads value by value and leaves what the file does not cover at zero */
      memset(row, 0, sizeof(float) * (size_t)w * nb);
      for (int r = i - 1; r >= 0; --r) memset(out + (size_t)r * w * nb, 0, sizeof(float) * (size_t)w * nb);
      break;
    }
    if (swap)
      for (int j = 0; j < w * nb; ++j) {
        unsigned char *p = (unsigned char *)&row[j], t;
        t = p[0]; p[0] = p[3]; p[3] = t; t = p[1]; p[1] = p[2]; p[2] = t;
      }
  }
  fclose(f);
  *width = w; *height = h; *bands = nb; *data = out;
  return 0;
}

/* ---- OpenCV FileStorage XML: <depth-frame type_id="opencv-matrix"><rows>R</rows><cols>C</cols><dt>s</dt><data> ... </data> ---- */
static const char *find_tag(const char *s, const char *tag) {
  const char *p = strstr(s, tag);
  return p ? p + strlen(tag) : 0;
}
int b200h_read_depth_xml(const char *path, int *width, int *height, int16_t **data) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)n + 1);
  if (!buf) { fclose(f); return -4; }
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return -2; }
  fclose(f);
  buf[n] = 0;
  const char *node = strstr(buf, "<depth-frame");
  if (!node) { free(buf); return -2; }
  const char *pr = find_tag(node, "<rows>"), *pc = find_tag(node, "<cols>"), *pt = find_tag(node, "<dt>"), *pd = find_tag(node, "<data>");
  if (!pr || !pc || !pt || !pd) { free(buf); return -2; }
  const int rows = atoi(pr), cols = atoi(pc);
  while (*pt == ' ' || *pt == '\n') ++pt;
  if (*pt != 's' || rows <= 0 || cols <= 0) { free(buf); return -2; }      /* "Precomputed depth map had the wrong format." (CV_16SC1 only) */
  int16_t *out = (int16_t *)malloc(sizeof(int16_t) * (size_t)rows * cols);
  if (!out) { free(buf); return -4; }
  char *end = (char *)pd;
  for (size_t i = 0; i < (size_t)rows * cols; ++i) {
    const long v = strtol(end, &end, 10);
    if (*end == '<' && i + 1 < (size_t)rows * cols) { free(buf); free(out); return -3; }
    out[i] = (int16_t)v;
  }
  free(buf);
  *width = cols; *height = rows; *data = out;
  return 0;
}

/* PrecomputedDepthProvider.cpp:53-72: depths beyond the provider's maximum become "no measurement" */
void b200h_clamp_max_depth_s16(int16_t *d, size_t n, float max_depth_m) {
  const float max_depth_mm_f = max_depth_m * 1000.0f;
  const int16_t max_depth_mm_s = (int16_t)round(max_depth_mm_f);
  for (size_t i = 0; i < n; ++i) if (d[i] > max_depth_mm_s) d[i] = 0;
}
void b200h_clamp_max_depth_f32(float *d, size_t n, float max_depth_m) {
  const float max_depth_mm_f = max_depth_m * 1000.0f;
  for (size_t i = 0; i < n; ++i) if (d[i] > max_depth_mm_f) d[i] = 0.0f;
}

/* ---- numpy text dump of a mask: one line per row, values separated by blanks, parsed as doubles and truncated to a byte ---- */
int b200h_read_mask_txt(const char *path, int width, int height, uint8_t *out) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  int lines = 0, rc = 0;
  char *line = 0;
  size_t cap = 0;
  ssize_t len;
  while ((len = getline(&line, &cap, f)) >= 0) {
    if (lines >= height) { rc = -3; break; }                    /* "Image height mismatch." */
    if (len > 0 && line[len - 1] == '\n') line[--len] = 0;
    /* the reference's loop `while (!line_ss.eof()) { line_ss >> val; ... }` stores one value per extraction attempt: a trailing
       blank yields one more (failed) extraction, which leaves 0 in val (C++11 num_get) */
    char *p = line;
    int col = 0;
    for (;;) {
      if (col >= width) { rc = -3; break; }                      /* "Image width mismatch." */
      char *e;
      while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
      double v = 0.0;
      if (*p) { v = strtod(p, &e); if (e == p) { v = 0.0; p += strlen(p); } else p = e; }
      out[(size_t)lines * width + col] = (uint8_t)v;
      ++col;
      if (!*p) break;
    }
    if (rc) break;
    ++lines;
  }
  free(line);
  fclose(f);
  return rc;
}

/* ---- ITMMesh::WriteOBJ: three "v x y z r g b" lines per triangle, then the faces with the winding reversed ---- */
typedef struct { float p0[3], p1[3], p2[3], c0[3], c1[3], c2[3]; } b200h_triangle;   /* == b200_triangle == ITMMesh::Triangle */
int b200h_write_obj(const char *path, const b200h_triangle *t, uint32_t noTotalTriangles, uint32_t noMaxTriangles) {
  if (noTotalTriangles > noMaxTriangles) return -3;               /* "Unable to save mesh to file [...]. Too many triangles" */
  FILE *f = fopen(path, "w+");
  if (!f) return -1;                                              /* "Could not open file for writing the mesh." */
  for (uint32_t i = 0; i < noTotalTriangles; i++) {
    fprintf(f, "v %f %f %f %f %f %f\n", t[i].p0[0], t[i].p0[1], t[i].p0[2], t[i].c0[0], t[i].c0[1], t[i].c0[2]);
    fprintf(f, "v %f %f %f %f %f %f\n", t[i].p1[0], t[i].p1[1], t[i].p1[2], t[i].c1[0], t[i].c1[1], t[i].c1[2]);
    fprintf(f, "v %f %f %f %f %f %f\n", t[i].p2[0], t[i].p2[1], t[i].p2[2], t[i].c2[0], t[i].c2[1], t[i].c2[2]);
  }
  for (uint32_t i = 0; i < noTotalTriangles; i++) fprintf(f, "f %d %d %d\n", i * 3 + 2 + 1, i * 3 + 1 + 1, i * 3 + 0 + 1);
  fclose(f);
  return 0;
}
