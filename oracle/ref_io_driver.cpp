// ref_io_driver.cpp — TEST INFRASTRUCTURE ONLY. Pins dynslam_b200/csrc/hostio.c (the on-disk formats) to the reference's own code.
//
//  * ReadFilePFM: the head of src/pfmLib/ImageIOpfm.cpp (helpers + the reader; the rest of the file opens HighGUI windows), cut
//    out at build time, compiled against oracle/stubs/opencv2 (a byte matrix with at<T>() and Mat::zeros, ours);
//  * ReadMask: the free function of DS/InstRecLib/PrecomputedSegmentationProvider.cpp:24-71, cut out at build time; it needs
//    dynslam::utils::Format only for its error texts (a two-line stand-in below);
//  * ITMMesh::WriteOBJ: the reference's own header (ITMLib/Objects/ITMMesh.h), compiled for the host.
// Nothing of the reference is stored in this repo: oracle/build_ref.sh writes the cut-outs into oracle/_ref/ (git-ignored).
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "ITMLib/Objects/ITMMesh.h"       // reference (host build: COMPILE_WITHOUT_CUDA)
#include <opencv2/opencv.hpp>             // stand-in

#include "_ref/pfm_extract.inc"           // skip_space, littleendian, swapBytes, ReadFilePFM — unmodified reference text

namespace refmask {
using namespace std;
static string Format(const char *fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  return string(buf);
}
#include "_ref/mask_extract.inc"          // uint8_t *ReadMask(std::istream &, int width, int height) — unmodified reference text
}   // namespace refmask

extern "C" {

int ref_read_pfm(const char *path, int *w, int *h, int *bands, float *out, size_t cap) {
  cv::Mat im;
  if (ReadFilePFM(im, std::string(path), false) < 0) return -2;
  if (im.rows <= 0 || im.cols <= 0) return -2;
  const int nb = (int)(im.bytes.size() / ((size_t)im.rows * im.cols * 4));
  if (im.bytes.size() > cap * sizeof(float)) return -3;
  memcpy(out, im.bytes.data(), im.bytes.size());
  *w = im.cols; *h = im.rows; *bands = nb;
  return 0;
}

int ref_read_mask(const char *path, int w, int h, uint8_t *out) {
  std::ifstream in(path);
  if (!in.is_open()) return -1;
  try {
    uint8_t *m = refmask::ReadMask(in, w, h);
    memcpy(out, m, (size_t)w * h);
    delete[] m;
  } catch (const std::runtime_error &) { return -3; }
  return 0;
}

int ref_write_obj(const char *path, const void *triangles, unsigned noTotal, long sdfLocalBlockNum) {
  ITMLib::Objects::ITMMesh mesh(MEMORYDEVICE_CPU, sdfLocalBlockNum);
  if (noTotal <= mesh.noMaxTriangles) memcpy(mesh.triangles->GetData(MEMORYDEVICE_CPU), triangles, sizeof(ITMLib::Objects::ITMMesh::Triangle) * (size_t)noTotal);
  mesh.noTotalTriangles = noTotal;
  try { mesh.WriteOBJ(path); } catch (const std::runtime_error &) { return -3; }
  return 0;
}

}   // extern "C"
