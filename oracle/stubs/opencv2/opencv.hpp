// Minimal stand-in for <opencv2/opencv.hpp> — TEST INFRASTRUCTURE ONLY (oracle/build_ref.sh).
// OpenCV is not installed here; the reference's InstRecLib/Utils/Mask.h only needs a byte matrix with at<T>(row, col), and the
// silhouette functions of InstanceReconstructor.cpp only read it. Nothing of OpenCV's implementation is reproduced.
#pragma once
#include <cassert>
#include <cstddef>
#include <vector>
typedef unsigned char uchar;
namespace cv {
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
class Mat {
 public:
  int rows, cols;
  std::vector<unsigned char> bytes;
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c) : rows(r), cols(c), bytes((size_t)r * c) {}
  template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(bytes.data())[(size_t)r * cols + c]; }
  template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(bytes.data())[(size_t)r * cols + c]; }
  Size size() const { return Size(cols, rows); }
  static Mat zeros(int r, int c, int type) { Mat m; m.rows = r; m.cols = c; m.bytes.assign((size_t)r * c * (type == 21 ? 12 : 4), 0); return m; }   // CV_32FC1 = 5, CV_32FC3 = 21
};
class Mat1b : public Mat {
 public:
  Mat1b() {}
  Mat1b(int r, int c) : Mat(r, c) {}
  explicit Mat1b(Size s) : Mat(s.height, s.width) {}
};
}   // namespace cv
namespace cv {
class Mat1s : public Mat {   // 16-bit signed single-channel matrix (the input depth map in millimetres, Evaluation.cpp:280)
 public:
  Mat1s() {}
  Mat1s(int r, int c) { rows = r; cols = c; bytes.resize((size_t)r * c * 2); }
};
}   // namespace cv
// what src/pfmLib/ImageIOpfm.cpp's ReadFilePFM needs on top (oracle/ref_io_driver.cpp): float matrices of 1 or 3 channels
#ifndef B200_STUB_CV_FLOAT
#define B200_STUB_CV_FLOAT
#define CV_32FC1 5
#define CV_32FC3 21
namespace cv {
struct Vec3f { float v[3]; float &operator[](int i) { return v[i]; } const float &operator[](int i) const { return v[i]; } };
}   // namespace cv
#endif
