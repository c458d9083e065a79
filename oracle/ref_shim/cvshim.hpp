// cvshim.hpp — TEST INFRASTRUCTURE.  A stand-in for the slice of the OpenCV C++ API that the
// reference's Source/Libraries/ORB_SLAM2/src/ORBextractor.cc uses, so that the reference's OWN,
// UNMODIFIED translation unit can be compiled (from where it lies under /root/reference) into
// oracle/_ref/libref_orb.so and run as a checker for oracle_cpu.cpp and the CUDA path.
//
// The image primitives (resize, FAST, GaussianBlur, fastAtan2, copyMakeBorder) are NOT OpenCV here:
// they forward to the restatements in liborb_oracle.so, which tests/test_oracle_vs_cv2.py pins
// bit-exact against cv2 4.13.0.  Everything else the reference does — cell loops, DivideNode,
// DistributeOctTree with its std::list discipline, IC_Angle, computeOrbDescriptor, operator() —
// is the reference's own code.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

extern "C" {
void orb_oracle_resize(const uint8_t *src, int sw, int sh, int sstep, uint8_t *dst, int dw, int dh, int dstep);
void orb_oracle_blur(const uint8_t *src, int w, int h, int sstep, uint8_t *dst, int dstep);
float orb_oracle_fast_atan2(float y, float x);
int orb_oracle_fast(const uint8_t *img, int w, int h, int step, int th, int *dst, int cap);
}

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  Point_ &operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

// Mat::zeros returns an expression; assigning it to an existing Mat of the same size fills that Mat IN PLACE
// (cv::MatExpr semantics: create() is a no-op for a matching buffer).  computeDescriptors relies on this to write
// into a row range of the output matrix (ORBextractor.cc:969-971).
struct ZerosExpr { int rows, cols; };

class Mat {
 public:
  struct Step { size_t v; operator size_t() const { return v; } };
  int rows = 0, cols = 0;
  Step step{0};
  uchar *data = nullptr;
  Mat() {}
  Mat(Size sz, int) { create(sz.height, sz.width, 0); }
  Mat(int r, int c, int) { create(r, c, 0); }
  Mat(int r, int c, int, void *ext, size_t s) : rows(r), cols(c), step{s}, data((uchar *)ext) {}
  void create(int r, int c, int) {
    if (data && r == rows && c == cols) return;     // cv::Mat::create keeps a matching buffer (ROI included)
    rows = r; cols = c; step.v = (size_t)c;
    own_ = std::shared_ptr<uchar>(new uchar[(size_t)r * c + 64], std::default_delete<uchar[]>());
    data = own_.get();
  }
  void release() { own_.reset(); data = nullptr; rows = cols = 0; step.v = 0; }
  bool empty() const { return !data || !rows || !cols; }
  int type() const { return CV_8UC1; }
  size_t step1() const { return step.v; }
  template <class T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step.v + c * sizeof(T)); }
  template <class T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step.v + c * sizeof(T)); }
  uchar *ptr(int r = 0) { return data + (size_t)r * step.v; }
  const uchar *ptr(int r = 0) const { return data + (size_t)r * step.v; }
  Mat roi(int x, int y, int w, int h) const { Mat m; m.rows = h; m.cols = w; m.step = step; m.data = data + (size_t)y * step.v + x; m.own_ = own_; return m; }
  Mat operator()(const Rect &r) const { return roi(r.x, r.y, r.width, r.height); }
  Mat rowRange(int a, int b) const { return roi(0, a, cols, b - a); }
  Mat colRange(int a, int b) const { return roi(a, 0, b - a, rows); }
  Mat clone() const { Mat m(rows, cols, 0); for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), cols); return m; }
  static ZerosExpr zeros(int r, int c, int) { return ZerosExpr{r, c}; }
  Mat(const ZerosExpr &e) { *this = e; }
  Mat &operator=(const ZerosExpr &e) {
    create(e.rows, e.cols, 0);
    for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, cols);
    return *this;
  }
 private:
  std::shared_ptr<uchar> own_;
};

class _InputArray {
 public:
  _InputArray() {}
  _InputArray(const Mat &m) : m_(&m) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  bool empty() const { return !m_ || m_->empty(); }
 private:
  const Mat *m_ = nullptr;
};
class _OutputArray {
 public:
  _OutputArray(Mat &m) : m_(&m) {}
  void create(int r, int c, int t) const { m_->create(r, c, t); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
 private:
  Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

inline float fastAtan2(float y, float x) { return orb_oracle_fast_atan2(y, x); }

// cv::resize into a pre-sized destination (the reference passes an ROI of its padded plane, ORBextractor.cc:1054)
inline void resize(const Mat &src, Mat &dst, Size dsize, double, double, int) {
  dst.create(dsize.height, dsize.width, 0);
  orb_oracle_resize(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
// cv::copyMakeBorder, BORDER_REFLECT_101 (+ISOLATED); in place when src is the centre ROI of dst (ORBextractor.cc:1057-1061)
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int) {
  dst.create(src.rows + top + bottom, src.cols + left + right, 0);
  for (int y = 0; y < src.rows; y++)
    if (dst.ptr(y + top) + left != src.ptr(y)) std::memmove(dst.ptr(y + top) + left, src.ptr(y), src.cols);
  for (int y = 0; y < dst.rows; y++) {
    const uchar *S = dst.ptr(reflect101(y - top, src.rows) + top) + left;
    uchar *D = dst.ptr(y);
    for (int x = 0; x < dst.cols; x++) {
      if (y >= top && y < top + src.rows && x >= left && x < left + src.cols) { x = left + src.cols - 1; continue; }
      D[x] = S[reflect101(x - left, src.cols)];
    }
  }
}

inline void GaussianBlur(const Mat &src, Mat &dst, Size, double, double, int) {
  Mat tmp = src.clone();
  dst.create(src.rows, src.cols, 0);
  orb_oracle_blur(tmp.data, tmp.cols, tmp.rows, (int)tmp.step.v, dst.data, (int)dst.step.v);
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression = true), TYPE_9_16
inline void FAST(const Mat &img, std::vector<KeyPoint> &kps, int threshold, bool /*nms*/) {
  std::vector<int> buf((size_t)(img.rows * img.cols / 2 + 16) * 3);
  const int n = orb_oracle_fast(img.data, img.cols, img.rows, (int)img.step.v, threshold, buf.data(), (int)buf.size() / 3);
  kps.clear();
  for (int i = 0; i < n; i++) kps.push_back(KeyPoint((float)buf[3 * i], (float)buf[3 * i + 1], 7.f, -1.f, (float)buf[3 * i + 2]));
}

struct KeyPointsFilter {   // only reached from ComputeKeyPointsOld, which the reference never calls (ORBextractor.cc:992)
  static void retainBest(std::vector<KeyPoint> &k, int n) {
    if (n >= 0 && (int)k.size() > n) {
      std::nth_element(k.begin(), k.begin() + n, k.end(), [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
      k.resize(n);
    }
  }
};

}  // namespace cv
