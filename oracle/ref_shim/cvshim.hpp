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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

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
struct ZerosExpr { int rows, cols, type; };

// Element types: only CV_8U (ORBextractor.cc, image patches) and CV_32F (poses, points, patches in Frame.cc /
// KeyFrame.cc / MapPoint.cc / ORBmatcher.cc) occur on the path.  Arithmetic is eager (no MatExpr fusion): products
// accumulate in double and round once to float, like cv::gemm's float kernels; dot / norm return double.
class Mat {
 public:
  struct Step { size_t v; operator size_t() const { return v; } };
  int rows = 0, cols = 0;
  Step step{0};
  uchar *data = nullptr;
  Mat() {}
  Mat(Size sz, int t) { create(sz.height, sz.width, t); }
  Mat(int r, int c, int t) { create(r, c, t); }
  Mat(int r, int c, int t, void *ext, size_t s) : rows(r), cols(c), step{s}, data((uchar *)ext), type_(t) {}
  Mat(int r, int c, int t, void *ext) : rows(r), cols(c), step{(size_t)c * esz(t)}, data((uchar *)ext), type_(t) {}
  static int esz(int t) { return t == CV_32F ? 4 : 1; }
  void create(int r, int c, int t) {
    if (data && r == rows && c == cols && t == type_) return;     // cv::Mat::create keeps a matching buffer (ROI included)
    rows = r; cols = c; type_ = t; step.v = (size_t)c * esz(t);
    own_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step.v + 64], std::default_delete<uchar[]>());
    data = own_.get();
  }
  void release() { own_.reset(); data = nullptr; rows = cols = 0; step.v = 0; }
  bool empty() const { return !data || !rows || !cols; }
  int type() const { return type_; }
  size_t step1() const { return step.v / esz(type_); }
  bool isContinuous() const { return rows <= 1 || step.v == (size_t)cols * esz(type_); }
  template <class T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step.v + c * sizeof(T)); }
  template <class T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step.v + c * sizeof(T)); }
  // single-index access as cv::Mat::at(int i0): linear for continuous / row vectors, strided for column vectors
  template <class T> T &at(int i) { return const_cast<T &>(static_cast<const Mat *>(this)->at<T>(i)); }
  template <class T> const T &at(int i) const {
    if (isContinuous() || rows == 1) return ((const T *)data)[i];
    if (cols == 1) return *(const T *)(data + (size_t)i * step.v);
    return *(const T *)(data + (size_t)(i / cols) * step.v + (i % cols) * sizeof(T));
  }
  uchar *ptr(int r = 0) { return data + (size_t)r * step.v; }
  const uchar *ptr(int r = 0) const { return data + (size_t)r * step.v; }
  template <class T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step.v); }
  template <class T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step.v); }
  Mat roi(int x, int y, int w, int h) const {
    Mat m; m.rows = h; m.cols = w; m.step = step; m.type_ = type_; m.data = data + (size_t)y * step.v + (size_t)x * esz(type_); m.own_ = own_;
    return m;
  }
  Mat operator()(const Rect &r) const { return roi(r.x, r.y, r.width, r.height); }
  Mat rowRange(int a, int b) const { return roi(0, a, cols, b - a); }
  Mat colRange(int a, int b) const { return roi(a, 0, b - a, rows); }
  Mat row(int r) const { return roi(0, r, cols, 1); }
  Mat col(int c) const { return roi(c, 0, 1, rows); }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_));
    return m;
  }
  void copyTo(Mat &dst) const { dst.create(rows, cols, type_); for (int r = 0; r < rows; r++) std::memmove(dst.ptr(r), ptr(r), (size_t)cols * esz(type_)); }
  void copyTo(Mat &&dst) const { copyTo(dst); }     // into an ROI header of matching size: written in place
  Mat reshape(int) const { return *this; }           // only on the cv::undistortPoints path, which this stand-in does not provide
  void convertTo(Mat &dst, int t) const {
    Mat out(rows, cols, t);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        const float v = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
        if (t == CV_32F) out.at<float>(r, c) = v; else out.at<uchar>(r, c) = (uchar)std::min(255.f, std::max(0.f, std::nearbyintf(v)));
      }
    dst = out;
  }
  Mat t() const { Mat m(cols, rows, type_); for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c); return m; }
  double dot(const Mat &o) const {
    double s = 0;
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
    return s;
  }
  static ZerosExpr zeros(int r, int c, int t) { return ZerosExpr{r, c, t}; }
  static Mat ones(int r, int c, int t) { Mat m(r, c, t); for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = 1.f; return m; }
  static Mat eye(int r, int c, int t) { Mat m = ZerosExpr{r, c, t}; for (int i = 0; i < std::min(r, c); i++) m.at<float>(i, i) = 1.f; return m; }
  Mat(const ZerosExpr &e) { *this = e; }
  Mat &operator=(const ZerosExpr &e) {
    create(e.rows, e.cols, e.type);
    for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, (size_t)cols * esz(type_));
    return *this;
  }
 private:
  std::shared_ptr<uchar> own_;
  int type_ = CV_8U;
};

template <class T> class Mat_ : public Mat {         // only the comma initialiser (cv::Mat_<float>(3,1) << x, y, z) is used
 public:
  Mat_(int r, int c) : Mat(r, c, CV_32F) {}
};
template <class T> struct MatCommaInit {
  Mat m; int i;
  MatCommaInit &operator,(double v) { m.at<T>(i / m.cols, i % m.cols) = (T)v; i++; return *this; }
  operator Mat() const { return m; }
};
template <class T> inline MatCommaInit<T> operator<<(const Mat_<T> &m, double v) { MatCommaInit<T> ci{m, 0}; return (ci, v); }

inline Mat matBinary(const Mat &a, const Mat &b, int op) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = op ? a.at<float>(r, c) - b.at<float>(r, c) : a.at<float>(r, c) + b.at<float>(r, c);
  return m;
}
inline Mat operator+(const Mat &a, const Mat &b) { return matBinary(a, b, 0); }
inline Mat operator-(const Mat &a, const Mat &b) { return matBinary(a, b, 1); }
inline Mat matScale(const Mat &a, double s) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = (float)((double)a.at<float>(r, c) * s);
  return m;
}
inline Mat operator-(const Mat &a) { return matScale(a, -1.0); }
inline Mat operator*(const Mat &a, double s) { return matScale(a, s); }
inline Mat operator*(double s, const Mat &a) { return matScale(a, s); }
inline Mat operator/(const Mat &a, double s) { return matScale(a, 1.0 / s); }
inline Mat operator*(const Mat &a, const Mat &b) {
  Mat m(a.rows, b.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < b.cols; c++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += (double)a.at<float>(r, k) * (double)b.at<float>(k, c);
      m.at<float>(r, c) = (float)s;
    }
  return m;
}
enum { NORM_L1 = 2, NORM_L2 = 4 };
inline double norm(const Mat &a) { return std::sqrt(a.dot(a)); }
inline double norm(const Mat &a, const Mat &b, int /*NORM_L1*/) {
  double s = 0;
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
  return s;
}
inline void undistortPoints(const Mat &, Mat &, const Mat &, const Mat &, const Mat &, const Mat &) {
  fprintf(stderr, "cvshim: cv::undistortPoints is not provided (use zero distortion coefficients)\n");
  abort();
}

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

// cv::FileStorage / cv::FileNode: only so that DBoW2's YAML save / load members (virtual, hence instantiated) compile;
// the checker loads vocabularies through the reference's own text / binary loaders (ORBVocabulary.cc) instead.
struct FileNode;
struct FileNodeIterator {
  int i = 0;
  FileNodeIterator &operator++() { i++; return *this; }
  bool operator!=(const FileNodeIterator &o) const { return i != o.i; }
  inline FileNode operator*() const;
};
struct FileNode {
  FileNodeIterator begin() const { return FileNodeIterator(); }
  FileNodeIterator end() const { return FileNodeIterator(); }
  FileNode operator[](const std::string &) const { return FileNode(); }
  FileNode operator[](const char *) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0; }
  operator std::string() const { return std::string(); }
};
inline FileNode FileNodeIterator::operator*() const { return FileNode(); }
struct FileStorage {
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string &, int) { fprintf(stderr, "cvshim: cv::FileStorage is not provided\n"); abort(); }
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string &) const { return FileNode(); }
  template <class T> FileStorage &operator<<(const T &) { return *this; }
};

struct KeyPointsFilter {   // only reached from ComputeKeyPointsOld, which the reference never calls (ORBextractor.cc:992)
  static void retainBest(std::vector<KeyPoint> &k, int n) {
    if (n >= 0 && (int)k.size() > n) {
      std::nth_element(k.begin(), k.begin() + n, k.end(), [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
      k.resize(n);
    }
  }
};

}  // namespace cv
