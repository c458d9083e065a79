// cv_compat.h — the handful of OpenCV types the ORBextractor / Frame interfaces mention, for
// building the shim where OpenCV's C++ headers are absent (this container).  With real OpenCV,
// compile WITHOUT -DORBGPU_CV_COMPAT and these definitions are never seen.  Layouts match
// OpenCV's where the ABI depends on them (cv::KeyPoint = 7 x 4 bytes).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

struct Point2f { float x = 0, y = 0; };

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

// Row-major, reference-counted 8-bit matrix with a row step: enough for images and descriptors.
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t *data = nullptr;
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
  Mat(int r, int c, int /*type*/, void *ext, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((uint8_t *)ext) {}
  void create(int r, int c, int /*type*/) {
    if (r == rows && c == cols && own_) return;
    rows = r; cols = c; step = (size_t)c;
    own_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c + 1], std::default_delete<uint8_t[]>());
    data = own_.get();
  }
  void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  uint8_t *ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t *ptr(int r = 0) const { return data + (size_t)r * step; }
  Mat row(int r) const { Mat m(1, cols, CV_8U, (void *)ptr(r), step); m.own_ = own_; return m; }
 private:
  std::shared_ptr<uint8_t> own_;
};

// InputArray / OutputArray as thin views over Mat (what the reference passes at every call site).
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
  Mat &ref() const { return *m_; }
 private:
  Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

}  // namespace cv
