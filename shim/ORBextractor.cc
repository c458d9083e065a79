// See ORBextractor.h.  Error behaviour: the reference asserts / lets OpenCV throw; this shim throws
// std::runtime_error carrying orbgpu_last_error() and never exits the process.
#include "ORBextractor.h"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint) && sizeof(cv::KeyPoint) == 28,
              "cv::KeyPoint must be the 7 x 4-byte record the C ABI writes");

static void Check(int rc, const char *what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + orbgpu_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  mvImagePyramid.resize(nlevels);
  // The scale / sigma / quota tables of ORBextractor.cc:411-441 come from the library, so that the
  // getters work before the first image exactly as in the reference.
  EnsureHandle(640, 480);
}

ORBextractor::~ORBextractor() { orbgpu_extractor_destroy(mpHandle); }

void ORBextractor::EnsureHandle(int width, int height) {
  if (mpHandle && width <= mMaxWidth && height <= mMaxHeight) return;
  mMaxWidth = std::max(mMaxWidth, width);
  mMaxHeight = std::max(mMaxHeight, height);
  orbgpu_extractor_destroy(mpHandle);
  mpHandle = nullptr;
  orbgpu_extractor_params p{};
  p.nfeatures = nfeatures; p.scale_factor = (float)scaleFactor; p.nlevels = nlevels;
  p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
  p.max_width = mMaxWidth; p.max_height = mMaxHeight; p.max_batch = 1; p.device = -1;
  Check(orbgpu_extractor_create(&p, &mpHandle), "orbgpu_extractor_create");
  mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
  Check(orbgpu_extractor_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "orbgpu_extractor_tables");
  const int cap = orbgpu_extractor_capacity(mpHandle);
  mKpBuffer.resize(cap);
  mDescBuffer.resize((size_t)cap * 32);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                              cv::OutputArray _descriptors) {
  if (_image.empty()) return;                                   // ORBextractor.cc:981-982
  cv::Mat image = _image.getMat();
  if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");   // :985 asserts
  EnsureHandle(image.cols, image.rows);
  int n = 0;
  Check(orbgpu_extract(mpHandle, image.ptr(0), image.cols, image.rows, image.step, mKpBuffer.data(), mDescBuffer.data(),
                       (int)mKpBuffer.size(), &n), "orbgpu_extract");
  _keypoints.clear();                                           // ORBextractor.cc:1006
  if (n == 0) {
    _descriptors.release();                                     // ORBextractor.cc:1000
  } else {
    _descriptors.create(n, 32, CV_8U);                          // ORBextractor.cc:1002
    cv::Mat d = _descriptors.getMat();
    for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &mDescBuffer[(size_t)i * 32], 32);
    _keypoints.resize(n);
    std::memcpy(static_cast<void *>(_keypoints.data()), mKpBuffer.data(), (size_t)n * sizeof(cv::KeyPoint));
  }
  if (mbDownloadPyramid) {
    for (int l = 0; l < nlevels; l++) {
      int w = 0, h = 0;
      Check(orbgpu_level_size(mpHandle, l, &w, &h), "orbgpu_level_size");
      mvImagePyramid[l].create(h, w, CV_8UC1);
      Check(orbgpu_get_pyramid_level(mpHandle, 0, l, 0, mvImagePyramid[l].ptr(0), mvImagePyramid[l].step),
            "orbgpu_get_pyramid_level");
    }
  }
}

}  // namespace ORB_SLAM2
