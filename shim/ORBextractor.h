// Drop-in replacement for Source/Libraries/ORB_SLAM2/include/ORBextractor.h of
// sjulier/Refactored_ORB_SLAM2: same class name, constructor, operator(), getters and public
// mvImagePyramid (ORBextractor.h:47-75), so Frame / Tracking compile unchanged.  The body forwards
// to liborbgpu.so (include/orbgpu.h); ExtractorNode and the protected helpers of the reference
// have no counterpart here — their work happens inside the CUDA kernels.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#ifdef ORBGPU_CV_COMPAT
#include "cv_compat.h"
#else
#include <opencv2/core.hpp>
#endif

#include "../include/orbgpu.h"

namespace ORB_SLAM2 {

class ORBextractor {
public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  ORBextractor(const ORBextractor &) = delete;
  ORBextractor &operator=(const ORBextractor &) = delete;

  // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                  cv::OutputArray descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // Host copy of the pyramid of the last image (ORBextractor.h:75).  Only
  // Frame::ComputeStereoMatches reads it in the reference (Frame.cc:483-587); when that is replaced
  // by ComputeStereoMatchesGPU (shim/FrameStereo.h) the download can be switched off.
  std::vector<cv::Mat> mvImagePyramid;
  void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }

  // The C-ABI handle, for ComputeStereoMatchesGPU and device-resident consumers.
  orbgpu_extractor *Handle() { return mpHandle; }

protected:
  void EnsureHandle(int width, int height);

  int nfeatures;
  double scaleFactor;
  int nlevels;
  int iniThFAST;
  int minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

  orbgpu_extractor *mpHandle = nullptr;
  int mMaxWidth = 0, mMaxHeight = 0;
  bool mbDownloadPyramid = true;
  std::vector<orbgpu_keypoint> mKpBuffer;
  std::vector<uint8_t> mDescBuffer;
};

}  // namespace ORB_SLAM2

#endif
