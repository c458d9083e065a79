// Replacement body for Frame::ComputeStereoMatches (Source/Libraries/ORB_SLAM2/src/Frame.cc:477-646):
//
//   void Frame::ComputeStereoMatches() {
//     ORB_SLAM2::ComputeStereoMatchesGPU(*mpORBextractorLeft, *mpORBextractorRight, mbf, mbf / fx, N, mvuRight, mvDepth);
//   }
//
// It runs on the keypoints / descriptors / pyramids the two extractors left in HBM, so it must be
// called after both operator() calls of the same frame (Frame.cc:87-99 already orders it so).
// `mb` is passed explicitly: the reference reads the member before assigning it (Frame.cc:505 vs :124).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBextractor.h"

namespace ORB_SLAM2 {

inline void ComputeStereoMatchesGPU(ORBextractor &left, ORBextractor &right, float mbf, float mb, int N,
                                    std::vector<float> &mvuRight, std::vector<float> &mvDepth) {
  mvuRight.assign(N, -1.0f);
  mvDepth.assign(N, -1.0f);
  if (N == 0) return;
  const int cap = orbgpu_extractor_capacity(left.Handle());
  std::vector<float> u(cap), z(cap);
  if (orbgpu_stereo_match(left.Handle(), right.Handle(), 0, 0, 1, mbf, mb, u.data(), z.data(), cap) < 0)
    throw std::runtime_error(std::string("orbgpu_stereo_match: ") + orbgpu_last_error());
  std::copy(u.begin(), u.begin() + N, mvuRight.begin());
  std::copy(z.begin(), z.begin() + N, mvDepth.begin());
}

}  // namespace ORB_SLAM2
