// Replacement bodies for Frame::UndistortKeyPoints and the distorted branch of Frame::ComputeImageBounds
// (Source/Libraries/ORB_SLAM2/src/Frame.cc:419-445, 447-475):
//
//   void Frame::UndistortKeyPoints() { ORB_SLAM2::UndistortKeyPointsGPU(mvKeys, mK, mDistCoef, mvKeysUn); }
//   ... in ComputeImageBounds:  ORB_SLAM2::UndistortPointsGPU(corners /* 4 x (x, y) */, 4, mK, mDistCoef);
//
// cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) becomes one call into liborbgpu.so (five double-precision
// fixed-point iterations per point in OpenCV 4.x's operation order, bit-exact against cv2).  A frame whose feature set is
// built with orbgpu_features_from_extractor_undistorted never brings the distorted keypoints to the host at all; this
// header is the source-compatible path for code that keeps mvKeysUn on the host (Tracking, Initializer, PnPsolver read it).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "orbgpu.h"

namespace ORB_SLAM2 {

inline orbgpu_camera CameraFrom(const cv::Mat &K, const cv::Mat &distCoef) {
  orbgpu_camera c{};
  c.fx = K.at<float>(0, 0); c.fy = K.at<float>(1, 1); c.cx = K.at<float>(0, 2); c.cy = K.at<float>(1, 2);
  c.ndist = distCoef.rows * distCoef.cols >= 5 ? 5 : 4;           // Tracking.cc resizes mDistCoef to 5 when k3 != 0
  for (int i = 0; i < c.ndist; i++) c.dist[i] = distCoef.at<float>(i);
  return c;
}

inline void UndistortKeyPointsGPU(const std::vector<cv::KeyPoint> &keys, const cv::Mat &K, const cv::Mat &distCoef,
                                  std::vector<cv::KeyPoint> &keysUn) {
  static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint), "cv::KeyPoint must be the 28-byte record of orbgpu.h");
  keysUn.resize(keys.size());
  if (keys.empty()) return;
  const orbgpu_camera c = CameraFrom(K, distCoef);
  if (orbgpu_undistort_keypoints(reinterpret_cast<const orbgpu_keypoint *>(keys.data()), (int)keys.size(), &c,
                                 reinterpret_cast<orbgpu_keypoint *>(keysUn.data()), -1) != ORBGPU_OK)
    throw std::runtime_error(std::string("orbgpu_undistort_keypoints: ") + orbgpu_last_error());
}

inline void UndistortPointsGPU(float *xy, int n, const cv::Mat &K, const cv::Mat &distCoef) {
  const orbgpu_camera c = CameraFrom(K, distCoef);
  if (orbgpu_undistort_points(xy, n, &c, xy, -1) != ORBGPU_OK)
    throw std::runtime_error(std::string("orbgpu_undistort_points: ") + orbgpu_last_error());
}

}  // namespace ORB_SLAM2
