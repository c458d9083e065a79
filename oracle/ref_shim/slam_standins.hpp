// slam_standins.hpp — TEST INFRASTRUCTURE, force-included (-include) when the reference's own Frame.cc, KeyFrame.cc,
// MapPoint.cc, Map.cc and ORBmatcher.cc are compiled into oracle/_ref/libref_slam.so.  The three headers below pull
// in g2o / Eigen / the DBoW2 vocabulary, which do not exist in this container; their include guards are pre-defined
// here and minimal declarations take their place.  None of them is on the ORB front-end path: the vocabulary
// transform is never called (the harness fills mFeatVec directly), the keyframe database is only told about erased
// keyframes, and Converter is only used by ComputeBoW.
#pragma once
#define CONVERTER_H
#define ORBVOCABULARY_H
#define KEYFRAMEDATABASE_H
#include <climits>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>
#include "DBoW2/DBoW2.h"
#include "cvshim.hpp"
namespace ORB_SLAM2 {
class KeyFrame;
class ORBVocabulary {
 public:
  void transform(const std::vector<cv::Mat> &, DBoW2::BowVector &, DBoW2::FeatureVector &, int) {}
  double score(const DBoW2::BowVector &, const DBoW2::BowVector &) const { return 0; }
};
class KeyFrameDatabase {
 public:
  void erase(KeyFrame *) {}
};
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat &d) {
    std::vector<cv::Mat> v;
    for (int i = 0; i < d.rows; i++) v.push_back(d.row(i));
    return v;
  }
};
}  // namespace ORB_SLAM2
