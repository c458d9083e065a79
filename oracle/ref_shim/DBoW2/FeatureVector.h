// TEST INFRASTRUCTURE — stand-in for DBoW2::FeatureVector (…/DBoW2/FeatureVector.h): std::map<NodeId, vector<unsigned>>,
// which is all ORBmatcher.cc relies on (ordered iteration, lower_bound; ORBmatcher.cc:178-250).
#pragma once
#include <map>
#include <vector>
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
 public:
  void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
