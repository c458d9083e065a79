// TEST INFRASTRUCTURE.  The slice of ORB_SLAM2::KeyFrame that src/KeyFrameDatabase.cc touches (include/KeyFrame.h:37-240):
// the BowVector, the loop / relocalisation query bookkeeping, and the two covisibility-graph getters, whose answers the test
// supplies (the graph itself is control plane, out of scope).  Lets the reference's own, unmodified KeyFrameDatabase.cc and
// the drop-in shim/KeyFrameDatabase.cc compile against the real DBoW2 without Frame / Map / MapPoint / g2o.
#ifndef KEYFRAME_H
#define KEYFRAME_H
#include <set>
#include <vector>

#include "DBoW2/BowVector.h"

namespace ORB_SLAM2 {
class KeyFrame {
public:
  long unsigned int mnId = 0;
  DBoW2::BowVector mBowVec;
  long unsigned int mnLoopQuery = 0;     // KeyFrame.cc:37-43 initialises all six to 0
  int mnLoopWords = 0;
  float mLoopScore = 0;
  long unsigned int mnRelocQuery = 0;
  int mnRelocWords = 0;
  float mRelocScore = 0;
  std::set<KeyFrame *> connected;
  std::vector<KeyFrame *> covisible;     // best first
  std::set<KeyFrame *> GetConnectedKeyFrames() { return connected; }
  std::vector<KeyFrame *> GetBestCovisibilityKeyFrames(const int &N) {
    return (int)covisible.size() < N ? covisible : std::vector<KeyFrame *>(covisible.begin(), covisible.begin() + N);
  }
};
}  // namespace ORB_SLAM2
#endif
