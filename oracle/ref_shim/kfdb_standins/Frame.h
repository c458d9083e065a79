// TEST INFRASTRUCTURE.  The slice of ORB_SLAM2::Frame that src/KeyFrameDatabase.cc touches (see KeyFrame.h beside this file).
#ifndef FRAME_H
#define FRAME_H
#include "DBoW2/BowVector.h"

namespace ORB_SLAM2 {
class Frame {
public:
  long unsigned int mnId = 0;
  DBoW2::BowVector mBowVec;
};
}  // namespace ORB_SLAM2
#endif
