// TEST INFRASTRUCTURE — stand-in for DBoW2::BowVector (Source/ThirdParty/DBoW2/DBoW2-local/include/DBoW2/BowVector.h):
// the type is a std::map<WordId, WordValue>; nothing on the ORB front-end path calls its methods.
#pragma once
#include <map>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
}  // namespace DBoW2
