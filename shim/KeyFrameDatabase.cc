// shim/KeyFrameDatabase.cc — drop-in replacement for the reference's src/KeyFrameDatabase.cc (Source/Libraries/ORB_SLAM2).
//
// Compiled against the reference's own, unchanged include/KeyFrameDatabase.h, so LoopClosing, Tracking (relocalisation) and
// LocalMapping link against it as they do against the original.  What changes is where the candidate scoring runs:
//
//   device   every keyframe's BowVector lives in HBM (orbgpu_bowdb, include/orbgpu.h); ONE orbgpu_bowdb_query per detection
//            returns for all keyframes the number of words shared with the query, the first shared word and DBoW2's L1 score —
//            the inverted-file walk (:81-99, :195-213) and the mpVoc->score loop (:118-132, :231-241) of the reference;
//   host     the reference's bookkeeping on the KeyFrame objects (mnLoopQuery / mnLoopWords / mLoopScore and the reloc
//            twins, including its quirks), the 0.8 x max-words filter, and the covisibility accumulation, which walks the
//            KeyFrame pointer graph (:140-185, :248-300).
//
// The order of lKFsSharingWords — the order in which the walk meets each keyframe first — is rebuilt from (first shared
// word, slot): words are visited ascending, and a word's list holds its keyframes in add() order = slot order.
// No CPU fallback: a failing device call throws std::runtime_error.  Only L1 scoring (the scoring of ORBvoc.txt and of
// ORBVocabulary's default) is implemented; another scoring type throws at construction.
//
// Line references ":NNN" are to the reference's src/KeyFrameDatabase.cc.
#include "KeyFrameDatabase.h"

#include <algorithm>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "DBoW2/BowVector.h"
#include "KeyFrame.h"
#include "orbgpu.h"

using std::vector;

namespace ORB_SLAM2 {

namespace {

// The reference's header has no room for the device handle: per-database state lives in a side table.  (The class has no
// destructor to hook either; a database's device memory is released by clear() and at process exit.)
struct DeviceState {
  orbgpu_bowdb *db = nullptr;
  std::map<KeyFrame *, int> slotOf;
  vector<KeyFrame *> bySlot;               // nullptr once erased
};
std::mutex g_tableMutex;
std::map<const KeyFrameDatabase *, DeviceState> g_table;

[[noreturn]] void fail(const char *what) {
  throw std::runtime_error(std::string("orbgpu KeyFrameDatabase: ") + what + ": " + orbgpu_last_error());
}

DeviceState &state_of(const KeyFrameDatabase *self) {
  std::lock_guard<std::mutex> lock(g_tableMutex);
  DeviceState &S = g_table[self];
  if (!S.db && orbgpu_bowdb_create(-1, &S.db) != ORBGPU_OK) fail("orbgpu_bowdb_create");
  return S;
}

void flatten(const DBoW2::BowVector &v, vector<uint32_t> &ids, vector<double> &weights) {
  ids.clear(); weights.clear();
  ids.reserve(v.size()); weights.reserve(v.size());
  for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it) { ids.push_back((uint32_t)it->first); weights.push_back(it->second); }
}

struct QueryResult { vector<int32_t> common; vector<uint32_t> first; vector<double> score; vector<int> order; };

// One device call; `order` = the slots that share a word with the query, as the inverted-file walk meets them first.
void query(DeviceState &S, const DBoW2::BowVector &bow, QueryResult &R) {
  vector<uint32_t> ids;
  vector<double> weights;
  flatten(bow, ids, weights);
  const int slots = orbgpu_bowdb_slots(S.db);
  R.common.assign(slots, 0); R.first.assign(slots, 0xffffffffu); R.score.assign(slots, 0.0);
  if (orbgpu_bowdb_query(S.db, ids.data(), weights.data(), (int)ids.size(), R.common.data(), R.first.data(), R.score.data(), slots) != ORBGPU_OK)
    fail("orbgpu_bowdb_query");
  R.order.clear();
  for (int s = 0; s < slots; s++)
    if (R.common[s] > 0 && S.bySlot[s]) R.order.push_back(s);
  std::sort(R.order.begin(), R.order.end(), [&](int a, int b) { return R.first[a] != R.first[b] ? R.first[a] < R.first[b] : a < b; });
}

}  // namespace

KeyFrameDatabase::KeyFrameDatabase(const ORBVocabulary &voc) : mpVoc(&voc) {
  mvInvertedFile.resize(voc.size());      // :34-36; kept for layout only, the words live on the device
  if (voc.getScoringType() != DBoW2::L1_NORM) throw std::runtime_error("orbgpu KeyFrameDatabase: only L1_NORM scoring is implemented on the device");
}

void KeyFrameDatabase::add(KeyFrame *pKF) {   // :38-46
  std::unique_lock<std::mutex> lock(mMutex);
  DeviceState &S = state_of(this);
  vector<uint32_t> ids;
  vector<double> weights;
  flatten(pKF->mBowVec, ids, weights);
  const int slot = orbgpu_bowdb_add(S.db, ids.data(), weights.data(), (int)ids.size());
  if (slot < 0) fail("orbgpu_bowdb_add");
  // the reference pushes the keyframe into its words' lists again if it is added twice; the device keeps one entry per add as
  // well, and erase() removes the first one, like list::erase on the first hit (:57-63)
  S.bySlot.resize(slot + 1, nullptr);
  S.bySlot[slot] = pKF;
  if (!S.slotOf.count(pKF)) S.slotOf[pKF] = slot;
}

void KeyFrameDatabase::erase(KeyFrame *pKF) {   // :48-66
  std::unique_lock<std::mutex> lock(mMutex);
  DeviceState &S = state_of(this);
  std::map<KeyFrame *, int>::iterator it = S.slotOf.find(pKF);
  if (it == S.slotOf.end()) return;
  const int slot = it->second;
  if (orbgpu_bowdb_erase(S.db, slot) != ORBGPU_OK) fail("orbgpu_bowdb_erase");
  S.bySlot[slot] = nullptr;
  S.slotOf.erase(it);
  for (int s = slot + 1; s < (int)S.bySlot.size(); s++)            // a second add of the same keyframe takes over
    if (S.bySlot[s] == pKF) { S.slotOf[pKF] = s; break; }
}

void KeyFrameDatabase::clear() {   // :68-71
  mvInvertedFile.clear();
  mvInvertedFile.resize(mpVoc->size());
  DeviceState &S = state_of(this);
  if (orbgpu_bowdb_clear(S.db) != ORBGPU_OK) fail("orbgpu_bowdb_clear");
  S.slotOf.clear();
  S.bySlot.clear();
}

vector<KeyFrame *> KeyFrameDatabase::DetectLoopCandidates(KeyFrame *pKF, float minScore) {
  std::set<KeyFrame *> spConnectedKeyFrames = pKF->GetConnectedKeyFrames();
  std::list<KeyFrame *> lKFsSharingWords;
  QueryResult R;
  vector<int> slotOfCandidate;
  {
    std::unique_lock<std::mutex> lock(mMutex);
    DeviceState &S = state_of(this);
    query(S, pKF->mBowVec, R);
    // :81-99 on the per-keyframe totals.  The reference resets mnLoopWords whenever mnLoopQuery differs, and only marks the
    // keyframe when it is not connected: a connected keyframe therefore ends with mnLoopWords = 1 (reset at every word it
    // shares), a keyframe already marked by an earlier query with the same id keeps counting.
    for (size_t i = 0; i < R.order.size(); i++) {
      KeyFrame *pKFi = S.bySlot[R.order[i]];
      const int words = R.common[R.order[i]];
      if (pKFi->mnLoopQuery != pKF->mnId) {
        if (!spConnectedKeyFrames.count(pKFi)) {
          pKFi->mnLoopQuery = pKF->mnId;
          pKFi->mnLoopWords = words;
          lKFsSharingWords.push_back(pKFi);
          slotOfCandidate.push_back(R.order[i]);
        } else {
          pKFi->mnLoopWords = 1;
        }
      } else {
        pKFi->mnLoopWords += words;
      }
    }
  }
  if (lKFsSharingWords.empty()) return vector<KeyFrame *>();

  int maxCommonWords = 0;                                            // :106-114
  for (std::list<KeyFrame *>::iterator lit = lKFsSharingWords.begin(); lit != lKFsSharingWords.end(); ++lit)
    maxCommonWords = std::max(maxCommonWords, (*lit)->mnLoopWords);
  const int minCommonWords = maxCommonWords * 0.8f;

  std::list<std::pair<float, KeyFrame *>> lScoreAndMatch;            // :118-132
  {
    size_t i = 0;
    for (std::list<KeyFrame *>::iterator lit = lKFsSharingWords.begin(); lit != lKFsSharingWords.end(); ++lit, ++i) {
      KeyFrame *pKFi = *lit;
      if (pKFi->mnLoopWords > minCommonWords) {
        const float si = R.score[slotOfCandidate[i]];                // float si = mpVoc->score(pKF->mBowVec, pKFi->mBowVec)
        pKFi->mLoopScore = si;
        if (si >= minScore) lScoreAndMatch.push_back(std::make_pair(si, pKFi));
      }
    }
  }
  if (lScoreAndMatch.empty()) return vector<KeyFrame *>();

  std::list<std::pair<float, KeyFrame *>> lAccScoreAndMatch;         // :137-166: accumulate over the best covisible keyframes
  float bestAccScore = minScore;
  for (std::list<std::pair<float, KeyFrame *>>::iterator it = lScoreAndMatch.begin(); it != lScoreAndMatch.end(); ++it) {
    KeyFrame *pKFi = it->second;
    vector<KeyFrame *> vpNeighs = pKFi->GetBestCovisibilityKeyFrames(10);
    float bestScore = it->first, accScore = it->first;
    KeyFrame *pBestKF = pKFi;
    for (size_t k = 0; k < vpNeighs.size(); k++) {
      KeyFrame *pKF2 = vpNeighs[k];
      if (pKF2->mnLoopQuery == pKF->mnId && pKF2->mnLoopWords > minCommonWords) {
        accScore += pKF2->mLoopScore;
        if (pKF2->mLoopScore > bestScore) { pBestKF = pKF2; bestScore = pKF2->mLoopScore; }
      }
    }
    lAccScoreAndMatch.push_back(std::make_pair(accScore, pBestKF));
    if (accScore > bestAccScore) bestAccScore = accScore;
  }

  const float minScoreToRetain = 0.75f * bestAccScore;               // :169-185
  std::set<KeyFrame *> spAlreadyAddedKF;
  vector<KeyFrame *> vpLoopCandidates;
  vpLoopCandidates.reserve(lAccScoreAndMatch.size());
  for (std::list<std::pair<float, KeyFrame *>>::iterator it = lAccScoreAndMatch.begin(); it != lAccScoreAndMatch.end(); ++it)
    if (it->first > minScoreToRetain && spAlreadyAddedKF.insert(it->second).second) vpLoopCandidates.push_back(it->second);
  return vpLoopCandidates;
}

vector<KeyFrame *> KeyFrameDatabase::DetectRelocalizationCandidates(Frame *F) {
  std::list<KeyFrame *> lKFsSharingWords;
  QueryResult R;
  vector<int> slotOfCandidate;
  {
    std::unique_lock<std::mutex> lock(mMutex);
    DeviceState &S = state_of(this);
    query(S, F->mBowVec, R);
    for (size_t i = 0; i < R.order.size(); i++) {                    // :195-213
      KeyFrame *pKFi = S.bySlot[R.order[i]];
      const int words = R.common[R.order[i]];
      if (pKFi->mnRelocQuery != F->mnId) {
        pKFi->mnRelocWords = words;
        pKFi->mnRelocQuery = F->mnId;
        lKFsSharingWords.push_back(pKFi);
        slotOfCandidate.push_back(R.order[i]);
      } else {
        pKFi->mnRelocWords += words;
      }
    }
  }
  if (lKFsSharingWords.empty()) return vector<KeyFrame *>();

  int maxCommonWords = 0;                                            // :218-226
  for (std::list<KeyFrame *>::iterator lit = lKFsSharingWords.begin(); lit != lKFsSharingWords.end(); ++lit)
    maxCommonWords = std::max(maxCommonWords, (*lit)->mnRelocWords);
  const int minCommonWords = maxCommonWords * 0.8f;

  std::list<std::pair<float, KeyFrame *>> lScoreAndMatch;            // :231-241
  {
    size_t i = 0;
    for (std::list<KeyFrame *>::iterator lit = lKFsSharingWords.begin(); lit != lKFsSharingWords.end(); ++lit, ++i) {
      KeyFrame *pKFi = *lit;
      if (pKFi->mnRelocWords > minCommonWords) {
        const float si = R.score[slotOfCandidate[i]];
        pKFi->mRelocScore = si;
        lScoreAndMatch.push_back(std::make_pair(si, pKFi));
      }
    }
  }
  if (lScoreAndMatch.empty()) return vector<KeyFrame *>();

  std::list<std::pair<float, KeyFrame *>> lAccScoreAndMatch;         // :246-275
  float bestAccScore = 0;
  for (std::list<std::pair<float, KeyFrame *>>::iterator it = lScoreAndMatch.begin(); it != lScoreAndMatch.end(); ++it) {
    KeyFrame *pKFi = it->second;
    vector<KeyFrame *> vpNeighs = pKFi->GetBestCovisibilityKeyFrames(10);
    float bestScore = it->first, accScore = bestScore;
    KeyFrame *pBestKF = pKFi;
    for (size_t k = 0; k < vpNeighs.size(); k++) {
      KeyFrame *pKF2 = vpNeighs[k];
      if (pKF2->mnRelocQuery != F->mnId) continue;
      accScore += pKF2->mRelocScore;
      if (pKF2->mRelocScore > bestScore) { pBestKF = pKF2; bestScore = pKF2->mRelocScore; }
    }
    lAccScoreAndMatch.push_back(std::make_pair(accScore, pBestKF));
    if (accScore > bestAccScore) bestAccScore = accScore;
  }

  const float minScoreToRetain = 0.75f * bestAccScore;               // :278-300
  std::set<KeyFrame *> spAlreadyAddedKF;
  vector<KeyFrame *> vpRelocCandidates;
  vpRelocCandidates.reserve(lAccScoreAndMatch.size());
  for (std::list<std::pair<float, KeyFrame *>>::iterator it = lAccScoreAndMatch.begin(); it != lAccScoreAndMatch.end(); ++it)
    if (it->first > minScoreToRetain && spAlreadyAddedKF.insert(it->second).second) vpRelocCandidates.push_back(it->second);
  return vpRelocCandidates;
}

}  // namespace ORB_SLAM2
