// shim/ORBmatcher.cc — drop-in replacement for the reference's src/ORBmatcher.cc (Source/Libraries/ORB_SLAM2).
//
// Compiled against the reference's own, unchanged include/ORBmatcher.h (and Frame.h / KeyFrame.h / MapPoint.h), so
// Tracking, LocalMapping, LoopClosing, MapPoint and Frame link against it as they do against the original.  Every
// search keeps its signature and its observable behaviour; what changes is where the candidate loops run:
//
//   host, before   what the reference computes per MapPoint / keypoint ahead of its candidate loop — projection with
//                  the caller's pose, image-bounds / depth / viewing-angle gates, PredictScale, the search radius —
//                  collected into one orbgpu_query per surviving query, in the reference's loop order;
//   device         window walk over the frame grid (GetFeaturesInArea) or vocabulary-node list, level / stereo /
//                  chi-square / epipolar gates, DescriptorDistance, best + second best, threshold and ratio test, and
//                  the "an accepted match takes its keypoint away from later queries" dependence — one orbgpu_search
//                  call on the frame's device-resident feature set (include/orbgpu.h);
//   host, after    the pointer-graph updates (mvpMapPoints, AddObservation, Replace …) and the rotation histogram.
//
// Feature sets are uploaded once per Frame / KeyFrame id and kept in a small cache.  There is no CPU fallback: if the
// device call fails the routine throws std::runtime_error.
//
// Line references ":NNN" below are to the reference's src/ORBmatcher.cc.
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <cstring>
#include <list>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

#include "orbgpu.h"

using std::vector;

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;        // :38-40
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

// ---- device-resident feature sets ---------------------------------------------------------------------------------
// Tracking, LocalMapping and LoopClosing call into this file concurrently.  Cache entries are therefore SHARED: a search
// holds a reference to its feature set for the whole device call, and an entry evicted meanwhile by another thread is only
// destroyed when its last user lets go.  (The set itself is read-only after creation; every thread searches through its own
// thread-local workspace and stream inside liborbgpu.)
using FeatureSet = std::shared_ptr<orbgpu_features>;
struct CacheEntry {
  int kind; unsigned long id; int n; uint64_t digest;
  FeatureSet set;
};
std::mutex g_cacheMutex;
std::list<CacheEntry> g_cache;                 // most recently used first
size_t g_cacheEntries = 48;                   // orbgpu_shim_set_feature_cache_capacity

// Digest of EVERYTHING the device copy is built from (every keypoint record, descriptor row and uRight, 64 bits at a time:
// ~140 KB per frame, microseconds): Frame::nNextId restarts at 0 on Tracking::Reset, so (id, n) alone would alias a stale set.
uint64_t content_digest(const vector<cv::KeyPoint> &keys, const cv::Mat &desc, const vector<float> &uRight, const float *geometry, int ngeo) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  auto mix = [&](const void *p, size_t nbytes) {
    const unsigned char *b = (const unsigned char *)p;
    size_t i = 0;
    for (; i + 8 <= nbytes; i += 8) {
      uint64_t w;
      std::memcpy(&w, b + i, 8);
      h = (h ^ w) * 0xff51afd7ed558ccdull;
      h ^= h >> 29;
    }
    uint64_t tail = nbytes - i;
    for (; i < nbytes; i++) tail = (tail << 8) | b[i];
    h = (h ^ tail) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 32;
  };
  const int n = (int)keys.size();
  if (n) mix(keys.data(), (size_t)n * sizeof(cv::KeyPoint));
  if (desc.isContinuous()) { if (n) mix(desc.ptr(0), (size_t)n * 32); }
  else for (int i = 0; i < n; i++) mix(desc.ptr(i), 32);
  if (!uRight.empty()) mix(uRight.data(), uRight.size() * sizeof(float));
  mix(geometry, (size_t)ngeo * sizeof(float));
  return h;
}

[[noreturn]] void fail(const char *what) {
  throw std::runtime_error(std::string("orbgpu ORBmatcher: ") + what + ": " + orbgpu_last_error());
}

FeatureSet resident(int kind, unsigned long id, const vector<cv::KeyPoint> &keysUn, const vector<float> &uRight,
                    const cv::Mat &desc, float minX, float minY, float wInv, float hInv, int cols, int rows, float qMinX,
                    float qMinY) {
  const int n = (int)keysUn.size();
  const float geometry[8] = {minX, minY, wInv, hInv, (float)cols, (float)rows, qMinX, qMinY};
  const uint64_t digest = content_digest(keysUn, desc, uRight, geometry, 8);
  {
    std::lock_guard<std::mutex> lock(g_cacheMutex);
    for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
      if (it->kind == kind && it->id == id && it->n == n && it->digest == digest) {
        g_cache.splice(g_cache.begin(), g_cache, it);
        return g_cache.front().set;
      }
  }
  // upload outside the lock (a device call): two threads that miss on the same frame at once both build a set, the second
  // insert simply shadows the first, whose users keep it alive until they are done
  static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint), "cv::KeyPoint must be the 28-byte record of orbgpu.h");
  vector<unsigned char> rows32((size_t)n * 32);
  for (int i = 0; i < n; i++) std::memcpy(&rows32[(size_t)i * 32], desc.ptr(i), 32);
  orbgpu_features_desc d{};
  d.n = n;
  d.keys_un = reinterpret_cast<const orbgpu_keypoint *>(keysUn.data());
  d.u_right = uRight.empty() ? nullptr : uRight.data();
  d.descriptors = rows32.data();
  d.min_x = minX; d.min_y = minY; d.grid_w_inv = wInv; d.grid_h_inv = hInv; d.grid_cols = cols; d.grid_rows = rows;
  d.query_min_x = qMinX; d.query_min_y = qMinY; d.device = -1;
  orbgpu_features *raw = nullptr;
  if (orbgpu_features_create(&d, &raw) != ORBGPU_OK) fail("orbgpu_features_create");
  FeatureSet set(raw, orbgpu_features_destroy);
  std::list<CacheEntry> evicted;               // destroyed (if unused) after the lock is released
  {
    std::lock_guard<std::mutex> lock(g_cacheMutex);
    g_cache.push_front(CacheEntry{kind, id, n, digest, set});
    while (g_cache.size() > g_cacheEntries) evicted.splice(evicted.begin(), g_cache, std::prev(g_cache.end()));
  }
  return set;
}

FeatureSet resident(const Frame &F) {
  return resident(0, F.mnId, F.mvKeysUn, F.mvuRight, F.mDescriptors, Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv,
                  Frame::mfGridElementHeightInv, FRAME_GRID_COLS, FRAME_GRID_ROWS, Frame::mnMinX, Frame::mnMinY);
}
// A KeyFrame's grid is a copy of its Frame's (KeyFrame.cc:56-61) — built with the Frame's float bounds — but its
// window walk reads the truncated integer bounds of KeyFrame.h:190-193.
FeatureSet resident(const KeyFrame *K) {
  return resident(1, K->mnId, K->mvKeysUn, K->mvuRight, K->mDescriptors, Frame::mnMinX, Frame::mnMinY, K->mfGridElementWidthInv,
                  K->mfGridElementHeightInv, K->mnGridCols, K->mnGridRows, (float)K->mnMinX, (float)K->mnMinY);
}

// ---- one batch of queries -------------------------------------------------------------------------------------------
struct Batch {
  vector<orbgpu_query> q;
  vector<unsigned char> desc;
  vector<int> tag;                 // caller's index for each query (iMP, i1, …)
  vector<int> cand;                // ORBGPU_SRC_LIST candidates
  vector<orbgpu_match> out;
  void add(const orbgpu_query &query, const cv::Mat &descriptor, int t) {
    q.push_back(query);
    const size_t o = desc.size();
    desc.resize(o + 32);
    std::memcpy(&desc[o], descriptor.ptr(0), 32);
    tag.push_back(t);
  }
  // `set` is held by value: the device copy outlives the call even if another thread evicts it from the cache meanwhile
  void run(FeatureSet set, const orbgpu_search_policy &P, const vector<unsigned char> *occupied) {
    out.resize(q.size());
    if (q.empty()) return;
    if (orbgpu_search(set.get(), q.data(), desc.data(), (int)q.size(), occupied ? occupied->data() : nullptr, cand.data(), (int)cand.size(),
                      &P, out.data(), nullptr) != ORBGPU_OK)
      fail("orbgpu_search");
  }
};

orbgpu_query window_query(float u, float v, float r, int minLevel, int maxLevel) {
  orbgpu_query q{};
  q.u = u; q.v = v; q.r = r; q.min_level = minLevel; q.max_level = maxLevel;
  return q;
}
orbgpu_query list_query(int begin, int end) {
  orbgpu_query q{};
  q.min_level = -1; q.max_level = -1; q.list_begin = begin; q.list_end = end;
  return q;
}
orbgpu_search_policy policy(int source, int ratioMode, int greedy, int acceptMax, float ratio) {
  orbgpu_search_policy P{};
  P.source = source; P.ratio_mode = ratioMode; P.greedy = greedy; P.accept_max = acceptMax; P.ratio = ratio;
  return P;
}

// ---- rotation-consistency histogram (:171-174, :232-239, :255-270) --------------------------------------------------
struct RotationHistogram {
  vector<int> bin[30];
  RotationHistogram() { for (auto &b : bin) b.reserve(500); }
  void vote(float angle1, float angle2, int value) {
    const float factor = 1.0f / 30;             // the reference's 1/HISTO_LENGTH (not HISTO_LENGTH/360), kept as written
    float rot = angle1 - angle2;
    if (rot < 0.0) rot += 360.0f;
    int b = (int)std::round(rot * factor);
    if (b == 30) b = 0;
    bin[b].push_back(value);
  }
};

// Walks two DBoW2 feature vectors in step and calls visit(indices1, indices2) for every shared node (:188-250).
template <class Visit>
void for_each_shared_node(const DBoW2::FeatureVector &a, const DBoW2::FeatureVector &b, Visit visit) {
  auto ia = a.begin(), ib = b.begin();
  while (ia != a.end() && ib != b.end()) {
    if (ia->first == ib->first) { visit(ia->second, ib->second); ++ia; ++ib; }
    else if (ia->first < ib->first) ia = a.lower_bound(ib->first);
    else ib = b.lower_bound(ia->first);
  }
}

// Decomposition of a similarity [sR | t] into the rigid part the projection uses (:284-289, :918-923).
struct RigidFromSim3 {
  cv::Mat Rcw, tcw, Ow;
  explicit RigidFromSim3(const cv::Mat &Scw) {
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw / scw;
    tcw = Scw.rowRange(0, 3).col(3) / scw;
    Ow = -Rcw.t() * tcw;
  }
};

// Projection of a MapPoint into a KeyFrame with the gates shared by SearchByProjection(KF, Scw), both Fuse variants and
// SearchBySim3: positive depth, inside the image, distance inside the scale-invariance range, viewing angle below 60
// degrees (optional), predicted level.  `invDouble` selects 1.0 / z (double quotient rounded to float, :939, :1094)
// over 1 / z (:311, :796).
struct KeyFrameProjection { float u, v, invz, dist3D; int level; };
bool project_into(KeyFrame *pKF, MapPoint *pMP, const cv::Mat &p3Dc, const cv::Mat *PO, bool invDouble, KeyFrameProjection &out) {
  const float z = p3Dc.at<float>(2);
  if (z < 0.0f) return false;
  const float invz = invDouble ? (float)(1.0 / z) : 1 / z;
  const float x = p3Dc.at<float>(0) * invz;
  const float y = p3Dc.at<float>(1) * invz;
  const float u = pKF->fx * x + pKF->cx;
  const float v = pKF->fy * y + pKF->cy;
  if (!pKF->IsInImage(u, v)) return false;
  const float maxDistance = pMP->GetMaxDistanceInvariance();
  const float minDistance = pMP->GetMinDistanceInvariance();
  const float dist3D = PO ? cv::norm(*PO) : cv::norm(p3Dc);
  if (dist3D < minDistance || dist3D > maxDistance) return false;
  if (PO) {
    cv::Mat Pn = pMP->GetNormal();
    if (PO->dot(Pn) < 0.5 * dist3D) return false;
  }
  out.u = u; out.v = v; out.invz = invz; out.dist3D = dist3D;
  out.level = pMP->PredictScale(dist3D, pKF);
  return true;
}

void copy_levels(float *dst, const vector<float> &src) {
  for (size_t i = 0; i < src.size() && i < 16; i++) dst[i] = src[i];
}

}  // namespace

// How many Frame / KeyFrame feature sets stay resident on the device (~70 B per keypoint each); at least 1.
extern "C" void orbgpu_shim_set_feature_cache_capacity(int entries) {
  std::list<CacheEntry> dropped;
  std::lock_guard<std::mutex> lock(g_cacheMutex);
  g_cacheEntries = (size_t)(entries < 1 ? 1 : entries);
  while (g_cache.size() > g_cacheEntries) dropped.splice(dropped.begin(), g_cache, std::prev(g_cache.end()));
}

extern "C" void orbgpu_shim_clear_feature_cache() {
  std::list<CacheEntry> dropped;               // sets still in use by a running search die with their last user
  std::lock_guard<std::mutex> lock(g_cacheMutex);
  dropped.swap(g_cache);
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// :1542-1556.  Called pair by pair from host loops (MapPoint::ComputeDistinctiveDescriptors, Frame::ComputeStereoMatches when
// the host version of the latter is kept), so it stays a host function; bulk distances go through orbgpu_search.
int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {
  const int32_t *pa = a.ptr<int32_t>();
  const int32_t *pb = b.ptr<int32_t>();
  int dist = 0;
  for (int w = 0; w < 8; w++) dist += __builtin_popcount((unsigned)(pa[w] ^ pb[w]));
  return dist;
}

float ORBmatcher::RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5 : 4.0; }   // :130-135

// :137-159 — kept for source compatibility (protected member); the device applies the same test inside
// SearchForTriangulation's candidate loop.
bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, const KeyFrame *pKF2) {
  float l[3];
  for (int j = 0; j < 3; j++) l[j] = kp1.pt.x * F12.at<float>(0, j) + kp1.pt.y * F12.at<float>(1, j) + F12.at<float>(2, j);
  const float num = l[0] * kp2.pt.x + l[1] * kp2.pt.y + l[2];
  const float den = l[0] * l[0] + l[1] * l[1];
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

// :1506-1538
void ORBmatcher::ComputeThreeMaxima(vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3) {
  int top[3] = {0, 0, 0};
  int *ind[3] = {&ind1, &ind2, &ind3};
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    int rank = 3;
    for (int k = 0; k < 3; k++) if (s > top[k]) { rank = k; break; }
    if (rank == 3) continue;
    for (int k = 2; k > rank; k--) { top[k] = top[k - 1]; *ind[k] = *ind[k - 1]; }
    top[rank] = s; *ind[rank] = i;
  }
  if (top[1] < 0.1f * (float)top[0]) { ind2 = -1; ind3 = -1; }
  else if (top[2] < 0.1f * (float)top[0]) ind3 = -1;
}

// ---- Tracking::SearchLocalPoints (:45-128) ----------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint *> &vpMapPoints, const float th) {
  const bool bFactor = th != 1.0;
  Batch B;
  for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
    MapPoint *pMP = vpMapPoints[iMP];
    if (!pMP->mbTrackInView || pMP->isBad()) continue;
    const int level = pMP->mnTrackScaleLevel;
    float r = RadiusByViewingCos(pMP->mTrackViewCos);
    if (bFactor) r *= th;
    const float window = r * F.mvScaleFactors[level];
    orbgpu_query q = window_query(pMP->mTrackProjX, pMP->mTrackProjY, window, level - 1, level);
    q.ur = pMP->mTrackProjXR; q.ur_tol = window;
    q.flags = pMP->Observations() > 0 ? ORBGPU_Q_BLOCKS : 0;
    B.add(q, pMP->GetDescriptor(), (int)iMP);
  }
  vector<unsigned char> occupied(F.N, 0);
  for (int i = 0; i < F.N; i++) occupied[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
  orbgpu_search_policy P = policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_SAME_LEVEL, ORBGPU_GREEDY_TAKEN, TH_HIGH, mfNNratio);
  P.stereo_gate = 1;
  B.run(resident(F), P, &occupied);
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++)
    if (B.out[k].accepted) { F.mvpMapPoints[B.out[k].idx] = vpMapPoints[B.tag[k]]; nmatches++; }
  return nmatches;
}

// ---- Tracking::TrackReferenceKeyFrame / Relocalization (:161-273) ---------------------------------------------------
int ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint *> &vpMapPointMatches) {
  const vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = vector<MapPoint *>(F.N, static_cast<MapPoint *>(NULL));
  Batch B;
  for_each_shared_node(pKF->mFeatVec, F.mFeatVec, [&](const vector<unsigned int> &inKF, const vector<unsigned int> &inF) {
    const int begin = (int)B.cand.size();
    B.cand.insert(B.cand.end(), inF.begin(), inF.end());
    const int end = (int)B.cand.size();
    for (unsigned int idxKF : inKF) {
      MapPoint *pMP = vpMapPointsKF[idxKF];
      if (!pMP || pMP->isBad()) continue;
      orbgpu_query q = list_query(begin, end);
      q.flags = ORBGPU_Q_BLOCKS;
      B.add(q, pKF->mDescriptors.row(idxKF), (int)idxKF);
    }
  });
  B.run(resident(F), policy(ORBGPU_SRC_LIST, ORBGPU_RATIO_LESS, ORBGPU_GREEDY_TAKEN, TH_LOW, mfNNratio), nullptr);
  RotationHistogram hist;
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int idxF = B.out[k].idx, idxKF = B.tag[k];
    vpMapPointMatches[idxF] = vpMapPointsKF[idxKF];
    if (mbCheckOrientation) hist.vote(pKF->mvKeysUn[idxKF].angle, F.mvKeys[idxF].angle, idxF);
    nmatches++;
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idxF : hist.bin[i]) { vpMapPointMatches[idxF] = static_cast<MapPoint *>(NULL); nmatches--; }
    }
  }
  return nmatches;
}

// ---- LoopClosing::ComputeSim3 (:275-386) ----------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, vector<MapPoint *> &vpMatched, int th) {
  const RigidFromSim3 T(Scw);
  std::set<MapPoint *> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPoint *>(NULL));
  Batch B;
  for (int iMP = 0, iend = (int)vpPoints.size(); iMP < iend; iMP++) {
    MapPoint *pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Dc = T.Rcw * p3Dw + T.tcw;
    cv::Mat PO = p3Dw - T.Ow;
    KeyFrameProjection pr;
    if (!project_into(pKF, pMP, p3Dc, &PO, false, pr)) continue;
    orbgpu_query q = window_query(pr.u, pr.v, th * pKF->mvScaleFactors[pr.level], pr.level - 1, pr.level);
    q.flags = ORBGPU_Q_BLOCKS;
    B.add(q, pMP->GetDescriptor(), iMP);
  }
  vector<unsigned char> occupied(vpMatched.size(), 0);
  for (size_t i = 0; i < vpMatched.size(); i++) occupied[i] = vpMatched[i] != NULL;
  occupied.resize(pKF->N, 0);
  B.run(resident(pKF), policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_TAKEN, TH_LOW, 0.f), &occupied);
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++)
    if (B.out[k].accepted) { vpMatched[B.out[k].idx] = vpPoints[B.tag[k]]; nmatches++; }
  return nmatches;
}

// ---- Tracking::MonocularInitialization (:388-492) -----------------------------------------------------------------------
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize) {
  vnMatches12 = vector<int>(F1.mvKeysUn.size(), -1);
  Batch B;
  for (size_t i1 = 0, iend1 = F1.mvKeysUn.size(); i1 < iend1; i1++) {
    if (F1.mvKeysUn[i1].octave > 0) continue;
    B.add(window_query(vbPrevMatched[i1].x, vbPrevMatched[i1].y, (float)windowSize, 0, 0), F1.mDescriptors.row((int)i1), (int)i1);
  }
  B.run(resident(F2), policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_LESS, ORBGPU_GREEDY_DISTANCE, TH_LOW, mfNNratio), nullptr);
  RotationHistogram hist;
  vector<int> vnMatches21(F2.mvKeysUn.size(), -1);
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int i1 = B.tag[k], i2 = B.out[k].idx;
    if (vnMatches21[i2] >= 0) { vnMatches12[vnMatches21[i2]] = -1; nmatches--; }     // a closer match displaces the earlier one
    vnMatches12[i1] = i2;
    vnMatches21[i2] = i1;
    nmatches++;
    if (mbCheckOrientation) hist.vote(F1.mvKeysUn[i1].angle, F2.mvKeysUn[i2].angle, i1);
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int i1 : hist.bin[i]) if (vnMatches12[i1] >= 0) { vnMatches12[i1] = -1; nmatches--; }
    }
  }
  for (size_t i1 = 0, iend1 = vnMatches12.size(); i1 < iend1; i1++)
    if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
  return nmatches;
}

// ---- LoopClosing::ComputeSim3 (:494-612) ----------------------------------------------------------------------------
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12) {
  const vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches();
  const vector<MapPoint *> vpMapPoints2 = pKF2->GetMapPointMatches();
  vpMatches12 = vector<MapPoint *>(vpMapPoints1.size(), static_cast<MapPoint *>(NULL));
  vector<unsigned char> unavailable(pKF2->N, 0);
  for (size_t i = 0; i < vpMapPoints2.size(); i++) unavailable[i] = !vpMapPoints2[i] || vpMapPoints2[i]->isBad();
  Batch B;
  for_each_shared_node(pKF1->mFeatVec, pKF2->mFeatVec, [&](const vector<unsigned int> &in1, const vector<unsigned int> &in2) {
    const int begin = (int)B.cand.size();
    B.cand.insert(B.cand.end(), in2.begin(), in2.end());
    const int end = (int)B.cand.size();
    for (unsigned int idx1 : in1) {
      MapPoint *pMP1 = vpMapPoints1[idx1];
      if (!pMP1 || pMP1->isBad()) continue;
      orbgpu_query q = list_query(begin, end);
      q.flags = ORBGPU_Q_BLOCKS;
      B.add(q, pKF1->mDescriptors.row(idx1), (int)idx1);
    }
  });
  B.run(resident(pKF2), policy(ORBGPU_SRC_LIST, ORBGPU_RATIO_LESS, ORBGPU_GREEDY_TAKEN, TH_LOW - 1, mfNNratio), &unavailable);   // '<' TH_LOW (:564)
  RotationHistogram hist;
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int idx1 = B.tag[k], idx2 = B.out[k].idx;
    vpMatches12[idx1] = vpMapPoints2[idx2];
    if (mbCheckOrientation) hist.vote(pKF1->mvKeysUn[idx1].angle, pKF2->mvKeysUn[idx2].angle, idx1);
    nmatches++;
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : hist.bin[i]) { vpMatches12[idx1] = static_cast<MapPoint *>(NULL); nmatches--; }
    }
  }
  return nmatches;
}

// ---- LocalMapping::CreateNewMapPoints (:614-764) ----------------------------------------------------------------------
int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<std::pair<size_t, size_t>> &vMatchedPairs,
                                       const bool bOnlyStereo) {
  // epipole of camera 1 in image 2 (:620-627)
  cv::Mat Cw = pKF1->GetCameraCenter();
  cv::Mat R2w = pKF2->GetRotation();
  cv::Mat t2w = pKF2->GetTranslation();
  cv::Mat C2 = R2w * Cw + t2w;
  const float invz = 1.0f / C2.at<float>(2);
  orbgpu_search_policy P = policy(ORBGPU_SRC_LIST, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_NONE, TH_LOW, 0.f);
  P.gate = ORBGPU_GATE_EPIPOLAR;
  P.only_stereo = bOnlyStereo ? 1 : 0;
  P.ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
  P.ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) P.f12[3 * r + c] = F12.at<float>(r, c);
  copy_levels(P.level_sigma2, pKF2->mvLevelSigma2);
  copy_levels(P.scale_factors, pKF2->mvScaleFactors);

  vector<unsigned char> unavailable(pKF2->N, 0);
  for (int i = 0; i < pKF2->N; i++) unavailable[i] = pKF2->GetMapPoint(i) != NULL;
  Batch B;
  for_each_shared_node(pKF1->mFeatVec, pKF2->mFeatVec, [&](const vector<unsigned int> &in1, const vector<unsigned int> &in2) {
    const int begin = (int)B.cand.size();
    B.cand.insert(B.cand.end(), in2.begin(), in2.end());
    const int end = (int)B.cand.size();
    for (unsigned int idx1 : in1) {
      if (pKF1->GetMapPoint(idx1)) continue;
      const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
      if (bOnlyStereo && !bStereo1) continue;
      orbgpu_query q = list_query(begin, end);
      q.u = pKF1->mvKeysUn[idx1].pt.x; q.v = pKF1->mvKeysUn[idx1].pt.y;
      q.flags = bStereo1 ? ORBGPU_Q_STEREO : 0;
      B.add(q, pKF1->mDescriptors.row(idx1), (int)idx1);
    }
  });
  B.run(resident(pKF2), P, &unavailable);
  vector<int> vMatches12(pKF1->N, -1);
  RotationHistogram hist;
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int idx1 = B.tag[k], idx2 = B.out[k].idx;
    vMatches12[idx1] = idx2;
    nmatches++;
    if (mbCheckOrientation) hist.vote(pKF1->mvKeysUn[idx1].angle, pKF2->mvKeysUn[idx2].angle, idx1);
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : hist.bin[i]) { vMatches12[idx1] = -1; nmatches--; }
    }
  }
  vMatchedPairs.clear();
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
    if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
  return nmatches;
}

// ---- LocalMapping::SearchInNeighbors (:766-907) ---------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th) {
  cv::Mat Rcw = pKF->GetRotation();
  cv::Mat tcw = pKF->GetTranslation();
  cv::Mat Ow = pKF->GetCameraCenter();
  const float bf = pKF->mbf;
  Batch B;
  for (int i = 0, n = (int)vpMapPoints.size(); i < n; i++) {
    MapPoint *pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Dc = Rcw * p3Dw + tcw;
    cv::Mat PO = p3Dw - Ow;
    KeyFrameProjection pr;
    if (!project_into(pKF, pMP, p3Dc, &PO, false, pr)) continue;
    orbgpu_query q = window_query(pr.u, pr.v, th * pKF->mvScaleFactors[pr.level], pr.level - 1, pr.level);
    q.ur = pr.u - bf * pr.invz;
    B.add(q, pMP->GetDescriptor(), i);
  }
  orbgpu_search_policy P = policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_NONE, TH_LOW, 0.f);
  P.gate = ORBGPU_GATE_FUSE;
  copy_levels(P.inv_level_sigma2, pKF->mvInvLevelSigma2);
  B.run(resident(pKF), P, nullptr);
  // The candidate loop of one MapPoint reads nothing an earlier iteration's Replace / AddObservation can change, so
  // the matches are those of the sequential run; the bad / already-in-keyframe tests and the map updates are replayed
  // here in order, on the state each iteration would have seen (:780-781, :886-900).
  int nFused = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    MapPoint *pMP = vpMapPoints[B.tag[k]];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF) || !B.out[k].accepted) continue;
    const int bestIdx = B.out[k].idx;
    MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// ---- LoopClosing::SearchAndFuse (:909-1027) ---------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint) {
  const RigidFromSim3 T(Scw);
  const std::set<MapPoint *> spAlreadyFound = pKF->GetMapPoints();
  Batch B;
  for (int iMP = 0, n = (int)vpPoints.size(); iMP < n; iMP++) {
    MapPoint *pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Dc = T.Rcw * p3Dw + T.tcw;
    cv::Mat PO = p3Dw - T.Ow;
    KeyFrameProjection pr;
    if (!project_into(pKF, pMP, p3Dc, &PO, true, pr)) continue;
    B.add(window_query(pr.u, pr.v, th * pKF->mvScaleFactors[pr.level], pr.level - 1, pr.level), pMP->GetDescriptor(), iMP);
  }
  B.run(resident(pKF), policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_NONE, TH_LOW, 0.f), nullptr);
  int nFused = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    MapPoint *pMP = vpPoints[B.tag[k]];
    if (pMP->isBad() || !B.out[k].accepted) continue;
    const int bestIdx = B.out[k].idx;
    MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[B.tag[k]] = pMPinKF;
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// ---- LoopClosing::ComputeSim3 (:1029-1245) --------------------------------------------------------------------------------
int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12, const float &s12, const cv::Mat &R12,
                             const cv::Mat &t12, const float th) {
  cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
  cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  cv::Mat sR12 = s12 * R12;
  cv::Mat sR21 = (1.0 / s12) * R12.t();
  cv::Mat t21 = -sR21 * t12;

  const vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches();
  const vector<MapPoint *> vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
  vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
  for (int i = 0; i < N1; i++) {
    MapPoint *pMP = vpMatches12[i];
    if (!pMP) continue;
    vbAlreadyMatched1[i] = true;
    const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
    if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
  }

  // one direction: the MapPoints seen in `from` are moved into `into`'s camera and searched there
  auto search = [&](const vector<MapPoint *> &points, const vector<bool> &done, const cv::Mat &Rw, const cv::Mat &tw, const cv::Mat &sR,
                    const cv::Mat &t, KeyFrame *into, vector<int> &match) {
    Batch B;
    for (int i = 0, n = (int)points.size(); i < n; i++) {
      MapPoint *pMP = points[i];
      if (!pMP || done[i] || pMP->isBad()) continue;
      cv::Mat p3Dw = pMP->GetWorldPos();
      cv::Mat inOwn = Rw * p3Dw + tw;
      cv::Mat inOther = sR * inOwn + t;
      KeyFrameProjection pr;
      if (!project_into(into, pMP, inOther, nullptr, true, pr)) continue;
      B.add(window_query(pr.u, pr.v, th * into->mvScaleFactors[pr.level], pr.level - 1, pr.level), pMP->GetDescriptor(), i);
    }
    B.run(resident(into), policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_NONE, TH_HIGH, 0.f), nullptr);
    for (size_t k = 0; k < B.q.size(); k++) if (B.out[k].accepted) match[B.tag[k]] = B.out[k].idx;
  };
  vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
  search(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1);
  search(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2);

  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
  }
  return nFound;
}

// ---- Tracking::TrackWithMotionModel (:1247-1383) --------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono) {
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat twc = -Rcw.t() * tcw;
  const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat tlc = Rlw * twc + tlw;
  const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
  const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

  Batch B;
  for (int i = 0; i < LastFrame.N; i++) {
    MapPoint *pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    cv::Mat x3Dw = pMP->GetWorldPos();
    cv::Mat x3Dc = Rcw * x3Dw + tcw;
    const float xc = x3Dc.at<float>(0);
    const float yc = x3Dc.at<float>(1);
    const float invzc = 1.0 / x3Dc.at<float>(2);
    if (invzc < 0) continue;
    const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
    const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
    if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
    if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
    const int octave = LastFrame.mvKeys[i].octave;
    const float radius = th * CurrentFrame.mvScaleFactors[octave];
    orbgpu_query q = bForward    ? window_query(u, v, radius, octave, -1)
                     : bBackward ? window_query(u, v, radius, 0, octave)
                                 : window_query(u, v, radius, octave - 1, octave + 1);
    q.ur = u - CurrentFrame.mbf * invzc; q.ur_tol = radius;
    q.flags = pMP->Observations() > 0 ? ORBGPU_Q_BLOCKS : 0;
    B.add(q, pMP->GetDescriptor(), i);
  }
  vector<unsigned char> occupied(CurrentFrame.N, 0);
  for (int i = 0; i < CurrentFrame.N; i++) occupied[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;
  orbgpu_search_policy P = policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_TAKEN, TH_HIGH, 0.f);
  P.stereo_gate = 1;
  B.run(resident(CurrentFrame), P, &occupied);
  RotationHistogram hist;
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int i = B.tag[k], i2 = B.out[k].idx;
    CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[i];
    nmatches++;
    if (mbCheckOrientation) hist.vote(LastFrame.mvKeysUn[i].angle, CurrentFrame.mvKeysUn[i2].angle, i2);
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int i2 : hist.bin[i]) { CurrentFrame.mvpMapPoints[i2] = static_cast<MapPoint *>(NULL); nmatches--; }
    }
  }
  return nmatches;
}

// ---- Tracking::Relocalization (:1385-1504) ----------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th,
                                   const int ORBdist) {
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat Ow = -Rcw.t() * tcw;
  const vector<MapPoint *> vpMPs = pKF->GetMapPointMatches();
  Batch B;
  for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
    MapPoint *pMP = vpMPs[i];
    if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
    cv::Mat x3Dw = pMP->GetWorldPos();
    cv::Mat x3Dc = Rcw * x3Dw + tcw;
    const float xc = x3Dc.at<float>(0);
    const float yc = x3Dc.at<float>(1);
    const float invzc = 1.0 / x3Dc.at<float>(2);
    const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
    const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
    if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
    if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
    cv::Mat PO = x3Dw - Ow;
    const float dist3D = cv::norm(PO);
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const int level = pMP->PredictScale(dist3D, &CurrentFrame);
    orbgpu_query q = window_query(u, v, th * CurrentFrame.mvScaleFactors[level], level - 1, level + 1);
    q.flags = ORBGPU_Q_BLOCKS;
    B.add(q, pMP->GetDescriptor(), (int)i);
  }
  vector<unsigned char> occupied(CurrentFrame.N, 0);
  for (int i = 0; i < CurrentFrame.N; i++) occupied[i] = CurrentFrame.mvpMapPoints[i] != NULL;
  B.run(resident(CurrentFrame), policy(ORBGPU_SRC_WINDOW, ORBGPU_RATIO_NONE, ORBGPU_GREEDY_TAKEN, ORBdist, 0.f), &occupied);
  RotationHistogram hist;
  int nmatches = 0;
  for (size_t k = 0; k < B.q.size(); k++) {
    if (!B.out[k].accepted) continue;
    const int i = B.tag[k], i2 = B.out[k].idx;
    CurrentFrame.mvpMapPoints[i2] = vpMPs[i];
    nmatches++;
    if (mbCheckOrientation) hist.vote(pKF->mvKeysUn[i].angle, CurrentFrame.mvKeysUn[i2].angle, i2);
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(hist.bin, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int i2 : hist.bin[i]) { CurrentFrame.mvpMapPoints[i2] = NULL; nmatches--; }
    }
  }
  return nmatches;
}

}  // namespace ORB_SLAM2
