// slam_api.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN data model and matcher.
//
// Built three times by the Makefile beside it, always together with the reference's unmodified Frame.cc, KeyFrame.cc,
// MapPoint.cc, Map.cc and ORBextractor.cc (compiled from where they lie under /root/reference against the OpenCV
// API stand-in cvshim.hpp and the three stand-ins of slam_standins.hpp):
//
//   oracle/_ref/libref_slam.so        + the reference's own src/ORBmatcher.cc                  -> the checker
//   oracle/_ref/libshim_slam_cpu.so   + shim/ORBmatcher.cc on the CPU oracle backend           -> pins host policy + oracle
//   oracle/_ref/libshim_slam_gpu.so   + shim/ORBmatcher.cc AND shim/ORBextractor.cc (instead of the reference's) on
//                                       liborbgpu.so: the reference linked against the drop-in -> GPU parity tests
//
// A "scene" is a small map built from flat arrays (frames with keypoints / descriptors / right coordinates, keyframes
// made from them, MapPoints with observations).  Every run_* function calls one public ORBmatcher method exactly as
// Tracking / LocalMapping / LoopClosing do (call sites in SURVEY.md section 3.3) and the dump functions flatten the
// state it left behind, so two builds can be compared output by output from Python (tests/test_matcher_ref.py).
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"
#ifdef SLAM_SHIM_EXTRACTOR
#include "FrameStereo.h"
#endif
#undef private
#undef protected

using namespace ORB_SLAM2;

struct RefKp { float x, y, size, angle, response; int octave, class_id; };
static_assert(sizeof(cv::KeyPoint) == sizeof(RefKp), "KeyPoint layout");

extern "C" void orbgpu_shim_clear_feature_cache() __attribute__((weak));

namespace {

struct Scene {
  Map map;
  KeyFrameDatabase db;
  ORBextractor *ex = nullptr;            // only for its scale tables (ORBextractor.cc:407-425)
  std::vector<Frame *> frames;
  std::vector<KeyFrame *> kfs;
  std::vector<MapPoint *> mps;
  std::map<MapPoint *, int> mpId;
  cv::Mat K;
  ~Scene() {
    for (auto *p : mps) delete p;
    for (auto *p : kfs) delete p;
    for (auto *p : frames) delete p;
    delete ex;
  }
  int idOf(MapPoint *p) const { if (!p) return -1; auto it = mpId.find(p); return it == mpId.end() ? -2 : it->second; }
};

cv::Mat mat_from(const float *v, int rows, int cols) {
  cv::Mat m(rows, cols, CV_32F);
  for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(r, c) = v[r * cols + c];
  return m;
}

}  // namespace

extern "C" {

void *slam_scene_create(float fx, float fy, float cx, float cy, float minX, float maxX, float minY, float maxY, int nlevels, float scaleFactor) {
  Scene *S = new Scene;
  S->ex = new ORBextractor(1000, scaleFactor, nlevels, 20, 7);
  Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy; Frame::invfx = 1.0f / fx; Frame::invfy = 1.0f / fy;
  Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (maxX - minX);       // Frame.cc:108-111
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (maxY - minY);
  Frame::mbInitialComputations = false;
  const float k[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  S->K = mat_from(k, 3, 3);
  return S;
}
void slam_scene_destroy(void *h) {
  delete (Scene *)h;
  if (orbgpu_shim_clear_feature_cache) orbgpu_shim_clear_feature_cache();
}

int slam_add_frame(void *h, int n, const RefKp *keys, const float *uRight, const float *depth, const uint8_t *desc, float mbf, float thDepth,
                   const float *Tcw) {
  Scene *S = (Scene *)h;
  Frame *F = new Frame();
  F->mpORBvocabulary = nullptr; F->mpORBextractorLeft = S->ex; F->mpORBextractorRight = nullptr;
  F->mTimeStamp = (double)S->frames.size();
  F->mK = S->K.clone();
  F->mDistCoef = cv::Mat::zeros(4, 1, CV_32F);
  F->mbf = mbf; F->mb = mbf / Frame::fx; F->mThDepth = thDepth;
  F->N = n;
  F->mvKeys.resize(n);
  for (int i = 0; i < n; i++) std::memcpy(&F->mvKeys[i], &keys[i], sizeof(RefKp));
  F->mvKeysUn = F->mvKeys;                                     // Frame.cc:420-423 (zero distortion)
  F->mvuRight.assign(uRight, uRight + n);
  F->mvDepth.assign(depth, depth + n);
  F->mDescriptors = cv::Mat(n, 32, CV_8U);
  for (int i = 0; i < n; i++) std::memcpy(F->mDescriptors.ptr(i), desc + (size_t)i * 32, 32);
  F->mvpMapPoints = std::vector<MapPoint *>(n, static_cast<MapPoint *>(NULL));
  F->mvbOutlier = std::vector<bool>(n, false);
  F->mnId = Frame::nNextId++;
  F->mpReferenceKF = nullptr;
  F->mnScaleLevels = S->ex->GetLevels();                       // Frame.cc:79-85
  F->mfScaleFactor = S->ex->GetScaleFactor();
  F->mfLogScaleFactor = log(F->mfScaleFactor);
  F->mvScaleFactors = S->ex->GetScaleFactors();
  F->mvInvScaleFactors = S->ex->GetInverseScaleFactors();
  F->mvLevelSigma2 = S->ex->GetScaleSigmaSquares();
  F->mvInvLevelSigma2 = S->ex->GetInverseScaleSigmaSquares();
  F->AssignFeaturesToGrid();                                    // the reference's own grid (Frame.cc:250-263)
  if (Tcw) F->SetPose(mat_from(Tcw, 4, 4));
  S->frames.push_back(F);
  return (int)S->frames.size() - 1;
}
void slam_frame_set_pose(void *h, int f, const float *Tcw) { ((Scene *)h)->frames[f]->SetPose(mat_from(Tcw, 4, 4)); }
int slam_frame_id(void *h, int f) { return (int)((Scene *)h)->frames[f]->mnId; }

int slam_add_keyframe(void *h, int f) {
  Scene *S = (Scene *)h;
  KeyFrame *K = new KeyFrame(*S->frames[f], &S->map, &S->db);
  S->map.AddKeyFrame(K);
  S->kfs.push_back(K);
  return (int)S->kfs.size() - 1;
}
int slam_add_mappoint(void *h, const float *pos, int refKF) {
  Scene *S = (Scene *)h;
  MapPoint *p = new MapPoint(mat_from(pos, 3, 1), S->kfs[refKF], &S->map);
  S->map.AddMapPoint(p);
  S->mpId[p] = (int)S->mps.size();
  S->mps.push_back(p);
  return (int)S->mps.size() - 1;
}
void slam_observe(void *h, int mp, int kf, int idx) {
  Scene *S = (Scene *)h;
  S->mps[mp]->AddObservation(S->kfs[kf], idx);
  S->kfs[kf]->AddMapPoint(S->mps[mp], idx);
}
void slam_mappoint_update(void *h, int mp) {
  Scene *S = (Scene *)h;
  S->mps[mp]->ComputeDistinctiveDescriptors();
  S->mps[mp]->UpdateNormalAndDepth();
}
// ComputeDistinctiveDescriptors walks a std::map keyed by KeyFrame POINTERS (MapPoint.cc:229-321), so with two
// observations (equal medians) the chosen descriptor depends on heap addresses; scenes pin it explicitly.
void slam_mappoint_set_descriptor(void *h, int mp, const uint8_t *d) {
  cv::Mat m(1, 32, CV_8U);
  std::memcpy(m.ptr(0), d, 32);
  ((Scene *)h)->mps[mp]->mDescriptor = m;
}
// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:229-321), the reference's own: returns the observation descriptors in
// the order the reference visits them (its std::map<KeyFrame*, size_t>, bad keyframes skipped) and the descriptor it chose.
int slam_mappoint_distinctive(void *h, int mp, uint8_t *rows, int cap, uint8_t *chosen) {
  Scene *S = (Scene *)h;
  MapPoint *p = S->mps[mp];
  int n = 0;
  for (auto &ob : p->mObservations) {
    if (ob.first->isBad()) continue;
    if (n < cap) std::memcpy(rows + (size_t)n * 32, ob.first->mDescriptors.ptr((int)ob.second), 32);
    n++;
  }
  p->ComputeDistinctiveDescriptors();
  cv::Mat d = p->GetDescriptor();
  std::memcpy(chosen, d.ptr(0), 32);
  return n;
}
void slam_mappoint_set_bad(void *h, int mp) { ((Scene *)h)->mps[mp]->SetBadFlag(); }
void slam_mappoint_descriptor(void *h, int mp, uint8_t *out) { cv::Mat d = ((Scene *)h)->mps[mp]->GetDescriptor(); std::memcpy(out, d.ptr(0), 32); }
void slam_frame_assoc(void *h, int f, int idx, int mp) { Scene *S = (Scene *)h; S->frames[f]->mvpMapPoints[idx] = mp < 0 ? nullptr : S->mps[mp]; }
void slam_frame_outlier(void *h, int f, int idx, int flag) { ((Scene *)h)->frames[f]->mvbOutlier[idx] = flag != 0; }

void slam_set_featvec(void *h, int isKF, int id, int nnodes, const int *nodeIds, const int *offsets, const int *indices) {
  Scene *S = (Scene *)h;
  DBoW2::FeatureVector &fv = isKF ? S->kfs[id]->mFeatVec : S->frames[id]->mFeatVec;
  fv.clear();
  for (int k = 0; k < nnodes; k++)
    for (int j = offsets[k]; j < offsets[k + 1]; j++) fv.addFeature((DBoW2::NodeId)nodeIds[k], (unsigned)indices[j]);
}

// ---- state dumps ------------------------------------------------------------------------------------------------------
void slam_frame_mappoints(void *h, int f, int *out) { Scene *S = (Scene *)h; Frame *F = S->frames[f]; for (int i = 0; i < F->N; i++) out[i] = S->idOf(F->mvpMapPoints[i]); }
void slam_kf_mappoints(void *h, int k, int *out) {
  Scene *S = (Scene *)h;
  const std::vector<MapPoint *> v = S->kfs[k]->GetMapPointMatches();
  for (size_t i = 0; i < v.size(); i++) out[i] = S->idOf(v[i]);
}
void slam_mappoint_state(void *h, int mp, int *out) {   // bad, observations, replaced-by id, index in each keyframe follows
  Scene *S = (Scene *)h;
  MapPoint *p = S->mps[mp];
  out[0] = p->isBad() ? 1 : 0; out[1] = p->Observations(); out[2] = S->idOf(p->GetReplaced());
}
int slam_mappoint_index_in_kf(void *h, int mp, int k) { Scene *S = (Scene *)h; return S->mps[mp]->GetIndexInKeyFrame(S->kfs[k]); }
void slam_frame_grid(void *h, int f, int *cellStart, int *items) {
  Frame *F = ((Scene *)h)->frames[f];
  int k = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
      cellStart[ix * FRAME_GRID_ROWS + iy] = k;
      for (size_t v : F->mGrid[ix][iy]) items[k++] = (int)v;
    }
  cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
}
int slam_features_in_area(void *h, int isKF, int id, float x, float y, float r, int minLevel, int maxLevel, int *out, int cap) {
  Scene *S = (Scene *)h;
  const std::vector<size_t> v = isKF ? S->kfs[id]->GetFeaturesInArea(x, y, r) : S->frames[id]->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}

// ---- the eleven searches, called as their callers call them ----------------------------------------------------------------
// Tracking::SearchLocalPoints (Tracking.cc:1029-1078): isInFrustum fills the MapPoints' tracking fields, then the search.
int slam_search_local_points(void *h, int f, const int *mps, int n, float th, float nnratio, int checkOri, int *inView) {
  Scene *S = (Scene *)h;
  Frame &F = *S->frames[f];
  std::vector<MapPoint *> v;
  for (int i = 0; i < n; i++) {
    MapPoint *p = S->mps[mps[i]];
    const bool in = !p->isBad() && F.isInFrustum(p, 0.5);
    if (inView) inView[i] = in;
    v.push_back(p);
  }
  ORBmatcher matcher(nnratio, checkOri != 0);
  return matcher.SearchByProjection(F, v, th);
}
int slam_search_last_frame(void *h, int cur, int last, float th, int mono, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  ORBmatcher matcher(nnratio, checkOri != 0);
  return matcher.SearchByProjection(*S->frames[cur], *S->frames[last], th, mono != 0);
}
int slam_search_reloc(void *h, int cur, int kf, const int *found, int nfound, float th, int orbDist, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  std::set<MapPoint *> s;
  for (int i = 0; i < nfound; i++) s.insert(S->mps[found[i]]);
  ORBmatcher matcher(nnratio, checkOri != 0);
  return matcher.SearchByProjection(*S->frames[cur], S->kfs[kf], s, th, orbDist);
}
int slam_search_sim3_projection(void *h, int kf, const float *Scw, const int *mps, int n, int *matched, int th, float nnratio) {
  Scene *S = (Scene *)h;
  KeyFrame *K = S->kfs[kf];
  std::vector<MapPoint *> pts, m(K->N, static_cast<MapPoint *>(NULL));
  for (int i = 0; i < n; i++) pts.push_back(S->mps[mps[i]]);
  for (int i = 0; i < K->N; i++) if (matched[i] >= 0) m[i] = S->mps[matched[i]];
  ORBmatcher matcher(nnratio, true);
  const int r = matcher.SearchByProjection(K, mat_from(Scw, 4, 4), pts, m, th);
  for (int i = 0; i < K->N; i++) matched[i] = S->idOf(m[i]);
  return r;
}
// Concurrency stress for the drop-in's feature-set cache: `nthreads` host threads (std::thread, no interpreter lock in the way)
// run SearchByProjection(KF, Scw, ...) over ALL keyframes `rounds` times, thread t starting at keyframe t * nkf / nthreads.
// counts[t * nkf + k] = matches thread t found in keyframe k in its last round (all threads must agree).
void slam_stress_sim3_projection(void *h, const float *Scws, const int *mps, int n, int th, float nnratio, int nthreads, int rounds, int *counts) {
  Scene *S = (Scene *)h;
  const int nkf = (int)S->kfs.size();
  std::vector<MapPoint *> pts;
  for (int i = 0; i < n; i++) pts.push_back(S->mps[mps[i]]);
  auto worker = [&](int t) {
    for (int r = 0; r < rounds; r++)
      for (int j = 0; j < nkf; j++) {
        const int k = (j + t * nkf / nthreads) % nkf;
        KeyFrame *K = S->kfs[k];
        std::vector<MapPoint *> m(K->N, static_cast<MapPoint *>(NULL));
        ORBmatcher matcher(nnratio, true);
        counts[t * nkf + k] = matcher.SearchByProjection(K, mat_from(Scws + 16 * k, 4, 4), pts, m, th);
      }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back(worker, t);
  for (auto &t : pool) t.join();
}
int slam_search_bow_frame(void *h, int kf, int f, int *out, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  std::vector<MapPoint *> m;
  ORBmatcher matcher(nnratio, checkOri != 0);
  const int r = matcher.SearchByBoW(S->kfs[kf], *S->frames[f], m);
  for (size_t i = 0; i < m.size(); i++) out[i] = S->idOf(m[i]);
  return r;
}
int slam_search_bow_kf(void *h, int kf1, int kf2, int *out, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  std::vector<MapPoint *> m;
  ORBmatcher matcher(nnratio, checkOri != 0);
  const int r = matcher.SearchByBoW(S->kfs[kf1], S->kfs[kf2], m);
  for (size_t i = 0; i < m.size(); i++) out[i] = S->idOf(m[i]);
  return r;
}
int slam_search_initialization(void *h, int f1, int f2, float *prevMatched, int *matches12, int windowSize, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  Frame &F1 = *S->frames[f1];
  std::vector<cv::Point2f> prev(F1.mvKeysUn.size());
  for (size_t i = 0; i < prev.size(); i++) prev[i] = cv::Point2f(prevMatched[2 * i], prevMatched[2 * i + 1]);
  std::vector<int> m;
  ORBmatcher matcher(nnratio, checkOri != 0);
  const int r = matcher.SearchForInitialization(F1, *S->frames[f2], prev, m, windowSize);
  for (size_t i = 0; i < prev.size(); i++) { prevMatched[2 * i] = prev[i].x; prevMatched[2 * i + 1] = prev[i].y; matches12[i] = m[i]; }
  return r;
}
int slam_search_triangulation(void *h, int kf1, int kf2, const float *F12, int onlyStereo, int *pairs, int cap, float nnratio, int checkOri) {
  Scene *S = (Scene *)h;
  std::vector<std::pair<size_t, size_t>> v;
  ORBmatcher matcher(nnratio, checkOri != 0);
  matcher.SearchForTriangulation(S->kfs[kf1], S->kfs[kf2], mat_from(F12, 3, 3), v, onlyStereo != 0);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) { pairs[2 * i] = (int)v[i].first; pairs[2 * i + 1] = (int)v[i].second; }
  return (int)v.size();
}
int slam_fuse(void *h, int kf, const int *mps, int n, float th, float nnratio) {
  Scene *S = (Scene *)h;
  std::vector<MapPoint *> v;
  for (int i = 0; i < n; i++) v.push_back(mps[i] < 0 ? nullptr : S->mps[mps[i]]);
  ORBmatcher matcher(nnratio, true);
  return matcher.Fuse(S->kfs[kf], v, th);
}
int slam_fuse_sim3(void *h, int kf, const float *Scw, const int *mps, int n, float th, int *replace, float nnratio) {
  Scene *S = (Scene *)h;
  std::vector<MapPoint *> v, rep(n, static_cast<MapPoint *>(NULL));
  for (int i = 0; i < n; i++) v.push_back(S->mps[mps[i]]);
  ORBmatcher matcher(nnratio, true);
  const int r = matcher.Fuse(S->kfs[kf], mat_from(Scw, 4, 4), v, th, rep);
  for (int i = 0; i < n; i++) replace[i] = S->idOf(rep[i]);
  return r;
}
int slam_search_by_sim3(void *h, int kf1, int kf2, int *matches12, float s12, const float *R12, const float *t12, float th, float nnratio) {
  Scene *S = (Scene *)h;
  KeyFrame *K1 = S->kfs[kf1];
  std::vector<MapPoint *> m(K1->N, static_cast<MapPoint *>(NULL));
  for (int i = 0; i < K1->N; i++) if (matches12[i] >= 0) m[i] = S->mps[matches12[i]];
  ORBmatcher matcher(nnratio, true);
  const int r = matcher.SearchBySim3(K1, S->kfs[kf2], m, s12, mat_from(R12, 3, 3), mat_from(t12, 3, 1), th);
  for (int i = 0; i < K1->N; i++) matches12[i] = S->idOf(m[i]);
  return r;
}
int slam_descriptor_distance(const uint8_t *a, const uint8_t *b) {
  cv::Mat A(1, 32, CV_8U, (void *)a, 32), B(1, 32, CV_8U, (void *)b, 32);
  return ORBmatcher::DescriptorDistance(A, B);
}

#ifndef SLAM_SHIM_EXTRACTOR
// ---- Frame::ComputeStereoMatches (Frame.cc:477-646), the reference's own code ----------------------------------------------
// Keypoints and descriptors are given (the octree's tie-break makes the reference's own extraction allocator-dependent,
// see ref_api.cpp); the pyramids are built by the reference's ORBextractor::ComputePyramid from the two images.
int slam_compute_stereo(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t *left, const uint8_t *right, int w,
                        int hgt, const RefKp *kL, const uint8_t *dL, int nL, const RefKp *kR, const uint8_t *dR, int nR, float mbf, float mb,
                        float *uRight, float *depth) {
  ORBextractor exL(nfeatures, scaleFactor, nlevels, iniTh, minTh), exR(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  exL.ComputePyramid(cv::Mat(hgt, w, CV_8UC1, (void *)left, (size_t)w));
  exR.ComputePyramid(cv::Mat(hgt, w, CV_8UC1, (void *)right, (size_t)w));
  Frame F;
  F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
  F.N = nL;
  F.mvKeys.resize(nL); F.mvKeysRight.resize(nR);
  for (int i = 0; i < nL; i++) std::memcpy(&F.mvKeys[i], &kL[i], sizeof(RefKp));
  for (int i = 0; i < nR; i++) std::memcpy(&F.mvKeysRight[i], &kR[i], sizeof(RefKp));
  F.mDescriptors = cv::Mat(nL, 32, CV_8U); F.mDescriptorsRight = cv::Mat(nR, 32, CV_8U);
  for (int i = 0; i < nL; i++) std::memcpy(F.mDescriptors.ptr(i), dL + (size_t)i * 32, 32);
  for (int i = 0; i < nR; i++) std::memcpy(F.mDescriptorsRight.ptr(i), dR + (size_t)i * 32, 32);
  F.mvScaleFactors = exL.GetScaleFactors(); F.mvInvScaleFactors = exL.GetInverseScaleFactors();
  F.mbf = mbf; F.mb = mb;                  // Frame.cc:505 reads mb before Frame.cc:124 assigns it; the caller supplies mbf / fx
  F.ComputeStereoMatches();
  int matched = 0;
  for (int i = 0; i < nL; i++) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; matched += F.mvuRight[i] >= 0; }
  return matched;
}

#endif

// ---- the reference's own stereo Frame constructor (Frame.cc:66-127) -------------------------------------------------------
// Two extractor threads, UndistortKeyPoints, ComputeStereoMatches, image bounds, AssignFeaturesToGrid — all the
// reference's code; which ORBextractor / ORBmatcher it runs on depends on the build (reference's own, or the drop-in
// shims on liborbgpu.so).  The constructor reads `mb` before assigning it (Frame.cc:505 vs :124), so the object is
// constructed in storage where that member already holds the intended mbf / fx.
int slam_stereo_frame(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t *left, const uint8_t *right, int w,
                      int hgt, float fx, float fy, float cx, float cy, float bf, float thDepth, int cap, RefKp *keys, uint8_t *desc,
                      float *uRight, float *depth, int *cellStart, int *items, float *uRightDevice, float *depthDevice) {
  ORBextractor exL(nfeatures, scaleFactor, nlevels, iniTh, minTh), exR(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  const float k[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  cv::Mat K = mat_from(k, 3, 3), dist = cv::Mat::zeros(4, 1, CV_32F);
  cv::Mat imL(hgt, w, CV_8UC1, (void *)left, (size_t)w), imR(hgt, w, CV_8UC1, (void *)right, (size_t)w);
  Frame::mbInitialComputations = true;
  std::vector<char> storage(sizeof(Frame) + 64, 0);
  char *raw = storage.data() + (64 - (uintptr_t)storage.data() % 64) % 64;
  *(float *)(raw + offsetof(Frame, mb)) = bf / fx;
  Frame *F = new (raw) Frame(imL, imR, 0.0, &exL, &exR, nullptr, K, dist, bf, thDepth);
  const int n = F->N;
  if (n <= cap) {
    for (int i = 0; i < n; i++) {
      std::memcpy(&keys[i], &F->mvKeysUn[i], sizeof(RefKp));
      std::memcpy(desc + (size_t)i * 32, F->mDescriptors.ptr(i), 32);
      uRight[i] = F->mvuRight[i]; depth[i] = F->mvDepth[i];
    }
    int c = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
      for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
        cellStart[ix * FRAME_GRID_ROWS + iy] = c;
        for (size_t v : F->mGrid[ix][iy]) items[c++] = (int)v;
      }
    cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = c;
#ifdef SLAM_SHIM_EXTRACTOR
    // the same frame through the device-side replacement of the method's body (shim/FrameStereo.h)
    std::vector<float> u, z;
    ComputeStereoMatchesGPU(exL, exR, bf, bf / fx, n, u, z);
    for (int i = 0; i < n; i++) { uRightDevice[i] = u[i]; depthDevice[i] = z[i]; }
#endif
  }
  F->~Frame();
  return n;
}

}  // extern "C"
