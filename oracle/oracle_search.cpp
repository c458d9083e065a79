// oracle_search.cpp — TEST INFRASTRUCTURE (CPU oracle; never linked into the product library).
//
// Sequential restatement of the candidate loops of the reference's ORBmatcher searches and of the frame grid
// they walk, on flat arrays, parametrised exactly like include/orbgpu.h's orbgpu_search so that the CUDA path
// can be compared record by record.  Reference paths are relative to Source/Libraries/ORB_SLAM2/.
//
//   grid            Frame::AssignFeaturesToGrid / PosInGrid        src/Frame.cc:250-263, 399-411
//   window walk     Frame::GetFeaturesInArea                       src/Frame.cc:341-397
//                   KeyFrame::GetFeaturesInArea                    src/KeyFrame.cc:526-567
//   distance        ORBmatcher::DescriptorDistance                 src/ORBmatcher.cc:1542-1556
//   loops           SearchByProjection(F, MPs)                     src/ORBmatcher.cc:45-128
//                   SearchByBoW(KF, F) / (KF, KF)                  :161-273 / :494-612
//                   SearchByProjection(KF, Scw, ...)               :275-386
//                   SearchForInitialization                        :388-492
//                   SearchForTriangulation + CheckDistEpipolarLine :614-764, :137-159
//                   Fuse x2                                        :766-907, :909-1027
//                   SearchBySim3                                   :1029-1245
//                   SearchByProjection(Cur, Last) / (Cur, KF)      :1247-1383 / :1385-1504
//
// Pinned by tests/test_matcher_ref.py: the host policy of shim/ORBmatcher.cc running on THIS backend produces the
// same outputs as the reference's own ORBmatcher.cc (oracle/_ref/libref_slam.so) on the same scenes.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/orbgpu.h"

namespace {

// ORBmatcher.cc:1542-1556 — the SWAR popcount over 8 x 32-bit words, as written.
inline int descriptor_distance(const uint8_t *a, const uint8_t *b) {
  int32_t pa[8], pb[8];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    unsigned int v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

}  // namespace

struct OracleFeatures {
  orbgpu_features_desc d;
  std::vector<orbgpu_keypoint> keys;
  std::vector<float> uRight;
  std::vector<uint8_t> desc;
  std::vector<std::vector<int>> grid;     // [ix * rows + iy]
};

extern "C" {

OracleFeatures *orb_oracle_features_create(const orbgpu_features_desc *desc) {
  OracleFeatures *F = new OracleFeatures;
  F->d = *desc;
  const int n = desc->n;
  F->keys.assign(desc->keys_un, desc->keys_un + n);
  if (desc->u_right) F->uRight.assign(desc->u_right, desc->u_right + n); else F->uRight.assign(n, -1.f);
  F->desc.assign(desc->descriptors, desc->descriptors + (size_t)n * 32);
  F->d.keys_un = nullptr; F->d.u_right = nullptr; F->d.descriptors = nullptr;
  F->grid.assign((size_t)desc->grid_cols * desc->grid_rows, std::vector<int>());
  for (int i = 0; i < n; i++) {                                       // Frame.cc:256-262
    const orbgpu_keypoint &kp = F->keys[i];
    const int posX = (int)roundf((kp.x - desc->min_x) * desc->grid_w_inv);   // Frame.cc:400-401 (round, half away from zero)
    const int posY = (int)roundf((kp.y - desc->min_y) * desc->grid_h_inv);
    if (posX < 0 || posX >= desc->grid_cols || posY < 0 || posY >= desc->grid_rows) continue;   // Frame.cc:405-407
    F->grid[(size_t)posX * desc->grid_rows + posY].push_back(i);
  }
  return F;
}
void orb_oracle_features_destroy(OracleFeatures *F) { delete F; }
int orb_oracle_features_count(const OracleFeatures *F) { return F->d.n; }

void orb_oracle_features_grid(const OracleFeatures *F, int32_t *cellStart, int32_t *cellItems, int32_t *nItems) {
  int k = 0;
  for (size_t c = 0; c < F->grid.size(); c++) {
    cellStart[c] = k;
    for (int i : F->grid[c]) cellItems[k++] = i;
  }
  cellStart[F->grid.size()] = k;
  if (nItems) *nItems = k;
}

// Frame.cc:341-397 / KeyFrame.cc:526-567.  Level rule: Frame.cc:369,377-383 — written so that the KeyFrame variant
// (no level test) is min_level = -1, max_level = -1.
static void features_in_area(const OracleFeatures *F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int> &out) {
  out.clear();
  const orbgpu_features_desc &d = F->d;
  const int nMinCellX = std::max(0, (int)floorf((x - d.query_min_x - r) * d.grid_w_inv));
  if (nMinCellX >= d.grid_cols) return;
  const int nMaxCellX = std::min(d.grid_cols - 1, (int)ceilf((x - d.query_min_x + r) * d.grid_w_inv));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)floorf((y - d.query_min_y - r) * d.grid_h_inv));
  if (nMinCellY >= d.grid_rows) return;
  const int nMaxCellY = std::min(d.grid_rows - 1, (int)ceilf((y - d.query_min_y + r) * d.grid_h_inv));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
      for (int idx : F->grid[(size_t)ix * d.grid_rows + iy]) {
        const orbgpu_keypoint &kp = F->keys[idx];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (fabsf(distx) < r && fabsf(disty) < r) out.push_back(idx);
      }
}

int orb_oracle_features_in_area(const OracleFeatures *F, const orbgpu_query *q, int nq, int32_t *offsets, int32_t *items, int cap) {
  std::vector<int> v;
  int k = 0;
  bool overflow = false;
  for (int i = 0; i < nq; i++) {
    offsets[i] = k;
    features_in_area(F, q[i].u, q[i].v, q[i].r, q[i].min_level, q[i].max_level, v);
    for (int idx : v) { if (k < cap) items[k] = idx; else overflow = true; k++; }
  }
  offsets[nq] = k;
  return overflow ? ORBGPU_E_CAPACITY : ORBGPU_OK;
}

// ORBmatcher.cc:137-159 with kp1 = (q.u, q.v), kp2 = candidate; sigma2 = mvLevelSigma2[kp2.octave].
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, const float *F12, float sigma2) {
  const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
  const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
  const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
  const float num = a * x2 + b * y2 + c;
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * sigma2;
}

// The candidate loops.  `qdesc` [nq][32]; occupied [n] or NULL; result [nq].
int orb_oracle_search(const OracleFeatures *F, const orbgpu_query *queries, const uint8_t *qdesc, int nq, const uint8_t *occupied,
                      const int32_t *cand, int ncand, const orbgpu_search_policy *P, orbgpu_match *result) {
  const int n = F->d.n;
  std::vector<uint8_t> taken(n, 0);                    // greedy TAKEN: keypoint unavailable to later queries
  std::vector<int> matchedDistance(n, INT_MAX);        // greedy DISTANCE: vMatchedDistance (ORBmatcher.cc:404)
  if (occupied) for (int i = 0; i < n; i++) taken[i] = occupied[i] ? 1 : 0;
  std::vector<int> window;
  for (int qi = 0; qi < nq; qi++) {
    const orbgpu_query &q = queries[qi];
    const uint8_t *dq = qdesc + (size_t)qi * 32;
    const int *list; int nlist;
    if (P->source == ORBGPU_SRC_WINDOW) {
      features_in_area(F, q.u, q.v, q.r, q.min_level, q.max_level, window);
      list = window.data(); nlist = (int)window.size();
    } else {
      if (q.list_begin < 0 || q.list_end > ncand || q.list_begin > q.list_end) return ORBGPU_E_INVALID;
      list = cand + q.list_begin; nlist = q.list_end - q.list_begin;
    }
    int bestDist = 256, bestDist2 = 256, bestLevel = -1, bestLevel2 = -1, bestIdx = -1;
    for (int k = 0; k < nlist; k++) {
      const int idx = list[k];
      if (idx < 0 || idx >= n) return ORBGPU_E_INVALID;
      if (taken[idx]) continue;                                                        // :84-86, :208, :362, :538, :667, :1302, :1452
      const orbgpu_keypoint &kp = F->keys[idx];
      if (P->source == ORBGPU_SRC_LIST) {       // explicit lists carry no level gate in the reference; (-1, -1) disables it
        if (kp.octave < q.min_level) continue;
        if (q.max_level >= 0 && kp.octave > q.max_level) continue;
      }
      if (P->stereo_gate && F->uRight[idx] > 0) {                                      // :88-92, :1306-1311
        const float er = fabsf(q.ur - F->uRight[idx]);
        if (er > q.ur_tol) continue;
      }
      if (P->gate == ORBGPU_GATE_FUSE) {                                               // :843-866
        const float ex = q.u - kp.x, ey = q.v - kp.y;
        if (F->uRight[idx] >= 0) {
          const float er = q.ur - F->uRight[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * P->inv_level_sigma2[kp.octave] > 7.8) continue;
        } else {
          const float e2 = ex * ex + ey * ey;
          if (e2 * P->inv_level_sigma2[kp.octave] > 5.99) continue;
        }
      }
      bool bStereo2 = false;
      if (P->gate == ORBGPU_GATE_EPIPOLAR) {                                           // :669-673
        bStereo2 = F->uRight[idx] >= 0;
        if (P->only_stereo && !bStereo2) continue;
      }
      const int dist = descriptor_distance(dq, F->desc.data() + (size_t)idx * 32);
      if (P->greedy == ORBGPU_GREEDY_DISTANCE && matchedDistance[idx] <= dist) continue;   // :447-448
      if (P->gate == ORBGPU_GATE_EPIPOLAR) {
        if (dist > P->accept_max || dist > std::min(bestDist, P->accept_max)) continue;    // :678-679 (bestDist starts at TH_LOW)
        const bool bStereo1 = (q.flags & ORBGPU_Q_STEREO) != 0;
        if (!bStereo1 && !bStereo2) {                                                  // :683-689
          const float distex = P->ex - kp.x, distey = P->ey - kp.y;
          if (distex * distex + distey * distey < 100 * P->scale_factors[kp.octave]) continue;
        }
        if (check_dist_epipolar_line(q.u, q.v, kp.x, kp.y, P->f12, P->level_sigma2[kp.octave])) {   // :691-694
          bestIdx = idx; bestDist = dist; bestLevel = kp.octave;
        }
        continue;
      }
      if (dist < bestDist) {                                                           // :102-111
        bestDist2 = bestDist; bestLevel2 = bestLevel;
        bestDist = dist; bestLevel = kp.octave; bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = kp.octave; bestDist2 = dist;
      }
    }
    bool accepted = bestIdx >= 0 && bestDist <= P->accept_max;
    if (accepted && P->ratio_mode == ORBGPU_RATIO_SAME_LEVEL && bestLevel == bestLevel2 && bestDist > P->ratio * bestDist2) accepted = false;   // :116-118
    if (accepted && P->ratio_mode == ORBGPU_RATIO_LESS && !((float)bestDist < P->ratio * (float)bestDist2)) accepted = false;                   // :225
    orbgpu_match &m = result[qi];
    m.idx = bestIdx; m.dist = bestDist; m.dist2 = bestDist2; m.level = (int8_t)bestLevel; m.level2 = (int8_t)bestLevel2;
    m.accepted = accepted ? 1 : 0; m.pad = 0;
    if (accepted) {
      if (P->greedy == ORBGPU_GREEDY_TAKEN && (q.flags & ORBGPU_Q_BLOCKS)) taken[bestIdx] = 1;     // :120, :229, :380, :567, :1325, :1466
      if (P->greedy == ORBGPU_GREEDY_DISTANCE) matchedDistance[bestIdx] = bestDist;                // :462
    }
  }
  return ORBGPU_OK;
}

}  // extern "C"

// ---- DBoW2 vocabulary tree descent -----------------------------------------------------------------------------------
// TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
//   (Source/ThirdParty/DBoW2/DBoW2-local/include/DBoW2/TemplatedVocabulary.h:1216-1258) with FORB::distance
//   (…/src/FORB.cpp:77-100: 64-bit SWAR popcount over 4 words, returned as double).
// Where the reference leaves *nid uninitialised (leaf above level L - levelsup) the leaf's node id is reported.
namespace {
inline double forb_distance(const uint8_t *a, const uint8_t *b) {
  uint64_t pa[4], pb[4];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  uint64_t v, ret = 0;
  for (size_t i = 0; i < 4; ++i) {
    v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & (uint64_t) ~(uint64_t)0 / 3);
    v = (v & (uint64_t) ~(uint64_t)0 / 15 * 3) + ((v >> 2) & (uint64_t) ~(uint64_t)0 / 15 * 3);
    v = (v + (v >> 4)) & (uint64_t) ~(uint64_t)0 / 255 * 15;
    ret += (uint64_t)(v * ((uint64_t) ~(uint64_t)0 / 255)) >> (sizeof(uint64_t) - 1) * 8;
  }
  return static_cast<double>(ret);
}
}  // namespace

extern "C" int orb_oracle_voc_transform(int nodes, const int32_t *childStart, const int32_t *children, const uint8_t *nodeDesc, const int32_t *wordId,
                                        int depthLevels, const uint8_t *features, int n, int levelsup, int32_t *wordIds, int32_t *nodeIds) {
  (void)nodes;
  const int nid_level = depthLevels - levelsup;
  for (int f = 0; f < n; f++) {
    const uint8_t *feature = features + (size_t)f * 32;
    bool nidSet = false;
    int nid = 0;
    if (nid_level <= 0) { nid = 0; nidSet = true; }                 // :1225
    int final_id = 0, current_level = 0;
    while (childStart[final_id] != childStart[final_id + 1]) {      // do { ... } while (!isLeaf); an empty tree is handled by the caller (:1131)
      ++current_level;
      const int cs = childStart[final_id], ce = childStart[final_id + 1];
      final_id = children[cs];
      double best_d = forb_distance(feature, nodeDesc + (size_t)final_id * 32);
      for (int c = cs + 1; c < ce; c++) {
        const int id = children[c];
        const double d = forb_distance(feature, nodeDesc + (size_t)id * 32);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) { nid = final_id; nidSet = true; }
    }
    wordIds[f] = wordId[final_id];
    nodeIds[f] = nidSet ? nid : final_id;
  }
  return ORBGPU_OK;
}

// ---- input staging: OpenCV's fixed-point cvtColor(*2GRAY) and remap(INTER_LINEAR, BORDER_CONSTANT) ----
// Call sites: src/Tracking.cc:166-182 (cvtColor), Examples/Stereo/stereo_euroc.cc:159-160 (remap).  OpenCV is not under
// /root/reference; both restatements are pinned bit-exact against cv2 4.13.0 by tests/test_staging.py.
extern "C" void orb_oracle_cvt_gray(const uint8_t *src, int channels, int code, int width, int height, uint8_t *dst) {
  const int c0 = code == ORBGPU_RGB2GRAY ? 9798 : 3735, c2 = code == ORBGPU_RGB2GRAY ? 3735 : 9798;
  for (size_t i = 0; i < (size_t)width * height; i++) {
    const uint8_t *p = src + i * channels;
    dst[i] = (uint8_t)((p[0] * c0 + p[1] * 19235 + p[2] * c2 + (1 << 14)) >> 15);
  }
}

extern "C" void orb_oracle_remap(const uint8_t *src, int sw, int sh, const float *mapX, const float *mapY, int width, int height, uint8_t *dst) {
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const size_t m = (size_t)y * width + x;
      const int sx = (int)lrintf(mapX[m] * 32.f), sy = (int)lrintf(mapY[m] * 32.f);      // cvRound(v * INTER_TAB_SIZE)
      const int fx = sx & 31, fy = sy & 31;
      const int ix = std::max(-32768, std::min(32767, sx >> 5)), iy = std::max(-32768, std::min(32767, sy >> 5));
      auto tap = [&](int yy, int xx) -> int { return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sw + xx] : 0; };
      const int v = (32 - fx) * (32 - fy) * 32 * tap(iy, ix) + fx * (32 - fy) * 32 * tap(iy, ix + 1) + (32 - fx) * fy * 32 * tap(iy + 1, ix) +
                    fx * fy * 32 * tap(iy + 1, ix + 1);
      dst[m] = (uint8_t)((v + (1 << 14)) >> 15);
    }
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:229-321), batched over groups of descriptor rows ----
extern "C" void orb_oracle_distinctive(const uint8_t *descriptors, const int32_t *offsets, int groups, int32_t *bestIndex, int32_t *bestMedian) {
  for (int g = 0; g < groups; g++) {
    const int lo = offsets[g], N = offsets[g + 1] - lo;
    if (N <= 0) { bestIndex[g] = -1; bestMedian[g] = 0; continue; }                       // :243-244, :257-258: nothing to do
    std::vector<std::vector<int>> Distances(N, std::vector<int>(N, 0));                     // :287-296
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        const int distij = descriptor_distance(descriptors + (size_t)(lo + i) * 32, descriptors + (size_t)(lo + j) * 32);
        Distances[i][j] = distij; Distances[j][i] = distij;
      }
    int BestMedian = INT_MAX, BestIdx = 0;                                                   // :299-310
    for (int i = 0; i < N; i++) {
      std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
      std::sort(vDists.begin(), vDists.end());
      const int median = vDists[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    bestIndex[g] = BestIdx; bestMedian[g] = BestMedian;
  }
}
