// oracle/oracle_cpu.cpp
//
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Dependency-free C++17 restatement of the reference's CPU ORB front-end:
//   ORBextractor        Source/Libraries/ORB_SLAM2/src/ORBextractor.cc:72-146, 407-815, 969-1065
//   DescriptorDistance  Source/Libraries/ORB_SLAM2/src/ORBmatcher.cc:1542-1556
//   SearchByBoW policy  Source/Libraries/ORB_SLAM2/src/ORBmatcher.cc:201-243, 1506-1538
//   ComputeStereoMatches Source/Libraries/ORB_SLAM2/src/Frame.cc:477-646
// (paths relative to /root/reference).  The OpenCV primitives the reference
// calls (cv::resize INTER_LINEAR, copyMakeBorder REFLECT_101, cv::FAST 9/16 +
// NMS, GaussianBlur 7x7 sigma 2, fastAtan2, cvRound) are NOT in the reference
// tree and OpenCV is un-pinned there; they are restated here from OpenCV's
// published fixed-point algorithms and pinned bit-exact against cv2 4.13.0 by
// tests/test_oracle_vs_cv2.py.
//
// PINNING.  The reference ships no tests, golden vectors or fixtures for this
// path and its build (CMake + OpenCV C++, Eigen, Pangolin) is not possible
// here.  The oracle is pinned three ways instead:
//  1. every OpenCV primitive bit-exact vs real cv2 4.13.0;
//  2. the whole extractor vs an independent cv2-driven Python restatement
//     (oracle/oracle_cv2.py) and the golden vectors it produced (tests/golden);
//  3. the whole extractor vs OUTPUTS OF THE REFERENCE ITSELF RUN HERE: its own,
//     unmodified ORBextractor.cc compiled from /root/reference against an
//     OpenCV API stand-in (oracle/ref_shim -> oracle/_ref/libref_orb.so;
//     primitives = the restatements of (1), monotone allocator) — keypoints,
//     order and descriptors byte-identical (tests/test_ref_vs_oracle.py).
// ORBmatcher.cc / Frame.cc pull in the whole SLAM object graph and cannot be
// built that way: DescriptorDistance, the best/second-best scan and
// ComputeStereoMatches remain restatements (parity unpinned for those two).
//
// Two reference behaviours are not functions of the input and are fixed here
// (see DESIGN.md "Determinisation"):
//  * DistributeOctTree sorts (count, ExtractorNode*) — ties fall to heap
//    addresses (ORBextractor.cc:659-661).  Rule here: ties ordered by node
//    creation sequence (what a monotone allocator yields).
//  * Frame::ComputeStereoMatches reads `mb` before it is assigned
//    (Frame.cc:505 vs :124).  `mb` is an explicit argument here.
//
// Build: g++ -O3 -std=c++17 -shared -fPIC (no -march: the reference sets none,
// so float expressions are plain SSE2 mul/add, never fused).

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>

namespace {

const int kPatchSize = 31;      // ORBextractor.cc:72
const int kHalfPatch = 15;      // :73
const int kEdge = 19;           // :74

const int8_t kPattern[1024] = {
#include "../include/orb_pattern.inc"
};

// cvRound(float)/cvRound(double) on x86-64 = cvtss2si / cvtsd2si: round half to even.
inline int round_f(float v) { return (int)lrintf(v); }
inline int round_d(double v) { return (int)lrint(v); }

struct KeyPointRec {  // cv::KeyPoint layout, 28 bytes
  float x, y, size, angle, response;
  int octave, class_id;
};
static_assert(sizeof(KeyPointRec) == 28, "cv::KeyPoint is 7 x 4 bytes");

// One pyramid level: interior w x h inside a plane padded by kEdge on every side.
struct Plane {
  int w = 0, h = 0, stride = 0;
  std::vector<uint8_t> buf;
  void alloc(int w_, int h_) {
    w = w_; h = h_; stride = w + 2 * kEdge;
    buf.assign((size_t)stride * (h + 2 * kEdge), 0);
  }
  uint8_t *px(int x, int y) { return &buf[(size_t)(y + kEdge) * stride + x + kEdge]; }
  const uint8_t *px(int x, int y) const { return &buf[(size_t)(y + kEdge) * stride + x + kEdge]; }
};

// ---------------------------------------------------------------- OpenCV primitives

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) on CV_8UC1 (called at ORBextractor.cc:1054).
// OpenCV's fixed-point path: 11-bit coefficients, horizontal pass in int32,
// vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
void resize_axis_tables(int srcN, int dstN, std::vector<int> &ofs, std::vector<short> &coef) {
  const double inv_scale = (double)dstN / srcN;
  const double scale = 1. / inv_scale;
  ofs.resize(dstN);
  coef.resize(2 * dstN);
  for (int d = 0; d < dstN; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { f = 0; s = 0; }
    if (s >= srcN - 1) { f = 0; s = srcN - 1; }
    ofs[d] = s;
    coef[2 * d] = (short)round_f((1.f - f) * 2048.f);
    coef[2 * d + 1] = (short)round_f(f * 2048.f);
  }
}

void resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                      int dstride) {
  std::vector<int> xo, yo;
  std::vector<short> xa, ya;
  resize_axis_tables(sw, dw, xo, xa);
  resize_axis_tables(sh, dh, yo, ya);
  std::vector<int> row0(dw), row1(dw);
  int have0 = -1, have1 = -1;
  auto hpass = [&](int sy, std::vector<int> &out) {
    const uint8_t *S = src + (size_t)sy * sstride;
    for (int x = 0; x < dw; x++) {
      int s = xo[x];
      int s1 = std::min(s + 1, sw - 1);
      out[x] = S[s] * xa[2 * x] + S[s1] * xa[2 * x + 1];
    }
  };
  for (int y = 0; y < dh; y++) {
    int s0 = yo[y], s1 = std::min(s0 + 1, sh - 1);
    if (have0 != s0) {
      if (have1 == s0) { row0.swap(row1); std::swap(have0, have1); }
      else { hpass(s0, row0); have0 = s0; }
    }
    if (have1 != s1) { hpass(s1, row1); have1 = s1; }
    const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
    uint8_t *D = dst + (size_t)y * dstride;
    for (int x = 0; x < dw; x++)
      D[x] = (uint8_t)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

// cv::copyMakeBorder(..., 19,19,19,19, BORDER_REFLECT_101) in place on a padded plane
// (ORBextractor.cc:1057-1061): index -k -> k, n-1+k -> n-1-k.
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
void fill_border(Plane &pl) {
  for (int y = -kEdge; y < pl.h + kEdge; y++) {
    int sy = reflect101(y, pl.h);
    for (int x = -kEdge; x < pl.w + kEdge; x++) {
      if (y >= 0 && y < pl.h && x >= 0 && x < pl.w) { x = pl.w - 1; continue; }
      *pl.px(x, y) = *pl.px(reflect101(x, pl.w), sy);
    }
  }
}

// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8UC1 (ORBextractor.cc:1018-1019):
// OpenCV's bit-exact fixed-point separable filter, kernel {18,34,48,56,48,34,18}/256,
// horizontal in 8.8, vertical in 16.16, round-to-nearest at the end.
const int kGauss[7] = {18, 34, 48, 56, 48, 34, 18};
void gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
  std::vector<uint16_t> tmp((size_t)w * h);
  std::vector<uint8_t> line(w + 6);
  for (int y = 0; y < h; y++) {
    const uint8_t *S = src + (size_t)y * sstride;
    for (int k = 0; k < 3; k++) { line[k] = S[reflect101(k - 3, w)]; line[w + 3 + k] = S[reflect101(w + k, w)]; }
    std::memcpy(&line[3], S, w);
    uint16_t *T = &tmp[(size_t)y * w];
    for (int x = 0; x < w; x++) {
      const uint8_t *p = &line[x];
      T[x] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
  }
  for (int y = 0; y < h; y++) {
    const uint16_t *R[7];
    for (int k = -3; k <= 3; k++) R[k + 3] = &tmp[(size_t)reflect101(y + k, h) * w];
    uint8_t *D = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++) {
      const uint32_t acc = 18u * (R[0][x] + R[6][x]) + 34u * (R[1][x] + R[5][x]) + 48u * (R[2][x] + R[4][x]) + 56u * R[3][x];
      D[x] = (uint8_t)((acc + 0x8000u) >> 16);
    }
  }
}

// cv::fastAtan2 scalar path (called at ORBextractor.cc:99); degrees in [0,360), no FMA.
float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16 (ORBextractor.cc:774-780).
// Ring order as in OpenCV: starts at (0,3) and runs towards (+x).
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

struct CellKp { int x, y, resp; };

// Largest t for which the pixel is still a FAST-9/16 corner, plus one:
// m = max over the 16 arcs of 9 of min(+-(ring - centre)).  corner(t) <=> m > t;
// OpenCV's cornerScore() of a detected corner is m - 1.
inline int fast_arc_margin(const uint8_t *p, const int *off) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = (int)p[off[k]] - v;
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -255;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

void fast9_cell(const uint8_t *img, int stride, int cw, int ch, int th, std::vector<CellKp> &out) {
  out.clear();
  if (cw < 7 || ch < 7) return;
  int off[16];
  for (int k = 0; k < 16; k++) off[k] = kRingDy[k] * stride + kRingDx[k];
  // three rolling rows of scores, as cv::FAST keeps them; zero outside [3, n-3)
  std::vector<int> rows[3];
  for (auto &r : rows) r.assign(cw, 0);
  std::vector<int> pos[3];
  for (int y = 3; y <= ch - 3; y++) {
    std::vector<int> &cur = rows[y % 3];
    std::vector<int> &cpos = pos[y % 3];
    std::fill(cur.begin(), cur.end(), 0);
    cpos.clear();
    if (y < ch - 3) {
      const uint8_t *row = img + (size_t)y * stride;
      for (int x = 3; x < cw - 3; x++) {
        const uint8_t *p = row + x;
        const int v = p[0], hi = v + th, lo = v - th;
        // high-speed test on the four compass pairs before the full ring
        int c = 0;
        {
          int a = p[off[0]], b = p[off[8]];
          c = ((a > hi) | (b > hi)) | (((a < lo) | (b < lo)) << 1);
          if (!c) continue;
          a = p[off[4]]; b = p[off[12]];
          c &= ((a > hi) | (b > hi)) | (((a < lo) | (b < lo)) << 1);
          if (!c) continue;
          a = p[off[2]]; b = p[off[10]];
          c &= ((a > hi) | (b > hi)) | (((a < lo) | (b < lo)) << 1);
          if (!c) continue;
          a = p[off[6]]; b = p[off[14]];
          c &= ((a > hi) | (b > hi)) | (((a < lo) | (b < lo)) << 1);
          if (!c) continue;
        }
        int m = fast_arc_margin(p, off);
        if (m > th) { cur[x] = m - 1; cpos.push_back(x); }
      }
    }
    if (y == 3) continue;
    // non-max suppression of row y-1 against rows y-2, y-1, y
    const std::vector<int> &prv = rows[(y - 1) % 3], &pp = rows[(y - 2) % 3];
    for (int x : pos[(y - 1) % 3]) {
      int s = prv[x];
      if (s > prv[x - 1] && s > prv[x + 1] && s > pp[x - 1] && s > pp[x] && s > pp[x + 1] &&
          s > cur[x - 1] && s > cur[x] && s > cur[x + 1])
        out.push_back({x, y - 1, s});
    }
  }
}

// ---------------------------------------------------------------- extractor

struct Cand { int x, y, resp; };  // octree coordinates: level coordinates minus minBorder

struct LevelTrace {               // stage outputs kept for stage-parity tests
  std::vector<Cand> cands;        // vToDistributeKeys, in reference order
  std::vector<Cand> picked;       // DistributeOctTree result, in list order
};

struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;             // ORBextractor.h:91 — a double member
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> quota;
  int umax[kHalfPatch + 1];
  std::vector<Plane> pyr, blurred;
  std::vector<LevelTrace> trace;
  int tie_lo = 0;                 // 0: later-created node first among equal counts (default); 1: earlier first

  Extractor(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    // ORBextractor.cc:411-425
    scale.resize(nl); sigma2.resize(nl); invScale.resize(nl); invSigma2.resize(nl);
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);
      sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nl; i++) {
      invScale[i] = 1.0f / scale[i];
      invSigma2[i] = 1.0f / sigma2[i];
    }
    // ORBextractor.cc:429-441
    quota.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float want = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      quota[l] = round_f(want);
      sum += quota[l];
      want *= factor;
    }
    quota[nl - 1] = std::max(nf - sum, 0);
    // ORBextractor.cc:449-463
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = round_d(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    pyr.resize(nl); blurred.resize(nl); trace.resize(nl);
  }

  // ORBextractor.cc:1041-1065
  void compute_pyramid(const uint8_t *img, int w, int h, int step) {
    for (int l = 0; l < nlevels; l++) {
      float s = invScale[l];
      int lw = round_f((float)w * s), lh = round_f((float)h * s);
      pyr[l].alloc(lw, lh);
      if (l == 0) {
        for (int y = 0; y < h; y++) std::memcpy(pyr[0].px(0, y), img + (size_t)y * step, w);
      } else {
        resize_linear_u8(pyr[l - 1].px(0, 0), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].stride,
                         pyr[l].px(0, 0), lw, lh, pyr[l].stride);
      }
      fill_border(pyr[l]);
    }
  }

  // ORBextractor.cc:733-799 — candidates of one level in reference order
  void detect_level(int l, std::vector<Cand> &out) const {
    const Plane &P = pyr[l];
    out.clear();
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = P.w - kEdge + 3, maxBY = P.h - kEdge + 3;
    const float W = 30;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;  // reference divides by zero here; treat as "no cells"
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<CellKp> cell;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
        fast9_cell(P.px(x0, y0), P.stride, cw, ch, iniTh, cell);
        if (cell.empty()) fast9_cell(P.px(x0, y0), P.stride, cw, ch, minTh, cell);
        for (const CellKp &k : cell) out.push_back({k.x + j * wCell, k.y + i * hCell, k.resp});
      }
    }
  }

  // ORBextractor.cc:475-731 (DivideNode + DistributeOctTree)
  struct Node {
    int x0, x1, y0, y1;            // UL.x, UR.x, UL.y, BL.y (nodes stay axis-aligned rectangles)
    std::vector<int> keys;         // indices into the candidate list, input order preserved
    bool noMore = false;
    long seq = 0;                  // creation sequence = deterministic stand-in for the heap address
    std::list<Node>::iterator self;
  };

  static void divide(const Node &n, const std::vector<Cand> &c, Node ch[4]) {
    const int hx = (int)std::ceil((float)(n.x1 - n.x0) / 2);
    const int hy = (int)std::ceil((float)(n.y1 - n.y0) / 2);
    const int mx = n.x0 + hx, my = n.y0 + hy;
    ch[0].x0 = n.x0; ch[0].x1 = mx;   ch[0].y0 = n.y0; ch[0].y1 = my;
    ch[1].x0 = mx;   ch[1].x1 = n.x1; ch[1].y0 = n.y0; ch[1].y1 = my;
    ch[2].x0 = n.x0; ch[2].x1 = mx;   ch[2].y0 = my;   ch[2].y1 = n.y1;
    ch[3].x0 = mx;   ch[3].x1 = n.x1; ch[3].y0 = my;   ch[3].y1 = n.y1;
    for (int k : n.keys) {
      const float px = (float)c[k].x, py = (float)c[k].y;
      int q = (px < (float)mx) ? ((py < (float)my) ? 0 : 2) : ((py < (float)my) ? 1 : 3);
      ch[q].keys.push_back(k);
    }
    for (int q = 0; q < 4; q++) ch[q].noMore = ch[q].keys.size() == 1;
  }

  std::vector<Cand> distribute(const std::vector<Cand> &c, int minX, int maxX, int minY, int maxY,
                               int N) const {
    std::vector<Cand> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return result;  // reference: division by zero (aspect < 0.5)
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> L;
    long seq = 0;
    std::vector<Node *> roots(nIni);
    for (int i = 0; i < nIni; i++) {
      Node n;
      n.x0 = (int)(hX * (float)i);
      n.x1 = (int)(hX * (float)(i + 1));
      n.y0 = 0; n.y1 = maxY - minY;
      n.seq = seq++;
      L.push_back(n);
      roots[i] = &L.back();
    }
    for (size_t k = 0; k < c.size(); k++) roots[(size_t)((float)c[k].x / hX)]->keys.push_back((int)k);
    for (auto it = L.begin(); it != L.end();) {
      if (it->keys.size() == 1) { it->noMore = true; ++it; }
      else if (it->keys.empty()) it = L.erase(it);
      else ++it;
    }

    typedef std::pair<int, Node *> SizeNode;
    auto by_size_then_seq = [this](const SizeNode &a, const SizeNode &b) {
      if (a.first != b.first) return a.first < b.first;
      return tie_lo ? a.second->seq > b.second->seq : a.second->seq < b.second->seq;
    };
    std::vector<SizeNode> expandable;
    // Splits *it, pushes its non-empty children to the front (n1..n4), erases it.
    auto split = [&](std::list<Node>::iterator it, int *nToExpand) {
      Node ch[4];
      divide(*it, c, ch);
      for (int q = 0; q < 4; q++) {
        if (ch[q].keys.empty()) continue;
        ch[q].seq = seq++;
        L.push_front(ch[q]);
        L.front().self = L.begin();
        if (ch[q].keys.size() > 1) {
          if (nToExpand) ++*nToExpand;
          expandable.push_back(SizeNode((int)ch[q].keys.size(), &L.front()));
        }
      }
      return L.erase(it);
    };

    bool finish = false;
    while (!finish) {
      int prevSize = (int)L.size();
      int nToExpand = 0;
      expandable.clear();
      for (auto it = L.begin(); it != L.end();) {
        if (it->noMore) { ++it; continue; }
        it = split(it, &nToExpand);
      }
      if ((int)L.size() >= N || (int)L.size() == prevSize) {
        finish = true;
      } else if ((int)L.size() + nToExpand * 3 > N) {
        while (!finish) {
          prevSize = (int)L.size();
          std::vector<SizeNode> prev = expandable;
          expandable.clear();
          std::sort(prev.begin(), prev.end(), by_size_then_seq);
          for (int j = (int)prev.size() - 1; j >= 0; j--) {
            split(prev[j].second->self, nullptr);
            if ((int)L.size() >= N) break;
          }
          if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
        }
      }
    }
    result.reserve(L.size());
    for (const Node &n : L) {
      int best = n.keys[0];
      for (size_t k = 1; k < n.keys.size(); k++)
        if (c[n.keys[k]].resp > c[best].resp) best = n.keys[k];
      result.push_back(c[best]);
    }
    return result;
  }

  // ORBextractor.cc:76-100
  float ic_angle(const Plane &P, int x, int y) const {
    int m01 = 0, m10 = 0;
    const uint8_t *c = P.px(x, y);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    const int step = P.stride;
    for (int v = 1; v <= kHalfPatch; ++v) {
      int vsum = 0, d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int a = c[u + v * step], b = c[u - v * step];
        vsum += a - b;
        m10 += u * (a + b);
      }
      m01 += v * vsum;
    }
    return fast_atan2_deg((float)m01, (float)m10);
  }

  // ORBextractor.cc:102-146
  static void describe(const Plane &B, int x, int y, float angleDeg, uint8_t *desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angleDeg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t *c = B.px(x, y);
    const int step = B.stride;
    const int8_t *p = kPattern;
    for (int i = 0; i < 32; i++) {
      int val = 0;
      for (int k = 0; k < 8; k++, p += 4) {
        int t0 = c[round_f(p[0] * b + p[1] * a) * step + round_f(p[0] * a - p[1] * b)];
        int t1 = c[round_f(p[2] * b + p[3] * a) * step + round_f(p[2] * a - p[3] * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // ORBextractor.cc:978-1039; returns the keypoint count (<= cap, or -needed if cap too small)
  int extract(const uint8_t *img, int w, int h, int step, KeyPointRec *kps, uint8_t *desc, int cap) {
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(img, w, h, step);
    int total = 0;
    for (int l = 0; l < nlevels; l++) {
      const Plane &P = pyr[l];
      const int minBX = kEdge - 3, minBY = minBX, maxBX = P.w - kEdge + 3, maxBY = P.h - kEdge + 3;
      detect_level(l, trace[l].cands);
      trace[l].picked = distribute(trace[l].cands, minBX, maxBX, minBY, maxBY, quota[l]);
      total += (int)trace[l].picked.size();
    }
    if (total > cap) return -total;
    int n = 0;
    for (int l = 0; l < nlevels; l++) {
      const std::vector<Cand> &sel = trace[l].picked;
      if (sel.empty()) continue;
      const Plane &P = pyr[l];
      blurred[l].alloc(P.w, P.h);
      gaussian7_u8(P.px(0, 0), P.w, P.h, P.stride, blurred[l].px(0, 0), blurred[l].stride);
      const int patch = (int)(kPatchSize * scale[l]);
      for (const Cand &k : sel) {
        const int x = k.x + kEdge - 3, y = k.y + kEdge - 3;
        KeyPointRec &r = kps[n];
        r.angle = ic_angle(P, x, y);
        describe(blurred[l], x, y, r.angle, desc + (size_t)n * 32);
        r.x = (float)x; r.y = (float)y;
        if (l != 0) { r.x *= scale[l]; r.y *= scale[l]; }
        r.size = (float)patch;
        r.response = (float)k.resp;
        r.octave = l;
        r.class_id = -1;
        n++;
      }
    }
    return n;
  }
};

// ---------------------------------------------------------------- matcher

// ORBmatcher.cc:1542-1556 (same SWAR popcount over 8 x 32 bit)
inline int descriptor_distance(const uint8_t *a, const uint8_t *b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// Best / second-best scan of ORBmatcher.cc:201-221 over a candidate range.
inline void scan_top2(const uint8_t *q, const uint8_t *train, const int *cand, int ncand, int base,
                      int &d1, int &i1, int &d2) {
  for (int k = 0; k < ncand; k++) {
    const int idx = cand ? cand[k] : base + k;
    const int d = descriptor_distance(q, train + (size_t)idx * 32);
    if (d < d1) { d2 = d1; d1 = d; i1 = idx; }
    else if (d < d2) d2 = d;
  }
}

// ---------------------------------------------------------------- stereo

// Read-only view of one pyramid level (interior pointer + row stride) so that the stereo restatement can run on the
// oracle's own planes or on the reference-built checker's cv::Mat pyramids (oracle/ref_shim).
struct LevelView {
  const uint8_t *p; int w, h, stride;
  const uint8_t *px(int x, int y) const { return p + (size_t)y * stride + x; }
};
struct PyramidView {
  std::vector<LevelView> lv;
  std::vector<float> scale, invScale;
};
PyramidView view_of(const Extractor &e) {
  PyramidView v;
  for (int l = 0; l < e.nlevels; l++) v.lv.push_back({e.pyr[l].px(0, 0), e.pyr[l].w, e.pyr[l].h, e.pyr[l].stride});
  v.scale = e.scale; v.invScale = e.invScale;
  return v;
}

// Frame.cc:477-646
int stereo_matches_views(const PyramidView &VL, const PyramidView &VR, const KeyPointRec *kL,
                         const uint8_t *dL, int nL, const KeyPointRec *kR, const uint8_t *dR, int nR,
                         float mbf, float mb, float *uRight, float *depth, int *bestDistOut) {
  for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; if (bestDistOut) bestDistOut[i] = -1; }
  const int thOrbDist = (100 + 50) / 2;
  const int nRows = VL.lv[0].h;
  std::vector<std::vector<int>> rowIdx(nRows);
  for (int iR = 0; iR < nR; iR++) {
    const float y = kR[iR].y;
    const float r = 2.0f * VL.scale[kR[iR].octave];
    const int maxr = (int)std::ceil(y + r), minr = (int)std::floor(y - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);  // reference: unchecked index
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < nL; iL++) {
    const KeyPointRec &kp = kL[iL];
    const int levelL = kp.octave;
    const float vL = kp.y, uL = kp.x;
    const std::vector<int> &cands = rowIdx[(size_t)vL];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = 100;
    int bestR = 0;
    for (int iR : cands) {
      const KeyPointRec &kr = kR[iR];
      if (kr.octave < levelL - 1 || kr.octave > levelL + 1) continue;
      const float uR = kr.x;
      if (uR >= minU && uR <= maxU) {
        int d = descriptor_distance(dL + (size_t)iL * 32, dR + (size_t)iR * 32);
        if (d < bestDist) { bestDist = d; bestR = iR; }
      }
    }
    if (bestDist >= thOrbDist) continue;
    const float uR0 = kR[bestR].x;
    const float sf = VL.invScale[levelL];
    const float suL = std::round(kp.x * sf), svL = std::round(kp.y * sf), suR0 = std::round(uR0 * sf);
    const int w = 5, Lw = 5;
    const LevelView &PL = VL.lv[levelL], &PR = VR.lv[levelL];
    const float iniu = suR0 + Lw - w, endu = suR0 + Lw + w + 1;
    if (iniu < 0 || endu >= PR.w) continue;
    const int cy = (int)svL, cxL = (int)suL, cxR = (int)suR0;
    const int centreL = *PL.px(cxL, cy);
    int best = INT_MAX, bestInc = 0;
    float dists[2 * 5 + 1];
    for (int inc = -Lw; inc <= Lw; inc++) {
      const int centreR = *PR.px(cxR + inc, cy);
      int sad = 0;
      for (int dy = -w; dy <= w; dy++)
        for (int dx = -w; dx <= w; dx++)
          sad += std::abs((*PL.px(cxL + dx, cy + dy) - centreL) - (*PR.px(cxR + inc + dx, cy + dy) - centreR));
      const float dist = (float)sad;
      if (dist < best) { best = (int)dist; bestInc = inc; }
      dists[Lw + inc] = dist;
    }
    if (bestInc == -Lw || bestInc == Lw) continue;
    const float d1 = dists[Lw + bestInc - 1], d2 = dists[Lw + bestInc], d3 = dists[Lw + bestInc + 1];
    const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = VL.scale[levelL] * ((float)suR0 + (float)bestInc + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)(uL - 0.01); }
      depth[iL] = mbf / disparity;
      uRight[iL] = bestuR;
      if (bestDistOut) bestDistOut[iL] = best;
      distIdx.push_back(std::make_pair(best, iL));
    }
  }
  if (distIdx.empty()) return 0;  // reference: undefined behaviour (Frame.cc:635); defined as "no filtering"
  std::sort(distIdx.begin(), distIdx.end());
  const float median = (float)distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)distIdx.size();
  for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
    if (distIdx[i].first < thDist) break;
    uRight[distIdx[i].second] = -1;
    depth[distIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

int stereo_matches(const Extractor &EL, const Extractor &ER, const KeyPointRec *kL, const uint8_t *dL, int nL,
                   const KeyPointRec *kR, const uint8_t *dR, int nR, float mbf, float mb, float *uRight, float *depth,
                   int *bestDistOut) {
  return stereo_matches_views(view_of(EL), view_of(ER), kL, dL, nL, kR, dR, nR, mbf, mb, uRight, depth, bestDistOut);
}

template <class F>
void parallel_for(int n, int nthreads, F f) {
  nthreads = std::max(1, std::min(nthreads, n));
  if (nthreads == 1) { for (int i = 0; i < n; i++) f(i, 0); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t] { for (int i; (i = next.fetch_add(1)) < n;) f(i, t); });
  for (auto &x : th) x.join();
}

}  // namespace

// ---------------------------------------------------------------- C interface (ctypes)
extern "C" {

void *orb_oracle_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orb_oracle_destroy(void *h) { delete (Extractor *)h; }
void orb_oracle_set_tie_rule(void *h, int lo) { ((Extractor *)h)->tie_lo = lo; }

void orb_oracle_tables(void *h, float *scale, float *invScale, float *sigma2, float *invSigma2,
                       int *quota, int *umax16) {
  Extractor *e = (Extractor *)h;
  for (int i = 0; i < e->nlevels; i++) {
    scale[i] = e->scale[i]; invScale[i] = e->invScale[i];
    sigma2[i] = e->sigma2[i]; invSigma2[i] = e->invSigma2[i];
    quota[i] = e->quota[i];
  }
  for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

int orb_oracle_extract(void *h, const uint8_t *img, int w, int hgt, int step, void *kps,
                       uint8_t *desc, int cap) {
  return ((Extractor *)h)->extract(img, w, hgt, step, (KeyPointRec *)kps, desc, cap);
}

// One frame per thread, each thread owning its extractor (the reference's threading model,
// Frame.cc:87-90).  counts[i] receives the keypoint count of frame i.
void orb_oracle_extract_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                              const uint8_t *imgs, int nimg, int w, int hgt, int step,
                              long img_stride, void *kps, uint8_t *desc, int cap, int *counts,
                              int nthreads) {
  nthreads = std::max(1, std::min(nthreads, nimg));
  std::vector<Extractor *> ex(nthreads);
  for (auto &e : ex) e = new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  parallel_for(nimg, nthreads, [&](int i, int t) {
    counts[i] = ex[t]->extract(imgs + (size_t)i * img_stride, w, hgt, step,
                               (KeyPointRec *)kps + (size_t)i * cap, desc + (size_t)i * cap * 32, cap);
  });
  for (auto e : ex) delete e;
}

int orb_oracle_level_size(void *h, int level, int *w, int *hgt) {
  Extractor *e = (Extractor *)h;
  if (level < 0 || level >= e->nlevels) return -1;
  *w = e->pyr[level].w; *hgt = e->pyr[level].h;
  return 0;
}
// which: 0 = pyramid interior, 1 = pyramid with the 19 px border, 2 = blurred interior
int orb_oracle_get_level(void *h, int level, int which, uint8_t *dst, int dst_step) {
  Extractor *e = (Extractor *)h;
  if (level < 0 || level >= e->nlevels) return -1;
  const Plane &P = which == 2 ? e->blurred[level] : e->pyr[level];
  if (P.buf.empty()) return -2;
  const int b = which == 1 ? kEdge : 0;
  for (int y = -b; y < P.h + b; y++) std::memcpy(dst + (size_t)(y + b) * dst_step, P.px(-b, y), P.w + 2 * b);
  return 0;
}
// stage 0 = FAST candidates (reference order), 1 = octree-selected; rows of (x, y, response)
// in octree coordinates (level coordinates - 16).  Returns the row count.
int orb_oracle_get_stage(void *h, int level, int stage, int *dst, int cap) {
  Extractor *e = (Extractor *)h;
  if (level < 0 || level >= e->nlevels) return -1;
  const std::vector<Cand> &v = stage == 0 ? e->trace[level].cands : e->trace[level].picked;
  if (dst)
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) {
      dst[3 * i] = v[i].x; dst[3 * i + 1] = v[i].y; dst[3 * i + 2] = v[i].resp;
    }
  return (int)v.size();
}

// Stand-alone primitives (pinned against cv2 by tests/test_oracle_vs_cv2.py)
void orb_oracle_resize(const uint8_t *src, int sw, int sh, int sstep, uint8_t *dst, int dw, int dh, int dstep) {
  resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orb_oracle_blur(const uint8_t *src, int w, int h, int sstep, uint8_t *dst, int dstep) {
  gaussian7_u8(src, w, h, sstep, dst, dstep);
}
float orb_oracle_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
int orb_oracle_fast(const uint8_t *img, int w, int h, int step, int th, int *dst, int cap) {
  std::vector<CellKp> v;
  fast9_cell(img, step, w, h, th, v);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) {
    dst[3 * i] = v[i].x; dst[3 * i + 1] = v[i].y; dst[3 * i + 2] = v[i].resp;
  }
  return (int)v.size();
}
void orb_oracle_sincos(float angle, float *c, float *s) { *c = cosf(angle); *s = sinf(angle); }
// Octree alone on an arbitrary candidate list (rows of x, y, response; octree coordinates).
int orb_oracle_distribute(void *h, const int *cands, int n, int minX, int maxX, int minY, int maxY,
                          int N, int *dst, int cap) {
  std::vector<Cand> c(n);
  for (int i = 0; i < n; i++) c[i] = {cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]};
  std::vector<Cand> r = ((Extractor *)h)->distribute(c, minX, maxX, minY, maxY, N);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) {
    dst[3 * i] = r[i].x; dst[3 * i + 1] = r[i].y; dst[3 * i + 2] = r[i].resp;
  }
  return (int)r.size();
}

int orb_oracle_hamming(const uint8_t *a, const uint8_t *b) { return descriptor_distance(a, b); }

// Dense best / second-best over a whole bank (the SearchByBoW inner loop with one giant node,
// ORBmatcher.cc:201-221): strict '<', so the lowest index wins ties; d1/d2 start at 256.
void orb_oracle_top2_dense(const uint8_t *q, int nq, const uint8_t *bank, int nb, int *idx1, int *d1,
                           int *d2, int nthreads) {
  parallel_for(nq, nthreads, [&](int i, int) {
    int a = 256, b = 256, ia = -1;
    scan_top2(q + (size_t)i * 32, bank, nullptr, nb, 0, a, ia, b);
    idx1[i] = ia; d1[i] = a; d2[i] = b;
  });
}
// Candidate-list form (projection / BoW searches): query i scans cand[offsets[i] .. offsets[i+1]).
void orb_oracle_top2_csr(const uint8_t *q, int nq, const uint8_t *train, const int *offsets,
                         const int *cand, int *idx1, int *d1, int *d2) {
  for (int i = 0; i < nq; i++) {
    int a = 256, b = 256, ia = -1;
    scan_top2(q + (size_t)i * 32, train, cand + offsets[i], offsets[i + 1] - offsets[i], 0, a, ia, b);
    idx1[i] = ia; d1[i] = a; d2[i] = b;
  }
}
// Acceptance test of ORBmatcher.cc:224-226.
int orb_oracle_accept(int d1, int d2, float nnratio, int th_low) {
  return d1 <= th_low && (float)d1 < nnratio * (float)d2;
}
// Rotation-consistency filter of ORBmatcher.cc:232-239, 255-270, 1506-1538.  rot[i] = angle1 - angle2
// of match i; keep[i] receives 1 if the match survives.
void orb_oracle_rotation_filter(const float *angle1, const float *angle2, int n, uint8_t *keep) {
  const int HISTO = 30;
  const float factor = 1.0f / HISTO;
  std::vector<int> hist[HISTO];
  for (int i = 0; i < n; i++) {
    float rot = angle1[i] - angle2[i];
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO) bin = 0;
    hist[bin].push_back(i);
  }
  int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
  for (int i = 0; i < HISTO; i++) {
    const int s = (int)hist[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
  for (int i = 0; i < HISTO; i++) {
    const bool ok = i == ind1 || i == ind2 || i == ind3;
    for (int m : hist[i]) keep[m] = ok;
  }
}

// Frame::ComputeStereoMatches on the pyramids held by two extractors after their extract() calls.
int orb_oracle_stereo(void *hL, void *hR, const void *kL, const uint8_t *dL, int nL, const void *kR,
                      const uint8_t *dR, int nR, float mbf, float mb, float *uRight, float *depth,
                      int *bestDist) {
  return stereo_matches(*(Extractor *)hL, *(Extractor *)hR, (const KeyPointRec *)kL, dL, nL,
                        (const KeyPointRec *)kR, dR, nR, mbf, mb, uRight, depth, bestDist);
}

// Whole stereo frames, one pair per thread: extraction L + R, then stereo matching (bench baseline).
void orb_oracle_stereo_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                             const uint8_t *left, const uint8_t *right, int npairs, int w, int hgt,
                             long img_stride, float mbf, float mb, int cap, int *countsL,
                             int *countsR, int *matched, int nthreads) {
  nthreads = std::max(1, std::min(nthreads, npairs));
  struct Ctx { Extractor *L, *R; std::vector<KeyPointRec> kL, kR; std::vector<uint8_t> dL, dR; std::vector<float> u, z; };
  std::vector<Ctx> ctx(nthreads);
  for (auto &c : ctx) {
    c.L = new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    c.R = new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    c.kL.resize(cap); c.kR.resize(cap); c.dL.resize((size_t)cap * 32); c.dR.resize((size_t)cap * 32);
    c.u.resize(cap); c.z.resize(cap);
  }
  parallel_for(npairs, nthreads, [&](int i, int t) {
    Ctx &c = ctx[t];
    int nl = c.L->extract(left + (size_t)i * img_stride, w, hgt, w, c.kL.data(), c.dL.data(), cap);
    int nr = c.R->extract(right + (size_t)i * img_stride, w, hgt, w, c.kR.data(), c.dR.data(), cap);
    countsL[i] = nl; countsR[i] = nr;
    matched[i] = (nl > 0 && nr > 0)
                     ? stereo_matches(*c.L, *c.R, c.kL.data(), c.dL.data(), nl, c.kR.data(), c.dR.data(), nr,
                                      mbf, mb, c.u.data(), c.z.data(), nullptr)
                     : 0;
  });
  for (auto &c : ctx) { delete c.L; delete c.R; }
}

// Frame::ComputeStereoMatches on caller-provided pyramids: per level an interior pointer, width, height and row stride.
int orb_oracle_stereo_views(int nlevels, const uint8_t *const *planesL, const uint8_t *const *planesR, const int *w, const int *h,
                            const int *strideL, const int *strideR, const float *scale, const float *invScale, const void *kL,
                            const uint8_t *dL, int nL, const void *kR, const uint8_t *dR, int nR, float mbf, float mb,
                            float *uRight, float *depth) {
  PyramidView VL, VR;
  for (int l = 0; l < nlevels; l++) {
    VL.lv.push_back({planesL[l], w[l], h[l], strideL[l]});
    VR.lv.push_back({planesR[l], w[l], h[l], strideR[l]});
  }
  VL.scale.assign(scale, scale + nlevels); VL.invScale.assign(invScale, invScale + nlevels);
  VR.scale = VL.scale; VR.invScale = VL.invScale;
  return stereo_matches_views(VL, VR, (const KeyPointRec *)kL, dL, nL, (const KeyPointRec *)kR, dR, nR, mbf, mb, uRight, depth, nullptr);
}

// cv::undistortPoints(src, dst, K, distCoeffs, R = empty, P = K) as Frame::UndistortKeyPoints (Frame.cc:419-445) and
// Frame::ComputeImageBounds (Frame.cc:447-475) call it.  OpenCV 4.x (calib3d undistort.dispatch.cpp, cvUndistortPointsInternal)
// with the default TermCriteria(MAX_ITER, 5, 0.01): K and the distortion coefficients (k1 k2 p1 p2 [k3], float in the
// reference) widened to double, five fixed-point iterations in double, re-projection with P, ONE rounding to float at the
// end.  The operation order below is the source's; pinned bit for bit against cv2 4.13 in tests/test_undistort.py.
void orb_oracle_undistort_points(const float *xy, int n, float fxf, float fyf, float cxf, float cyf, const float *dist, int ndist,
                                 float *out) {
  double k[14] = {0};
  for (int i = 0; i < ndist && i < 14; i++) k[i] = (double)dist[i];
  const double fx = fxf, fy = fyf, cx = cxf, cy = cyf, ifx = 1. / fx, ify = 1. / fy;
  for (int i = 0; i < n; i++) {
    double x = xy[2 * i], y = xy[2 * i + 1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    // P = K, R = identity: RR = [fx 0 cx; 0 fy cy; 0 0 1]
    const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
    out[2 * i] = (float)(xx * ww);
    out[2 * i + 1] = (float)(yy * ww);
  }
}

// KeyFrameDatabase candidate scoring (KeyFrameDatabase.cc:81-132, 195-241) for every keyframe of a CSR database at once:
// words shared with the query, the smallest shared word, and DBoW2's L1 score (ScoringObject.cpp:23-67: merge of the two
// ascending BowVectors, score += fabs(vi - wi) - fabs(vi) - fabs(wi) on equal ids with v = query, w = keyframe, then
// -score / 2).  Pinned to the reference's own KeyFrameDatabase.cc + DBoW2 in tests/test_kfdb_ref.py.
void orb_oracle_bow_query(int slots, const int *start, const uint32_t *words, const double *weights, const uint8_t *alive,
                          const uint32_t *qWords, const double *qWeights, int nq, int32_t *common, uint32_t *firstWord, double *score) {
  for (int s = 0; s < slots; s++) {
    int n = 0;
    uint32_t first = 0xffffffffu;
    double sc = 0;
    if (alive[s]) {
      int a = 0, b = start[s];
      const int bend = start[s + 1];
      while (a < nq && b < bend) {
        if (qWords[a] == words[b]) {
          const double vi = qWeights[a], wi = weights[b];
          sc += std::fabs(vi - wi) - std::fabs(vi) - std::fabs(wi);
          if (!n) first = words[b];
          n++; a++; b++;
        } else if (qWords[a] < words[b]) {
          a++;
        } else {
          b++;
        }
      }
    }
    common[s] = n; firstWord[s] = first; score[s] = -sc / 2.0;
  }
}

// Host side of orbgpu_selftest_sincos: the same per-chunk checksum with THIS host's libm cosf / sinf (what the reference's
// computeOrbDescriptor calls, ORBextractor.cc:104), over `count` consecutive float32 bit patterns, on `nthreads` threads.
void orb_oracle_sincos_checksum(uint32_t firstBits, uint64_t count, int chunkLog2, uint64_t *sums, int nthreads) {
  const uint64_t chunk = 1ull << chunkLog2, chunks = (count + chunk - 1) >> chunkLog2;
  parallel_for((int)chunks, nthreads, [&](int c, int) {
    uint64_t h = 0;
    const uint64_t lo = (uint64_t)c << chunkLog2, hi = std::min(count, lo + chunk);
    for (uint64_t i = lo; i < hi; i++) {
      const uint32_t bits = firstBits + (uint32_t)i;
      float x;
      std::memcpy(&x, &bits, 4);
      const float sn = sinf(x), cs = cosf(x);
      uint32_t sb, cb;
      std::memcpy(&sb, &sn, 4); std::memcpy(&cb, &cs, 4);
      h += (((uint64_t)sb << 32) | cb) * (0x9E3779B97F4A7C15ull + 2ull * bits);
    }
    sums[c] = h;
  });
}

}  // extern "C"
