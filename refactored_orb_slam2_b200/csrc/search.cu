// search.cu — device-resident feature sets (frame grid) and the candidate loops of the ORBmatcher searches.
//
// Reference (paths relative to Source/Libraries/ORB_SLAM2): Frame::AssignFeaturesToGrid / PosInGrid
// (src/Frame.cc:250-263, 399-411), Frame::GetFeaturesInArea (src/Frame.cc:341-397), KeyFrame::GetFeaturesInArea
// (src/KeyFrame.cc:526-567) and the candidate loops of the eleven searches in src/ORBmatcher.cc (line numbers in
// include/orbgpu.h and oracle/oracle_search.cpp).
//
// Layout: a feature set keeps keypoints (28-byte records), right coordinates, descriptors (2 x uint4 per row) and the
// grid in CSR form with cell (ix, iy) at row ix * rows + iy — the order of mGrid[ix][iy] — so that the window walk's
// inner loop "for iy in [minY, maxY]" over one column ix is ONE contiguous range of cell_items.
//
// k_search: one warp per query.  Lanes stride over the candidates of a column range (or of a vocabulary-node list);
// every candidate carries its position `ord` in the reference's enumeration order, and best / second best are the
// lexicographic minima of (distance, ord) — which is exactly what the reference's strict-'<' update produces.  Keys are
// packed into 64 bits (distance | ord | index | level) and merged across lanes with shuffles.
//
// Greedy searches ("an accepted match takes its keypoint away from later queries"): the sequential result is the
// unique fixed point of  R[q] = search(q | keypoints claimed by accepted queries q' < q in R).  Pass 0 runs with no
// claims; each further pass links the accepted queries of the previous pass into per-keypoint lists (k_link) and
// re-runs the search against them; after pass k the first k queries are final, and the loop stops at the first pass
// that changes nothing (2-4 passes in practice, one host read of a flag per pass).
#include <algorithm>
#include <cstring>
#include <mutex>

#include "internal.cuh"

namespace orbgpu {

struct FeatureView {
  int n;
  const orbgpu_keypoint *kps;
  const float *uRight;
  const uint4 *desc;
  const int *cellStart;
  const int *cellItems;
  float minX, minY, wInv, hInv, qMinX, qMinY;
  int cols, rows;
};

namespace {

constexpr unsigned long long kNoKey = ~0ull;
constexpr int kOrdBits = 24, kIdxBits = 20;

__device__ __forceinline__ unsigned long long pack_key(int dist, unsigned ord, int idx, int level) {
  return ((unsigned long long)dist << 52) | ((unsigned long long)(ord & 0xffffffu) << 28) | ((unsigned long long)idx << 8) |
         (unsigned long long)(level & 0xff);
}
__device__ __forceinline__ int key_dist(unsigned long long k) { return (int)(k >> 52); }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)((k >> 8) & 0xfffffu); }
__device__ __forceinline__ int key_level(unsigned long long k) { return (int)(signed char)(k & 0xff); }

__global__ void k_grid_count(FeatureView F, int *cellOf, int *counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F.n) return;
  // Frame.cc:400-401: round((kp.pt.x - mnMinX) * mfGridElementWidthInv), half away from zero
  const int px = (int)roundf(__fmul_rn(__fsub_rn(F.kps[i].x, F.minX), F.wInv));
  const int py = (int)roundf(__fmul_rn(__fsub_rn(F.kps[i].y, F.minY), F.hInv));
  int cell = -1;
  if (px >= 0 && px < F.cols && py >= 0 && py < F.rows) { cell = px * F.rows + py; atomicAdd(&counts[cell], 1); }
  cellOf[i] = cell;
}

// exclusive scan of the cell counts (<= 64 * 48 + 1 entries), one block
__global__ void k_grid_scan(const int *counts, int *cellStart, int ncells) {
  __shared__ int warpSum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < ncells; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < ncells ? counts[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = threadIdx.x < (blockDim.x >> 5) ? warpSum[threadIdx.x] : 0;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
      warpSum[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = carry + ((threadIdx.x >> 5) ? warpSum[(threadIdx.x >> 5) - 1] : 0) + s - v;
    if (i < ncells) cellStart[i] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) cellStart[ncells] = carry;
}

__global__ void k_grid_fill(int n, const int *cellOf, const int *cellStart, int *cursor, int *items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cellOf[i];
  if (c >= 0) items[cellStart[c] + atomicAdd(&cursor[c], 1)] = i;
}

// AssignFeaturesToGrid pushes keypoints in ascending index: sort each (short) cell list
__global__ void k_grid_sort(const int *cellStart, int *items, int ncells) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int lo = cellStart[c], hi = cellStart[c + 1];
  for (int a = lo + 1; a < hi; a++) {
    const int v = items[a];
    int b = a - 1;
    while (b >= lo && items[b] > v) { items[b + 1] = items[b]; b--; }
    items[b + 1] = v;
  }
}

// cv::undistortPoints(src, dst, K, distCoeffs, R = empty, P = K) as Frame::UndistortKeyPoints (src/Frame.cc:419-445) and
// Frame::ComputeImageBounds (:447-475) call it: OpenCV 4.x cvUndistortPointsInternal with its default TermCriteria
// (MAX_ITER 5): float K and (k1 k2 p1 p2 [k3]) widened to double, five fixed-point iterations in double in the source's
// operation order (round-to-nearest intrinsics: nothing is fused), ONE rounding to float.  One thread per point; `stride`
// floats between consecutive points (2 for bare points, 7 for the pt field of 28-byte cv::KeyPoint records).
struct CameraDev { double fx, fy, cx, cy, ifx, ify, k[5]; };
__global__ void k_undistort(const float *src, float *dst, int n, int stride, const CameraDev C) {   // src may alias dst
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double u = (double)src[(size_t)i * stride], v = (double)src[(size_t)i * stride + 1];
  double x = __dmul_rn(__dsub_rn(u, C.cx), C.ifx), y = __dmul_rn(__dsub_rn(v, C.cy), C.ify);
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int j = 0; j < 5; j++) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    // (1 + ((k7 r2 + k6) r2 + k5) r2) with k5 = k6 = k7 = 0 is exactly 1
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(C.k[4], r2), C.k[1]), r2), C.k[0]), r2));
    const double icdist = __ddiv_rn(1.0, den);
    if (icdist < 0) { x = x0; y = y0; break; }            // (u - cx) ifx again: the value x started from
    // deltaX = 2 k2 x y + k3 (r2 + 2 x x) [+ k8 r2 + k9 r2 r2 = +0], deltaY = k2 (r2 + 2 y y) + 2 k3 x y [+ 0]
    const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, C.k[2]), x), y), __dmul_rn(C.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    const double dY = __dadd_rn(__dmul_rn(C.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(__dmul_rn(2.0, C.k[3]), x), y));
    x = __dmul_rn(__dsub_rn(x0, dX), icdist);
    y = __dmul_rn(__dsub_rn(y0, dY), icdist);
  }
  // P = K, R = identity: xx = fx x + 0 y + cx, ww = 1 / (0 x + 0 y + 1) = 1
  dst[(size_t)i * stride] = (float)__dadd_rn(__dmul_rn(C.fx, x), C.cx);
  dst[(size_t)i * stride + 1] = (float)__dadd_rn(__dmul_rn(C.fy, y), C.cy);
}
CameraDev to_device_camera(const orbgpu_camera &c) {
  CameraDev C;
  C.fx = c.fx; C.fy = c.fy; C.cx = c.cx; C.cy = c.cy; C.ifx = 1. / C.fx; C.ify = 1. / C.fy;
  for (int i = 0; i < 5; i++) C.k[i] = i < c.ndist ? (double)c.dist[i] : 0.0;
  return C;
}

__global__ void k_fill_float(float *p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct Window { int minCX, maxCX, minCY, maxCY; bool empty; };

// Frame.cc:348-368 / KeyFrame.cc:531-549, float expressions in the reference's order, never fused
__device__ __forceinline__ Window window_of(const FeatureView &F, float x, float y, float r) {
  Window w;
  w.empty = true;
  w.minCX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.qMinX), r), F.wInv)));
  if (w.minCX >= F.cols) return w;
  w.maxCX = min(F.cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.qMinX), r), F.wInv)));
  if (w.maxCX < 0) return w;
  w.minCY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.qMinY), r), F.hInv)));
  if (w.minCY >= F.rows) return w;
  w.maxCY = min(F.rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.qMinY), r), F.hInv)));
  if (w.maxCY < 0) return w;
  w.empty = false;
  return w;
}

__device__ __forceinline__ bool level_ok(int octave, int minLevel, int maxLevel) {   // Frame.cc:369, 377-383
  if (octave < minLevel) return false;
  if (maxLevel >= 0 && octave > maxLevel) return false;
  return true;
}

struct PolicyDev {
  int source, ratioMode, greedy, gate, acceptMax, stereoGate, onlyStereo;
  float ratio, ex, ey;
  float f12[9];
  float levelSigma2[16], invLevelSigma2[16], scaleFactors[16];
};

__device__ __forceinline__ int hamming256(const uint4 &a0, const uint4 &a1, const uint4 *row) {
  const uint4 b0 = row[0], b1 = row[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Gates + distance for one candidate; returns the packed key or kNoKey.
__device__ __forceinline__ unsigned long long evaluate(const FeatureView &F, const PolicyDev &P, const orbgpu_query &q, int qi, const uint4 &a0,
                                                       const uint4 &a1, int idx, unsigned ord, bool windowTest, const uint8_t *occupied,
                                                       const orbgpu_match *prev, const int *head, const int *next) {
  if (occupied && occupied[idx]) return kNoKey;
  if (P.greedy == ORBGPU_GREEDY_TAKEN && prev) {
    for (int t = head[idx]; t >= 0; t = next[t]) if (t < qi) return kNoKey;      // claimed by an earlier accepted query
  }
  const orbgpu_keypoint kp = F.kps[idx];
  if (!level_ok(kp.octave, q.min_level, q.max_level)) return kNoKey;
  if (windowTest) {                                                                 // Frame.cc:386-390
    const float distx = __fsub_rn(kp.x, q.u), disty = __fsub_rn(kp.y, q.v);
    if (!(fabsf(distx) < q.r && fabsf(disty) < q.r)) return kNoKey;
  }
  const float uR = F.uRight[idx];
  if (P.stereoGate && uR > 0) {                                                     // ORBmatcher.cc:88-92, 1306-1311
    if (fabsf(__fsub_rn(q.ur, uR)) > q.ur_tol) return kNoKey;
  }
  if (P.gate == ORBGPU_GATE_FUSE) {                                                 // ORBmatcher.cc:843-866
    const float ex = __fsub_rn(q.u, kp.x), ey = __fsub_rn(q.v, kp.y);
    float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    double limit = 5.99;
    if (uR >= 0) { const float er = __fsub_rn(q.ur, uR); e2 = __fadd_rn(e2, __fmul_rn(er, er)); limit = 7.8; }
    if ((double)__fmul_rn(e2, P.invLevelSigma2[kp.octave]) > limit) return kNoKey;
  }
  const bool stereo2 = uR >= 0;
  if (P.gate == ORBGPU_GATE_EPIPOLAR && P.onlyStereo && !stereo2) return kNoKey;     // ORBmatcher.cc:669-673
  const int dist = hamming256(a0, a1, F.desc + 2 * (size_t)idx);
  if (P.greedy == ORBGPU_GREEDY_DISTANCE && prev) {                                 // ORBmatcher.cc:447
    for (int t = head[idx]; t >= 0; t = next[t]) if (t < qi && prev[t].dist <= dist) return kNoKey;
  }
  if (P.gate == ORBGPU_GATE_EPIPOLAR) {
    if (dist > P.acceptMax) return kNoKey;                                          // ORBmatcher.cc:678
    if (!(q.flags & ORBGPU_Q_STEREO) && !stereo2) {                                 // :683-689
      const float dx = __fsub_rn(P.ex, kp.x), dy = __fsub_rn(P.ey, kp.y);
      if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, P.scaleFactors[kp.octave])) return kNoKey;
    }
    // CheckDistEpipolarLine, ORBmatcher.cc:137-159
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(q.u, P.f12[0]), __fmul_rn(q.v, P.f12[3])), P.f12[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(q.u, P.f12[1]), __fmul_rn(q.v, P.f12[4])), P.f12[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(q.u, P.f12[2]), __fmul_rn(q.v, P.f12[5])), P.f12[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp.x), __fmul_rn(b, kp.y)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0) return kNoKey;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    if (!((double)dsqr < 3.84 * (double)P.levelSigma2[kp.octave])) return kNoKey;
    ord = 0xffffffu - ord;                                                          // '<=' update: the LAST minimum wins (:678)
  }
  return pack_key(dist, ord, idx, kp.octave);
}

__device__ __forceinline__ void keep_two(unsigned long long k, unsigned long long &k1, unsigned long long &k2) {
  if (k < k1) { k2 = k1; k1 = k; }
  else if (k < k2) k2 = k;
}

__global__ void __launch_bounds__(128) k_search(FeatureView F, PolicyDev P, const orbgpu_query *__restrict__ queries, const uint4 *__restrict__ qdesc,
                                                 int nq, const uint8_t *__restrict__ occupied, const int *__restrict__ cand,
                                                 const orbgpu_match *__restrict__ prev, const int *__restrict__ head, const int *__restrict__ next,
                                                 orbgpu_match *__restrict__ out, int *changed) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  const orbgpu_query q = queries[qi];
  const uint4 a0 = qdesc[2 * (size_t)qi], a1 = qdesc[2 * (size_t)qi + 1];
  unsigned long long k1 = kNoKey, k2 = kNoKey;
  if (P.source == ORBGPU_SRC_WINDOW) {
    const Window w = window_of(F, q.u, q.v, q.r);
    if (!w.empty) {
      unsigned base = 0;
      for (int ix = w.minCX; ix <= w.maxCX; ix++) {
        const int lo = F.cellStart[ix * F.rows + w.minCY], hi = F.cellStart[ix * F.rows + w.maxCY + 1];
        for (int p = lo + lane; p < hi; p += 32)
          keep_two(evaluate(F, P, q, qi, a0, a1, F.cellItems[p], base + (unsigned)(p - lo), true, occupied, prev, head, next), k1, k2);
        base += (unsigned)(hi - lo);
      }
    }
  } else {
    for (int p = q.list_begin + lane; p < q.list_end; p += 32)
      keep_two(evaluate(F, P, q, qi, a0, a1, cand[p], (unsigned)(p - q.list_begin), false, occupied, prev, head, next), k1, k2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long e1 = __shfl_xor_sync(0xffffffffu, k1, o), e2 = __shfl_xor_sync(0xffffffffu, k2, o);
    if (e1 < k1) { k2 = min(k1, e2); k1 = e1; }
    else k2 = min(k2, e1);
  }
  if (lane != 0) return;
  orbgpu_match m;
  m.idx = -1; m.dist = 256; m.dist2 = 256; m.level = -1; m.level2 = -1; m.accepted = 0; m.pad = 0;
  if (k1 != kNoKey) { m.idx = key_idx(k1); m.dist = key_dist(k1); m.level = (int8_t)key_level(k1); }
  if (k2 != kNoKey && P.gate != ORBGPU_GATE_EPIPOLAR) { m.dist2 = key_dist(k2); m.level2 = (int8_t)key_level(k2); }
  bool ok = m.idx >= 0 && m.dist <= P.acceptMax;
  // ORBmatcher.cc:116-118: int > float * int; :225: (float) < float * (float)
  if (ok && P.ratioMode == ORBGPU_RATIO_SAME_LEVEL && m.level == m.level2 && (float)m.dist > __fmul_rn(P.ratio, (float)m.dist2)) ok = false;
  if (ok && P.ratioMode == ORBGPU_RATIO_LESS && !((float)m.dist < __fmul_rn(P.ratio, (float)m.dist2))) ok = false;
  m.accepted = ok ? 1 : 0;
  if (prev) {
    const orbgpu_match o = prev[qi];
    if (o.idx != m.idx || o.dist != m.dist || o.dist2 != m.dist2 || o.level != m.level || o.level2 != m.level2 || o.accepted != m.accepted) *changed = 1;
  }
  out[qi] = m;
}

// accepted queries of one pass, linked per claimed keypoint
__global__ void k_link(const orbgpu_match *res, const orbgpu_query *queries, int nq, int greedy, int *head, int *next) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const orbgpu_match m = res[qi];
  if (!m.accepted) return;
  if (greedy == ORBGPU_GREEDY_TAKEN && !(queries[qi].flags & ORBGPU_Q_BLOCKS)) return;
  next[qi] = atomicExch(&head[m.idx], qi);
}

// GetFeaturesInArea: count pass (items == nullptr) and fill pass; one warp per query, output in the reference's order
__global__ void k_in_area(FeatureView F, const orbgpu_query *queries, int nq, const int *offsets, int *counts, int *items) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  const orbgpu_query q = queries[qi];
  const Window w = window_of(F, q.u, q.v, q.r);
  int total = 0;
  if (!w.empty) {
    for (int ix = w.minCX; ix <= w.maxCX; ix++) {
      const int lo = F.cellStart[ix * F.rows + w.minCY], hi = F.cellStart[ix * F.rows + w.maxCY + 1];
      for (int p0 = lo; p0 < hi; p0 += 32) {
        const int p = p0 + lane;
        bool keep = false;
        int idx = -1;
        if (p < hi) {
          idx = F.cellItems[p];
          const orbgpu_keypoint kp = F.kps[idx];
          keep = level_ok(kp.octave, q.min_level, q.max_level) && fabsf(__fsub_rn(kp.x, q.u)) < q.r && fabsf(__fsub_rn(kp.y, q.v)) < q.r;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep && items) items[offsets[qi] + total + __popc(mask & ((1u << lane) - 1))] = idx;
        total += __popc(mask);
      }
    }
  }
  if (lane == 0 && counts) counts[qi] = total;
}

// per host thread: staging + device scratch + stream (feature sets themselves are read-only during a search)
struct Workspace {
  int device = -1;
  cudaStream_t stream = nullptr;
  uint8_t *hStage = nullptr; size_t hCap = 0;       // pinned
  uint8_t *dStage = nullptr; size_t dCap = 0;
  orbgpu_match *dRes[2] = {nullptr, nullptr}; int *dNext = nullptr; size_t qCap = 0;
  int *dHead = nullptr; size_t headCap = 0;
  int *dChanged = nullptr; int *hChanged = nullptr;
  orbgpu_match *hRes = nullptr; size_t hResCap = 0;
  // never freed: the owning thread may end after the CUDA runtime has shut down; the context releases everything
};
thread_local Workspace t_ws;

cudaError_t ensure(Workspace &W, int device, size_t stageBytes, size_t nq, size_t nfeat) {
  cudaError_t e;
  if (W.device != device) {
    if (W.device >= 0) {      // the calling thread moved to another device: drop the old scratch
      cudaFree(W.dStage); cudaFree(W.dRes[0]); cudaFree(W.dRes[1]); cudaFree(W.dNext); cudaFree(W.dHead); cudaFree(W.dChanged);
      if (W.stream) cudaStreamDestroy(W.stream);
      W.dStage = nullptr; W.dRes[0] = W.dRes[1] = nullptr; W.dNext = W.dHead = W.dChanged = nullptr; W.stream = nullptr;
      W.dCap = W.qCap = W.headCap = 0;
    }
    if ((e = cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&W.dChanged, sizeof(int))) != cudaSuccess) return e;
    if (!W.hChanged && (e = cudaMallocHost(&W.hChanged, sizeof(int))) != cudaSuccess) return e;
    W.device = device;
  }
  if (stageBytes > W.hCap) {
    if (W.hStage) cudaFreeHost(W.hStage);
    W.hCap = std::max(stageBytes * 2, (size_t)1 << 18);
    if ((e = cudaMallocHost(&W.hStage, W.hCap)) != cudaSuccess) { W.hStage = nullptr; W.hCap = 0; return e; }
  }
  if (stageBytes > W.dCap) {
    cudaFree(W.dStage);
    W.dCap = std::max(stageBytes * 2, (size_t)1 << 18);
    if ((e = cudaMalloc(&W.dStage, W.dCap)) != cudaSuccess) { W.dStage = nullptr; W.dCap = 0; return e; }
  }
  if (nq > W.qCap) {
    cudaFree(W.dRes[0]); cudaFree(W.dRes[1]); cudaFree(W.dNext);
    W.qCap = std::max(nq * 2, (size_t)4096);
    if ((e = cudaMalloc(&W.dRes[0], W.qCap * sizeof(orbgpu_match))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&W.dRes[1], W.qCap * sizeof(orbgpu_match))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&W.dNext, W.qCap * sizeof(int))) != cudaSuccess) return e;
  }
  if (nq > W.hResCap) {
    if (W.hRes) cudaFreeHost(W.hRes);
    W.hResCap = std::max(nq * 2, (size_t)4096);
    if ((e = cudaMallocHost(&W.hRes, W.hResCap * sizeof(orbgpu_match))) != cudaSuccess) { W.hRes = nullptr; W.hResCap = 0; return e; }
  }
  if (nfeat > W.headCap) {
    cudaFree(W.dHead);
    W.headCap = std::max(nfeat * 2, (size_t)8192);
    if ((e = cudaMalloc(&W.dHead, W.headCap * sizeof(int))) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

}  // namespace orbgpu

using namespace orbgpu;

struct orbgpu_features {
  int device = 0;
  int n = 0;
  orbgpu_features_desc geom{};
  orbgpu_keypoint *dKps = nullptr;
  float *dURight = nullptr;
  uint8_t *dDesc = nullptr;
  int *dCellStart = nullptr, *dCellItems = nullptr;
  int ncells = 0, nItems = 0;
  FeatureView view() const {
    FeatureView V;
    V.n = n; V.kps = dKps; V.uRight = dURight; V.desc = reinterpret_cast<const uint4 *>(dDesc); V.cellStart = dCellStart; V.cellItems = dCellItems;
    V.minX = geom.min_x; V.minY = geom.min_y; V.wInv = geom.grid_w_inv; V.hInv = geom.grid_h_inv; V.qMinX = geom.query_min_x; V.qMinY = geom.query_min_y;
    V.cols = geom.grid_cols; V.rows = geom.grid_rows;
    return V;
  }
};

namespace {

struct DevGuard {
  int prev = -1;
  bool ok = true;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_geometry(const orbgpu_features_desc *d) {
  if (!d || d->n < 0 || d->grid_cols < 1 || d->grid_rows < 1 || (long long)d->grid_cols * d->grid_rows > (1 << 20)) { set_error("bad feature-set geometry"); return ORBGPU_E_INVALID; }
  if (d->n >= (1 << kIdxBits)) { set_error("feature sets are limited to %d keypoints", (1 << kIdxBits) - 1); return ORBGPU_E_CAPACITY; }
  return ORBGPU_OK;
}

int allocate(orbgpu_features *f) {
  const size_t n = std::max(f->n, 1);
  ORBGPU_CUDA_CHECK(cudaMalloc(&f->dKps, n * sizeof(orbgpu_keypoint)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&f->dURight, n * sizeof(float)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&f->dDesc, n * 32));
  ORBGPU_CUDA_CHECK(cudaMalloc(&f->dCellStart, (size_t)(f->ncells + 1) * sizeof(int)));
  ORBGPU_CUDA_CHECK(cudaMalloc(&f->dCellItems, n * sizeof(int)));
  return ORBGPU_OK;
}

// grid build on `s`; scratch: cellOf [n] + counts [ncells] + cursor [ncells]
int build_grid(orbgpu_features *f, cudaStream_t s) {
  const int n = f->n, nc = f->ncells;
  int *scratch = nullptr;
  ORBGPU_CUDA_CHECK(cudaMalloc(&scratch, ((size_t)std::max(n, 1) + 2 * (size_t)nc) * sizeof(int)));
  int *cellOf = scratch, *counts = scratch + std::max(n, 1), *cursor = counts + nc;
  cudaError_t e = cudaMemsetAsync(counts, 0, 2 * (size_t)nc * sizeof(int), s);
  if (e == cudaSuccess && n > 0) { k_grid_count<<<(n + 255) / 256, 256, 0, s>>>(f->view(), cellOf, counts); count_launch(); e = cudaGetLastError(); }
  if (e == cudaSuccess) { k_grid_scan<<<1, 1024, 0, s>>>(counts, f->dCellStart, nc); count_launch(); e = cudaGetLastError(); }
  if (e == cudaSuccess && n > 0) {
    k_grid_fill<<<(n + 255) / 256, 256, 0, s>>>(n, cellOf, f->dCellStart, cursor, f->dCellItems);
    k_grid_sort<<<(nc + 255) / 256, 256, 0, s>>>(f->dCellStart, f->dCellItems, nc);
    count_launch(2);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&f->nItems, f->dCellStart + nc, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(scratch);
  if (e != cudaSuccess) { set_error("grid build failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  return ORBGPU_OK;
}

PolicyDev to_device_policy(const orbgpu_search_policy &p) {
  PolicyDev P;
  P.source = p.source; P.ratioMode = p.ratio_mode; P.greedy = p.greedy; P.gate = p.gate; P.acceptMax = p.accept_max; P.stereoGate = p.stereo_gate;
  P.onlyStereo = p.only_stereo; P.ratio = p.ratio; P.ex = p.ex; P.ey = p.ey;
  std::memcpy(P.f12, p.f12, sizeof(P.f12));
  std::memcpy(P.levelSigma2, p.level_sigma2, sizeof(P.levelSigma2));
  std::memcpy(P.invLevelSigma2, p.inv_level_sigma2, sizeof(P.invLevelSigma2));
  std::memcpy(P.scaleFactors, p.scale_factors, sizeof(P.scaleFactors));
  return P;
}

}  // namespace

// defined in capi.cu: device pointers of image `image` of the extractor's last batch (+ stereo results of `pair`)
extern "C" int orbgpu_internal_extractor_view(orbgpu_extractor *e, int image, int pair, const orbgpu_keypoint **kps, const uint8_t **desc,
                                              int *count, const float **uRight, int *device, void **stream);

extern "C" {

int orbgpu_features_create(const orbgpu_features_desc *desc, orbgpu_features **out) {
  if (!out) { set_error("null output"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  int rc = check_geometry(desc);
  if (rc != ORBGPU_OK) return rc;
  if (desc->n > 0 && (!desc->keys_un || !desc->descriptors)) { set_error("null keypoints / descriptors"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DevGuard guard(desc->device);
  if (!guard.ok) { set_error("cannot select device %d", desc->device); return ORBGPU_E_CUDA; }
  orbgpu_features *f = new orbgpu_features;
  cudaGetDevice(&f->device);
  f->n = desc->n; f->geom = *desc; f->geom.keys_un = nullptr; f->geom.u_right = nullptr; f->geom.descriptors = nullptr;
  f->ncells = desc->grid_cols * desc->grid_rows;
  rc = allocate(f);
  if (rc == ORBGPU_OK && (rc = (ensure(t_ws, f->device, 0, 0, 0) == cudaSuccess ? ORBGPU_OK : ORBGPU_E_CUDA)) != ORBGPU_OK) set_error("workspace allocation failed");
  if (rc == ORBGPU_OK && f->n > 0) {
    cudaStream_t s = t_ws.stream;
    cudaError_t e = cudaMemcpyAsync(f->dKps, desc->keys_un, (size_t)f->n * sizeof(orbgpu_keypoint), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->dDesc, desc->descriptors, (size_t)f->n * 32, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
      if (desc->u_right) e = cudaMemcpyAsync(f->dURight, desc->u_right, (size_t)f->n * sizeof(float), cudaMemcpyHostToDevice, s);
      else { k_fill_float<<<(f->n + 255) / 256, 256, 0, s>>>(f->dURight, f->n, -1.f); count_launch(); e = cudaGetLastError(); }   // monocular: mvuRight = -1
    }
    if (e != cudaSuccess) { set_error("feature upload failed: %s", cudaGetErrorString(e)); rc = ORBGPU_E_CUDA; }
  }
  if (rc == ORBGPU_OK) rc = build_grid(f, t_ws.stream);
  if (rc != ORBGPU_OK) { orbgpu_features_destroy(f); return rc; }
  *out = f;
  return ORBGPU_OK;
}

static int features_from_extractor(orbgpu_extractor *e, int image, int pair, const orbgpu_features_desc *geometry, const orbgpu_camera *cam,
                                   orbgpu_features **out) {
  if (!out) { set_error("null output"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  if (!e || !geometry) { set_error("null argument"); return ORBGPU_E_INVALID; }
  const orbgpu_keypoint *kps; const uint8_t *desc; const float *uRight; int count, device; void *stream;
  int rc = orbgpu_internal_extractor_view(e, image, pair, &kps, &desc, &count, &uRight, &device, &stream);
  if (rc != ORBGPU_OK) return rc;
  orbgpu_features_desc d = *geometry;
  d.n = count; d.device = device;
  if ((rc = check_geometry(&d)) != ORBGPU_OK) return rc;
  DevGuard guard(device);
  orbgpu_features *f = new orbgpu_features;
  f->device = device; f->n = count; f->geom = d; f->geom.keys_un = nullptr; f->geom.u_right = nullptr; f->geom.descriptors = nullptr;
  f->ncells = d.grid_cols * d.grid_rows;
  rc = allocate(f);
  if (rc == ORBGPU_OK && count > 0) {
    cudaStream_t s = (cudaStream_t)stream;     // the extractor's own stream: ordered after the extraction / stereo kernels
    cudaError_t ce = cudaMemcpyAsync(f->dKps, kps, (size_t)count * sizeof(orbgpu_keypoint), cudaMemcpyDeviceToDevice, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(f->dDesc, desc, (size_t)count * 32, cudaMemcpyDeviceToDevice, s);
    if (ce == cudaSuccess) {
      if (uRight) ce = cudaMemcpyAsync(f->dURight, uRight, (size_t)count * sizeof(float), cudaMemcpyDeviceToDevice, s);
      else { k_fill_float<<<(count + 255) / 256, 256, 0, s>>>(f->dURight, count, -1.f); count_launch(); ce = cudaGetLastError(); }
    }
    // distorted input: mvKeysUn = cv::undistortPoints(mvKeys) in place, before the grid is built from it (Frame.cc:122, 419-445)
    if (ce == cudaSuccess && cam && cam->dist[0] != 0.0f) {
      k_undistort<<<(count + 127) / 128, 128, 0, s>>>(reinterpret_cast<const float *>(f->dKps), reinterpret_cast<float *>(f->dKps), count,
                                                      (int)(sizeof(orbgpu_keypoint) / sizeof(float)), to_device_camera(*cam));
      count_launch();
      ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) { set_error("feature copy failed: %s", cudaGetErrorString(ce)); rc = ORBGPU_E_CUDA; }
  }
  if (rc == ORBGPU_OK) rc = build_grid(f, (cudaStream_t)stream);
  if (rc != ORBGPU_OK) { orbgpu_features_destroy(f); return rc; }
  *out = f;
  return ORBGPU_OK;
}

int orbgpu_features_from_extractor(orbgpu_extractor *e, int image, int pair, const orbgpu_features_desc *geometry, orbgpu_features **out) {
  return features_from_extractor(e, image, pair, geometry, nullptr, out);
}
int orbgpu_features_from_extractor_undistorted(orbgpu_extractor *e, int image, int pair, const orbgpu_features_desc *geometry,
                                               const orbgpu_camera *camera, orbgpu_features **out) {
  if (!camera) { set_error("null camera"); return ORBGPU_E_INVALID; }
  if (camera->ndist < 4 || camera->ndist > 5 || !(camera->fx != 0.f) || !(camera->fy != 0.f)) { set_error("camera: need fx, fy != 0 and 4 or 5 distortion coefficients"); return ORBGPU_E_INVALID; }
  return features_from_extractor(e, image, pair, geometry, camera, out);
}

static int undistort_host(const float *src, float *dst, int n, int stride, const orbgpu_camera *camera, int device) {
  if (n < 0 || !camera || (n > 0 && (!src || !dst))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (camera->ndist < 4 || camera->ndist > 5 || !(camera->fx != 0.f) || !(camera->fy != 0.f)) { set_error("camera: need fx, fy != 0 and 4 or 5 distortion coefficients"); return ORBGPU_E_INVALID; }
  if (n == 0) return ORBGPU_OK;
  if (camera->dist[0] == 0.0f) {               // Frame.cc:420-423: mvKeysUn = mvKeys
    if (dst != src) std::memcpy(dst, src, (size_t)n * stride * sizeof(float));
    return ORBGPU_OK;
  }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DevGuard guard(device);
  if (!guard.ok) { set_error("cannot select device %d", device); return ORBGPU_E_CUDA; }
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t bytes = (size_t)n * stride * sizeof(float);
  Workspace &W = t_ws;
  cudaError_t e = ensure(W, dev, bytes, 0, 0);
  if (e != cudaSuccess) { set_error("workspace allocation failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  std::memcpy(W.hStage, src, bytes);
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.dStage, W.hStage, bytes, cudaMemcpyHostToDevice, W.stream));
  k_undistort<<<(n + 127) / 128, 128, 0, W.stream>>>(reinterpret_cast<const float *>(W.dStage), reinterpret_cast<float *>(W.dStage), n, stride, to_device_camera(*camera));
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.hStage, W.dStage, bytes, cudaMemcpyDeviceToHost, W.stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(W.stream));
  std::memcpy(dst, W.hStage, bytes);
  return ORBGPU_OK;
}
int orbgpu_undistort_points(const float *xy, int n, const orbgpu_camera *camera, float *xy_undistorted, int device) {
  return undistort_host(xy, xy_undistorted, n, 2, camera, device);
}
int orbgpu_undistort_keypoints(const orbgpu_keypoint *keys, int n, const orbgpu_camera *camera, orbgpu_keypoint *keys_un, int device) {
  return undistort_host(reinterpret_cast<const float *>(keys), reinterpret_cast<float *>(keys_un), n, (int)(sizeof(orbgpu_keypoint) / sizeof(float)), camera, device);
}

void orbgpu_features_destroy(orbgpu_features *f) {
  if (!f) return;
  DevGuard guard(f->device);
  cudaFree(f->dKps); cudaFree(f->dURight); cudaFree(f->dDesc); cudaFree(f->dCellStart); cudaFree(f->dCellItems);
  delete f;
}

int orbgpu_features_count(const orbgpu_features *f) { return f ? f->n : 0; }

int orbgpu_features_grid(orbgpu_features *f, int32_t *cell_start, int32_t *cell_items, int32_t *n_items) {
  if (!f || !cell_start || !cell_items) { set_error("null argument"); return ORBGPU_E_INVALID; }
  DevGuard guard(f->device);
  ORBGPU_CUDA_CHECK(cudaMemcpy(cell_start, f->dCellStart, (size_t)(f->ncells + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (f->nItems > 0) ORBGPU_CUDA_CHECK(cudaMemcpy(cell_items, f->dCellItems, (size_t)f->nItems * sizeof(int), cudaMemcpyDeviceToHost));
  if (n_items) *n_items = f->nItems;
  return ORBGPU_OK;
}

int orbgpu_features_download(orbgpu_features *f, orbgpu_keypoint *keys_un, float *u_right, uint8_t *descriptors) {
  if (!f) { set_error("null feature set"); return ORBGPU_E_INVALID; }
  DevGuard guard(f->device);
  if (f->n == 0) return ORBGPU_OK;
  if (keys_un) ORBGPU_CUDA_CHECK(cudaMemcpy(keys_un, f->dKps, (size_t)f->n * sizeof(orbgpu_keypoint), cudaMemcpyDeviceToHost));
  if (u_right) ORBGPU_CUDA_CHECK(cudaMemcpy(u_right, f->dURight, (size_t)f->n * sizeof(float), cudaMemcpyDeviceToHost));
  if (descriptors) ORBGPU_CUDA_CHECK(cudaMemcpy(descriptors, f->dDesc, (size_t)f->n * 32, cudaMemcpyDeviceToHost));
  return ORBGPU_OK;
}

int orbgpu_search(orbgpu_features *f, const orbgpu_query *queries, const uint8_t *query_descriptors, int nq, const uint8_t *occupied,
                  const int32_t *cand, int ncand, const orbgpu_search_policy *policy, orbgpu_match *result, int *iterations) {
  if (iterations) *iterations = 0;
  if (!f || !policy || nq < 0 || ncand < 0 || (nq > 0 && (!queries || !query_descriptors || !result))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (policy->source == ORBGPU_SRC_LIST && ncand > 0 && !cand) { set_error("null candidate list"); return ORBGPU_E_INVALID; }
  if (nq == 0) return ORBGPU_OK;
  if (policy->source == ORBGPU_SRC_LIST) {
    for (int i = 0; i < nq; i++)
      if (queries[i].list_begin < 0 || queries[i].list_end > ncand || queries[i].list_begin > queries[i].list_end) { set_error("query %d: bad list range", i); return ORBGPU_E_INVALID; }
    for (int i = 0; i < ncand; i++)
      if (cand[i] < 0 || cand[i] >= f->n) { set_error("candidate %d out of range", cand[i]); return ORBGPU_E_INVALID; }
  }
  if (f->n == 0) {
    for (int i = 0; i < nq; i++) { result[i].idx = -1; result[i].dist = result[i].dist2 = 256; result[i].level = result[i].level2 = -1; result[i].accepted = 0; result[i].pad = 0; }
    return ORBGPU_OK;
  }
  DevGuard guard(f->device);
  if (!guard.ok) { set_error("cannot select device %d", f->device); return ORBGPU_E_CUDA; }
  // one staging buffer: queries | descriptors | occupied | candidates
  const size_t oQ = 0, oD = align16(oQ + (size_t)nq * sizeof(orbgpu_query)), oO = align16(oD + (size_t)nq * 32),
               oC = align16(oO + (occupied ? (size_t)f->n : 0)), total = align16(oC + (size_t)ncand * sizeof(int));
  Workspace &W = t_ws;
  cudaError_t e = ensure(W, f->device, total, (size_t)nq, (size_t)f->n);
  if (e != cudaSuccess) { set_error("workspace allocation failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  std::memcpy(W.hStage + oQ, queries, (size_t)nq * sizeof(orbgpu_query));
  std::memcpy(W.hStage + oD, query_descriptors, (size_t)nq * 32);
  if (occupied) std::memcpy(W.hStage + oO, occupied, (size_t)f->n);
  if (ncand) std::memcpy(W.hStage + oC, cand, (size_t)ncand * sizeof(int));
  cudaStream_t s = W.stream;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.dStage, W.hStage, total, cudaMemcpyHostToDevice, s));
  const orbgpu_query *dQ = reinterpret_cast<const orbgpu_query *>(W.dStage + oQ);
  const uint4 *dD = reinterpret_cast<const uint4 *>(W.dStage + oD);
  const uint8_t *dO = occupied ? W.dStage + oO : nullptr;
  const int *dC = reinterpret_cast<const int *>(W.dStage + oC);
  const FeatureView V = f->view();
  const PolicyDev P = to_device_policy(*policy);
  const int blocks = (nq + 3) / 4;
  int cur = 0, passes = 1;
  k_search<<<blocks, 128, 0, s>>>(V, P, dQ, dD, nq, dO, dC, nullptr, nullptr, nullptr, W.dRes[0], W.dChanged);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  if (policy->greedy != ORBGPU_GREEDY_NONE) {
    for (;;) {
      ORBGPU_CUDA_CHECK(cudaMemsetAsync(W.dHead, 0xff, (size_t)f->n * sizeof(int), s));
      ORBGPU_CUDA_CHECK(cudaMemsetAsync(W.dChanged, 0, sizeof(int), s));
      k_link<<<(nq + 255) / 256, 256, 0, s>>>(W.dRes[cur], dQ, nq, policy->greedy, W.dHead, W.dNext);
      k_search<<<blocks, 128, 0, s>>>(V, P, dQ, dD, nq, dO, dC, W.dRes[cur], W.dHead, W.dNext, W.dRes[cur ^ 1], W.dChanged);
      count_launch(2);
      ORBGPU_CUDA_CHECK(cudaGetLastError());
      ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.hChanged, W.dChanged, sizeof(int), cudaMemcpyDeviceToHost, s));
      ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
      cur ^= 1;
      passes++;
      if (!*W.hChanged) break;
      if (passes > nq + 2) { set_error("greedy search did not converge"); return ORBGPU_E_CUDA; }
    }
  }
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.hRes, W.dRes[cur], (size_t)nq * sizeof(orbgpu_match), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
  std::memcpy(result, W.hRes, (size_t)nq * sizeof(orbgpu_match));
  if (iterations) *iterations = passes;
  return ORBGPU_OK;
}

int orbgpu_features_in_area(orbgpu_features *f, const orbgpu_query *queries, int nq, int32_t *offsets, int32_t *items, int cap) {
  if (!f || nq < 0 || !offsets || (nq > 0 && !queries) || cap < 0 || (cap > 0 && !items)) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  offsets[0] = 0;
  if (nq == 0) return ORBGPU_OK;
  if (f->n == 0) { for (int i = 0; i <= nq; i++) offsets[i] = 0; return ORBGPU_OK; }
  DevGuard guard(f->device);
  orbgpu_query *dQ = nullptr; int *dCounts = nullptr, *dOff = nullptr, *dItems = nullptr;
  auto cleanup = [&] { cudaFree(dQ); cudaFree(dCounts); cudaFree(dOff); cudaFree(dItems); };
  cudaError_t e = cudaMalloc(&dQ, (size_t)nq * sizeof(orbgpu_query));
  if (e == cudaSuccess) e = cudaMalloc(&dCounts, (size_t)nq * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dOff, (size_t)(nq + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(dQ, queries, (size_t)nq * sizeof(orbgpu_query), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { k_in_area<<<(nq + 3) / 4, 128>>>(f->view(), dQ, nq, nullptr, dCounts, nullptr); count_launch(); e = cudaGetLastError(); }
  std::vector<int> counts(nq);
  if (e == cudaSuccess) e = cudaMemcpy(counts.data(), dCounts, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error("features_in_area: %s", cudaGetErrorString(e)); cleanup(); return ORBGPU_E_CUDA; }
  for (int i = 0; i < nq; i++) offsets[i + 1] = offsets[i] + counts[i];
  const int total = offsets[nq];
  if (total > cap) { cleanup(); set_error("features_in_area: %d items, capacity %d", total, cap); return ORBGPU_E_CAPACITY; }
  if (total > 0) {
    e = cudaMalloc(&dItems, (size_t)total * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(dOff, offsets, (size_t)(nq + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { k_in_area<<<(nq + 3) / 4, 128>>>(f->view(), dQ, nq, dOff, nullptr, dItems); count_launch(); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(items, dItems, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (e != cudaSuccess) { set_error("features_in_area: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  return ORBGPU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// DBoW2 vocabulary tree descent (TemplatedVocabulary.h:1216-1258, FORB.cpp:77-100).  Half a warp per feature: at every
// level the 16 lanes take the children in strides of 16, the lexicographic minimum of (distance, child position) is the
// reference's strict-'<' choice.  The upper levels of the tree (k + k^2 + k^3 nodes = 35 KB of descriptors for k = 10)
// stay in L1 / L2; the leaves are one 32-byte row each.
namespace orbgpu {
namespace {

__global__ void __launch_bounds__(128) k_voc_transform(const int *__restrict__ childStart, const int *__restrict__ children,
                                                        const uint4 *__restrict__ nodeDesc, const int *__restrict__ wordId,
                                                        const uint4 *__restrict__ feat, int n, int nidLevel, int *__restrict__ words,
                                                        int *__restrict__ nodes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = t >> 4, sub = t & 15;
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  const bool live = f < n;
  const int fi = live ? f : n - 1;
  const uint4 a0 = feat[2 * (size_t)fi], a1 = feat[2 * (size_t)fi + 1];
  int node = 0, level = 0, nid = nidLevel <= 0 ? 0 : -1;
  for (;;) {
    const int cs = childStart[node], ce = childStart[node + 1];
    if (cs == ce) break;                                    // leaf
    level++;
    unsigned best = 0xffffffffu;                            // distance << 16 | child position
    for (int c = cs + sub; c < ce; c += 16) {
      const int id = children[c];
      const unsigned key = ((unsigned)hamming256(a0, a1, nodeDesc + 2 * (size_t)id) << 16) | (unsigned)(c - cs);
      best = min(best, key);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(mask, best, o, 16));
    node = children[cs + (int)(best & 0xffffu)];
    if (level == nidLevel) nid = node;
  }
  if (live && sub == 0) {
    words[f] = wordId[node];
    nodes[f] = nid >= 0 ? nid : node;                      // leaf above the requested level: see include/orbgpu.h
  }
}

}  // namespace
}  // namespace orbgpu

struct orbgpu_vocabulary {
  int device = 0, nodes = 0, depth = 0;
  int *dChildStart = nullptr, *dChildren = nullptr, *dWordId = nullptr;
  uint8_t *dDesc = nullptr;
};

namespace {
int voc_run(orbgpu_vocabulary *v, const uint4 *dFeat, int n, int levelsup, int32_t *word_ids, int32_t *node_ids, cudaStream_t s, Workspace &W) {
  // results reuse the workspace's match buffer: [words n | nodes n]
  int *dOut = reinterpret_cast<int *>(W.dRes[0]);
  k_voc_transform<<<(n * 16 + 127) / 128, 128, 0, s>>>(v->dChildStart, v->dChildren, reinterpret_cast<const uint4 *>(v->dDesc), v->dWordId, dFeat, n,
                                                      v->depth - levelsup, dOut, dOut + n);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  int *hOut = reinterpret_cast<int *>(W.hRes);
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(hOut, dOut, (size_t)2 * n * sizeof(int), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
  std::memcpy(word_ids, hOut, (size_t)n * sizeof(int));
  std::memcpy(node_ids, hOut + n, (size_t)n * sizeof(int));
  return ORBGPU_OK;
}
}  // namespace

extern "C" {

int orbgpu_vocabulary_create(int32_t nodes, const int32_t *child_start, const int32_t *children, const uint8_t *descriptors,
                             const int32_t *word_id, int32_t depth_levels, int device, orbgpu_vocabulary **out) {
  if (!out) { set_error("null output"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  if (nodes < 1 || !child_start || !descriptors || !word_id || depth_levels < 0 || child_start[0] != 0) { set_error("bad vocabulary arguments"); return ORBGPU_E_INVALID; }
  const int nchild = child_start[nodes];
  if (nchild > 0 && !children) { set_error("null children"); return ORBGPU_E_INVALID; }
  for (int i = 0; i < nodes; i++) {
    if (child_start[i] > child_start[i + 1] || child_start[i + 1] - child_start[i] > 65535) { set_error("node %d: bad child range", i); return ORBGPU_E_INVALID; }
    // children must come after their parent (DBoW2 appends nodes as it grows the tree): guarantees the descent terminates
    for (int c = child_start[i]; c < child_start[i + 1]; c++)
      if (children[c] <= i || children[c] >= nodes) { set_error("node %d: child %d out of order", i, children[c]); return ORBGPU_E_INVALID; }
  }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DevGuard guard(device);
  if (!guard.ok) { set_error("cannot select device %d", device); return ORBGPU_E_CUDA; }
  orbgpu_vocabulary *v = new orbgpu_vocabulary;
  cudaGetDevice(&v->device);
  v->nodes = nodes; v->depth = depth_levels;
  cudaError_t e = cudaMalloc(&v->dChildStart, (size_t)(nodes + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&v->dChildren, (size_t)std::max(nchild, 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&v->dWordId, (size_t)nodes * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&v->dDesc, (size_t)nodes * 32);
  if (e == cudaSuccess) e = cudaMemcpy(v->dChildStart, child_start, (size_t)(nodes + 1) * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nchild) e = cudaMemcpy(v->dChildren, children, (size_t)nchild * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dWordId, word_id, (size_t)nodes * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->dDesc, descriptors, (size_t)nodes * 32, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { set_error("vocabulary upload failed: %s", cudaGetErrorString(e)); orbgpu_vocabulary_destroy(v); return ORBGPU_E_CUDA; }
  *out = v;
  return ORBGPU_OK;
}

void orbgpu_vocabulary_destroy(orbgpu_vocabulary *v) {
  if (!v) return;
  DevGuard guard(v->device);
  cudaFree(v->dChildStart); cudaFree(v->dChildren); cudaFree(v->dWordId); cudaFree(v->dDesc);
  delete v;
}

int orbgpu_vocabulary_transform(orbgpu_vocabulary *v, const uint8_t *descriptors, int n, int levelsup, int32_t *word_ids, int32_t *node_ids) {
  if (!v || n < 0 || (n > 0 && (!descriptors || !word_ids || !node_ids))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (n == 0) return ORBGPU_OK;
  DevGuard guard(v->device);
  Workspace &W = t_ws;
  cudaError_t e = ensure(W, v->device, (size_t)n * 32, (size_t)n, 0);
  if (e != cudaSuccess) { set_error("workspace allocation failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  std::memcpy(W.hStage, descriptors, (size_t)n * 32);
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.dStage, W.hStage, (size_t)n * 32, cudaMemcpyHostToDevice, W.stream));
  return voc_run(v, reinterpret_cast<const uint4 *>(W.dStage), n, levelsup, word_ids, node_ids, W.stream, W);
}

int orbgpu_vocabulary_transform_features(orbgpu_vocabulary *v, orbgpu_features *f, int levelsup, int32_t *word_ids, int32_t *node_ids) {
  if (!v || !f || (f->n > 0 && (!word_ids || !node_ids))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (f->device != v->device) { set_error("vocabulary and feature set live on different devices"); return ORBGPU_E_INVALID; }
  if (f->n == 0) return ORBGPU_OK;
  DevGuard guard(v->device);
  Workspace &W = t_ws;
  cudaError_t e = ensure(W, v->device, 0, (size_t)f->n, 0);
  if (e != cudaSuccess) { set_error("workspace allocation failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  return voc_run(v, reinterpret_cast<const uint4 *>(f->dDesc), f->n, levelsup, word_ids, node_ids, W.stream, W);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:229-321), one warp per MapPoint.  For row i the lanes take the
// columns j in strides of 32; the median (k-th smallest of the N distances, k = floor((N - 1) / 2), the 0 to itself
// included) is found by bisection on the threshold t over [0, 256]: smallest t with |{j : d(i, j) <= t}| >= k + 1.
namespace orbgpu {
namespace {
__global__ void __launch_bounds__(128) k_distinctive(const uint4 *__restrict__ desc, const int *__restrict__ offsets, int groups, int *__restrict__ bestIndex,
                                                      int *__restrict__ bestMedian) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= groups) return;
  const int lo0 = offsets[g], n = offsets[g + 1] - lo0;
  int best = -1, bestMed = 0x7fffffff;
  const int k = (n - 1) >> 1;                              // (size_t)(0.5 * (N - 1))
  for (int i = 0; i < n; i++) {
    const uint4 a0 = desc[2 * (size_t)(lo0 + i)], a1 = desc[2 * (size_t)(lo0 + i) + 1];
    int lo = 0, hi = 256;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int cnt = 0;
      for (int j = lane; j < n; j += 32) cnt += hamming256(a0, a1, desc + 2 * (size_t)(lo0 + j)) <= mid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (cnt >= k + 1) hi = mid; else lo = mid + 1;
    }
    if (lo < bestMed) { bestMed = lo; best = i; }          // strict '<': the first minimum wins (MapPoint.cc:306)
  }
  if (lane == 0) { bestIndex[g] = best; bestMedian[g] = n > 0 ? bestMed : 0; }
}
}  // namespace
}  // namespace orbgpu

extern "C" int orbgpu_distinctive_descriptors(const uint8_t *descriptors, const int32_t *offsets, int groups, int32_t *best_index, int32_t *best_median,
                                              int device) {
  if (groups < 0 || (groups > 0 && (!offsets || !best_index || !best_median))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  if (groups == 0) return ORBGPU_OK;
  for (int g = 0; g < groups; g++)
    if (offsets[g] > offsets[g + 1] || offsets[g] < 0) { set_error("offsets must be non-decreasing"); return ORBGPU_E_INVALID; }
  const int rows = offsets[groups];
  if (rows > 0 && !descriptors) { set_error("null descriptors"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DevGuard guard(device);
  if (!guard.ok) { set_error("cannot select device %d", device); return ORBGPU_E_CUDA; }
  int dev = 0;
  cudaGetDevice(&dev);
  Workspace &W = t_ws;
  const size_t oD = 0, oO = align16((size_t)std::max(rows, 1) * 32), total = align16(oO + (size_t)(groups + 1) * sizeof(int));
  cudaError_t e = ensure(W, dev, total, (size_t)groups, 0);
  if (e != cudaSuccess) { set_error("workspace allocation failed: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  if (rows) std::memcpy(W.hStage + oD, descriptors, (size_t)rows * 32);
  std::memcpy(W.hStage + oO, offsets, (size_t)(groups + 1) * sizeof(int));
  cudaStream_t s = W.stream;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(W.dStage, W.hStage, total, cudaMemcpyHostToDevice, s));
  int *dOut = reinterpret_cast<int *>(W.dRes[0]);          // [index groups | median groups]
  k_distinctive<<<(groups + 3) / 4, 128, 0, s>>>(reinterpret_cast<const uint4 *>(W.dStage + oD), reinterpret_cast<const int *>(W.dStage + oO), groups, dOut,
                                                 dOut + groups);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  int *hOut = reinterpret_cast<int *>(W.hRes);
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(hOut, dOut, (size_t)2 * groups * sizeof(int), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
  std::memcpy(best_index, hOut, (size_t)groups * sizeof(int));
  std::memcpy(best_median, hOut + groups, (size_t)groups * sizeof(int));
  return ORBGPU_OK;
}
