// Frame::ComputeStereoMatches (Frame.cc:477-646) on the device-resident outputs of two extract calls.
//
// k_stereo_rows — one CTA per pair: counting sort of the right keypoints by image row (int)yR
//     (histogram, scan, scatter), the device form of the reference's vRowIndices table.  The sorted
//     table holds 16-byte RECORDS (x, row reach minr / maxr, index, octave), everything the candidate
//     gates need, so the match kernel reads one record per candidate instead of chasing
//     index -> keypoint.
// k_stereo_match — one warp per LEFT keypoint:
//   * candidate scan.  The reference scans the bucket of row (size_t)vL in ascending iR; a right
//     keypoint belongs to that bucket iff floor(yR - r) <= row <= ceil(yR + r), r = 2*scale[octR]
//     (Frame.cc:493-502).  Lanes stride over the sorted right keypoints whose own row lies within
//     +-(ceil(2*scale_max) + 2) of the left row (a superset), apply that exact predicate plus the
//     octave and u-range gates (Frame.cc:540-545) on up to 128 records per trip (four independent
//     128-bit loads per lane), compact the survivors (about one candidate in five) into the warp's
//     shared-memory list by ballot, and only then read descriptors — one per lane, all in flight at
//     once; a shuffle reduction on (distance, iR) lexicographic order reproduces the sequential
//     strict-'<' scan.  The kernel is bound by memory latency: this keeps the dependent loads per
//     keypoint at bucket bounds -> records -> descriptors -> patches.
//   * 11 x 11 SAD at 11 offsets on the keypoint's pyramid level (Frame.cc:557-600): patches are
//     centre-subtracted integers, the L1 norm is exact in int; lanes split the 121 pixels.
//   * parabola fit, disparity and depth in the reference's float order with _rn intrinsics
//     (Frame.cc:606-629).
// k_stereo_filter — one CTA per pair: median of the matched SADs by a two-level (high byte, low
//     byte) histogram select — the same value as sort()[size/2].first — then every match with
//     SAD >= 1.5f*1.4f*median is withdrawn (Frame.cc:634-645).
#include <cmath>

#include "internal.cuh"

namespace orbgpu {

namespace {

constexpr int kThHigh = 100, kThLow = 50;   // ORBmatcher.cc:38-39

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// This translation unit is compiled twice: as is (throughput shape) and through *_lat.cu with a second thread shape for
// frame-at-a-time calls (a stereo pair keeps only 16-32 CTAs busy, so the latency build gives each more threads);
// ORBGPU_VARIANT suffixes the exported launchers.
#ifndef ORBGPU_VARIANT
#define ORBGPU_VARIANT
#endif
#define ORBGPU_CAT2(a, b) a##b
#define ORBGPU_CAT(a, b) ORBGPU_CAT2(a, b)
#define VNAME(x) ORBGPU_CAT(x, ORBGPU_VARIANT)
#ifndef ORBGPU_STEREO_WARPS
#define ORBGPU_STEREO_WARPS 2
#endif
constexpr int kWarps = ORBGPU_STEREO_WARPS;

// rowStart[pair][rows + 1], recs[pair][kpCapR]: right keypoints grouped by (int)y as
// (x bits, minr | maxr << 16, iR, octave) with minr / maxr = floor(yR - r) / ceil(yR + r), r = 2 * scale[octave] (Frame.cc:493-502)
__global__ void __launch_bounds__(256)
k_stereo_rows(const LevelTable *__restrict__ T, StereoArgs A, int rows) {
  extern __shared__ int sHist[];                 // rows + 1 counters, then a running cursor per row
  __shared__ int sWarpSum[8];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int imgR = A.firstR + pair;
  const int nR = A.cntR[imgR];
  const orbgpu_keypoint *kR = A.kpsR + (size_t)imgR * A.kpCapR;
  int *rowStart = A.rowStart + (size_t)pair * (rows + 1);
  int4 *recs = A.recs + (size_t)pair * A.kpCapR;
  for (int i = tid; i <= rows; i += 256) sHist[i] = 0;
  __syncthreads();
  for (int i = tid; i < nR; i += 256) atomicAdd(&sHist[min(max((int)kR[i].y, 0), rows - 1)], 1);
  __syncthreads();
  // exclusive scan of rows + 1 counters: contiguous chunk per thread, warp scan, scan of warp totals
  const int per = (rows + 1 + 255) / 256;
  const int lo = min(tid * per, rows + 1), hi = min(lo + per, rows + 1);
  int sum = 0;
  for (int i = lo; i < hi; i++) sum += sHist[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sWarpSum[warp] = inc;
  __syncthreads();
  int warpBase = 0;
  for (int w = 0; w < warp; w++) warpBase += sWarpSum[w];
  int run = warpBase + inc - sum;
  for (int i = lo; i < hi; i++) { const int v = sHist[i]; sHist[i] = run; rowStart[i] = run; run += v; }
  __syncthreads();
  for (int i = tid; i < nR; i += 256) {
    const orbgpu_keypoint kr = kR[i];
    const int row = min(max((int)kr.y, 0), rows - 1);
    const float r = __fmul_rn(2.0f, T->lv[kr.octave].scale);
    const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));
    recs[atomicAdd(&sHist[row], 1)] = make_int4(__float_as_int(kr.x), (minr & 0xffff) | (maxr << 16), i, kr.octave);
  }
}

#ifndef ORBGPU_STEREO_MIN_CTAS
#define ORBGPU_STEREO_MIN_CTAS 20     // 48 registers, 40 resident warps: 16 / 20 / 24 CTAs per SM (64 / 48 / 40 registers, the last with spills) measured 0.523 / 0.483 / 0.494 ms per 320 pairs
#endif
__global__ void __launch_bounds__(kWarps * 32, ORBGPU_STEREO_MIN_CTAS)
k_stereo_match(const LevelTable *__restrict__ T, StereoArgs A, int rows, int rowReach) {
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int iL = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int imgL = A.firstL + pair, imgR = A.firstR + pair;
  const int nL = A.cntL[imgL];
  if (iL >= nL) return;
  const size_t outIdx = (size_t)pair * A.kpCapL + iL;
  const orbgpu_keypoint *kL = A.kpsL + (size_t)imgL * A.kpCapL;
  const uint4 *dL = reinterpret_cast<const uint4 *>(A.descL + ((size_t)imgL * A.kpCapL + iL) * 32);
  const uint4 *dR = reinterpret_cast<const uint4 *>(A.descR + (size_t)imgR * A.kpCapR * 32);
  const int *rowStart = A.rowStart + (size_t)pair * (rows + 1);
  const int4 *recs = A.recs + (size_t)pair * A.kpCapR;
  __shared__ int2 sCand[kWarps][128];            // gate survivors of one trip: (iR, x bits)
  int2 *cand = sCand[threadIdx.x >> 5];

  float uRight = -1.0f, depth = -1.0f;
  int sadOut = -1;

  const orbgpu_keypoint kp = kL[iL];
  const int levelL = kp.octave;
  const float vL = kp.y, uL = kp.x;
  const int rowL = (int)vL;
  const float minD = 0.f, maxD = __fdiv_rn(A.mbf, A.mb);
  const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);

  int bestDist = kThHigh, bestR = 0x7fffffff, bestX = 0;
  if (!(maxU < 0)) {
    const uint4 a0 = dL[0], a1 = dL[1];
    const int cLo = rowStart[min(max(rowL - rowReach, 0), rows)], cHi = rowStart[min(max(rowL + rowReach + 1, 0), rows)];
    const unsigned below = (1u << lane) - 1u;
    for (int c0 = cLo; c0 < cHi; c0 += 128) {
      int n = 0;                                  // survivors of this trip (warp-uniform)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = c0 + 32 * u + lane;
        bool ok = false;
        int4 rc = make_int4(0, 0, 0, 0);
        if (c < cHi) {
          rc = recs[c];
          const int minr = (int)(short)(rc.y & 0xffff), maxr = rc.y >> 16;
          const float x = __int_as_float(rc.x);
          ok = rowL >= minr && rowL <= maxr && rc.w >= levelL - 1 && rc.w <= levelL + 1 && x >= minU && x <= maxU;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) cand[n + __popc(m & below)] = make_int2(rc.z, rc.x);
        n += __popc(m);
      }
      __syncwarp();
      for (int j = lane; j < n; j += 32) {
        const int2 cd = cand[j];
        const int d = hamming256(a0, a1, dR[2 * cd.x], dR[2 * cd.x + 1]);
        if (d < bestDist || (d == bestDist && cd.x < bestR)) { bestDist = d; bestR = cd.x; bestX = cd.y; }
      }
      __syncwarp();
    }
  }
  const int myDist = bestDist, myR = bestR;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int od = __shfl_xor_sync(0xffffffffu, bestDist, o);
    const int orr = __shfl_xor_sync(0xffffffffu, bestR, o);
    if (od < bestDist || (od == bestDist && orr < bestR)) { bestDist = od; bestR = orr; }
  }
  // x of the winner from the lane that found it (bestR = 0x7fffffff on every lane when nothing passed the gates)
  const unsigned owner = __ballot_sync(0xffffffffu, myDist == bestDist && myR == bestR);
  const float uR0 = __int_as_float(__shfl_sync(0xffffffffu, bestX, __ffs((int)owner) - 1));

  if (bestDist < (kThHigh + kThLow) / 2) {
    const LevelGeom &G = T->lv[levelL];
    const float sf = G.invScale;
    const float suL = roundf(__fmul_rn(kp.x, sf)), svL = roundf(__fmul_rn(kp.y, sf));
    const float suR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, L = 5;
    const float iniu = suR0 + (float)(L - w), endu = suR0 + (float)(L + w + 1);
    if (!(iniu < 0 || endu >= (float)G.w)) {
      const int cy = (int)svL, cxL = (int)suL, cxR = (int)suR0;
      const uint8_t *PL = A.pyrL + G.pyrOffset + (size_t)imgL * G.planeBytes + (size_t)(cy + kEdge) * G.pitch + (cxL + kEdge);
      const uint8_t *PR = A.pyrR + G.pyrOffset + (size_t)imgR * G.planeBytes + (size_t)(cy + kEdge) * G.pitch + (cxR + kEdge);
      const int centreL = PL[0];
      // left patch values of this lane's pixels (4 strides over the 121 pixels)
      int lv[4], off[4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int i = lane + 32 * t;
        const int dy = i / 11 - w, dx = i % 11 - w;
        off[t] = dy * G.pitch + dx;
        lv[t] = i < 121 ? (int)PL[off[t]] - centreL : 0;
      }
      // per-lane partial SADs of the 11 shifts: |lv - (R - centreR)| = |(lv + centreR) - R| as one add + one absolute-difference-
      // accumulate per pixel; the warp sums are exact integers <= 121 * 510 < 2^16, so two shifts share one 32-bit warp reduction
      unsigned part[11];
#pragma unroll
      for (int inc = -L; inc <= L; inc++) {
        const int centreR = PR[inc];
        unsigned sad = 0;
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (lane + 32 * t < 121) sad = __sad(lv[t] + centreR, (int)PR[off[t] + inc], sad);
        part[inc + L] = sad;
      }
      // best shift (first minimum, Frame.cc:586-590) and its two neighbours for the parabola, tracked while the sums arrive:
      // no indexed array, everything stays in registers
      int best = 0x7fffffff, bestInc = 0;
      float prev = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      bool justBest = false;
      auto visit = [&](int k, int sv) {
        const float f = (float)sv;
        if (justBest) d3 = f;
        justBest = sv < best;
        if (justBest) { best = sv; bestInc = k - L; d1 = prev; d2 = f; }
        prev = f;
      };
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const unsigned both = __reduce_add_sync(0xffffffffu, part[2 * k] | (k < 5 ? part[2 * k + 1] << 16 : 0u));
        visit(2 * k, (int)(both & 0xffffu));
        if (k < 5) visit(2 * k + 1, (int)(both >> 16));
      }
      if (!(bestInc == -L || bestInc == L)) {
        const float deltaR = __fdiv_rn(__fsub_rn(d1, d3),
                                       __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
        if (!(deltaR < -1 || deltaR > 1)) {
          float bestuR = __fmul_rn(G.scale, __fadd_rn(__fadd_rn(suR0, (float)bestInc), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) {
              disparity = (float)0.01;
              bestuR = (float)((double)uL - 0.01);
            }
            depth = __fdiv_rn(A.mbf, disparity);
            uRight = bestuR;
            sadOut = best;
          }
        }
      }
    }
  }
  if (lane == 0) {
    A.uRight[outIdx] = uRight;
    A.depth[outIdx] = depth;
    A.bestSad[outIdx] = sadOut;
  }
}

__global__ void __launch_bounds__(256)
k_stereo_filter(StereoArgs A) {
  __shared__ int sHist[256];
  __shared__ int sBin, sRank, sCount;
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int nL = A.cntL[A.firstL + pair];
  const int *sad = A.bestSad + (size_t)pair * A.kpCapL;
  float *uRight = A.uRight + (size_t)pair * A.kpCapL, *depth = A.depth + (size_t)pair * A.kpCapL;

  sHist[tid] = 0;
  if (tid == 0) sCount = 0;
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < nL; i += 256) {
    const int s = sad[i];
    if (s >= 0) { atomicAdd(&sHist[min(s >> 8, 255)], 1); mine++; }
  }
  if (mine) atomicAdd(&sCount, mine);
  __syncthreads();
  const int count = sCount;
  if (count == 0) return;             // reference: undefined behaviour (Frame.cc:635); nothing to filter
  if (tid == 0) {
    int rank = count / 2, b = 0;
    while (rank >= sHist[b]) { rank -= sHist[b]; b++; }
    sBin = b; sRank = rank;
  }
  __syncthreads();
  const int bin = sBin;
  __syncthreads();
  sHist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < nL; i += 256) {
    const int s = sad[i];
    if (s >= 0 && min(s >> 8, 255) == bin) atomicAdd(&sHist[s & 255], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int rank = sRank, b = 0;
    while (rank >= sHist[b]) { rank -= sHist[b]; b++; }
    sBin = (bin << 8) | b;
  }
  __syncthreads();
  const float median = (float)sBin;
  const float thDist = __fmul_rn(1.5f * 1.4f, median);
  for (int i = tid; i < nL; i += 256) {
    const int s = sad[i];
    if (s >= 0 && !((float)s < thDist)) { uRight[i] = -1.0f; depth[i] = -1.0f; }
  }
}

}  // namespace

cudaError_t VNAME(stereo_configure)(int rows) {
  return cudaFuncSetAttribute(k_stereo_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (rows + 1) * (int)sizeof(int));
}

cudaError_t VNAME(launch_stereo)(const LevelTable &T, const LevelTable *dT, const StereoArgs &A, int maxKp, cudaStream_t s) {
  const int rows = T.lv[0].h;
  // a right keypoint at row yR reaches rows floor(yR - r) .. ceil(yR + r), r <= 2 * scale of the top level
  const int rowReach = (int)std::ceil(2.0f * T.lv[T.nlevels - 1].scale) + 2;
  k_stereo_rows<<<A.pairs, 256, (rows + 1) * sizeof(int), s>>>(dT, A, rows);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid((maxKp + kWarps - 1) / kWarps, A.pairs);
  k_stereo_match<<<grid, kWarps * 32, 0, s>>>(dT, A, rows, rowReach);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_stereo_filter<<<A.pairs, 256, 0, s>>>(A);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
