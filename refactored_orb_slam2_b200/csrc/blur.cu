// cv::GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as called at ORBextractor.cc:1018-1019.
//
// OpenCV's 8-bit path is a fixed-point separable filter: kernel {18,34,48,56,48,34,18}/256,
// horizontal sums kept exactly in 8.8 (<= 65280), vertical sums in 16.16, one rounding at the end
// ((v + 0x8000) >> 16).  Nothing is rounded before that, so the result is the exact 2-D integer
// convolution and any evaluation order reproduces it.  The source plane already carries the
// reflect-101 apron, so reading 3 px past the interior reproduces the border mode.
//
// Work unit = one WARP on a 128 x 64 tile, warps fully independent (no __syncthreads):
//   * one TMA tensor copy (cp.async.bulk.tensor.3d, per-warp mbarrier) stages the (128+16) x (64+6)
//     source window in the warp's shared-memory slot;
//   * a lane owns 4 adjacent columns and walks down the rows: 3 LDS.32 per row, horizontal taps on
//     byte pairs packed 16x2 (one IMAD filters two pixels; sums <= 65280 never carry across halves),
//     then 7 rolling 32-bit accumulators per column take the vertical taps; every input row
//     finishes one output row, which leaves as one 32-bit store per lane.
// Tiles are anchored at padded x = 4: the window then starts on a 16-byte boundary (a TMA
// requirement for the innermost coordinate) and loads / stores are 4-byte aligned; the 15 columns
// left of the interior receive valid blur values of the mirrored image and are never read.
#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

namespace {

#ifndef ORBGPU_BLUR_TH
#define ORBGPU_BLUR_TH 64
#endif
#ifndef ORBGPU_BLUR_WARPS
#define ORBGPU_BLUR_WARPS 4
#endif
constexpr int kTW = 128, kTH = ORBGPU_BLUR_TH;
constexpr int kBoxW = kTW + 16, kBoxH = kTH + 6;
constexpr int kWarpsPerCta = ORBGPU_BLUR_WARPS;
constexpr int kSlotBytes = (kBoxW * kBoxH + 127) & ~127;

// bytes O, O+1 of the 12-byte window (w0, w1, w2) zero-extended into packed 16x2
template <int O>
__device__ __forceinline__ unsigned pair_at(unsigned w0, unsigned w1, unsigned w2) {
  static_assert(O >= 0 && O <= 10, "window offset");
  if constexpr (O <= 2) return __byte_perm(w0, 0u, 0x4040u | O | ((O + 1) << 8));
  else if constexpr (O == 3) return __byte_perm(w0, w1, 0x4443u) & 0x00ff00ffu;
  else if constexpr (O <= 6) return __byte_perm(w1, 0u, 0x4040u | (O - 4) | ((O - 3) << 8));
  else if constexpr (O == 7) return __byte_perm(w1, w2, 0x4443u) & 0x00ff00ffu;
  else return __byte_perm(w2, 0u, 0x4040u | (O - 8) | ((O - 7) << 8));
}

// tiles[t] = (level, tile x, tile y, 0)
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_blur(const __grid_constant__ TmaMaps maps, const LevelTable *__restrict__ T, const int4 *__restrict__ tiles, int ntiles,
       uint8_t *__restrict__ blur, int imgBase) {
  extern __shared__ uint8_t smemRaw[];
  __shared__ __align__(8) unsigned long long sBar[kWarpsPerCta];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * kWarpsPerCta + warp;
  if (t >= ntiles) return;
  const int4 tile = tiles[t];
  const LevelGeom &G = T->lv[tile.x];
  const int img = imgBase + blockIdx.y;
  const int x0 = 4 + kTW * tile.y;                  // padded coordinates of the tile's first output column / row
  const int y0 = kEdge + kTH * tile.z;
  uint8_t *slot = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u) + warp * kSlotBytes;

  if (lane == 0) {
    mbar_init(&sBar[warp], 1);
    mbar_fence_init();
    mbar_expect_tx(&sBar[warp], kBoxW * kBoxH);
    tma_load_3d(slot, &maps.m[tile.x], x0 - 4, y0 - 3, img, &sBar[warp]);
  }
  __syncwarp();
  mbar_wait(&sBar[warp], 0);

  const int xEnd = 4 + 4 * ((G.w + 15 + 3) / 4);    // first padded column not produced
  const int x = x0 + 4 * lane;
  const int rowsOut = min(kTH, G.h + kEdge - y0);
  const bool colOk = x < xEnd;
  uint8_t *dst = blur + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)y0 * G.pitch + x;

  unsigned acc[7][4];
#pragma unroll
  for (int j = 0; j < 7; j++)
#pragma unroll
    for (int p = 0; p < 4; p++) acc[j][p] = 0;

  // 7 input rows per loop trip: the accumulator rotation has period 7, so registers stay static while the code stays
  // small enough for the instruction cache (a full 70-row unroll stalled on instruction fetch)
  static_assert(kBoxH % 7 == 0, "row loop is unrolled by the accumulator period");
  for (int r0 = 0; r0 < kBoxH; r0 += 7) {
#pragma unroll
    for (int rr = 0; rr < 7; rr++) {
      const int r = r0 + rr;
      // horizontal taps of input row r (tile row r-3): h for pixels x..x+3, two pixels per register
      const unsigned *row = reinterpret_cast<const unsigned *>(slot + r * kBoxW) + lane;
      const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
      const unsigned p1 = pair_at<1>(w0, w1, w2), p2 = pair_at<2>(w0, w1, w2), p3 = pair_at<3>(w0, w1, w2),
                     p4 = pair_at<4>(w0, w1, w2), p5 = pair_at<5>(w0, w1, w2), p6 = pair_at<6>(w0, w1, w2),
                     p7 = pair_at<7>(w0, w1, w2), p8 = pair_at<8>(w0, w1, w2), p9 = pair_at<9>(w0, w1, w2);
      // pixel pair (x, x+1): window bytes 1..8 ; pair (x+2, x+3): bytes 3..10
      const unsigned hA = 18u * (p1 + p7) + 34u * (p2 + p6) + 48u * (p3 + p5) + 56u * p4;
      const unsigned hB = 18u * (p3 + p9) + 34u * (p4 + p8) + 48u * (p5 + p7) + 56u * p6;
      const unsigned h[4] = {hA & 0xffffu, hA >> 16, hB & 0xffffu, hB >> 16};
      // vertical taps: input row r feeds output rows r-6 .. r; accumulator (rr + j) % 7 belongs to output row r-6+j
      constexpr unsigned kW[7] = {18u, 34u, 48u, 56u, 48u, 34u, 18u};
#pragma unroll
      for (int j = 0; j < 7; j++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[(rr + j) % 7][p] += kW[6 - j] * h[p];
      // accumulator rr now holds output row r-6 complete
      const int orow = r - 6;
      if (orow >= 0 && colOk && orow < rowsOut) {
        const unsigned out = ((acc[rr][0] + 0x8000u) >> 16) | (((acc[rr][1] + 0x8000u) >> 16) << 8) |
                             (((acc[rr][2] + 0x8000u) >> 16) << 16) | (((acc[rr][3] + 0x8000u) >> 16) << 24);
        *reinterpret_cast<unsigned *>(dst + (size_t)orow * G.pitch) = out;
      }
#pragma unroll
      for (int p = 0; p < 4; p++) acc[rr][p] = 0;       // it restarts as the accumulator of output row r+1
    }
  }
}

}  // namespace

// Per-level TMA box of the blur kernel is the same for every level.
int blur_box_w() { return kBoxW; }
int blur_box_h() { return kBoxH; }
int blur_smem_bytes() { return 128 + kWarpsPerCta * kSlotBytes; }
cudaError_t blur_configure() {
  return cudaFuncSetAttribute(k_blur, cudaFuncAttributeMaxDynamicSharedMemorySize, blur_smem_bytes());
}

void blur_make_tiles(const LevelTable &T, std::vector<int4> &tiles) {
  tiles.clear();
  for (int l = 0; l < T.nlevels; l++) {
    const LevelGeom &G = T.lv[l];
    const int tx = (G.w + 15 + kTW - 1) / kTW, ty = (G.h + kTH - 1) / kTH;
    for (int y = 0; y < ty; y++)
      for (int x = 0; x < tx; x++) tiles.push_back(make_int4(l, x, y, 0));
  }
}

cudaError_t launch_blur(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &maps, const int4 *tiles, int ntiles,
                        int imgBase, int batch, cudaStream_t s) {
  dim3 grid((ntiles + kWarpsPerCta - 1) / kWarpsPerCta, batch);
  k_blur<<<grid, kWarpsPerCta * 32, blur_smem_bytes(), s>>>(maps, B.dT, tiles, ntiles, B.blur, imgBase);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
