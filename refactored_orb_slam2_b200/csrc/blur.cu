// cv::GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as called at ORBextractor.cc:1018-1019.
//
// OpenCV's 8-bit path is a fixed-point separable filter: kernel {18,34,48,56,48,34,18}/256,
// horizontal sums kept exactly in 8.8 (<= 65280), vertical sums in 16.16, one rounding at the end
// ((v + 0x8000) >> 16).  Nothing is rounded before that, so the result is the exact 2-D integer
// convolution and any evaluation order reproduces it.  The source plane already carries the
// reflect-101 apron, so reading 3 px past the interior reproduces the border mode.
//
// Work unit = one WARP on a 128 x 66 tile, warps fully independent (no __syncthreads):
//   * one TMA tensor copy (cp.async.bulk.tensor.3d, per-warp mbarrier) stages the (128+16) x (66+6)
//     source window in the warp's shared-memory slot;
//   * a lane owns 4 adjacent columns and walks down the rows: 3 LDS.32 per row, per pixel two funnel-shifted
//     words and two DP4A for the seven horizontal taps (sums <= 65280 fit 16 bits),
//     then the vertical taps run as IDP.2A on vertically packed row pairs (see the kernel); every input
//     row finishes one output row, which leaves as one 32-bit store per lane.
// Tiles are anchored at padded x = 4: the window then starts on a 16-byte boundary (a TMA
// requirement for the innermost coordinate) and loads / stores are 4-byte aligned; the 15 columns
// left of the interior receive valid blur values of the mirrored image and are never read.
#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

namespace {

#ifndef ORBGPU_BLUR_TH
#define ORBGPU_BLUR_TH 66
#endif
#ifndef ORBGPU_BLUR_WARPS
#define ORBGPU_BLUR_WARPS 4
#endif
constexpr int kTW = 128, kTH = ORBGPU_BLUR_TH;
constexpr int kBoxW = kTW + 16, kBoxH = kTH + 6;
constexpr int kWarpsPerCta = ORBGPU_BLUR_WARPS;
constexpr int kSlotBytes = (kBoxW * kBoxH + 127) & ~127;

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

  // Vertical pass on VERTICALLY packed pairs: P(r)[p] = (h of pixel p on input row r-1, h on input row r) as two 16-bit halves,
  // so that one IDP.2A applies two vertical taps of one pixel.  The output row centred on input row r-3 is complete when input
  // row r arrives:
  //   18 h[r-6] + 34 h[r-5] | 48 h[r-4] + 56 h[r-3] | 48 h[r-2] + 34 h[r-1] | 18 h[r]
  //   = dp2a(P(r-5), 18, 34) + dp2a(P(r-3), 48, 56) + dp2a(P(r-1), 48, 34) + dp2a(P(r), 0, 18)      (+ 0x8000, then >> 16)
  // Six rows of pairs (24 registers, rotated by the six-fold unroll) replace the seven rolling 32-bit accumulators: per row and
  // lane 4 PRMT + 16 IDP.2A + 3 PRMT (byte 2 of each sum IS the rounded result) instead of 4 extractions + 28 IMAD + 11.
  unsigned P[6][4];
  unsigned hPrev[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < 6; k++)
#pragma unroll
    for (int p = 0; p < 4; p++) P[k][p] = 0u;

  static_assert(kBoxH % 6 == 0, "row loop is unrolled by the pair-history depth");
  const int rEnd = min(kBoxH, (rowsOut + 6 + 5) / 6 * 6);     // a level's last tile row stops after its own rows (+ the 6 rows of history)
  for (int r0 = 0; r0 < rEnd; r0 += 6) {
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
      const int r = r0 + rr;
      // horizontal taps of input row r (tile row r-3): h for pixels x..x+3, two pixels per register
      const unsigned *row = reinterpret_cast<const unsigned *>(slot + r * kBoxW) + lane;
      const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
      // pixel x+j reads window bytes j+1 .. j+7 (w0 = bytes x-4 .. x-1): two funnel-shifted words per pixel hold them, and the
      // seven taps are two byte dot products  dp4a(bytes j+1..j+4, 18 34 48 56) + dp4a(bytes j+5..j+8, 48 34 18 0)
      unsigned h[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned a = j == 3 ? w1 : __funnelshift_r(w0, w1, 8 * (j + 1));
        const unsigned b = j == 3 ? w2 : __funnelshift_r(w1, w2, 8 * (j + 1));
        h[j] = __dp4a(b, 0x00122230u, __dp4a(a, 0x38302212u, 0u));
      }
      unsigned (&Pn)[4] = P[rr];                         // P(r) takes the slot of P(r-6)
#pragma unroll
      for (int j = 0; j < 4; j++) { Pn[j] = __byte_perm(hPrev[j], h[j], 0x5410u); hPrev[j] = h[j]; }
      const int orow = r - 6;
      if (orow >= 0 && colOk && orow < rowsOut) {
        const unsigned (&P1)[4] = P[(rr + 5) % 6], (&P3)[4] = P[(rr + 3) % 6], (&P5)[4] = P[(rr + 1) % 6];
        unsigned sum[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          unsigned a = __dp2a_lo(P5[p], 0x00002212u, 0x8000u);             // 18 h[r-6] + 34 h[r-5] + rounding
          a = __dp2a_lo(P3[p], 0x00003830u, a);                            // 48 h[r-4] + 56 h[r-3]
          a = __dp2a_lo(P1[p], 0x00002230u, a);                            // 48 h[r-2] + 34 h[r-1]
          sum[p] = __dp2a_lo(Pn[p], 0x00001200u, a);                       //  0 h[r-1] + 18 h[r]
        }
        // (sum >> 16) fits a byte: byte 2 of each sum, packed
        const unsigned lo = __byte_perm(sum[0], sum[1], 0x0062u), hi = __byte_perm(sum[2], sum[3], 0x0062u);
        *reinterpret_cast<unsigned *>(dst + (size_t)orow * G.pitch) = __byte_perm(lo, hi, 0x5410u);
      }
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
