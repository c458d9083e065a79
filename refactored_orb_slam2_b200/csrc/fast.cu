// Per-cell FAST-9/16 with non-max suppression: the detection half of
// ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:733-799), i.e. the reference's
//   for every ~30x30 cell: cv::FAST(cell, iniThFAST, nms) ; if empty: cv::FAST(cell, minThFAST, nms)
//
// Formulation (equivalent to the per-cell cv::FAST calls, see DESIGN.md "FAST"):
//   m(p)  = max over the 16 arcs of 9 contiguous ring pixels of min(+-(ring - centre))
//           -> p is a corner at threshold t  <=>  m(p) > t ;  cv::FAST's response is m(p) - 1
//   A cell's detection rectangle is the cell image minus cv::FAST's 3 px frame; the rectangles of
//   neighbouring cells tile the level without gaps or overlap.  Inside its rectangle a pixel
//   survives NMS  <=>  m(p) > t  and  m(p) > m(q) for the 8 neighbours q INSIDE the same
//   rectangle (cv::FAST scores pixels outside the rectangle, and non-corners, as 0).
//   t = iniThFAST if the cell has at least one survivor at iniThFAST, else minThFAST.
//
// One CTA = up to kCellsPerTile horizontally adjacent cells of one cell row of one level of one
// image.  The source window (+3 px ring apron) is staged in shared memory; arc margins are
// evaluated for two horizontally adjacent pixels at a time in packed s16x2 registers with the
// 3-input VIMNMX3.S16x2 min/max of sm_100a (a sliding min3-of-min3 gives every 9-arc in 32 ops);
// a cheap upper bound from the 8 opposite ring pairs rejects most pixel pairs first.  Scores go to
// a shared u8 tile, NMS + the per-cell threshold fallback run on that tile, survivors are staged in
// shared memory and appended to the level's candidate list with one global atomic per CTA.
// Candidate order in HBM is arbitrary: the octree only needs each candidate's position in the
// reference's cell-major order to break response ties, and that is a function of (x, y).
#include "internal.cuh"

namespace orbgpu {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned pack2(unsigned lo, unsigned hi) { return lo | (hi << 16); }
__device__ __forceinline__ unsigned mn3(unsigned a, unsigned b, unsigned c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mx3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mn2(unsigned a, unsigned b) { return __vmins2(a, b); }
__device__ __forceinline__ unsigned mx2(unsigned a, unsigned b) { return __vmaxs2(a, b); }

// Arc margins of the two pixels at p[0], p[1] (row pitch `sp`), packed s16x2; returns 0x80008000-ish
// negative halves when the pair cannot reach `minTh` (prefilter), exact margins otherwise.
__device__ __forceinline__ unsigned fast_margin_pair(const uint8_t *p, int sp, int minTh) {
  unsigned d[16];
  {
    const uint8_t *r;
    const unsigned negc = __vneg2(pack2(p[0], p[1]));
    r = p + 3 * sp;
    { unsigned a = r[-1], b = r[0], c = r[1], e = r[2];
      d[15] = __vadd2(pack2(a, b), negc); d[0] = __vadd2(pack2(b, c), negc); d[1] = __vadd2(pack2(c, e), negc); }
    r = p + 2 * sp;
    { unsigned a = r[-2], b = r[-1], c = r[2], e = r[3];
      d[14] = __vadd2(pack2(a, b), negc); d[2] = __vadd2(pack2(c, e), negc); }
    r = p + sp;
    { unsigned a = r[-3], b = r[-2], c = r[3], e = r[4];
      d[13] = __vadd2(pack2(a, b), negc); d[3] = __vadd2(pack2(c, e), negc); }
    r = p;
    { unsigned a = r[-3], b = r[-2], c = r[3], e = r[4];
      d[12] = __vadd2(pack2(a, b), negc); d[4] = __vadd2(pack2(c, e), negc); }
    r = p - sp;
    { unsigned a = r[-3], b = r[-2], c = r[3], e = r[4];
      d[11] = __vadd2(pack2(a, b), negc); d[5] = __vadd2(pack2(c, e), negc); }
    r = p - 2 * sp;
    { unsigned a = r[-2], b = r[-1], c = r[2], e = r[3];
      d[10] = __vadd2(pack2(a, b), negc); d[6] = __vadd2(pack2(c, e), negc); }
    r = p - 3 * sp;
    { unsigned a = r[-1], b = r[0], c = r[1], e = r[2];
      d[9] = __vadd2(pack2(a, b), negc); d[8] = __vadd2(pack2(b, c), negc); d[7] = __vadd2(pack2(c, e), negc); }
  }
  // Prefilter: every 9-arc contains one pixel of each opposite pair (k, k+8), so
  //   bright margin <= min_k max(d_k, d_k+8)   and   dark margin <= -max_k min(d_k, d_k+8).
  {
    unsigned hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { hi[k] = mx2(d[k], d[k + 8]); lo[k] = mn2(d[k], d[k + 8]); }
    const unsigned ubB = mn3(mn3(hi[0], hi[1], hi[2]), mn3(hi[3], hi[4], hi[5]), mn2(hi[6], hi[7]));
    const unsigned lbD = mx3(mx3(lo[0], lo[1], lo[2]), mx3(lo[3], lo[4], lo[5]), mx2(lo[6], lo[7]));
    const unsigned ub = mx2(ubB, __vneg2(lbD));
    const int u0 = (short)(ub & 0xffff), u1 = (short)(ub >> 16);
    if (max(u0, u1) <= minTh) return 0x80008000u;   // both halves = -32768
  }
  // min / max over every window of 3, then of 9 = three windows of 3 spaced by 3
  unsigned a3[16], b3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    a3[k] = mn3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    b3[k] = mx3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  }
  unsigned bright = 0x80008000u, darkNeg = 0x7fff7fffu;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const unsigned a9 = mn3(a3[k], a3[(k + 3) & 15], a3[(k + 6) & 15]);
    const unsigned a9b = mn3(a3[k + 1], a3[(k + 4) & 15], a3[(k + 7) & 15]);
    bright = mx3(bright, a9, a9b);
    const unsigned b9 = mx3(b3[k], b3[(k + 3) & 15], b3[(k + 6) & 15]);
    const unsigned b9b = mx3(b3[k + 1], b3[(k + 4) & 15], b3[(k + 7) & 15]);
    darkNeg = mn3(darkNeg, b9, b9b);
  }
  return mx2(bright, __vneg2(darkNeg));
}

// tiles[t] = (level, cell row i, first cell column j0, cell count)
__global__ void __launch_bounds__(kThreads)
k_fast_cells(const LevelTable *__restrict__ T, const uint8_t *__restrict__ pyr, const int4 *__restrict__ tiles,
             Cand *__restrict__ cand, int *__restrict__ candCount, unsigned *__restrict__ status,
             int imgPitchMax, int scorePitchMax, int scoreBytes, int stageCap) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int sCellCnt[8];
  __shared__ int sEmit, sBase;

  const int4 tile = tiles[blockIdx.x];
  const int level = tile.x, ci = tile.y, cj0 = tile.z, nc = tile.w;
  const LevelGeom &G = T->lv[level];
  const int img = blockIdx.y;
  const int iniTh = T->iniTh, minTh = T->minTh;
  const int tid = threadIdx.x;

  // detection rectangle of the tile, level coordinates
  const int X0 = kEdge + cj0 * G.wCell, Y0 = kEdge + ci * G.hCell;
  const int X1 = min(kEdge + (cj0 + nc) * G.wCell, G.w - kEdge);
  const int Y1 = min(Y0 + G.hCell, G.h - kEdge);
  const int TW = X1 - X0, TH = Y1 - Y0;
  if (TW <= 0 || TH <= 0) return;

  // ---- stage the source window: rows Y0-3 .. Y1+2, columns from the 4-aligned byte at or before X0-3
  const int px0 = (X0 - 3 + kEdge) & ~3;            // padded column of staged byte 0
  const int lead = (X0 - 3 + kEdge) - px0;          // staged column of level x = X0-3
  const int sp = imgPitchMax;                       // staged row pitch (multiple of 4)
  const int words = (lead + TW + 10) >> 2;           // covers the phantom pixel of an odd-width tile
  uint8_t *sImg = smem;
  uint8_t *sScore = smem + (size_t)sp * (G.hCell + 6);          // score tile, pitch scorePitchMax
  Cand *sStage = reinterpret_cast<Cand *>(sScore + scoreBytes);
  const uint8_t *S = pyr + G.pyrOffset + (size_t)img * G.planeBytes;
  for (int i = tid; i < (TH + 6) * words; i += kThreads) {
    const int r = i / words, c4 = i - r * words;
    const int x = px0 + 4 * c4;
    unsigned v = 0;
    if (x + 3 < G.pitch) v = *reinterpret_cast<const unsigned *>(S + (size_t)(Y0 - 3 + r + kEdge) * G.pitch + x);
    *reinterpret_cast<unsigned *>(sImg + r * sp + 4 * c4) = v;
  }
  if (tid < 8) sCellCnt[tid] = 0;
  if (tid == 0) sEmit = 0;
  __syncthreads();

  // ---- arc margins, two pixels per step
  const int pairsPerRow = (TW + 1) >> 1;
  for (int i = tid; i < pairsPerRow * TH; i += kThreads) {
    const int y = i / pairsPerRow, x = (i - y * pairsPerRow) * 2;
    const uint8_t *p = sImg + (y + 3) * sp + lead + 3 + x;
    const unsigned m = fast_margin_pair(p, sp, minTh);
    const int m0 = (short)(m & 0xffff), m1 = (short)(m >> 16);
    sScore[y * scorePitchMax + x] = (uint8_t)(m0 > minTh ? m0 : 0);
    if (x + 1 < TW) sScore[y * scorePitchMax + x + 1] = (uint8_t)(m1 > minTh ? m1 : 0);
  }
  __syncthreads();

  // ---- NMS inside each cell's rectangle; count survivors at iniTh per cell
  unsigned long long surv = 0;
  int it = 0;
  for (int i = tid; i < TW * TH; i += kThreads, it++) {
    const int y = i / TW, x = i - y * TW;
    const int s = sScore[y * scorePitchMax + x];
    if (s == 0) continue;
    const int c = x / G.wCell;
    const int cx0 = c * G.wCell, cx1 = min(cx0 + G.wCell, TW);
    bool keep = true;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const int ny = y + dy;
      if (ny < 0 || ny >= TH) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        const int nx = x + dx;
        if ((dx | dy) == 0 || nx < cx0 || nx >= cx1) continue;
        keep = keep && s > sScore[ny * scorePitchMax + nx];
      }
    }
    if (keep) {
      surv |= 1ull << it;
      if (s > iniTh) atomicAdd(&sCellCnt[c], 1);
    }
  }
  __syncthreads();

  // ---- emit: cells with no survivor at iniTh fall back to minTh (ORBextractor.cc:776-780)
  it = 0;
  for (int i = tid; i < TW * TH; i += kThreads, it++) {
    if (!((surv >> it) & 1ull)) continue;
    const int y = i / TW, x = i - y * TW;
    const int s = sScore[y * scorePitchMax + x];
    const int c = x / G.wCell;
    if (s > iniTh || sCellCnt[c] == 0) {
      const int slot = atomicAdd(&sEmit, 1);
      if (slot < stageCap) {
        Cand k;
        k.xy = (unsigned)(X0 + x - kMinBorder) | ((unsigned)(Y0 + y - kMinBorder) << 16);
        k.resp = (unsigned)(s - 1);
        sStage[slot] = k;
      }
    }
  }
  __syncthreads();
  const int n = min(sEmit, stageCap);
  if (tid == 0) {
    sBase = n > 0 ? atomicAdd(&candCount[img * kMaxLevels + level], n) : 0;
    if (sEmit > stageCap) atomicOr(status, kStatusCandOverflow);
  }
  __syncthreads();
  const int base = sBase;
  if (base + n > G.candCap && tid == 0) atomicOr(status, kStatusCandOverflow);
  Cand *out = cand + G.candOffset + (size_t)img * G.candCap;
  for (int i = tid; i < n; i += kThreads)
    if (base + i < G.candCap) out[base + i] = sStage[i];
}

}  // namespace

cudaError_t fast_configure(int smemBytes) {
  return cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
}

FastConfig fast_make_config(const LevelTable &T) {
  int wMax = 0, hMax = 0;
  for (int l = 0; l < T.nlevels; l++) { wMax = max(wMax, T.lv[l].wCell); hMax = max(hMax, T.lv[l].hCell); }
  FastConfig c;
  const int tw = kFastCellsPerTile * wMax;
  c.imgPitch = (3 + tw + 10 + 3) & ~3;
  c.scorePitch = (tw + 3) & ~3;
  c.scoreBytes = (c.scorePitch * hMax + 15) & ~15;
  c.stageCap = (tw / 2 + 1) * (hMax / 2 + 1);      // bound on strict 8-neighbour maxima in a tile
  c.smemBytes = c.imgPitch * (hMax + 6) + c.scoreBytes + c.stageCap * (int)sizeof(Cand);
  c.maxTilePixels = tw * hMax;
  return c;
}

cudaError_t launch_fast(const LevelTable &T, const ExtractBuffers &B, const int4 *tiles, int ntiles,
                        const FastConfig &c, int batch, cudaStream_t s) {
  dim3 grid(ntiles, batch);
  k_fast_cells<<<grid, kThreads, c.smemBytes, s>>>(B.dT, B.pyr, tiles, B.cand, B.candCount, B.status, c.imgPitch,
                                                    c.scorePitch, c.scoreBytes, c.stageCap);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
