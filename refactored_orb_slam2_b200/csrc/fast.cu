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
// One CTA = up to kFastCellsPerTile horizontally adjacent cells of one cell row of one level of one
// image.  The source window (+3 px ring apron) is brought into shared memory by ONE TMA tensor copy
// (cp.async.bulk.tensor.3d on the level's [batch][rows][pitch] u8 tensor map, completion on an
// mbarrier).  TMA needs the window's first column on a 16-byte boundary, so pixels are grouped by
// fours on the PADDED plane's 4-byte grid (groups sit on 32-bit words in shared memory) and the up
// to 3 + 3 pixels a row of groups overhangs the tile are masked out.  The u8 window is unpacked ONCE per tile into two
// u16 planes — plane A holds pixel pairs (2j, 2j+1) per 32-bit word, plane B the pairs (2j+1, 2j+2) — so that every ring
// pixel pair of a thread's two adjacent pixels, at any horizontal offset, is one aligned shared-memory word that is already
// the packed u16x2 operand of the arithmetic (no per-use PRMT / mask: the unpack costs 1 ALU op per pixel instead of ~13).
// A thread owns 4 horizontally adjacent pixels = two such pairs: 22 LDS (64-bit where the second pair's word sits next to
// the first's) fetch both rings and centres, and the arc margins come from the 3-input VIMNMX3.S16x2 min/max of sm_100a
// (groups of four arcs around their shared six pixels, 68 ops per pixel pair for both polarities).  The TMA copy of the next
// tile is issued as soon as the unpack has released the u8 window, so it overlaps the whole margin / NMS / emit phase.  Scores (margin - minTh, 0 = not a corner at minTh) go to a
// shared u8 tile and the coordinates of the non-zero ones to a shared queue (warp-aggregated
// reservation); NMS and the per-cell threshold fallback run over that queue only; survivors are
// staged in shared memory and appended to the level's candidate list with one global atomic per CTA.
// Candidate order in HBM is arbitrary: the octree only needs each candidate's position in the
// reference's cell-major order to break response ties, and that is a function of (x, y).
#include <cstdlib>

#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

namespace {

constexpr int kThreads = ORBGPU_FAST_THREADS;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ unsigned mn3(unsigned a, unsigned b, unsigned c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mx3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mx2(unsigned a, unsigned b) { return __vmaxs2(a, b); }

// Arc margins (packed s16x2) of two horizontally adjacent pixels from their 16 ring values d[k] (packed u16x2, ring order
// of cv::FAST 9/16: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)) and
// their centres c.  min / max commute with subtracting the centre, so the sliding-window networks run on the RAW ring pixels
// and the centre is subtracted once from the two extrema:
//   bright margin = max_arcs min_arc(ring) - centre ,  dark margin = centre - min_arcs max_arc(ring).
// Returns max(margin - minTh, 0) per halfword; k1 = (1 - minTh) in both halves.  With ~v = -v - 1 per halfword,
//   margin - minTh = max(bright + (~c + k1), (c + k1) + ~dark), and the final max with 0 rides on the same instruction.
__device__ __forceinline__ unsigned score_of_ring(const unsigned (&d)[16], unsigned c, unsigned k1) {
  // The four arcs that start at k .. k+3 share the six pixels k+3 .. k+8, and max(min(S, x), min(S, y)) = min(S, max(x, y)):
  //   max_j min(arc k+j) = min(ring[k+3 .. k+8], max(T[k], min(r[k+1], r[k+2], r[k+9]), min(r[k+2], r[k+9], r[k+10]), T[k+9]))
  // with T[j] = min(r[j], r[j+1], r[j+2]).  Eight T's serve all four groups (T[k+4] is also half of group k's shared
  // part): 34 three-input operations per polarity instead of the 40 of a plain sliding window.
  unsigned tn[16], tx[16];
#pragma unroll
  for (int k = 0; k < 16; k += 4) {
    tn[k] = mn3(d[k], d[k + 1], d[k + 2]);
    tx[k] = mx3(d[k], d[k + 1], d[k + 2]);
    tn[k + 1] = mn3(d[k + 1], d[k + 2], d[k + 3]);
    tx[k + 1] = mx3(d[k + 1], d[k + 2], d[k + 3]);
  }
  unsigned gb[4], gd[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int k = 4 * q;
    const unsigned r1 = d[(k + 1) & 15], r2 = d[(k + 2) & 15], r9 = d[(k + 9) & 15], r10 = d[(k + 10) & 15];
    const unsigned r3 = d[(k + 3) & 15], r7 = d[(k + 7) & 15], r8 = d[(k + 8) & 15];
    const unsigned sideB = mx2(mx3(tn[k], mn3(r1, r2, r9), mn3(r2, r9, r10)), tn[(k + 9) & 15]);
    gb[q] = mn3(tn[(k + 4) & 15], mn3(r3, r7, r8), sideB);
    const unsigned sideD = __vmins2(mn3(tx[k], mx3(r1, r2, r9), mx3(r2, r9, r10)), tx[(k + 9) & 15]);
    gd[q] = mx3(tx[(k + 4) & 15], mx3(r3, r7, r8), sideD);
  }
  const unsigned bright = mx2(mx3(gb[0], gb[1], gb[2]), gb[3]);
  const unsigned dark = __vmins2(mn3(gd[0], gd[1], gd[2]), gd[3]);
  return __viaddmax_s16x2_relu(bright, __vadd2(~c, k1), __vadd2(__vadd2(c, k1), ~dark));
}

// Geometry of one work item = one tile of one image.
struct TileGeom {
  int level, img, X0, Y0, TW, TH, gx0, offL, groups, bx, lead;
};
// (ti, img) = tile index within an image and image of the work item; the caller steps them without dividing
__device__ __forceinline__ TileGeom tile_geom(const LevelTable *T, const int4 *tiles, int ti, int img) {
  TileGeom t;
  t.img = img;
  const int4 tile = tiles[ti];                            // (level, cell row i, first cell column j0, cell count)
  t.level = tile.x;
  const LevelGeom &G = T->lv[t.level];
  // detection rectangle of the tile, level coordinates
  t.X0 = kEdge + tile.z * G.wCell; t.Y0 = kEdge + tile.y * G.hCell;
  const int X1 = min(kEdge + (tile.z + tile.w) * G.wCell, G.w - kEdge);
  const int Y1 = min(t.Y0 + G.hCell, G.h - kEdge);
  t.TW = X1 - t.X0; t.TH = Y1 - t.Y0;
  // 4-pixel groups on the padded plane's 4-byte grid
  t.gx0 = (t.X0 + kEdge) & ~3;                            // padded column of the first group
  t.offL = (t.X0 + kEdge) - t.gx0;                        // tile x = score column - offL
  t.groups = (t.offL + t.TW + 3) >> 2;
  t.bx = (t.gx0 - 4) & ~15;                               // window origin: 16-byte aligned for TMA
  t.lead = (t.gx0 - 4) - t.bx;                            // 0, 4, 8 or 12
  return t;
}

// Persistent CTAs: each walks items blockIdx.x, blockIdx.x + gridDim.x, ...
__global__ void __launch_bounds__(kThreads, ORBGPU_FAST_CTAS_PER_SM)
k_fast_cells(const __grid_constant__ TmaMaps maps, const LevelTable *__restrict__ T, const int4 *__restrict__ tiles,
             int ntiles, int totalItems, int imgBase, Cand *__restrict__ cand, int *__restrict__ candCount,
             unsigned *__restrict__ status, const FastConfig cfg) {
  extern __shared__ uint8_t smemRaw[];
  __shared__ __align__(8) unsigned long long sBar;
  __shared__ int sCellCnt[kFastCellsPerTile];
  __shared__ int sQueued, sKept;

  const int iniTh = T->iniTh, minTh = T->minTh;
  const int tid = threadIdx.x;
  uint8_t *smem = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u);   // TMA destination: 128-byte aligned
  // [u8 window: imgBytes][plane A: 2 x imgBytes][plane B: 2 x imgBytes][scores][queue]; the survivor stage list aliases
  // plane A, which is dead once the margins are computed
  unsigned *sPlaneA = reinterpret_cast<unsigned *>(smem + cfg.imgBytes);
  unsigned *sPlaneB = reinterpret_cast<unsigned *>(smem + 3 * cfg.imgBytes);
  // u8 scores, pitch cfg.scorePitch, column 0 = padded column gx0; one halo row above / below (kept zero) and one halo
  // column on each side (always masked by the cell test) so that NMS reads its 8 neighbours without bounds checks
  uint8_t *sScore = smem + 5 * cfg.imgBytes + cfg.scorePitch + 4;
  unsigned short *sQueue = reinterpret_cast<unsigned short *>(smem + 5 * cfg.imgBytes + cfg.scoreBytes);   // (row << 8 | column) of non-zero scores
  Cand *sStage = reinterpret_cast<Cand *>(sPlaneA);
  const int sp = cfg.scorePitch;

  if (tid == 0) {
    mbar_init(&sBar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < totalItems) {
    const TileGeom n = tile_geom(T, tiles, (int)blockIdx.x % ntiles, imgBase + (int)blockIdx.x / ntiles);
    const LevelGeom &Gn = T->lv[n.level];
    mbar_expect_tx(&sBar, (unsigned)(Gn.fastBoxW * Gn.fastBoxH));
    tma_load_3d(smem, &maps.m[n.level], n.bx, n.Y0 - 3 + kEdge, n.img, &sBar);
  }
  // work item = (tile ti of image imgCur); stepping by gridDim.x items without a division per tile
  int ti = (int)blockIdx.x % ntiles, imgCur = imgBase + (int)blockIdx.x / ntiles;
  const int stepImg = (int)gridDim.x / ntiles, stepTi = (int)gridDim.x - stepImg * ntiles;

  int iter = 0;
  for (int item = blockIdx.x; item < totalItems; item += gridDim.x, iter++) {
    const TileGeom tg = tile_geom(T, tiles, ti, imgCur);
    int tiNext = ti + stepTi, imgNext = imgCur + stepImg;
    if (tiNext >= ntiles) { tiNext -= ntiles; imgNext++; }
    const LevelGeom &G = T->lv[tg.level];
    const int level = tg.level, img = tg.img, X0 = tg.X0, Y0 = tg.Y0, TW = tg.TW, TH = tg.TH;
    const int offL = tg.offL, groups = tg.groups, lead = tg.lead;
    const int bw = G.fastBoxW;
    if (tid == 0) { sQueued = 0; sKept = 0; }
    if (tid < kFastCellsPerTile) sCellCnt[tid] = 0;
    for (int i = tid; i < 2 * (sp >> 2); i += kThreads) {   // zero the halo rows -1 and TH
      const int half = sp >> 2;
      unsigned *rowp = reinterpret_cast<unsigned *>(sScore + (i < half ? -1 : TH) * sp - 4);
      rowp[i < half ? i : i - half] = 0u;
    }
    mbar_wait(&sBar, iter & 1);

    // ---- unpack the u8 window into the two u16 planes (the window's rows are contiguous: word i + 1 follows word i, and the
    // pair that straddles a row end is never read)
    {
      const unsigned *src = reinterpret_cast<const unsigned *>(smem);
      uint2 *dA = reinterpret_cast<uint2 *>(sPlaneA), *dB = reinterpret_cast<uint2 *>(sPlaneB);
      const int nwords = (bw >> 2) * (TH + 6);
      for (int i = tid; i < nwords; i += kThreads) {
        const unsigned w0 = src[i], w1 = src[i + 1];
        dA[i] = make_uint2(__byte_perm(w0, 0u, 0x4140u), __byte_perm(w0, 0u, 0x4342u));
        dB[i] = make_uint2(__byte_perm(w0, 0u, 0x4241u), __byte_perm(w0, w1, 0x4443u) & 0x00ff00ffu);
      }
    }
    __syncthreads();                                       // planes complete; the u8 window and last item's stage list are free
    if (tid == 0) {
      const int next = item + gridDim.x;
      if (next < totalItems) {
        const TileGeom n = tile_geom(T, tiles, tiNext, imgNext);
        const LevelGeom &Gn = T->lv[n.level];
        mbar_expect_tx(&sBar, (unsigned)(Gn.fastBoxW * Gn.fastBoxH));
        tma_load_3d(smem, &maps.m[n.level], n.bx, n.Y0 - 3 + kEdge, n.img, &sBar);
      }
    }

    // ---- arc margins: work items = (row, 4-pixel group) pairs, dealt round-robin to the CTA's threads so that every
    // lane is busy whatever the tile shape; corners at minTh are queued
    const unsigned k1 = ((unsigned)(1 - minTh) & 0xffffu) * 0x00010001u;
    const int items = TH * groups;
    const int stepY = kThreads / groups, stepG = kThreads - stepY * groups;
    const int pw = bw >> 1;                                // plane pitch in 32-bit words
    int y = tid / groups, g = tid - y * groups;
    for (int i0 = 0; i0 < items; i0 += kThreads) {
      unsigned packed = 0;
      if (i0 + tid < items) {
        // word index of the thread's first pixel pair (window column lead + 4g + 4, always even); P = that pair, Q = the next
        const int a = y * pw + ((lead + 4) >> 1) + 2 * g;
        const unsigned *A = sPlaneA + a, *B = sPlaneB + a;
        unsigned dP[16], dQ[16];
        {   // rows y-3 / y+3 (ring 7, 8, 9 / 15, 0, 1): dx = +1, 0, -1
          const uint2 a0 = *reinterpret_cast<const uint2 *>(A), b0 = *reinterpret_cast<const uint2 *>(B);
          const unsigned bm = B[-1];
          dP[8] = a0.x; dQ[8] = a0.y; dP[7] = b0.x; dQ[7] = b0.y; dP[9] = bm; dQ[9] = b0.x;
          const uint2 a6 = *reinterpret_cast<const uint2 *>(A + 6 * pw), b6 = *reinterpret_cast<const uint2 *>(B + 6 * pw);
          const unsigned bm6 = B[6 * pw - 1];
          dP[0] = a6.x; dQ[0] = a6.y; dP[1] = b6.x; dQ[1] = b6.y; dP[15] = bm6; dQ[15] = b6.x;
        }
        {   // rows y-2 / y+2 (ring 6, 10 / 2, 14): dx = +2, -2
          const uint2 a1 = *reinterpret_cast<const uint2 *>(A + pw);
          dP[6] = a1.y; dQ[6] = A[pw + 2]; dP[10] = A[pw - 1]; dQ[10] = a1.x;
          const uint2 a5 = *reinterpret_cast<const uint2 *>(A + 5 * pw);
          dP[2] = a5.y; dQ[2] = A[5 * pw + 2]; dP[14] = A[5 * pw - 1]; dQ[14] = a5.x;
        }
        {   // rows y-1 / y+1 (ring 5, 11 / 3, 13): dx = +3, -3
          const uint2 b2 = *reinterpret_cast<const uint2 *>(B + 2 * pw - 2);
          dP[5] = B[2 * pw + 1]; dQ[5] = B[2 * pw + 2]; dP[11] = b2.x; dQ[11] = b2.y;
          const uint2 b4 = *reinterpret_cast<const uint2 *>(B + 4 * pw - 2);
          dP[3] = B[4 * pw + 1]; dQ[3] = B[4 * pw + 2]; dP[13] = b4.x; dQ[13] = b4.y;
        }
        // row y (ring 4, 12): dx = +3, -3, and the centres
        const uint2 b3 = *reinterpret_cast<const uint2 *>(B + 3 * pw - 2);
        dP[4] = B[3 * pw + 1]; dQ[4] = B[3 * pw + 2]; dP[12] = b3.x; dQ[12] = b3.y;
        const uint2 c = *reinterpret_cast<const uint2 *>(A + 3 * pw);
        // score = max(margin, minTh) - minTh: 0 when the pixel is not a corner at minTh
        const unsigned sA = score_of_ring(dP, c.x, k1);
        const unsigned sB = score_of_ring(dQ, c.y, k1);
        // (the up to 3 + 3 pixels of the first / last group outside the tile's detection columns are scored and queued like
        // the rest: NMS never reads them as neighbours — the cell test masks those columns — and drops their own entries)
        packed = __byte_perm(sA, sB, 0x6420u);
        *reinterpret_cast<unsigned *>(sScore + y * sp + 4 * g) = packed;
      }
      // queue the non-zero scores as (row << 8 | score column); one shared-memory reservation per lane that has any
      if (packed != 0u) {
        const int mine = (packed & 0xffu ? 1 : 0) + (packed & 0xff00u ? 1 : 0) + (packed & 0xff0000u ? 1 : 0) + (packed >> 24 ? 1 : 0);
        int at = atomicAdd(&sQueued, mine);
        const unsigned short e = (unsigned short)((y << 8) | (4 * g));
        if (packed & 0xffu) sQueue[at++] = e;
        if (packed & 0xff00u) sQueue[at++] = e + 1;
        if (packed & 0xff0000u) sQueue[at++] = e + 2;
        if (packed >> 24) sQueue[at] = e + 3;
      }
      y += stepY; g += stepG;
      if (g >= groups) { g -= groups; y++; }
    }
    __syncthreads();

  // ---- NMS inside each cell's rectangle over the queue; survivors go to the stage list, cells count survivors at iniTh
  const int iniS = iniTh - minTh;                         // score > iniS  <=>  margin > iniTh
  const int nq = sQueued;
  for (int i = tid; i < nq; i += kThreads) {
    const int e = sQueue[i];
    const int y = e >> 8, sx = e & 0xff;
    const uint8_t *c0 = sScore + y * sp + sx;
    const int s = c0[0];
    const int x = sx - offL;                               // tile coordinate
    if (x < 0 || x >= TW) continue;                        // group overhang next to the tile
    const int c = (x >= G.wCell) + (x >= 2 * G.wCell) + (x >= 3 * G.wCell);
    const int cx0 = c * G.wCell, cx1 = min(cx0 + G.wCell, TW);
    // neighbours outside the cell's rectangle count as 0: rows -1 / TH are the zero halo, columns are masked
    const bool left = x > cx0, right = x + 1 < cx1;
    const int ml = max(max((int)c0[-1], (int)c0[-sp - 1]), (int)c0[sp - 1]);
    const int mr = max(max((int)c0[1], (int)c0[-sp + 1]), (int)c0[sp + 1]);
    const int m = max(max((int)c0[-sp], (int)c0[sp]), max(left ? ml : 0, right ? mr : 0));
    const bool keep = s > m;
    if (keep) {
      const int slot = atomicAdd(&sKept, 1);
      if (slot < cfg.stageCap) {
        Cand o;
        o.xy = (unsigned)(X0 + x - kMinBorder) | ((unsigned)(Y0 + y - kMinBorder) << 16);
        o.resp = (unsigned)(s + minTh - 1) | ((unsigned)c << 16);     // cv::FAST response = margin - 1; cell in the high half
        sStage[slot] = o;
      }
      if (s > iniS) atomicAdd(&sCellCnt[c], 1);
    }
  }
  __syncthreads();

  // ---- emit: cells with no survivor at iniTh fall back to minTh (ORBextractor.cc:776-780).  Every warp appends its share
  // of the stage list on its own: ballot, one global atomic per warp round, no CTA-wide barrier on the critical path.
  const int nk = min(sKept, cfg.stageCap);
  if (tid == 0 && sKept > cfg.stageCap) atomicOr(status, kStatusCandOverflow);
  {
    const int lane = tid & 31;
    int *counter = &candCount[img * kMaxLevels + level];
    Cand *out = cand + G.candOffset + (size_t)img * G.candCap;
    for (int i0 = (tid >> 5) * 32; i0 < nk; i0 += kThreads) {
      const int i = i0 + lane;
      Cand o;
      bool emit = false;
      if (i < nk) {
        o = sStage[i];
        const int s = (int)(o.resp & 0xffffu) + 1 - minTh, c = (int)(o.resp >> 16);
        emit = s > iniS || sCellCnt[c] == 0;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (mask == 0u) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(counter, __popc(mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (emit) {
        const int slot = base + __popc(mask & ((1u << lane) - 1u));
        o.resp &= 0xffffu;
        if (slot < G.candCap) out[slot] = o; else atomicOr(status, kStatusCandOverflow);
      }
    }
  }
    __syncthreads();   // stage list (= plane A), queue and counters are free again
    ti = tiNext; imgCur = imgNext;
  }
}

}  // namespace

cudaError_t fast_configure(int smemBytes) {
  return cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
}

// Shared-memory budget of the FAST kernel for a level table (whose fastBoxW / fastBoxH are already set).
FastConfig fast_make_config(const LevelTable &T) {
  int wMax = 0, hMax = 0, img = 0;
  for (int l = 0; l < T.nlevels; l++) {
    wMax = max(wMax, T.lv[l].wCell); hMax = max(hMax, T.lv[l].hCell);
    img = max(img, T.lv[l].fastBoxW * T.lv[l].fastBoxH);
  }
  FastConfig c;
  const int tw = kFastCellsPerTile * wMax;
  c.imgBytes = (img + 4 + 127) & ~127;               // + the word the unpack reads past the last one
  c.scorePitch = (tw + 3 + 3 + 3) & ~3;             // + the group overhang on both sides
  c.scoreBytes = (c.scorePitch * (hMax + 2) + 8 + 15) & ~15;   // + halo rows and the 4-byte lead of the halo column
  c.stageCap = (tw / 2 + kFastCellsPerTile) * (hMax / 2 + 1);   // strict 8-neighbour maxima per cell rectangle: ceil(cw / 2) * ceil(ch / 2)
  c.queueCap = c.scorePitch * hMax;
  // window + two u16 planes (the stage list lives in plane A) + scores + queue
  if (c.stageCap * (int)sizeof(Cand) > 2 * c.imgBytes) c.stageCap = 2 * c.imgBytes / (int)sizeof(Cand);
  c.smemBytes = 128 + 5 * c.imgBytes + c.scoreBytes + c.queueCap * 2;
  c.maxIters = ((hMax + kWarps - 1) / kWarps) * (((tw + 6 + 3) / 4 + 31) / 32);
  return c;
}

cudaError_t launch_fast(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &maps, const int4 *tiles, int ntiles,
                        const FastConfig &c, int imgBase, int batch, cudaStream_t s) {
  const int total = ntiles * batch;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const int ctasPerSm = [] { const char *e = getenv("ORBGPU_FAST_CTAS"); const int v = e ? atoi(e) : ORBGPU_FAST_CTAS_PER_SM; return v >= 1 && v <= ORBGPU_FAST_CTAS_PER_SM ? v : ORBGPU_FAST_CTAS_PER_SM; }();
  const int grid = total < ctasPerSm * sms ? total : ctasPerSm * sms;
  k_fast_cells<<<grid, kThreads, c.smemBytes, s>>>(maps, B.dT, tiles, ntiles, total, imgBase, B.cand, B.candCount, B.status, c);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
