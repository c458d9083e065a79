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
// Execution: ONE WARP = ONE STRIP = up to four horizontally adjacent cells of one cell row of one level of one image, at
// most 128 pixels wide; the warp marches down the strip's rows and never meets another warp (no shared-memory pixel
// tiles, no CTA barrier).  A lane owns 4 horizontally adjacent pixels = two packed s16x2 pixel pairs P, Q on the padded
// plane's 4-byte grid and keeps the 7 x 10 pixel neighbourhood of its pixels IN REGISTERS as a rolling window of seven
// rows x nine packed pair words (pairs starting at x-3 .. x+5): every step loads one new row (three coalesced 32-bit
// words per lane, issued one step ahead), unpacks it ONCE with nine PRMT, and all 2 x 16 ring operands of the
// step are then plain registers — the 3-input VIMNMX3.S16x2 arc networks (68 per pixel pair) are the only real work
// left on the ALU pipe.  The loop is unrolled seven-fold so that the window rotates by register renaming.
// Scores (margin - minTh, 0 = not a corner at minTh) stay in registers as well: NMS of row y-1 runs when row y is
// scored, horizontal neighbours come from the adjacent lanes by shuffle, cell borders are per-lane constant masks, and the
// row's decision leaves the loop as one word per lane (the four scores as bytes, zero where the pixel does not survive)
// stored without a branch to the warp's score rows in shared memory.  After the march the lanes collect the non-zero bytes
// into the warp's stage list (4 bytes per survivor), the warp applies the per-cell threshold fallback, claims its range of
// the level's candidate list with one global atomic and writes it.  Strips of four 32-pixel cells start two bytes off the
// 4-byte grid and run in the SHIFTED instantiation of the kernel (own launch, see unpack_row).
// Candidate order in HBM is arbitrary: the octree only needs each candidate's position in the
// reference's cell-major order to break response ties, and that is a function of (x, y).
#include <cstdlib>

#include "internal.cuh"

namespace orbgpu {

namespace {

constexpr int kWarps = ORBGPU_FAST_WARPS;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ unsigned mn3(unsigned a, unsigned b, unsigned c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mx3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ unsigned mx2(unsigned a, unsigned b) { return __vmaxs2(a, b); }
// The ALU pipe (VIMNMX3, LOP3, PRMT, IADD3, SHF: 64 lanes / clk / SM) is this kernel's bound; bitwise NOT and the 24-bit
// right shift are therefore written as integer multiply-adds, which issue on the otherwise idle FMA pipe:
// ~x = x * -1 + -1  and  x >> 24 = hi32(x * 256).  The multipliers arrive as kernel arguments (FmaConsts) — with literal
// constants ptxas folds both back into IADD3 / SHF on the ALU pipe.
struct FmaConsts { unsigned ones, two8, zero; };          // 0xffffffff, 256, 0
__device__ __forceinline__ unsigned not_fma(unsigned x, const FmaConsts &K) { return x * K.ones + K.ones; }
__device__ __forceinline__ unsigned top_byte_fma(unsigned x, const FmaConsts &K) { return __umulhi(x, K.two8); }

// Arc margins (packed s16x2) of two pixels from their 16 ring values d[k] (packed u16x2, ring order of cv::FAST 9/16:
// (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)) and their centres c.
// min / max commute with subtracting the centre, so the sliding-window networks run on the RAW ring pixels and the
// centre is subtracted once from the two extrema:
//   bright margin = max_arcs min_arc(ring) - centre ,  dark margin = centre - min_arcs max_arc(ring).
// The four arcs that start at k .. k+3 share the six pixels k+3 .. k+8, and max(min(S, x), min(S, y)) = min(S, max(x, y)):
//   max_j min(arc k+j) = min(ring[k+3 .. k+8], max(T[k], min(r[k+1], r[k+2], r[k+9]), min(r[k+2], r[k+9], r[k+10]), T[k+9]))
// with T[j] = min(r[j], r[j+1], r[j+2]); eight T's serve all four groups (T[k+4] is also half of group k's shared part):
// 34 three-input operations per polarity.  Returns max(margin - minTh, 0) per halfword; k1 = (1 - minTh) in both
// halves.  With ~v = -v - 1 per halfword, margin - minTh = max(bright + (~c + k1), (c + k1) + ~dark), and the final max
// with 0 rides on the same instruction.
__device__ __forceinline__ unsigned score_of_ring(const unsigned (&d)[16], unsigned c, unsigned k1, const FmaConsts &K) {
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
  return __viaddmax_s16x2_relu(bright, __vadd2(not_fma(c, K), k1), __vadd2(__vadd2(c, k1), not_fma(dark, K)));
}

// ring position k -> (dx, dy), dy positive = down
__host__ __device__ constexpr int ring_dx(int k) { return k <= 3 ? k : k <= 5 ? 3 : k <= 8 ? 8 - k : k <= 11 ? 8 - k : k <= 13 ? -3 : k - 16; }
__host__ __device__ constexpr int ring_dy(int k) { return k <= 1 ? 3 : k <= 4 ? 4 - k : k <= 7 ? 4 - k : k <= 9 ? -3 : k <= 12 ? k - 12 : k <= 14 ? k - 12 : 3; }
static_assert(ring_dx(0) == 0 && ring_dx(1) == 1 && ring_dx(2) == 2 && ring_dx(3) == 3 && ring_dx(4) == 3 && ring_dx(5) == 3 && ring_dx(6) == 2 &&
              ring_dx(7) == 1 && ring_dx(8) == 0 && ring_dx(9) == -1 && ring_dx(10) == -2 && ring_dx(11) == -3 && ring_dx(12) == -3 &&
              ring_dx(13) == -3 && ring_dx(14) == -2 && ring_dx(15) == -1, "cv::FAST 9/16 ring, x offsets");
static_assert(ring_dy(0) == 3 && ring_dy(1) == 3 && ring_dy(2) == 2 && ring_dy(3) == 1 && ring_dy(4) == 0 && ring_dy(5) == -1 && ring_dy(6) == -2 &&
              ring_dy(7) == -3 && ring_dy(8) == -3 && ring_dy(9) == -3 && ring_dy(10) == -2 && ring_dy(11) == -1 && ring_dy(12) == 0 &&
              ring_dy(13) == 1 && ring_dy(14) == 2 && ring_dy(15) == 3, "cv::FAST 9/16 ring, y offsets");

// Bytes x-4 .. x+7 (x+11 in the shifted form) of one row around a lane's 4-byte group at padded column x
struct Raw { unsigned w0, w1, w2, w3; };

template <bool kShift>
__device__ __forceinline__ Raw load_row(const uint8_t *p) {
  Raw r;
  r.w0 = __ldg(reinterpret_cast<const unsigned *>(p) - 1);
  r.w1 = __ldg(reinterpret_cast<const unsigned *>(p));
  r.w2 = __ldg(reinterpret_cast<const unsigned *>(p) + 1);
  r.w3 = kShift ? __ldg(reinterpret_cast<const unsigned *>(p) + 2) : 0u;
  return r;
}
// The nine packed u16x2 pixel pairs (x'+s, x'+s+1), s = -3 .. 5, of one row: win[s + 3].  x' = the lane's first pixel: byte
// x of the row in the plain form, byte x + 2 in the shifted form (a strip of four 32-pixel cells starts two bytes off the
// plane's 4-byte grid and would otherwise need a 33rd lane; it costs one more word per row and lane).  Pairs that straddle a
// word take their zero bytes from a shifted copy made on the FMA pipe.
template <bool kShift>
__device__ __forceinline__ void unpack_row(const Raw &r, unsigned (&win)[9], const FmaConsts &K) {
  if (!kShift) {
    win[0] = __byte_perm(r.w0, 0u, 0x4241u);
    win[1] = __byte_perm(r.w0, 0u, 0x4342u);
    win[2] = __byte_perm(top_byte_fma(r.w0, K), r.w1, 0x1410u);         // (w0.b3, w1.b0): the shifted word supplies the zero bytes
    win[3] = __byte_perm(r.w1, 0u, 0x4140u);
    win[4] = __byte_perm(r.w1, 0u, 0x4241u);
    win[5] = __byte_perm(r.w1, 0u, 0x4342u);
    win[6] = __byte_perm(top_byte_fma(r.w1, K), r.w2, 0x1410u);
    win[7] = __byte_perm(r.w2, 0u, 0x4140u);
    win[8] = __byte_perm(r.w2, 0u, 0x4241u);
  } else {
    win[0] = __byte_perm(top_byte_fma(r.w0, K), r.w1, 0x1410u);         // bytes x-1, x
    win[1] = __byte_perm(r.w1, 0u, 0x4140u);
    win[2] = __byte_perm(r.w1, 0u, 0x4241u);
    win[3] = __byte_perm(r.w1, 0u, 0x4342u);                            // the lane's own first pair: bytes x+2, x+3
    win[4] = __byte_perm(top_byte_fma(r.w1, K), r.w2, 0x1410u);
    win[5] = __byte_perm(r.w2, 0u, 0x4140u);
    win[6] = __byte_perm(r.w2, 0u, 0x4241u);
    win[7] = __byte_perm(r.w2, 0u, 0x4342u);
    win[8] = __byte_perm(top_byte_fma(r.w2, K), r.w3, 0x1410u);         // bytes x+7, x+8
  }
}

// Pulls the line(s) of a later row towards L1 without tying up a register: the three LDGs of that row then hit.  Rows past
// the strip still lie inside the padded plane (19 apron rows follow the last detection row).
constexpr int kPrefetchRows = ORBGPU_FAST_PREFETCH;
constexpr int kAhead = ORBGPU_FAST_AHEAD;             // rows between a row's register loads and its first use (1 .. 5)
__device__ __forceinline__ void prefetch_l1(const uint8_t *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ int atom_add_shared(int *p, int v) {   // plain ATOMS (keeps nvcc from expanding it into a warp scan)
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}

// NMS state a lane carries between rows: scores of the previous row, its left / right neighbour maximum, and the
// 3-wide row maxima of the previous two rows
struct NmsState { unsigned sP, sQ, sideP, sideQ, h3P, h3Q, h3ppP, h3ppQ; };
struct LaneMasks { unsigned mLP, mRP, mLQ, mRQ, vP, vQ; };

#if !ORBGPU_FAST_DEFER
// Decides row `yPrev` (scores in st.s*) now that the row below it is known (h3 of that row in h3P / h3Q; pass 0 below
// the strip), and appends the survivors to the warp's stage list as  x | y << 7 | score << 13.  About 2 % of the pixels
// survive, so the append is a short loop over the lane's set bits behind one rarely taken branch.
__device__ __forceinline__ void nms_row(const NmsState &st, unsigned h3P, unsigned h3Q, const LaneMasks &M, int x0, int yPrev,
                                        unsigned *stage, int *stageCount, int stageCap, const FmaConsts &K) {
  const unsigned nbP = mx3(st.h3ppP, st.sideP, h3P), nbQ = mx3(st.h3ppQ, st.sideQ, h3Q);
  // s + ~nb = s - nb - 1 per halfword: sign clear  <=>  s > nb (and then s > 0)
  const unsigned kP = ~__vadd2(st.sP, not_fma(nbP, K)) & M.vP, kQ = ~__vadd2(st.sQ, not_fma(nbQ, K)) & M.vQ;   // bit 15 / 31 per surviving pixel
  if ((kP | kQ) != 0u) {
    unsigned bits = ((kP >> 15) & 1u) | ((kP >> 30) & 2u) | ((kQ >> 13) & 4u) | ((kQ >> 28) & 8u);
    const unsigned row = (unsigned)yPrev << 7;
    do {
      const int j = __ffs((int)bits) - 1;
      bits &= bits - 1u;
      const unsigned w = (j & 2) ? st.sQ : st.sP;
      const unsigned s = (j & 1) ? (w >> 16) : (w & 0xffffu);
      const int at = atom_add_shared(stageCount, 1);
      if (at < stageCap) stage[at] = (unsigned)(x0 + j) | row | (s << 13);
    } while (bits != 0u);
  }
}
#endif

#if ORBGPU_FAST_DEFER
// Deferred form: the row's decision leaves the loop as ONE word per lane — the lane's four scores as bytes (score <= 255 -
// minTh), zero where the pixel does not survive — stored without a branch to the warp's score rows in shared memory; the
// survivors (about 2 % of the pixels, yet 9 of 10 warp rows have one) are collected after the march by a sweep that only
// visits the non-zero words, instead of a divergent append behind a branch that is taken in nearly every step.
__device__ __forceinline__ unsigned nms_row_scores(const NmsState &st, unsigned h3P, unsigned h3Q, const LaneMasks &M, const FmaConsts &K) {
  const unsigned nbP = mx3(st.h3ppP, st.sideP, h3P), nbQ = mx3(st.h3ppQ, st.sideQ, h3Q);
  const unsigned kP = ~__vadd2(st.sP, not_fma(nbP, K)) & M.vP, kQ = ~__vadd2(st.sQ, not_fma(nbQ, K)) & M.vQ;   // bit 15 / 31 per surviving pixel
  unsigned keep;                                                        // sign of bytes 1, 3, 5, 7 replicated: 0xff per surviving pixel
  asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(keep) : "r"(kP), "r"(kQ));     // (__byte_perm drops the selector's replicate bits)
  return __byte_perm(st.sP, st.sQ, 0x6420u) & keep;                     // the four scores' low bytes
}
#endif

// The march of one warp down its strip: rows Y0 - 3 .. Y1 + 2 through the rolling register window, scores, NMS, survivors to the stage list
// (deferred form: to the lane's column of the score rows, row y at rowsBuf[(y + 1) * 32]; row 0 takes the first step's empty decision).
template <bool kShift>
__device__ __forceinline__ void march_strip(const uint8_t *rowPtr, int pitch, int pfOff, int TH, const LaneMasks &M, int x0, unsigned k1,
                                            unsigned *stage, int *stageCount, int stageCap, unsigned *rowsBuf, const FmaConsts &K) {
  unsigned win[7][9];                                // rolling window: row r of the strip's source rows lives in slot r % 7
  NmsState st = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  const int nRows = TH + 6;                          // source rows Y0 - 3 .. Y1 + 2
#pragma unroll
  for (int i = 0; i < kPrefetchRows; i++) prefetch_l1(rowPtr + (kAhead + i) * pitch + pfOff);
  // raw rows in flight: row r is unpacked from raw[r % 7] while row r + kAhead is being loaded into raw[(r + kAhead) % 7] (a
  // ring indexed like the window, so no register is ever copied from a load's destination before its data is needed)
  Raw raw[7];
#pragma unroll
  for (int i = 0; i < kAhead; i++) raw[i] = load_row<kShift>(rowPtr + i * pitch);
  for (int r0 = 0; r0 < nRows; r0 += 7) {
#pragma unroll
    for (int k = 0; k < 7; k++) {
      const int r = r0 + k;
      if (r < nRows) {
        // The address of the row requested now is tied to the DATA of the row unpacked now (dep = w0 * 0, the zero a kernel argument
        // so that it survives ptxas).  ptxas tags the three loads of a row with one scoreboard and reuses it every other step; with
        // the new loads hoisted above the unpack, the unpack's wait for row r also waited for the loads just issued — a full memory
        // round trip in three of the seven unrolled steps, 15 % of the kernel's stall samples (profiles/r2_fast_stall_sites.txt).
        // With the dependence the request cannot be issued before row r has arrived (it has: it was requested a step ago).
        const unsigned dep = raw[k].w0 * K.zero;
        rowPtr += pitch + (int)dep;                                              // row r + 1
        raw[(k + kAhead) % 7] = load_row<kShift>(rowPtr + (kAhead - 1) * pitch);   // row r + kAhead (always inside the padded plane)
        prefetch_l1(rowPtr + (kAhead - 1 + kPrefetchRows) * pitch + pfOff);      // and later rows on their way from HBM / L2 into L1
        unpack_row<kShift>(raw[k], win[k], K);
        if (r >= 6) {
          // pixel row y = r - 6: its ring rows y-3 .. y+3 are the source rows r-6 .. r = slots (k + 1 + i) % 7
          unsigned dP[16], dQ[16];
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int slot = (k + 4 + ring_dy(i)) % 7;
            dP[i] = win[slot][ring_dx(i) + 3];
            dQ[i] = win[slot][ring_dx(i) + 5];
          }
          const unsigned sP = score_of_ring(dP, win[(k + 4) % 7][3], k1, K);
          const unsigned sQ = score_of_ring(dQ, win[(k + 4) % 7][5], k1, K);
          // horizontal neighbours: (x-1, x) / (x+1, x+2) / (x+3, x+4) as packed pairs, cell borders masked out
          const unsigned prevQ = __shfl_up_sync(0xffffffffu, sQ, 1), nextP = __shfl_down_sync(0xffffffffu, sP, 1);
          const unsigned mid = __byte_perm(sP, sQ, 0x5432u);
          const unsigned LP = __byte_perm(prevQ, sP, 0x5432u) & M.mLP, RP = mid & M.mRP;
          const unsigned LQ = mid & M.mLQ, RQ = __byte_perm(sQ, nextP, 0x5432u) & M.mRQ;
          const unsigned h3P = mx3(LP, sP, RP), h3Q = mx3(LQ, sQ, RQ);
#if ORBGPU_FAST_DEFER
          rowsBuf[(r - 6) * 32] = nms_row_scores(st, h3P, h3Q, M, K);              // row y - 1 (all zero for y = 0: the spare row)
#else
          nms_row(st, h3P, h3Q, M, x0, r - 7, stage, stageCount, stageCap, K);     // row y - 1 (nothing survives for y = 0)
#endif
          st.h3ppP = st.h3P; st.h3ppQ = st.h3Q; st.h3P = h3P; st.h3Q = h3Q;
          st.sP = sP; st.sQ = sQ; st.sideP = mx2(LP, RP); st.sideQ = mx2(LQ, RQ);
        }
      }
    }
  }
#if ORBGPU_FAST_DEFER
  rowsBuf[TH * 32] = nms_row_scores(st, 0u, 0u, M, K);                             // last row: nothing below it
  // sweep: a bit per non-zero word of the lane's column, then only those words are opened
  for (int h = 0; h < TH; h += 32) {
    const int n = min(32, TH - h);
    unsigned m = 0u;
    for (int y = 0; y < n; y++) m |= (rowsBuf[(h + y + 1) * 32] != 0u ? 1u : 0u) << y;
    while (m != 0u) {
      const int y = h + __ffs((int)m) - 1;
      m &= m - 1u;
      const unsigned v = rowsBuf[(y + 1) * 32];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned sc = (v >> (8 * j)) & 0xffu;
        if (sc != 0u) {
          const int at = atom_add_shared(stageCount, 1);
          if (at < stageCap) stage[at] = (unsigned)(x0 + j) | ((unsigned)y << 7) | (sc << 13);
        }
      }
    }
  }
#else
  nms_row(st, 0u, 0u, M, x0, TH - 1, stage, stageCount, stageCap, K);              // last row: nothing below it
#endif
}

// kShift: the launch's strips all run in the shifted form (fast_strip_shifted; the host keeps the two kinds in separate
// launches so that each kernel carries one copy of the unrolled march — both copies in one kernel thrash the instruction cache)
template <bool kShift>
__global__ void __launch_bounds__(kThreads, ORBGPU_FAST_CTAS_PER_SM)
k_fast_strips(const LevelTable *__restrict__ T, const int4 *__restrict__ strips, int nstrips, int totalItems, int imgBase,
              const uint8_t *__restrict__ pyr, Cand *__restrict__ cand, int *__restrict__ candCount, unsigned *__restrict__ status,
              int stageCap, int rowsCap, const FmaConsts K) {
  extern __shared__ unsigned sStageAll[];            // [kWarps][stageCap], then [kWarps][rowsCap][32] score rows (deferred form)
  __shared__ int sCount[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarps + warp;
  if (item >= totalItems) return;                    // warps are independent: no CTA barrier anywhere below
  const int imgRel = item / nstrips, img = imgBase + imgRel;
  const int4 strip = strips[item - imgRel * nstrips];     // (level, cell row i, first cell column j0, cell count)
  const int level = strip.x;
  const LevelGeom &G = T->lv[level];
  const int iniTh = T->iniTh, minTh = T->minTh;
  // detection rectangle of the strip, level coordinates
  const int X0 = kEdge + strip.z * G.wCell, Y0 = kEdge + strip.y * G.hCell;
  const int X1 = min(kEdge + (strip.z + strip.w) * G.wCell, G.w - kEdge), Y1 = min(Y0 + G.hCell, G.h - kEdge);
  const int TW = X1 - X0, TH = Y1 - Y0;
  const int gx0 = (X0 + kEdge) & ~3;                 // padded column of lane 0's 4-byte group
  // strip x = 4 * lane + j - offL.  A strip that would need a 33rd lane (four 32-pixel cells start two bytes off the 4-byte
  // grid) runs in the shifted form: every lane's pixels start at byte 2 of its group (one more word per row and lane)
  const int offL = (X0 + kEdge) - gx0 - (kShift ? 2 : 0);
  const int wCell = G.wCell;
  unsigned *stage = sStageAll + warp * stageCap;
  unsigned *rowsBuf = sStageAll + kWarps * stageCap + warp * rowsCap * 32 + lane;
  int *stageCount = &sCount[warp];
  if (lane == 0) *stageCount = 0;
  __syncwarp();

  // per-lane constants: which of the lane's four pixels lie in the strip, their cell, and whether their left / right
  // neighbour column belongs to the same cell rectangle
  const int x0 = 4 * lane - offL;
  LaneMasks M = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = x0 + j;
    const bool valid = x >= 0 && x < TW;
    const int c = (x >= wCell) + (x >= 2 * wCell) + (x >= 3 * wCell);
    const int cx0 = c * wCell, cx1 = min(cx0 + wCell, TW);
    const unsigned half = 0xffffu << (16 * (j & 1));
    const bool l = valid && x > cx0, r = valid && x + 1 < cx1;
    if (j < 2) { M.mLP |= l ? half : 0u; M.mRP |= r ? half : 0u; M.vP |= valid ? (half & 0x80008000u) : 0u; }
    else       { M.mLQ |= l ? half : 0u; M.mRQ |= r ? half : 0u; M.vQ |= valid ? (half & 0x80008000u) : 0u; }
  }

  const unsigned k1 = ((unsigned)(1 - minTh) & 0xffffu) * 0x00010001u;
  const int iniS = iniTh - minTh;                    // score > iniS  <=>  margin > iniTh
  const int pitch = G.pitch;
  const uint8_t *rowPtr = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)(Y0 - 3 + kEdge) * pitch + gx0 + 4 * lane;
  // prefetch addresses cover the whole 136 / 140-byte span the warp reads per row (lane 0 its left word, lane 31 its right word)
  const int pfOff = lane == 31 ? (kShift ? 8 : 4) : -4;

  march_strip<kShift>(rowPtr, pitch, pfOff, TH, M, x0, k1, stage, stageCount, stageCap, rowsBuf, K);

  // ---- emit: cells with no survivor at iniTh fall back to minTh (ORBextractor.cc:776-780)
  __syncwarp();
  const int kept = *stageCount;
  if (kept > stageCap && lane == 0) atomicOr(status, kStatusCandOverflow);
  const int nk = min(kept, stageCap);
  unsigned iniMask = 0u;                             // bit c: cell c of the strip has a survivor above iniTh
  for (int i = lane; i < nk; i += 32) {
    const unsigned e = stage[i];
    const int x = (int)(e & 127u);
    if ((int)((e >> 13) & 0xffu) > iniS) iniMask |= 1u << ((x >= wCell) + (x >= 2 * wCell) + (x >= 3 * wCell));
  }
  iniMask = __reduce_or_sync(0xffffffffu, iniMask);
  int *counter = &candCount[img * kMaxLevels + level];
  Cand *out = cand + G.candOffset + (size_t)img * G.candCap;
  // one global atomic per STRIP: the entries that qualify are counted first, the strip's range of the level's list is claimed
  // once, and a second pass writes (an atomic per 32 entries made every pass wait for its own round trip to L2)
  auto qualifies = [&](unsigned e) {
    const int x = (int)(e & 127u), c = (x >= wCell) + (x >= 2 * wCell) + (x >= 3 * wCell);
    return (int)((e >> 13) & 0xffu) > iniS || !((iniMask >> c) & 1u);
  };
  int mine = 0;
  for (int i = lane; i < nk; i += 32) mine += qualifies(stage[i]) ? 1 : 0;
  const int total = __reduce_add_sync(0xffffffffu, mine);
  if (total == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, total);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int i0 = 0; i0 < nk; i0 += 32) {
    const int i = i0 + lane;
    unsigned e = 0u;
    bool emit = false;
    if (i < nk) { e = stage[i]; emit = qualifies(e); }
    const unsigned mask = __ballot_sync(0xffffffffu, emit);
    if (emit) {
      const int slot = base + __popc(mask & ((1u << lane) - 1u));
      Cand o;
      o.xy = (unsigned)(X0 + (int)(e & 127u) - kMinBorder) | ((unsigned)(Y0 + (int)((e >> 7) & 63u) - kMinBorder) << 16);
      o.resp = ((e >> 13) & 0xffu) + (unsigned)minTh - 1u;             // cv::FAST response = margin - 1
      if (slot < G.candCap) out[slot] = o; else atomicOr(status, kStatusCandOverflow);
    }
    base += __popc(mask);
  }
}

}  // namespace

// Cells per strip of a level: as many (up to 4) as fit 128 pixels.  A strip starts 0 .. 3 bytes off the plane's 4-byte grid
// (padded column 38 + j * wCell); the shifted mode removes two of them, so at most one lane-pixel is lost: none when wCell
// is even (then the offset is even), one when it is odd.
bool fast_strip_shifted(const LevelGeom &G, int firstCell, int cells) {
  const int X0 = kEdge + firstCell * G.wCell, X1 = min(kEdge + (firstCell + cells) * G.wCell, G.w - kEdge);
  const int off = (X0 + kEdge) & 3;
  return ORBGPU_FAST_SHIFT && off >= 2 && X1 - X0 + off > 128;
}

int fast_cells_per_strip(int wCell) {
#if ORBGPU_FAST_SHIFT
  return min(4, max(1, (128 - (wCell & 1)) / wCell));
#else
  return wCell <= 31 ? 4 : wCell <= 41 ? 3 : wCell <= 62 ? 2 : 1;
#endif
}

// Stage-list entries per warp: NMS survivors are strict 8-neighbour maxima inside one cell rectangle, at most
// ceil(cw / 2) * ceil(ch / 2) per cell.
FastConfig fast_make_config(const LevelTable &T) {
  FastConfig c{};
  int cap = 64;
  for (int l = 0; l < T.nlevels; l++) {
    const LevelGeom &G = T.lv[l];
    cap = max(cap, fast_cells_per_strip(G.wCell) * ((G.wCell + 1) / 2) * ((G.hCell + 1) / 2));
  }
  c.stageCap = cap;
  int rows = 0;
#if ORBGPU_FAST_DEFER
  for (int l = 0; l < T.nlevels; l++) rows = max(rows, T.lv[l].hCell + 1);      // score rows: one per detection row + the spare row 0
#endif
  c.rowsCap = rows;
  c.smemBytes = kWarps * (cap + rows * 32) * (int)sizeof(unsigned);
  return c;
}

cudaError_t fast_configure(int smemBytes) {
  const cudaError_t e = cudaFuncSetAttribute(k_fast_strips<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
  return e != cudaSuccess ? e : cudaFuncSetAttribute(k_fast_strips<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
}

// strips[0 .. nPlain) run in the plain form, strips[nPlain .. nPlain + nShift) in the shifted form (fast_strip_shifted)
// (on stream sShift: frame-at-a-time calls run the two launches side by side, a march is ~8 us long whatever the batch)
cudaError_t launch_fast(const LevelTable &T, const ExtractBuffers &B, const int4 *strips, int nPlain, int nShift, const FastConfig &c,
                        int imgBase, int batch, cudaStream_t s, cudaStream_t sShift) {
  const FmaConsts K{0xffffffffu, 256u, 0u};
  if (nPlain > 0) {
    const int total = nPlain * batch;
    k_fast_strips<false><<<(total + kWarps - 1) / kWarps, kThreads, c.smemBytes, s>>>(B.dT, strips, nPlain, total, imgBase, B.pyr, B.cand, B.candCount,
                                                                                        B.status, c.stageCap, c.rowsCap, K);
    count_launch();
  }
  if (nShift > 0) {
    const int total = nShift * batch;
    k_fast_strips<true><<<(total + kWarps - 1) / kWarps, kThreads, c.smemBytes, sShift>>>(B.dT, strips + nPlain, nShift, total, imgBase, B.pyr, B.cand, B.candCount,
                                                                                       B.status, c.stageCap, c.rowsCap, K);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace orbgpu
