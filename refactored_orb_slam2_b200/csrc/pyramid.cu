// Image pyramid: ORBextractor::ComputePyramid (ORBextractor.cc:1041-1065).
//
// Every level lives in HBM as a [batch] array of padded planes: interior w x h surrounded by the
// reference's 19 px BORDER_REFLECT_101 apron, row pitch a multiple of 16 bytes.  The import and the
// generic resize write whole padded planes in one pass (every padded byte mapped through reflect-101
// to the interior pixel it mirrors); the tiled resize writes the interior rows with their side aprons
// and one small copy kernel adds the apron rows of all levels at the end.
//
// Level l >= 1 is cv::resize(level l-1, INTER_LINEAR) in OpenCV's 11-bit fixed point:
//   H(x)   = S[xo]*a0 + S[xo+1]*a1                      (int32, a0+a1 = 2048)
//   out    = ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// with per-axis (offset, coefficient) tables built on the host exactly like OpenCV builds them.
#include <algorithm>
#include <cmath>

#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

// BORDER_REFLECT_101 index: -k -> k, n-1+k -> n-1-k.  Excursions are at most kEdge (19) beyond the interior and every
// level is wider than that (the cell grid needs >= 62 pixels), so a single fold suffices.
__device__ __forceinline__ int reflect101(int p, int n) {
  p = p < 0 ? -p : p;
  return p >= n ? 2 * (n - 1) - p : p;
}

// Level 0 = the input image inside its apron.  One thread = 16 padded bytes = one 128-bit store.  Two launches of one kernel:
// BORDER = false covers the chunks that hold interior pixels only — their 16 source bytes come from five aligned 32-bit words
// re-aligned by funnel shifts (the source rows of a 1241-pixel image start on any byte); BORDER = true covers the few chunks
// per row that touch the left / right apron (thread = (row, border chunk), every lane on the byte path with the reflect-101
// index) so that no warp of the bulk copy ever executes the byte loop.
template <bool BORDER>
__global__ void __launch_bounds__(256)
k_import_level0(const __grid_constant__ LevelGeom G, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ src,
                size_t rowStep, size_t imgStep, int imgBase, int c0, int c1) {   // interior-only chunks are [c0, c1)
  const int img = imgBase + blockIdx.z;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  const int nchunks = G.pitch >> 4;
  if (BORDER) { if (chunk >= c0) chunk += c1 - c0; } else { chunk += c0; }
  if (py >= G.rows || chunk >= (BORDER ? nchunks : c1)) return;
  const int px16 = chunk * 16;
  const uint8_t *S = src + (size_t)img * imgStep + (size_t)reflect101(py - kEdge, G.h) * rowStep;
  uint4 out;
  if (!BORDER) {
    const uint8_t *p = S + (px16 - kEdge);
    const unsigned sh = (unsigned)((uintptr_t)p & 3) * 8;
    const unsigned *q = reinterpret_cast<const unsigned *>(p - ((uintptr_t)p & 3));
    const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);   // q + 4 holds pixels of this row (see c1)
    out = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
  } else {
    unsigned o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int px = px16 + k;
      if (px < G.w + 2 * kEdge) o[k >> 2] |= (unsigned)S[reflect101(px - kEdge, G.w)] << (8 * (k & 3));
    }
    out = make_uint4(o[0], o[1], o[2], o[3]);
  }
  uint8_t *D = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)py * G.pitch + px16;
  *reinterpret_cast<uint4 *>(D) = out;
}

__global__ void __launch_bounds__(256)
k_resize_level(const LevelTable *__restrict__ T, uint8_t *__restrict__ pyr, const int2 *__restrict__ tabs,
               int level, int imgBase) {
  const LevelGeom &G = T->lv[level];
  const LevelGeom &P = T->lv[level - 1];
  const int img = imgBase + blockIdx.z;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  const int px4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (py >= G.rows || px4 >= G.pitch) return;
  const int iy = reflect101(py - kEdge, G.h);
  const int2 ty = tabs[G.yTab + iy];           // .x = source row, .y = b0 | b1 << 16
  const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
  const uint8_t *S0 = pyr + P.pyrOffset + (size_t)img * P.planeBytes + (size_t)(ty.x + kEdge) * P.pitch + kEdge;
  const uint8_t *S1 = S0 + (ty.x + 1 < P.h ? P.pitch : 0);
  unsigned out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int px = px4 + k;
    unsigned v = 0;
    if (px < G.w + 2 * kEdge) {
      const int2 tx = tabs[G.xTab + reflect101(px - kEdge, G.w)];
      const int a0 = tx.y & 0xffff, a1 = tx.y >> 16;
      const int x0 = tx.x, x1 = min(tx.x + 1, P.w - 1);
      const int h0 = S0[x0] * a0 + S0[x1] * a1;
      const int h1 = S1[x0] * a0 + S1[x1] * a1;
      v = (unsigned)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
    out |= v << (8 * k);
  }
  uint8_t *D = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)py * G.pitch + px4;
  *reinterpret_cast<unsigned *>(D) = out;
}

// ---- level l >= 1, fast path: one WARP per 128 x kRH tile of the PADDED output plane, source window staged by TMA --------
// A lane owns 4 adjacent output columns (one aligned 32-bit store per row) and walks down the tile's rows.  Columns are
// padded-plane columns: a lane in the left / right apron maps its columns through reflect-101 to the interior pixel they
// mirror, i.e. reads the same source pixels as that pixel, so the kernel writes the 19-pixel side aprons itself; only the
// apron ROWS above / below the interior are left to k_fill_rows.
// Horizontal pass of one source row: three aligned 32-bit shared-memory loads per lane cover the <= 8 source bytes its four
// pixels need, two funnel shifts bring them to byte 0, and per pixel one PRMT (lane-constant selector) puts its two source
// bytes next to each other for one DP2A with the packed (a0, a1) weights: H = S[x0] a0 + S[x0 + 1] a1.  The result is kept
// as H >> 4 and reused by the next output row whenever both read the source row (4 of 5 rows at scale 1.2).
// Vertical pass: ((b0 (H0 >> 4)) >> 16) is the high word of (b0 << 16) * (H0 >> 4): one IMAD.HI per term.
namespace {
#ifndef ORBGPU_RESIZE_RH
#define ORBGPU_RESIZE_RH 24
#endif
#ifndef ORBGPU_RESIZE_WARPS
#define ORBGPU_RESIZE_WARPS 4
#endif
constexpr int kRW = 128, kRH = ORBGPU_RESIZE_RH, kResizeWarps = ORBGPU_RESIZE_WARPS;
static_assert(kRH <= 32, "the tile's row table is held one entry per lane");

__global__ void __launch_bounds__(kResizeWarps * 32)
k_resize_tiles(const __grid_constant__ TmaMaps maps, const LevelTable *__restrict__ T, const int2 *__restrict__ tabs,
               const int4 *__restrict__ tiles, int ntiles, uint8_t *__restrict__ pyr, int slotBytes, int imgBase) {
  extern __shared__ uint8_t smemRaw[];
  __shared__ __align__(8) unsigned long long sBar[kResizeWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * kResizeWarps + warp;
  if (t >= ntiles) return;
  const int4 tile = tiles[t];                    // (level, tile x, tile y, 0)
  const int level = tile.x;
  const LevelGeom &G = T->lv[level];
  const int img = imgBase + blockIdx.y;
  const int xp = kRW * tile.y + 4 * lane;        // padded column of this lane's first pixel
  const int y0 = kRH * tile.z;                   // interior row of the tile's first output row
  const int rows = min(kRH, G.h - y0);
  // column tables (source x, a0 | a1 << 16) of the lane's four pixels, and the tile's row table (one row per lane)
  int sx[4], wA[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int2 tx = tabs[G.xTab + reflect101(min(xp + p, G.w + 2 * kEdge - 1) - kEdge, G.w)];
    sx[p] = tx.x; wA[p] = tx.y;
  }
  const int2 myRow = tabs[G.yTab + min(y0 + lane, G.h - 1)];
  const int sxLane = min(min(sx[0], sx[1]), min(sx[2], sx[3]));
  const int sxMin = __reduce_min_sync(0xffffffffu, sxLane);
  const int syMin = __shfl_sync(0xffffffffu, myRow.x, 0);
  const int bx = (sxMin + kEdge) & ~15;          // TMA: innermost coordinate on a 16-byte boundary
  uint8_t *slot = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u) + warp * slotBytes;
  const int bw = G.rzBoxW;
  if (lane == 0) {
    mbar_init(&sBar[warp], 1);
    mbar_fence_init();
    mbar_expect_tx(&sBar[warp], (unsigned)(G.rzBoxW * G.rzBoxH));
    tma_load_3d(slot, &maps.m[level - 1], bx, syMin + kEdge, img, &sBar[warp]);
  }
  // per-lane constants of the horizontal pass: byte offset of the lane's first source pixel in a window row, and per pixel the
  // PRMT selector that moves its two source bytes (offsets o, o + 1 from that first pixel) into the low half word
  const int off0 = sxLane + kEdge - bx;
  const unsigned shift = (unsigned)(off0 & 3) * 8;
  unsigned sel[4];
#pragma unroll
  for (int p = 0; p < 4; p++) { const unsigned o = (unsigned)(sx[p] - sxLane); sel[p] = o | ((o + 1) << 4); }
  const unsigned *rowWords = reinterpret_cast<const unsigned *>(slot) + (off0 >> 2);
  const int bwWords = bw >> 2;
  const bool colOk = xp < G.pitch;
  uint8_t *dst = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)(y0 + kEdge) * G.pitch + xp;
  const size_t pitch = (size_t)G.pitch;
  __syncwarp();
  mbar_wait(&sBar[warp], 0);

  auto hpass = [&](int srow, unsigned (&h)[4]) {   // H >> 4 of the lane's four pixels on window row `srow`
    const unsigned *w = rowWords + srow * bwWords;
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
    const unsigned a0 = __funnelshift_r(w0, w1, shift), a1 = __funnelshift_r(w1, w2, shift);
#pragma unroll
    for (int p = 0; p < 4; p++) h[p] = __dp2a_lo((unsigned)wA[p], __byte_perm(a0, a1, sel[p]), 0u) >> 4;
  };
  unsigned hLo[4], hHi[4];
  int cur = -2;                                    // hLo / hHi hold window rows cur, cur + 1
  for (int r = 0; r < rows; r++) {
    const int s0 = __shfl_sync(0xffffffffu, myRow.x, r) - syMin;      // row s0 + 1 has weight 0 where the reference clamps it
    const unsigned bb = (unsigned)__shfl_sync(0xffffffffu, myRow.y, r);
    const unsigned b0 = bb << 16, b1 = bb & 0xffff0000u;              // (b0 << 16), (b1 << 16)
    if (s0 != cur) {
      if (s0 == cur + 1) {
#pragma unroll
        for (int p = 0; p < 4; p++) hLo[p] = hHi[p];
      } else {
        hpass(s0, hLo);
      }
      hpass(s0 + 1, hHi);
      cur = s0;
    }
    unsigned v[4];
#pragma unroll
    for (int p = 0; p < 4; p++) v[p] = (__umulhi(b0, hLo[p]) + __umulhi(b1, hHi[p]) + 2u) >> 2;
    const unsigned out = (v[1] * 256u + v[0]) + (v[3] * 256u + v[2]) * 65536u;
    if (colOk) *reinterpret_cast<unsigned *>(dst + (size_t)r * pitch) = out;
  }
}

// BORDER_REFLECT_101 apron ROWS of levels firstLevel .. nlevels-1 (ORBextractor.cc:1057-1059): the 19 rows above and below the
// interior are copies of interior rows, side aprons included (k_resize_tiles wrote those), one 128-bit word per thread.
__global__ void __launch_bounds__(256)
k_fill_rows(const LevelTable *__restrict__ T, uint8_t *__restrict__ pyr, int firstLevel, int imgBase) {
  const int level = firstLevel + blockIdx.y, img = imgBase + blockIdx.z;
  const LevelGeom &G = T->lv[level];
  uint8_t *plane = pyr + G.pyrOffset + (size_t)img * G.planeBytes;
  const int vecPerRow = G.pitch >> 4;
  const int total = 2 * kEdge * vecPerRow;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / vecPerRow, vx = i - r * vecPerRow;
    const int py = r < kEdge ? r : G.h + r;                           // rows 0 .. 18 and h + 19 .. h + 37
    const uint4 v = *reinterpret_cast<const uint4 *>(plane + (size_t)(reflect101(py - kEdge, G.h) + kEdge) * G.pitch + 16 * vx);
    *reinterpret_cast<uint4 *>(plane + (size_t)py * G.pitch + 16 * vx) = v;
  }
}
}  // namespace

static dim3 plane_grid(const LevelGeom &G, int batch, dim3 block, int bytesPerThread = 4) {
  return dim3((G.pitch / bytesPerThread + block.x - 1) / block.x, (G.rows + block.y - 1) / block.y, batch);
}

cudaError_t launch_import_level0(const LevelTable &T, const ExtractBuffers &B, const uint8_t *src,
                                 size_t rowStep, size_t imgStep, int imgBase, int batch, cudaStream_t s) {
  const LevelGeom &G = T.lv[0];
  // chunk c (padded bytes 16 c .. 16 c + 15) is interior-only when 16 c >= 19 and 16 c + 15 < w + 19; the bulk path also reads the
  // aligned word after its last byte, so its last chunk stays 4 pixels short of the row end
  const int c0 = (kEdge + 15) / 16, c1 = std::max(c0, (G.w + kEdge - 4) / 16), nchunks = G.pitch / 16;
  dim3 block(32, 8);
  if (c1 > c0) {
    dim3 grid((c1 - c0 + block.x - 1) / block.x, (G.rows + block.y - 1) / block.y, batch);
    k_import_level0<false><<<grid, block, 0, s>>>(G, B.pyr, src, rowStep, imgStep, imgBase, c0, c1);
    count_launch();
  }
  const int nborder = nchunks - (c1 - c0);
  dim3 bblock(8, 32);
  dim3 bgrid((nborder + bblock.x - 1) / bblock.x, (G.rows + bblock.y - 1) / bblock.y, batch);
  k_import_level0<true><<<bgrid, bblock, 0, s>>>(G, B.pyr, src, rowStep, imgStep, imgBase, c0, c1);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_resize_level(const LevelTable &T, const ExtractBuffers &B, int level, int imgBase, int batch,
                                cudaStream_t s) {
  dim3 block(64, 4);
  k_resize_level<<<plane_grid(T.lv[level], batch, block), block, 0, s>>>(B.dT, B.pyr, B.tabs, level, imgBase);
  count_launch();
  return cudaGetLastError();
}

// TMA box of level `l`'s source window (in level l-1): 128 output columns / kRH output rows worth of source pixels plus the
// interpolation neighbour, the 16-byte alignment lead, the third word of the last lane's loads and one spare row.
void resize_box(const LevelTable &T, int l, int *boxW, int *boxH) {
  const LevelGeom &G = T.lv[l], &P = T.lv[l - 1];
  const double sx = (double)P.w / G.w, sy = (double)P.h / G.h;
  *boxW = ((int)std::ceil((kRW - 1) * sx) + 3 + 15 + 12 + 15) & ~15;
  *boxH = (int)std::ceil((kRH - 1) * sy) + 4;
  if (sx > 2.0 || sy > 2.0) *boxW = *boxH = 1 << 20;                  // the PRMT selectors reach 8 source bytes: generic kernel beyond scale 2
}

int resize_smem_bytes(const LevelTable &T) {
  int slot = 0;
  for (int l = 1; l < T.nlevels; l++) slot = max(slot, (T.lv[l].rzBoxW * T.lv[l].rzBoxH + 127) & ~127);
  return slot;
}

cudaError_t resize_configure(int slotBytes) {
  return cudaFuncSetAttribute(k_resize_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + kResizeWarps * slotBytes);
}

void resize_make_tiles(const LevelTable &T, int level, std::vector<int4> &tiles) {
  tiles.clear();
  const LevelGeom &G = T.lv[level];
  const int tx = (G.w + 2 * kEdge + kRW - 1) / kRW, ty = (G.h + kRH - 1) / kRH;     // tiles of the padded width
  for (int y = 0; y < ty; y++)
    for (int x = 0; x < tx; x++) tiles.push_back(make_int4(level, x, y, 0));
}

cudaError_t launch_resize_tiles(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &maps, const int4 *tiles,
                                int ntiles, int slotBytes, int imgBase, int batch, cudaStream_t s) {
  dim3 grid((ntiles + kResizeWarps - 1) / kResizeWarps, batch);
  k_resize_tiles<<<grid, kResizeWarps * 32, 128 + kResizeWarps * slotBytes, s>>>(maps, B.dT, B.tabs, tiles, ntiles, B.pyr, slotBytes, imgBase);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_fill_borders(const LevelTable &T, const ExtractBuffers &B, int firstLevel, int imgBase, int batch, cudaStream_t s) {
  if (firstLevel >= T.nlevels) return cudaSuccess;
  dim3 grid(2, T.nlevels - firstLevel, batch);
  k_fill_rows<<<grid, 256, 0, s>>>(B.dT, B.pyr, firstLevel, imgBase);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
