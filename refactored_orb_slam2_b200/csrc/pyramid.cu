// Image pyramid: ORBextractor::ComputePyramid (ORBextractor.cc:1041-1065).
//
// Every level lives in HBM as a [batch] array of padded planes: interior w x h surrounded by the
// reference's 19 px BORDER_REFLECT_101 apron, row pitch a multiple of 16 bytes.  Both kernels
// write whole padded planes in one pass: a thread produces 4 consecutive padded bytes (one 32-bit
// store), mapping each through reflect-101 to the interior pixel it mirrors, so no second
// copyMakeBorder pass (and no re-read of the level) is needed.
//
// Level l >= 1 is cv::resize(level l-1, INTER_LINEAR) in OpenCV's 11-bit fixed point:
//   H(x)   = S[xo]*a0 + S[xo+1]*a1                      (int32, a0+a1 = 2048)
//   out    = ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// with per-axis (offset, coefficient) tables built on the host exactly like OpenCV builds them.
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

__global__ void __launch_bounds__(256)
k_import_level0(const __grid_constant__ LevelGeom G, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ src,
                size_t rowStep, size_t imgStep, int imgBase) {   // level-0 geometry by value: no dependent table loads in front of the copy
  const int img = imgBase + blockIdx.z;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  const int px4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (py >= G.rows || px4 >= G.pitch) return;
  const uint8_t *S = src + (size_t)img * imgStep + (size_t)reflect101(py - kEdge, G.h) * rowStep;
  unsigned out;
  if (px4 >= kEdge && px4 + 3 < G.w + kEdge) {      // four interior pixels: no mirroring
    const uint8_t *p = S + (px4 - kEdge);
    out = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
  } else {
    out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int px = px4 + k;
      unsigned v = 0;
      if (px < G.w + 2 * kEdge) v = S[reflect101(px - kEdge, G.w)];
      out |= v << (8 * k);
    }
  }
  uint8_t *D = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)py * G.pitch + px4;
  *reinterpret_cast<unsigned *>(D) = out;
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

// ---- level l >= 1, fast path: one WARP per 128 x 32 output tile, source window staged by TMA ------------------
// A lane owns 4 adjacent output columns (one aligned 32-bit store per row) and walks down the tile's rows; the
// horizontal interpolation of a source row is kept in registers and reused by the next output row whenever both
// read it (4 of 5 rows at scale 1.2).  Only the interior is produced here (words may spill garbage into the
// apron bytes next to it); k_fill_borders writes every apron afterwards.
namespace {
#ifndef ORBGPU_RESIZE_RH
#define ORBGPU_RESIZE_RH 16
#endif
#ifndef ORBGPU_RESIZE_WARPS
#define ORBGPU_RESIZE_WARPS 4
#endif
constexpr int kRW = 128, kRH = ORBGPU_RESIZE_RH, kResizeWarps = ORBGPU_RESIZE_WARPS;

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
  const int xp0 = 16 + kRW * tile.y;             // padded column of the tile's first output pixel (4-byte aligned)
  const int y0 = kRH * tile.z;                   // interior row of the tile's first output row
  const int xiFirst = min(max(xp0 - kEdge, 0), G.w - 1);
  const int sxMin = tabs[G.xTab + xiFirst].x, syMin = tabs[G.yTab + y0].x;
  const int bx = (sxMin + kEdge) & ~15;          // TMA: innermost coordinate on a 16-byte boundary
  const int lead = sxMin + kEdge - bx;
  uint8_t *slot = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u) + warp * slotBytes;
  const int bw = G.rzBoxW;
  if (lane == 0) {
    mbar_init(&sBar[warp], 1);
    mbar_fence_init();
    mbar_expect_tx(&sBar[warp], (unsigned)(G.rzBoxW * G.rzBoxH));
    tma_load_3d(slot, &maps.m[level - 1], bx, syMin + kEdge, img, &sBar[warp]);
  }
  // per-lane column tables while the copy is in flight
  const int xp = xp0 + 4 * lane;
  int rel[4], a0[4], a1[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int xi = min(max(xp + p - kEdge, 0), G.w - 1);
    const int2 tx = tabs[G.xTab + xi];
    rel[p] = tx.x - sxMin + lead;
    a0[p] = tx.y & 0xffff; a1[p] = tx.y >> 16;
  }
  const bool colOk = xp + 3 >= kEdge && xp < kEdge + G.w;
  uint8_t *dst = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)(y0 + kEdge) * G.pitch + xp;
  __syncwarp();
  mbar_wait(&sBar[warp], 0);

  int h0[4], h1[4], idx0 = -1, idx1 = -1;
  const int rows = min(kRH, G.h - y0);
  for (int r = 0; r < rows; r++) {
    const int2 ty = tabs[G.yTab + y0 + r];
    const int s0 = ty.x - syMin, s1 = s0 + 1;    // row s1 has weight 0 when it is clamped in the reference
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    if (s0 == idx1) {
#pragma unroll
      for (int p = 0; p < 4; p++) h0[p] = h1[p];
      idx0 = idx1;
    } else if (s0 != idx0) {
      const uint8_t *S = slot + s0 * bw;
#pragma unroll
      for (int p = 0; p < 4; p++) h0[p] = S[rel[p]] * a0[p] + S[rel[p] + 1] * a1[p];
      idx0 = s0;
    }
    if (s1 != idx1) {
      const uint8_t *S = slot + s1 * bw;
#pragma unroll
      for (int p = 0; p < 4; p++) h1[p] = S[rel[p]] * a0[p] + S[rel[p] + 1] * a1[p];
      idx1 = s1;
    }
    unsigned out = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
      out |= (unsigned)((((b0 * (h0[p] >> 4)) >> 16) + ((b1 * (h1[p] >> 4)) >> 16) + 2) >> 2) << (8 * p);
    if (colOk) *reinterpret_cast<unsigned *>(dst + (size_t)r * G.pitch) = out;
  }
}

// BORDER_REFLECT_101 apron of levels firstLevel .. nlevels-1 (ORBextractor.cc:1057-1059), one 32-bit word per item:
// the 19 rows above / below the interior in full, and on interior rows the words covering the left and right aprons.
__global__ void __launch_bounds__(256)
k_fill_borders(const LevelTable *__restrict__ T, uint8_t *__restrict__ pyr, int firstLevel, int imgBase) {
  const int level = firstLevel + blockIdx.y, img = imgBase + blockIdx.z;
  const LevelGeom &G = T->lv[level];
  uint8_t *plane = pyr + G.pyrOffset + (size_t)img * G.planeBytes;
  const int wordsPerRow = G.pitch >> 2;
  const int leftWords = (kEdge + 3) >> 2;                       // words 0 .. 4 hold padded columns 0 .. 19
  const int rightFirst = (G.w + kEdge) >> 2;                    // first word touching the right apron
  const int sideWords = leftWords + (wordsPerRow - rightFirst);
  const int total = 2 * kEdge * wordsPerRow + G.h * sideWords;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int py, wx;
    if (i < 2 * kEdge * wordsPerRow) {
      const int r = i / wordsPerRow;
      wx = i - r * wordsPerRow;
      py = r < kEdge ? r : G.h + r;                             // rows 0..18 and h+19 .. h+37
    } else {
      const int j = i - 2 * kEdge * wordsPerRow;
      const int r = j / sideWords, k = j - r * sideWords;
      py = kEdge + r;
      wx = k < leftWords ? k : rightFirst + (k - leftWords);
    }
    const uint8_t *S = plane + (size_t)(reflect101(py - kEdge, G.h) + kEdge) * G.pitch + kEdge;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int px = 4 * wx + k;
      if (px < G.w + 2 * kEdge) out |= (unsigned)S[reflect101(px - kEdge, G.w)] << (8 * k);
    }
    *reinterpret_cast<unsigned *>(plane + (size_t)py * G.pitch + 4 * wx) = out;
  }
}
}  // namespace

static dim3 plane_grid(const LevelGeom &G, int batch, dim3 block) {
  return dim3((G.pitch / 4 + block.x - 1) / block.x, (G.rows + block.y - 1) / block.y, batch);
}

cudaError_t launch_import_level0(const LevelTable &T, const ExtractBuffers &B, const uint8_t *src,
                                 size_t rowStep, size_t imgStep, int imgBase, int batch, cudaStream_t s) {
  dim3 block(64, 4);
  k_import_level0<<<plane_grid(T.lv[0], batch, block), block, 0, s>>>(T.lv[0], B.pyr, src, rowStep, imgStep, imgBase);
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

// TMA box of level `l`'s source window (in level l-1): 128 output columns / 32 output rows worth of source pixels
// plus the interpolation neighbour, the 16-byte alignment lead and one spare row.
void resize_box(const LevelTable &T, int l, int *boxW, int *boxH) {
  const LevelGeom &G = T.lv[l], &P = T.lv[l - 1];
  const double sx = (double)P.w / G.w, sy = (double)P.h / G.h;
  *boxW = ((int)std::ceil((kRW - 1) * sx) + 3 + 15 + 1 + 15) & ~15;
  *boxH = (int)std::ceil((kRH - 1) * sy) + 4;
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
  const int tx = (G.w + 3 + kRW - 1) / kRW, ty = (G.h + kRH - 1) / kRH;
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
  dim3 grid(8, T.nlevels - firstLevel, batch);
  k_fill_borders<<<grid, 256, 0, s>>>(B.dT, B.pyr, firstLevel, imgBase);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
