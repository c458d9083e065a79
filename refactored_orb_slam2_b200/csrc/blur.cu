// cv::GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as called at ORBextractor.cc:1018-1019.
//
// OpenCV's 8-bit path is a fixed-point separable filter: kernel {18,34,48,56,48,34,18}/256,
// horizontal sums kept exactly in 8.8 (<= 65280), vertical sums in 16.16, one rounding at the end
// ((v + 0x8000) >> 16).  The source plane already carries the reflect-101 apron, so reading 3 px
// past the interior reproduces the border mode with no special cases.
//
// One CTA = one 64 x 32 tile of one level of one image.  The (64+8) x (32+6) source window is
// staged in shared memory with aligned 32-bit loads, filtered horizontally into a u16 tile, then
// vertically; each thread emits 4 pixels per 32-bit store.  Tiles are anchored at padded x = 16 so
// that loads and stores are 4-byte aligned (the 3 extra columns left of the interior are valid
// blur values of the mirrored image and are never read).
#include "internal.cuh"

namespace orbgpu {

namespace {
constexpr int kTW = 64, kTH = 32;
constexpr int kInW = kTW + 8;     // bytes per staged row (x0-4 .. x0+67)
constexpr int kInH = kTH + 6;

__global__ void __launch_bounds__(256)
k_blur_level(const LevelTable *__restrict__ T, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur,
             int level) {
  __shared__ __align__(16) uint8_t sIn[kInH][kInW];
  __shared__ __align__(16) unsigned short sH[kInH][kTW];
  const LevelGeom &G = T->lv[level];
  const int img = blockIdx.z;
  const int x0 = 16 + kTW * blockIdx.x;          // padded coordinates of the tile origin
  const int y0 = kEdge + kTH * blockIdx.y;
  const size_t plane = G.pyrOffset + (size_t)img * G.planeBytes;
  const uint8_t *S = pyr + plane;
  const int tid = threadIdx.x;

  for (int i = tid; i < kInH * (kInW / 4); i += 256) {
    const int r = i / (kInW / 4), c4 = i % (kInW / 4);
    const int y = y0 - 3 + r, x = x0 - 4 + 4 * c4;
    unsigned v = 0;
    if (y < G.rows && x + 3 < G.pitch) v = *reinterpret_cast<const unsigned *>(S + (size_t)y * G.pitch + x);
    *reinterpret_cast<unsigned *>(&sIn[r][4 * c4]) = v;
  }
  __syncthreads();
  for (int i = tid; i < kInH * kTW; i += 256) {
    const int r = i / kTW, c = i % kTW;
    const uint8_t *p = &sIn[r][c + 1];
    const int acc = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
    sH[r][c] = (unsigned short)acc;
  }
  __syncthreads();
  const int xEnd = 16 + 4 * ((G.w + 3 + 3) / 4);   // first padded column not produced
  const int c4 = tid & 15, r0 = tid >> 4;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int r = r0 + 16 * j;
    const int y = y0 + r, x = x0 + 4 * c4;
    if (y >= G.h + kEdge || x >= xEnd) continue;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = 4 * c4 + k;
      const unsigned acc = 18u * (sH[r][c] + sH[r + 6][c]) + 34u * (sH[r + 1][c] + sH[r + 5][c]) +
                           48u * (sH[r + 2][c] + sH[r + 4][c]) + 56u * sH[r + 3][c];
      out |= ((acc + 0x8000u) >> 16) << (8 * k);
    }
    *reinterpret_cast<unsigned *>(blur + plane + (size_t)y * G.pitch + x) = out;
  }
}
}  // namespace

cudaError_t launch_blur(const LevelTable &T, const ExtractBuffers &B, int batch, cudaStream_t s) {
  for (int l = 0; l < T.nlevels; l++) {
    const LevelGeom &G = T.lv[l];
    dim3 grid((G.w + 3 + kTW - 1) / kTW + 0, (G.h + kTH - 1) / kTH, batch);
    k_blur_level<<<grid, 256, 0, s>>>(B.dT, B.pyr, B.blur, l);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace orbgpu
