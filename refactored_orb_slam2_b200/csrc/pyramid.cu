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
#include "internal.cuh"

namespace orbgpu {

__device__ __forceinline__ int reflect101(int p, int n) {
  // |p| excursions are at most kEdge (19) beyond the interior; a single fold suffices when n > 19,
  // the loop covers tiny levels.
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

__global__ void __launch_bounds__(256)
k_import_level0(const LevelTable *__restrict__ T, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ src,
                size_t rowStep, size_t imgStep) {
  const LevelGeom &G = T->lv[0];
  const int img = blockIdx.z;
  const int py = blockIdx.y * blockDim.y + threadIdx.y;
  const int px4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (py >= G.rows || px4 >= G.pitch) return;
  const uint8_t *S = src + (size_t)img * imgStep + (size_t)reflect101(py - kEdge, G.h) * rowStep;
  unsigned out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int px = px4 + k;
    unsigned v = 0;
    if (px < G.w + 2 * kEdge) v = S[reflect101(px - kEdge, G.w)];
    out |= v << (8 * k);
  }
  uint8_t *D = pyr + G.pyrOffset + (size_t)img * G.planeBytes + (size_t)py * G.pitch + px4;
  *reinterpret_cast<unsigned *>(D) = out;
}

__global__ void __launch_bounds__(256)
k_resize_level(const LevelTable *__restrict__ T, uint8_t *__restrict__ pyr, const int2 *__restrict__ tabs,
               int level) {
  const LevelGeom &G = T->lv[level];
  const LevelGeom &P = T->lv[level - 1];
  const int img = blockIdx.z;
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

static dim3 plane_grid(const LevelGeom &G, int batch, dim3 block) {
  return dim3((G.pitch / 4 + block.x - 1) / block.x, (G.rows + block.y - 1) / block.y, batch);
}

cudaError_t launch_import_level0(const LevelTable &T, const ExtractBuffers &B, const uint8_t *src,
                                 size_t rowStep, size_t imgStep, int batch, cudaStream_t s) {
  dim3 block(64, 4);
  k_import_level0<<<plane_grid(T.lv[0], batch, block), block, 0, s>>>(B.dT, B.pyr, src, rowStep, imgStep);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_resize_level(const LevelTable &T, const ExtractBuffers &B, int level, int batch,
                                cudaStream_t s) {
  dim3 block(64, 4);
  k_resize_level<<<plane_grid(T.lv[level], batch, block), block, 0, s>>>(B.dT, B.pyr, B.tabs, level);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
