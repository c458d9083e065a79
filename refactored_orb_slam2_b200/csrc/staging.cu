// staging.cu — the two image operations the reference applies right before ORBextractor::operator():
// cvtColor(..., COLOR_*2GRAY) (src/Tracking.cc:166-182, 200-211, 228-236) and, for unrectified stereo rigs,
// cv::remap(..., INTER_LINEAR) with the float maps of initUndistortRectifyMap (Examples/Stereo/stereo_euroc.cc:109-111,
// 159-160).  Both are HBM-bound byte kernels; arithmetic = OpenCV's fixed point (see include/orbgpu.h).
#include "internal.cuh"

namespace orbgpu {
namespace {

// One thread = 4 output pixels = 4 * CH source bytes, fetched as aligned 32-bit words (3-channel rows of odd width start
// on any byte, so the window is re-aligned in registers with funnel shifts) and written as one 32-bit word of gray pixels.
template <int CH>
__global__ void __launch_bounds__(128) k_cvt_gray(const uint8_t *__restrict__ src, int code, int width, int height, size_t srcStep, size_t srcImgStep,
                                                   uint8_t *__restrict__ dst, size_t dstStep, size_t dstImgStep) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, img = blockIdx.z;
  if (x4 >= width) return;
  const uint8_t *first = src + (size_t)img * srcImgStep + (size_t)y * srcStep + (size_t)x4 * CH;
  const int n = min(4, width - x4);
  const int mis = (int)((uintptr_t)first & 3);
  const unsigned *base = reinterpret_cast<const unsigned *>(first - mis);
  const int nwords = (mis + n * CH + 3) >> 2;              // never reads past the word holding the last needed byte
  unsigned w[CH + 1];
#pragma unroll
  for (int k = 0; k <= CH; k++) w[k] = k < nwords ? base[k] : 0u;
  unsigned a[CH];                                          // the 4 * CH bytes, re-aligned
#pragma unroll
  for (int k = 0; k < CH; k++) a[k] = __funnelshift_r(w[k], w[k + 1], 8 * mis);
  const int c0 = code == ORBGPU_RGB2GRAY ? 9798 : 3735, c2 = code == ORBGPU_RGB2GRAY ? 3735 : 9798;   // R2Y / B2Y (15-bit), G2Y = 19235
  auto byte_at = [&](int i) -> int { return (int)((a[i >> 2] >> (8 * (i & 3))) & 0xffu); };
  unsigned out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int v = (byte_at(k * CH) * c0 + byte_at(k * CH + 1) * 19235 + byte_at(k * CH + 2) * c2 + (1 << 14)) >> 15;
    out |= (unsigned)v << (8 * k);
  }
  uint8_t *d = dst + (size_t)img * dstImgStep + (size_t)y * dstStep + x4;
  if (n == 4 && (((uintptr_t)d) & 3) == 0) *reinterpret_cast<unsigned *>(d) = out;
  else for (int k = 0; k < n; k++) d[k] = (uint8_t)(out >> (8 * k));
}

// one thread = 4 consecutive output pixels: two 128-bit map loads, 16 byte taps (neighbouring outputs read neighbouring
// source pixels: L1 serves them), one 32-bit store
__global__ void __launch_bounds__(128) k_remap(const uint8_t *__restrict__ src, int sw, int sh, size_t srcStep, size_t srcImgStep,
                                                const float *__restrict__ mapX, const float *__restrict__ mapY, int width, int height,
                                                uint8_t *__restrict__ dst, size_t dstStep, size_t dstImgStep) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, img = blockIdx.z;
  if (x4 >= width) return;
  const size_t m = (size_t)y * width + x4;
  const int n = min(4, width - x4);
  float mx[4], my[4];
  if (n == 4 && ((((uintptr_t)(mapX + m)) | ((uintptr_t)(mapY + m))) & 15) == 0) {
    const float4 a = *reinterpret_cast<const float4 *>(mapX + m), b = *reinterpret_cast<const float4 *>(mapY + m);
    mx[0] = a.x; mx[1] = a.y; mx[2] = a.z; mx[3] = a.w; my[0] = b.x; my[1] = b.y; my[2] = b.z; my[3] = b.w;
  } else {
    for (int k = 0; k < 4; k++) { mx[k] = k < n ? mapX[m + k] : 0.f; my[k] = k < n ? mapY[m + k] : 0.f; }
  }
  const uint8_t *s = src + (size_t)img * srcImgStep;
  auto tap = [&](int yy, int xx) -> int {
    return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? (int)s[(size_t)yy * srcStep + xx] : 0;   // BORDER_CONSTANT, value 0
  };
  unsigned out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    // cvRound(map * INTER_TAB_SIZE): round half to even; the product by 32 is exact
    const int sx = __float2int_rn(__fmul_rn(mx[k], 32.f)), sy = __float2int_rn(__fmul_rn(my[k], 32.f));
    const int fx = sx & 31, fy = sy & 31;
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));   // saturate_cast<short>
    const int v = (32 - fx) * (32 - fy) * 32 * tap(iy, ix) + fx * (32 - fy) * 32 * tap(iy, ix + 1) + (32 - fx) * fy * 32 * tap(iy + 1, ix) +
                  fx * fy * 32 * tap(iy + 1, ix + 1);
    out |= (unsigned)((v + (1 << 14)) >> 15) << (8 * k);
  }
  uint8_t *d = dst + (size_t)img * dstImgStep + (size_t)y * dstStep + x4;
  if (n == 4 && (((uintptr_t)d) & 3) == 0) *reinterpret_cast<unsigned *>(d) = out;
  else for (int k = 0; k < n; k++) d[k] = (uint8_t)(out >> (8 * k));
}

}  // namespace
}  // namespace orbgpu

using namespace orbgpu;

extern "C" {

int orbgpu_cvt_gray_device(const uint8_t *d_src, int channels, int code, int width, int height, size_t src_step, size_t src_img_step,
                           uint8_t *d_dst, size_t dst_step, size_t dst_img_step, int batch, void *stream) {
  if (!d_src || !d_dst || (channels != 3 && channels != 4) || (code != ORBGPU_BGR2GRAY && code != ORBGPU_RGB2GRAY) || width < 1 || height < 1 ||
      batch < 1 || src_step < (size_t)width * channels || dst_step < (size_t)width || height > 65535 || batch > 65535) {
    set_error("cvt_gray: bad arguments");
    return ORBGPU_E_INVALID;
  }
  const dim3 grid(((width + 3) / 4 + 127) / 128, height, batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (channels == 3) k_cvt_gray<3><<<grid, 128, 0, s>>>(d_src, code, width, height, src_step, src_img_step, d_dst, dst_step, dst_img_step);
  else k_cvt_gray<4><<<grid, 128, 0, s>>>(d_src, code, width, height, src_step, src_img_step, d_dst, dst_step, dst_img_step);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  return ORBGPU_OK;
}

int orbgpu_remap_device(const uint8_t *d_src, int src_width, int src_height, size_t src_step, size_t src_img_step, const float *d_map_x,
                        const float *d_map_y, int width, int height, uint8_t *d_dst, size_t dst_step, size_t dst_img_step, int batch, void *stream) {
  if (!d_src || !d_dst || !d_map_x || !d_map_y || src_width < 1 || src_height < 1 || width < 1 || height < 1 || batch < 1 ||
      src_step < (size_t)src_width || dst_step < (size_t)width || height > 65535 || batch > 65535) {
    set_error("remap: bad arguments");
    return ORBGPU_E_INVALID;
  }
  const dim3 grid(((width + 3) / 4 + 127) / 128, height, batch);
  k_remap<<<grid, 128, 0, (cudaStream_t)stream>>>(d_src, src_width, src_height, src_step, src_img_step, d_map_x, d_map_y, width, height, d_dst, dst_step,
                                                 dst_img_step);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  return ORBGPU_OK;
}

namespace {
struct DeviceScope {
  int prev = -1;
  bool ok = true;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

int orbgpu_cvt_gray(const uint8_t *src, int channels, int code, int width, int height, uint8_t *dst, int batch, int device) {
  if (!src || !dst || width < 1 || height < 1 || batch < 1 || (channels != 3 && channels != 4)) { set_error("cvt_gray: bad arguments"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceScope scope(device);
  const size_t sb = (size_t)width * channels * height, db = (size_t)width * height;
  uint8_t *ds = nullptr, *dd = nullptr;
  cudaError_t e = cudaMalloc(&ds, sb * batch);
  if (e == cudaSuccess) e = cudaMalloc(&dd, db * batch);
  if (e == cudaSuccess) e = cudaMemcpy(ds, src, sb * batch, cudaMemcpyHostToDevice);
  int rc = ORBGPU_OK;
  if (e == cudaSuccess) rc = orbgpu_cvt_gray_device(ds, channels, code, width, height, (size_t)width * channels, sb, dd, (size_t)width, db, batch, nullptr);
  if (e == cudaSuccess && rc == ORBGPU_OK) e = cudaMemcpy(dst, dd, db * batch, cudaMemcpyDeviceToHost);
  cudaFree(ds); cudaFree(dd);
  if (e != cudaSuccess) { set_error("cvt_gray: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  return rc;
}

int orbgpu_remap(const uint8_t *src, int src_width, int src_height, const float *map_x, const float *map_y, int width, int height, uint8_t *dst,
                 int batch, int device) {
  if (!src || !dst || !map_x || !map_y || src_width < 1 || src_height < 1 || width < 1 || height < 1 || batch < 1) { set_error("remap: bad arguments"); return ORBGPU_E_INVALID; }
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DeviceScope scope(device);
  const size_t sb = (size_t)src_width * src_height, db = (size_t)width * height;
  uint8_t *ds = nullptr, *dd = nullptr;
  float *mx = nullptr, *my = nullptr;
  cudaError_t e = cudaMalloc(&ds, sb * batch);
  if (e == cudaSuccess) e = cudaMalloc(&dd, db * batch);
  if (e == cudaSuccess) e = cudaMalloc(&mx, db * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&my, db * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(ds, src, sb * batch, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mx, map_x, db * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(my, map_y, db * sizeof(float), cudaMemcpyHostToDevice);
  int rc = ORBGPU_OK;
  if (e == cudaSuccess) rc = orbgpu_remap_device(ds, src_width, src_height, (size_t)src_width, sb, mx, my, width, height, dd, (size_t)width, db, batch, nullptr);
  if (e == cudaSuccess && rc == ORBGPU_OK) e = cudaMemcpy(dst, dd, db * batch, cudaMemcpyDeviceToHost);
  cudaFree(ds); cudaFree(dd); cudaFree(mx); cudaFree(my);
  if (e != cudaSuccess) { set_error("remap: %s", cudaGetErrorString(e)); return ORBGPU_E_CUDA; }
  return rc;
}

}  // extern "C"
