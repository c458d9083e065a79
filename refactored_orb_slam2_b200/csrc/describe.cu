// Orientation, descriptor and output assembly: IC_Angle / computeOrientation
// (ORBextractor.cc:76-100, 466-473), computeOrbDescriptor (ORBextractor.cc:102-146) and the
// tail of ORBextractor::operator() (ORBextractor.cc:1023-1037).
//
// One warp = one keypoint.
//  * IC_Angle: 31 row reads of the staged disc of the UNBLURRED level (lane = column u, immediate offsets), integer moments
//    m10 / m01 reduced by shuffles (exact, order-free), then cv::fastAtan2's float polynomial with
//    round-to-nearest mul/add/div intrinsics (the reference build never fuses into FMAs).
//  * both neighbourhoods are staged in the warp's shared-memory slot by TWO TMA tensor copies per keypoint (the 31 x 31
//    orientation disc from the level, the 37 x 37 patch from its blurred copy; pattern reach <= 18 px) — the 512 rotated
//    sample positions are a scatter that would otherwise cost one 32-byte L2 sector per sample; then lane i builds
//    descriptor byte i from 16 pattern points rotated by (cos, sin) of the angle.  cosf/sinf replicate glibc's double-precision sincosf algorithm (see sincos_glibc) so
//    that cvRound(x*b + y*a) lands on the same pixel as on the host.
//  * The keypoint record (cv::KeyPoint layout) and the 32-byte descriptor row go straight to their
//    final, level-major slot: start = sum of the lower levels' octree counts.
#include <mutex>

#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

namespace {

__device__ const signed char d_pattern[1024] = {
#include "../../include/orb_pattern.inc"
};
// The 512 pattern points as float2, transposed so that the 32 lanes of a warp (lane = descriptor byte) read
// consecutive 8-byte entries: entry [j][lane] = point 16*lane + j.  Filled once per device by k_init_pattern; 4 KB,
// L1-resident for the descriptor loop.
__device__ float2 d_patternT[512];
__global__ void k_init_pattern() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 512) d_patternT[(i & 15) * 32 + (i >> 4)] = make_float2((float)d_pattern[2 * i], (float)d_pattern[2 * i + 1]);
}

// cv::fastAtan2(y, x) scalar path (called at ORBextractor.cc:99): degrees in [0, 360).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// glibc >= 2.28 sinf/cosf/sincosf (ARM optimized-routines algorithm): everything in double,
// Cody-Waite reduction by pi/2 for |x| < 120, degree-7/8 polynomials, one rounding to float.
// Only the |x| < 120 range is needed: angles are fastAtan2 degrees * pi/180 in [0, 2*pi].
__device__ __forceinline__ float sincos_poly(double x, double x2, bool neg, bool cosine) {
  const double C0 = neg ? -1.0 : 1.0;
  const double C1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
  const double C2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
  const double C3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
  const double C4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  if (!cosine) {
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
    const double x7 = __dmul_rn(x3, x2);
    const double s = __dadd_rn(x, __dmul_rn(x3, S1));
    return (float)__dadd_rn(s, __dmul_rn(x7, s1));
  }
  const double x4 = __dmul_rn(x2, x2);
  const double c2 = __dadd_rn(C3, __dmul_rn(x2, C4));
  const double c1 = __dadd_rn(C0, __dmul_rn(x2, C1));
  const double x6 = __dmul_rn(x4, x2);
  const double c = __dadd_rn(c1, __dmul_rn(x4, C2));
  return (float)__dadd_rn(c, __dmul_rn(x6, c2));
}

__device__ __forceinline__ void sincos_glibc(float y, float *sinp, float *cosp) {
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
  double x = (double)y;
  if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff)) {          // |y| < pi/4
    if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ff)) { *sinp = y; *cosp = 1.0f; return; }
    const double x2 = __dmul_rn(x, x);
    *sinp = sincos_poly(x, x2, false, false);
    *cosp = sincos_poly(x, x2, false, true);
    return;
  }
  const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
  const int n = ((int)r + 0x800000) >> 24;
  x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
  const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  const bool neg = (n & 2) != 0;
  const double xs = __dmul_rn(x, sgn), x2 = __dmul_rn(x, x);
  *sinp = sincos_poly(xs, x2, neg, (n & 1) != 0);
  *cosp = sincos_poly(xs, x2, neg, (n & 1) == 0);
}

// This translation unit is compiled twice: as is (throughput shape) and through *_lat.cu with a second thread shape for
// frame-at-a-time calls (a stereo pair keeps only 16-32 CTAs busy, so the latency build gives each more threads);
// ORBGPU_VARIANT suffixes the exported launchers.
#ifndef ORBGPU_VARIANT
#define ORBGPU_VARIANT
#endif
#define ORBGPU_CAT2(a, b) a##b
#define ORBGPU_CAT(a, b) ORBGPU_CAT2(a, b)
#define VNAME(x) ORBGPU_CAT(x, ORBGPU_VARIANT)
#ifndef ORBGPU_DESCRIBE_WARPS
#define ORBGPU_DESCRIBE_WARPS 4
#endif
constexpr int kWarpsPerCta = ORBGPU_DESCRIBE_WARPS;
constexpr int kPatchR = 18;                    // |rotated pattern coordinate| <= round(13 * sqrt 2) = 18 < EDGE_THRESHOLD
constexpr int kPatchRows = 2 * kPatchR + 1;    // 37
// Both neighbourhoods of a keypoint arrive as TMA boxes (cp.async.bulk.tensor.3d on the level's [batch][rows][pitch] maps,
// one per-warp mbarrier): the 31 x 31 orientation disc from the pyramid level and the 37 x 37 patch from the blurred level.
// Box columns start on the 16-byte boundary at or before the first needed pixel, so the boxes are 48 and 64 bytes wide.
constexpr int kDiscW = 48, kDiscRows = 2 * kHalfPatch + 1;        // 15 alignment bytes + 31 <= 48
constexpr int kPatchW = 64;                                        // 15 alignment bytes + 37 <= 64
constexpr int kDiscBytes = kDiscW * kDiscRows, kPatchBytes = kPatchW * kPatchRows;
constexpr int kDiscSlot = (kDiscBytes + 127) & ~127, kPatchSlot = (kPatchBytes + 127) & ~127;
constexpr int kWarpSlot = kDiscSlot + kPatchSlot;

// blockIdx.x = chunk of kWarpsPerCta output slots (level-major keypoint index), blockIdx.y = image
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_describe(const __grid_constant__ TmaMaps pyrMaps, const __grid_constant__ TmaMaps blurMaps, const LevelTable *__restrict__ T,
           const int4 *__restrict__ selAll, const int *__restrict__ selCount, orbgpu_keypoint *__restrict__ kps,
           uint8_t *__restrict__ desc, int *__restrict__ kpCount, unsigned *__restrict__ status, int imgBase) {
  extern __shared__ uint8_t smemRaw[];
  __shared__ __align__(8) unsigned long long sBar[kWarpsPerCta];

  const int img = imgBase + blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * kWarpsPerCta + warp;                    // output row of this warp's keypoint
  // the image's output rows are level-major: lane l holds the end (exclusive) of level l's rows after a warp scan of the
  // per-level octree counts; the keypoint's level is the number of levels that end at or before its slot
  const int nlevels = T->nlevels;
  int levelEnd = lane < nlevels ? selCount[img * kMaxLevels + lane] : 0;
#pragma unroll
  for (int o = 1; o < kMaxLevels; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, levelEnd, o);
    if (lane >= o) levelEnd += v;
  }
  const int total = __shfl_sync(0xffffffffu, levelEnd, nlevels - 1);
  const int level = min(__popc(__ballot_sync(0xffffffffu, lane < nlevels && slot >= levelEnd)), nlevels - 1);
  const int start = level ? __shfl_sync(0xffffffffu, levelEnd, level - 1) : 0;
  const int kpCap = T->kpCap;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    kpCount[img] = min(total, kpCap);
    if (total > kpCap) atomicOr(status, kStatusKpOverflow);
  }
  if (slot >= total || slot >= kpCap) return;
  const LevelGeom &G = T->lv[level];
  const int4 s = selAll[G.selOffset + (size_t)img * G.selCap + (slot - start)];   // x_level, y_level, response

  // ---- stage both neighbourhoods: two tensor copies per keypoint, no per-row address arithmetic, one memory latency
  uint8_t *disc = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u) + warp * kWarpSlot;
  uint8_t *patch = disc + kDiscSlot;
  const int dx0 = s.x + kEdge - kHalfPatch, px0 = s.x + kEdge - kPatchR;      // padded columns of u = -15 / of the patch's first pixel
  if (lane == 0) {
    mbar_init(&sBar[warp], 1);
    mbar_fence_init();
    mbar_expect_tx(&sBar[warp], kDiscBytes + kPatchBytes);
    tma_load_3d(disc, &pyrMaps.m[level], dx0 & ~15, s.y + kEdge - kHalfPatch, img, &sBar[warp]);
    tma_load_3d(patch, &blurMaps.m[level], px0 & ~15, s.y + kEdge - kPatchR, img, &sBar[warp]);
  }
  __syncwarp();
  static constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};   // ORBextractor.cc:449-463
  const int u = lane - kHalfPatch;
  const int au = abs(u);
  mbar_wait(&sBar[warp], 0);
  int ring[2 * kHalfPatch + 1];
  {
    const uint8_t *rp = disc + (dx0 & 15) + min(lane, 2 * kHalfPatch);     // lane = column u + 15 (lane 31 idle)
#pragma unroll
    for (int i = 0; i <= 2 * kHalfPatch; i++) {
      const int v = i - kHalfPatch;
      ring[i] = (lane < 31 && au <= kUmax[v < 0 ? -v : v]) ? (int)rp[i * kDiscW] : 0;
    }
  }
  const int lead = px0 & 15;

  // ---- IC_Angle
  int m10 = 0, m01 = 0;
#pragma unroll
  for (int i = 0; i <= 2 * kHalfPatch; i++) {
    m10 += ring[i];
    m01 += (i - kHalfPatch) * ring[i];
  }
  m10 *= u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- steered BRIEF on the staged blurred patch
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  float a, b;
  sincos_glibc(__fmul_rn(angle, factorPI), &b, &a);   // a = cos, b = sin
  // cvRound(x) = round-half-even = the low mantissa bits of x + 1.5 * 2^23 for |x| < 2^22 (an FADD on the FMA pipe instead of
  // an F2I on the quarter-rate XU pipe); the two biases are folded into the base offset (bits of 1.5 * 2^23 = 0x4B400000)
  const float kMagic = 12582912.0f;
  const uint8_t *pb = patch;
  const unsigned off0 = (unsigned)(kPatchR * kPatchW + lead + kPatchR) - 0x4B400000u * (unsigned)(kPatchW + 1);   // mod 2^32
  const float2 *p = d_patternT + lane;         // point 16*lane + j sits at p[32 * j]
  unsigned val = 0;
#pragma unroll
  for (int bit = 0; bit < 8; bit++) {
    int t[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float2 q = __ldg(&p[32 * (2 * bit + e)]);
      const unsigned ry = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(q.x, b), __fmul_rn(q.y, a)), kMagic));
      const unsigned rx = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(q.x, a), __fmul_rn(q.y, b)), kMagic));
      t[e] = pb[ry * (unsigned)kPatchW + rx + off0];
    }
    val |= (unsigned)(t[0] < t[1]) << bit;
  }
  const size_t row = (size_t)img * kpCap + slot;
  desc[row * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    orbgpu_keypoint o;
    o.x = (float)s.x; o.y = (float)s.y;
    if (level != 0) { o.x = __fmul_rn(o.x, G.scale); o.y = __fmul_rn(o.y, G.scale); }
    o.size = (float)G.patchSize;
    o.angle = angle;
    o.response = (float)s.z;
    o.octave = level;
    o.class_id = -1;
    kps[row] = o;
  }
}

}  // namespace

#ifndef ORBGPU_VARIANT_IS_SECONDARY
// ---- parity hook for the descriptor's trigonometry: the device sincos above over a RANGE of float32 bit patterns, reduced
// to one 64-bit checksum per 2^chunkLog2 consecutive patterns so that all ~1.09 G floats of [0, 2 pi] can be compared with the
// host libm without moving 8 GB (tests/test_gpu_sincos.py; the host side of the sum is oracle/orb_oracle_sincos_checksum)
namespace {
__device__ __forceinline__ unsigned long long sincos_hash(unsigned bits, float s, float c) {
  return (((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(c)) * (0x9E3779B97F4A7C15ull + 2ull * bits);
}
__global__ void __launch_bounds__(256) k_sincos_checksum(unsigned firstBits, unsigned long long count, int chunkLog2, unsigned long long *sums) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  // consecutive threads take consecutive patterns; a warp's 32 patterns fall into one chunk (chunks are multiples of 32)
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((count + 31) & ~31ull); i += stride) {
    unsigned long long h = 0;
    if (i < count) {
      const unsigned bits = firstBits + (unsigned)i;
      float sn, cs;
      sincos_glibc(__uint_as_float(bits), &sn, &cs);
      h = sincos_hash(bits, sn, cs);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sums[i >> chunkLog2], h);
  }
}
__global__ void k_sincos_values(unsigned firstBits, int n, float *s, float *c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sincos_glibc(__uint_as_float(firstBits + (unsigned)i), &s[i], &c[i]);
}
}  // namespace
cudaError_t launch_sincos_checksum(unsigned firstBits, unsigned long long count, int chunkLog2, unsigned long long *dSums, cudaStream_t st) {
  k_sincos_checksum<<<148 * 16, 256, 0, st>>>(firstBits, count, chunkLog2, dSums);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_sincos_values(unsigned firstBits, int n, float *dS, float *dC, cudaStream_t st) {
  k_sincos_values<<<(n + 255) / 256, 256, 0, st>>>(firstBits, n, dS, dC);
  count_launch();
  return cudaGetLastError();
}
#endif

// Fills this variant's d_patternT on the current device.  Called from orbgpu_extractor_create — outside any stream capture
// (small batches replay a captured graph, and a synchronize inside the capture would invalidate it) — and complete on
// return, so every stream that launches k_describe afterwards sees the table.
cudaError_t VNAME(describe_prepare)() {
  static std::mutex mu;
  static bool ready[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (ready[dev & 63]) return cudaSuccess;
  k_init_pattern<<<2, 256>>>();
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) ready[dev & 63] = true;
  return e;
}

int VNAME(describe_smem_bytes)() { return 128 + kWarpsPerCta * kWarpSlot; }
cudaError_t VNAME(describe_configure)() {
  return cudaFuncSetAttribute(k_describe, cudaFuncAttributeMaxDynamicSharedMemorySize, VNAME(describe_smem_bytes)());
}

cudaError_t VNAME(launch_describe)(const LevelTable &T, const ExtractBuffers &B, const TmaMaps &pyrMaps, const TmaMaps &blurMaps, int imgBase,
                                   int batch, cudaStream_t s) {
  dim3 grid((T.kpCap + kWarpsPerCta - 1) / kWarpsPerCta, batch);
  k_describe<<<grid, kWarpsPerCta * 32, VNAME(describe_smem_bytes)(), s>>>(pyrMaps, blurMaps, B.dT, B.sel, B.selCount, B.kps, B.desc, B.kpCount,
                                                                          B.status, imgBase);
  count_launch();
  return cudaGetLastError();
}

#ifndef ORBGPU_VARIANT_IS_SECONDARY
void describe_boxes(int *discW, int *discH, int *patchW, int *patchH) { *discW = kDiscW; *discH = kDiscRows; *patchW = kPatchW; *patchH = kPatchRows; }
#endif

}  // namespace orbgpu
