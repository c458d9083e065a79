// Orientation, descriptor and output assembly: IC_Angle / computeOrientation
// (ORBextractor.cc:76-100, 466-473), computeOrbDescriptor (ORBextractor.cc:102-146) and the
// tail of ORBextractor::operator() (ORBextractor.cc:1023-1037).
//
// One HALF-warp = one keypoint (two keypoints per warp): the per-keypoint scalar work — level lookup, cv::fastAtan2, the
// double-precision sincos, the output record — is issued once per warp instruction for two keypoints instead of one.
//  * both neighbourhoods are staged in the half-warp's shared-memory slot by TWO TMA tensor copies per keypoint (the 31 x 31
//    orientation disc from the level, the 37 x 37 patch from its blurred copy; pattern reach <= 18 px; box columns start on
//    a 16-byte boundary, so rows are 48 / 64 bytes) on one mbarrier per warp — the 512 rotated sample positions are a scatter
//    that would otherwise cost one 32-byte L2 sector per sample.
//  * IC_Angle: lane = disc ROW (two rows per lane): a row is eight words re-aligned to u = -15, masked to the circle (per-row byte masks,
//    L1-resident table) and reduced by two byte dot products per word — sum(u * I) with the constant weight words
//    -15 .. 15 and sum(I) with ones (IDP.4A) — instead of one byte load and two integer operations per pixel; integer
//    moments m10 / m01 reduced over the 16 lanes by shuffles (exact, order-free), then cv::fastAtan2's float polynomial with
//    round-to-nearest mul/add/div intrinsics (the reference build never fuses into FMAs).
//  * lane i of the half builds descriptor bytes i and i + 16 from 2 x 16 pattern points rotated by (cos, sin) of the angle.
//    cosf/sinf replicate glibc's double-precision sincosf algorithm (see sincos_glibc) so that cvRound(x*b + y*a) lands on
//    the same pixel as on the host.
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
// Byte masks of the orientation disc: word k of row i (v = i - 15) covers u = -15 + 4k .. -12 + 4k; a byte is 0xff where
// |u| <= umax[|v|] (ORBextractor.cc:449-463).  1 KB, filled with the pattern table.
__device__ __align__(16) unsigned d_discMask[31 * 8];
__global__ void k_init_pattern() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 512) d_patternT[(i & 15) * 32 + (i >> 4)] = make_float2((float)d_pattern[2 * i], (float)d_pattern[2 * i + 1]);
  if (i < 31 * 8) {
    const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int v = i / 8 - 15, k = i & 7, um = umax[v < 0 ? -v : v];
    unsigned m = 0u;
    for (int j = 0; j < 4; j++) {
      const int u = -15 + 4 * k + j;
      if (u >= -um && u <= um) m |= 0xffu << (8 * j);
    }
    d_discMask[i] = m;
  }
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
constexpr int kKpPerWarp = 2;
constexpr int kPatchR = 18;                    // |rotated pattern coordinate| <= round(13 * sqrt 2) = 18 < EDGE_THRESHOLD
constexpr int kPatchRows = 2 * kPatchR + 1;    // 37
// Both neighbourhoods of a keypoint arrive as TMA boxes (cp.async.bulk.tensor.3d on the level's [batch][rows][pitch] maps,
// one mbarrier per warp): the 31 x 31 orientation disc from the pyramid level and the 37 x 37 patch from the blurred level.
// Box columns start on the 16-byte boundary at or before the first needed pixel — a tensor copy's start address must be
// 16-byte aligned, which for these u8 planes means the innermost coordinate (boxes that started at the exact column failed on
// the B200, gpurun_out/r2n) — so the rows are 48 / 64 bytes and the disc's words are re-aligned by funnel shifts.
constexpr int kDiscW = 48, kPatchW = 64;       // 15 alignment bytes + 31 <= 48, + 37 <= 64
constexpr int kDiscRows = 2 * kHalfPatch + 1;
constexpr int kDiscBytes = kDiscW * kDiscRows, kPatchBytes = kPatchW * kPatchRows;
constexpr int kDiscSlot = (kDiscBytes + 127) & ~127, kPatchSlot = (kPatchBytes + 127) & ~127;
constexpr int kKpSlot = kDiscSlot + kPatchSlot;
constexpr int kWarpSlot = kKpPerWarp * kKpSlot;

__device__ __forceinline__ int dp4a_us(unsigned data, unsigned weights, int acc) {     // sum of u8 x s8 products
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(data), "r"(weights), "r"(acc));
  return d;
}

// CTAs of one image are numbered level by level: level l owns CTAs ctaStart[l] .. ctaStart[l + 1) (ceil(selCap / keypoints per
// CTA) each), so a warp knows its (level, node) from blockIdx alone and requests the node's record at once; the level's start
// in the level-major output only matters for the final stores.
struct DescribeCtas { int ctaStart[kMaxLevels + 1]; };

// blockIdx.x = CTA of the image (see DescribeCtas), blockIdx.y = image
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_describe(const __grid_constant__ TmaMaps pyrMaps, const __grid_constant__ TmaMaps blurMaps, const __grid_constant__ DescribeCtas ctas,
           const LevelTable *__restrict__ T, const int4 *__restrict__ selAll, const int *__restrict__ selCount,
           orbgpu_keypoint *__restrict__ kps, uint8_t *__restrict__ desc, int *__restrict__ kpCount, unsigned *__restrict__ status,
           int imgBase) {
  extern __shared__ uint8_t smemRaw[];
  __shared__ __align__(8) unsigned long long sBar[kWarpsPerCta];

  const int img = imgBase + blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int nlevels = T->nlevels;
  int level = 0;
#pragma unroll
  for (int l = 1; l < kMaxLevels; l++) level += (l < nlevels && (int)blockIdx.x >= ctas.ctaStart[l]) ? 1 : 0;
  const LevelGeom &G = T->lv[level];
  const int i0 = (((int)blockIdx.x - ctas.ctaStart[level]) * kWarpsPerCta + warp) * kKpPerWarp;    // the warp's first node of the level
  // two independent requests: the per-level octree counts (lane l holds level l's) and the node's record
  const int cnt = lane < nlevels ? selCount[img * kMaxLevels + lane] : 0;
  int4 s = selAll[G.selOffset + (size_t)img * G.selCap + min(i0 + half, G.selCap - 1)];   // x_level, y_level, response
  // the image's output rows are level-major: this level starts behind the lower levels' counts
  const int total = __reduce_add_sync(0xffffffffu, cnt);
  const int start = __reduce_add_sync(0xffffffffu, lane < level ? cnt : 0);
  const int levelCount = __shfl_sync(0xffffffffu, cnt, level);
  const int kpCap = T->kpCap;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    kpCount[img] = min(total, kpCap);
    if (total > kpCap) atomicOr(status, kStatusKpOverflow);
  }
  if (i0 >= levelCount || start + i0 >= kpCap) return;                  // warp-uniform
  const bool live = i0 + half < levelCount && start + i0 + half < kpCap;
  {
    // an idle second half repeats the first keypoint and stores nothing
    const int x0 = __shfl_sync(0xffffffffu, s.x, l16), y0 = __shfl_sync(0xffffffffu, s.y, l16), z0 = __shfl_sync(0xffffffffu, s.z, l16);
    if (!live) { s.x = x0; s.y = y0; s.z = z0; }
  }
  const int slot = start + i0 + (live ? half : 0);                      // output row of this half's keypoint

  // ---- stage both neighbourhoods: two tensor copies per keypoint, no per-row address arithmetic, one memory latency
  uint8_t *disc = smemRaw + ((128u - (smem_u32(smemRaw) & 127u)) & 127u) + warp * kWarpSlot + half * kKpSlot;
  uint8_t *patch = disc + kDiscSlot;
  const int dx0 = s.x + kEdge - kHalfPatch, px0 = s.x + kEdge - kPatchR;      // padded columns of u = -15 / of the patch's first pixel
  const int dBox = dx0 & ~15, pBox = px0 & ~15, lead = px0 & 15;
  if (lane == 0) {
    mbar_init(&sBar[warp], 1);
    mbar_fence_init();
    mbar_expect_tx(&sBar[warp], kKpPerWarp * (kDiscBytes + kPatchBytes));
  }
  __syncwarp();
  if (l16 == 0) {
    tma_load_3d(disc, &pyrMaps.m[level], dBox, s.y + kEdge - kHalfPatch, img, &sBar[warp]);
    tma_load_3d(patch, &blurMaps.m[level], pBox, s.y + kEdge - kPatchR, img, &sBar[warp]);
  }
  __syncwarp();
  mbar_wait(&sBar[warp], 0);

  // ---- IC_Angle: rows l16 and l16 + 16 of the disc
  int m10 = 0, m01 = 0;
#pragma unroll
  for (int rr = 0; rr < 2; rr++) {
    const int i = l16 + 16 * rr;
    if (i < kDiscRows) {
      unsigned D[8];
      const unsigned *rw = reinterpret_cast<const unsigned *>(disc + i * kDiscW) + ((dx0 & 15) >> 2);
      const unsigned sh = 8u * (unsigned)(dx0 & 3);
      unsigned w[9];
#pragma unroll
      for (int k = 0; k < 9; k++) w[k] = rw[k];
#pragma unroll
      for (int k = 0; k < 8; k++) D[k] = __funnelshift_r(w[k], w[k + 1], sh);
      const uint4 mlo = __ldg(reinterpret_cast<const uint4 *>(d_discMask + 8 * i)), mhi = __ldg(reinterpret_cast<const uint4 *>(d_discMask + 8 * i + 4));
      const unsigned mk[8] = {mlo.x, mlo.y, mlo.z, mlo.w, mhi.x, mhi.y, mhi.z, mhi.w};
      // weight words: the column offsets u = -15 + 4k + j as signed bytes (u = 16 is outside every mask)
      const unsigned wu[8] = {0xf4f3f2f1u, 0xf8f7f6f5u, 0xfcfbfaf9u, 0x00fffefdu, 0x04030201u, 0x08070605u, 0x0c0b0a09u, 0x000f0e0du};
      int rowSum = 0, rowM = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const unsigned d = D[k] & mk[k];
        rowM = dp4a_us(d, wu[k], rowM);
        rowSum = (int)__dp4a(d, 0x01010101u, (unsigned)rowSum);
      }
      m10 += rowM;
      m01 += (i - kHalfPatch) * rowSum;
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {                                      // within the half-warp
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
  unsigned val[2];
#pragma unroll
  for (int hb = 0; hb < 2; hb++) {
    const float2 *p = d_patternT + l16 + 16 * hb;         // point 16 * byte + j sits at p[32 * j], byte = l16 + 16 * hb
    unsigned v8 = 0;
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
      v8 |= (unsigned)(t[0] < t[1]) << bit;
    }
    val[hb] = v8;
  }
  if (!live) return;
  const size_t row = (size_t)img * kpCap + slot;
  desc[row * 32 + l16] = (uint8_t)val[0];
  desc[row * 32 + 16 + l16] = (uint8_t)val[1];
  if (l16 == 0) {
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
  DescribeCtas ctas{};
  const int perCta = kKpPerWarp * kWarpsPerCta;
  for (int l = 0; l < T.nlevels; l++) ctas.ctaStart[l + 1] = ctas.ctaStart[l] + (T.lv[l].selCap + perCta - 1) / perCta;
  for (int l = T.nlevels; l < kMaxLevels; l++) ctas.ctaStart[l + 1] = ctas.ctaStart[T.nlevels];
  dim3 grid(ctas.ctaStart[T.nlevels], batch);
  k_describe<<<grid, kWarpsPerCta * 32, VNAME(describe_smem_bytes)(), s>>>(pyrMaps, blurMaps, ctas, B.dT, B.sel, B.selCount, B.kps, B.desc,
                                                                          B.kpCount, B.status, imgBase);
  count_launch();
  return cudaGetLastError();
}

#ifndef ORBGPU_VARIANT_IS_SECONDARY
void describe_boxes(int *discW, int *discH, int *patchW, int *patchH) { *discW = kDiscW; *discH = kDiscRows; *patchW = kPatchW; *patchH = kPatchRows; }
#endif

}  // namespace orbgpu
