// ORBmatcher Hamming searches: DescriptorDistance (ORBmatcher.cc:1542-1556) and the
// best / second-best scan every ORBmatcher search runs over its candidates (ORBmatcher.cc:201-221:
// strict '<', so the first minimum in scan order wins; second = second-smallest of the multiset).
//
// Dense bank scan (BASELINE.json config 4: 2000 queries x 1M-row keyframe bank):
//   * a CTA owns a contiguous chunk of bank rows and a tile of 256*Q queries; each consumer thread
//     keeps Q query descriptors (8 words each) and their running (d1, idx1, d2) in registers;
//   * a producer warp streams the chunk through a 4-stage shared-memory ring with TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx); consumers read each 32-byte row as two broadcast
//     LDS.128 and release the stage through an "empty" mbarrier;
//   * distance = carry-save compression of the 8 XOR words (two levels of 3:2 adders in LOP3) down
//     to ones/ones/twos/fours words -> 4 POPC + 16 LOP3 per pair instead of 8 POPC + 8 LOP3, which
//     rebalances work from the quarter-rate POPC unit to the ALU pipe (variant 0 keeps the naive
//     8-POPC form for the roofline comparison);
//   * the top-2 update sits behind one compare (d < d2), which is rarely true after warm-up;
//   * every (chunk, query) emits a 16-byte record {d1, d2, idx}; a merge kernel folds chunks in
//     ascending bank order, which is exactly what the sequential scan would have produced.  The
//     same merge folds the per-GPU records of a row-sharded bank after the all-gather.
#include <cstdlib>

#include "internal.cuh"
#include "tma.cuh"

namespace orbgpu {

namespace {

constexpr int kConsumers = 256;
constexpr int kDenseThreads = kConsumers + 32;   // + one producer warp
constexpr int kTileRows = 128;
constexpr int kStages = 4;
constexpr int kTileBytes = kTileRows * 32;

template <int IMM>
__device__ __forceinline__ unsigned lop3(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(IMM));
  return d;
}

template <int VARIANT>
__device__ __forceinline__ int distance256(const unsigned (&q)[8], const uint4 b0, const uint4 b1) {
  const unsigned x0 = q[0] ^ b0.x, x1 = q[1] ^ b0.y, x2 = q[2] ^ b0.z, x3 = q[3] ^ b0.w;
  const unsigned x4 = q[4] ^ b1.x, x5 = q[5] ^ b1.y, x6 = q[6] ^ b1.z, x7 = q[7] ^ b1.w;
  if (VARIANT == 0) {          // 8 POPC + 8 LOP3: the reference's arithmetic, POPC-bound (quarter-rate unit)
    return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
  } else if (VARIANT == 1) {   // 4 POPC + 16 LOP3: two levels of 3:2 compression, ALU-bound
    const unsigned s0 = lop3<0x96>(x0, x1, x2), c0 = lop3<0xe8>(x0, x1, x2);
    const unsigned s1 = lop3<0x96>(x3, x4, x5), c1 = lop3<0xe8>(x3, x4, x5);
    const unsigned s2 = lop3<0x96>(s0, s1, x6), c2 = lop3<0xe8>(s0, s1, x6);
    const unsigned s4 = lop3<0x96>(c0, c1, c2), c4 = lop3<0xe8>(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(s4) + 4 * __popc(c4);
  } else if (VARIANT == 2) {   // 5 POPC + 14 LOP3: POPC (5/16 clk) and ALU (15/64 clk) nearly balanced
    const unsigned s0 = lop3<0x96>(x0, x1, x2), c0 = lop3<0xe8>(x0, x1, x2);
    const unsigned s1 = lop3<0x96>(x3, x4, x5), c1 = lop3<0xe8>(x3, x4, x5);
    const unsigned s2 = lop3<0x96>(s0, s1, x6), c2 = lop3<0xe8>(s0, s1, x6);
    return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
  } else {                     // 6 POPC + 12 LOP3
    const unsigned s0 = lop3<0x96>(x0, x1, x2), c0 = lop3<0xe8>(x0, x1, x2);
    const unsigned s1 = lop3<0x96>(x3, x4, x5), c1 = lop3<0xe8>(x3, x4, x5);
    return __popc(s0) + __popc(s1) + __popc(x6) + __popc(x7) + 2 * (__popc(c0) + __popc(c1));
  }
}

template <int Q, int VARIANT>
__global__ void __launch_bounds__(kDenseThreads, Q >= 8 ? 1 : 3)
k_hamming_dense(const uint8_t *__restrict__ bank, int rows, int chunkRows, long long indexBase,
                const uint8_t *__restrict__ queries, int nq, Top2Rec *__restrict__ partial) {
  __shared__ __align__(128) uint8_t sTile[kStages][kTileBytes];
  __shared__ __align__(8) unsigned long long sFull[kStages], sEmpty[kStages];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x;
  const int row0 = chunk * chunkRows;
  const int nrows = max(0, min(chunkRows, rows - row0));
  const int ntiles = (nrows + kTileRows - 1) / kTileRows;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(&sFull[s], 1); mbar_init(&sEmpty[s], kConsumers / 32); }
    mbar_fence_init();
  }
  __syncthreads();

  if (tid >= kConsumers) {
    // ---- producer warp: one elected lane keeps the ring full
    if (tid == kConsumers) {
      for (int t = 0; t < ntiles; t++) {
        const int s = t % kStages;
        if (t >= kStages) mbar_wait(&sEmpty[s], ((t / kStages) - 1) & 1);
        const int r = min(kTileRows, nrows - t * kTileRows);
        mbar_expect_tx(&sFull[s], (unsigned)r * 32u);
        bulk_g2s(sTile[s], bank + ((size_t)row0 + (size_t)t * kTileRows) * 32, (unsigned)r * 32u, &sFull[s]);
      }
    }
    return;
  }

  // ---- consumers
  const int qBase = blockIdx.y * (kConsumers * Q);
  unsigned q[Q][8];
  int d1[Q], d2[Q], i1[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int qi = qBase + j * kConsumers + tid;
    d1[j] = 256; d2[j] = 256; i1[j] = -1;
    if (qi < nq) {
      const uint4 a = reinterpret_cast<const uint4 *>(queries)[2 * (size_t)qi];
      const uint4 b = reinterpret_cast<const uint4 *>(queries)[2 * (size_t)qi + 1];
      q[j][0] = a.x; q[j][1] = a.y; q[j][2] = a.z; q[j][3] = a.w;
      q[j][4] = b.x; q[j][5] = b.y; q[j][6] = b.z; q[j][7] = b.w;
    } else {
#pragma unroll
      for (int w = 0; w < 8; w++) q[j][w] = 0;
    }
  }
  for (int t = 0; t < ntiles; t++) {
    const int s = t % kStages;
    mbar_wait(&sFull[s], (t / kStages) & 1);
    const uint4 *tile = reinterpret_cast<const uint4 *>(sTile[s]);
    const int r = min(kTileRows, nrows - t * kTileRows);
    const int rowIdx0 = t * kTileRows;
#pragma unroll 2
    for (int i = 0; i < r; i++) {
      const uint4 b0 = tile[2 * i], b1 = tile[2 * i + 1];
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const int d = distance256<VARIANT>(q[j], b0, b1);
        if (d < d2[j]) {
          if (d < d1[j]) { d2[j] = d1[j]; d1[j] = d; i1[j] = rowIdx0 + i; }
          else d2[j] = d;
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sEmpty[s]);
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int qi = qBase + j * kConsumers + tid;
    if (qi < nq) {
      Top2Rec o;
      o.d1 = d1[j]; o.d2 = d2[j];
      o.idx = i1[j] < 0 ? -1ll : indexBase + row0 + i1[j];
      partial[(size_t)chunk * nq + qi] = o;
    }
  }
}

// Folds `nparts` records per query in ascending part order == ascending bank index.  One warp per query:
// lanes stride over the parts, then a shuffle tree folds lane results; (d1, idx) lexicographic order stands in
// for "earlier part first" because bank indices grow with the part number.
__device__ __forceinline__ void fold_top2(int &d1, int &d2, long long &idx, int e1, int e2, long long eidx) {
  if (e1 < d1 || (e1 == d1 && eidx >= 0 && (idx < 0 || eidx < idx))) { d2 = min(d1, e2); d1 = e1; idx = eidx; }
  else d2 = min(d2, e1);
}

__global__ void __launch_bounds__(256)
k_top2_merge(const Top2Rec *__restrict__ parts, int nparts, int nq, Top2Rec *__restrict__ result,
             float nnratio, int thLow, long long *__restrict__ match) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  int d1 = 256, d2 = 256;
  long long idx = -1;
  for (int p = lane; p < nparts; p += 32) {
    const Top2Rec r = parts[(size_t)p * nq + qi];
    fold_top2(d1, d2, idx, r.d1, r.d2, r.idx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int e1 = __shfl_xor_sync(0xffffffffu, d1, o);
    const int e2 = __shfl_xor_sync(0xffffffffu, d2, o);
    const long long eidx = __shfl_xor_sync(0xffffffffu, idx, o);
    fold_top2(d1, d2, idx, e1, e2, eidx);
  }
  if (lane == 0) {
    if (result) { Top2Rec o; o.d1 = d1; o.d2 = d2; o.idx = idx; result[qi] = o; }
    if (match) {
      // ORBmatcher.cc:224-226
      const bool ok = d1 <= thLow && (float)d1 < __fmul_rn(nnratio, (float)d2);
      match[qi] = ok ? idx : -1ll;
    }
  }
}

__device__ __forceinline__ Top2Rec *window_records(char *base, const ExchangeView &x, unsigned epoch, int rank) {
  return reinterpret_cast<Top2Rec *>(base) + ((size_t)(epoch & 1u) * x.world + rank) * x.maxQueries;
}
__device__ __forceinline__ unsigned *window_flags(char *base, const ExchangeView &x, unsigned epoch) {
  return reinterpret_cast<unsigned *>(base + (size_t)2 * x.world * x.maxQueries * sizeof(Top2Rec) + 128) + (size_t)(epoch & 1u) * x.world;
}

// Folds this shard's chunk partials per query (as k_top2_merge) and stores the record into EVERY rank's window — peer
// stores over NVLink, 16 bytes per query and peer; the last CTA to finish publishes this rank's arrival flag everywhere.
__global__ void __launch_bounds__(256)
k_top2_push(const Top2Rec *__restrict__ parts, int nparts, int nq, const ExchangeView x, unsigned epoch) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi < nq) {
    int d1 = 256, d2 = 256;
    long long idx = -1;
    for (int p = lane; p < nparts; p += 32) {
      const Top2Rec r = parts[(size_t)p * nq + qi];
      fold_top2(d1, d2, idx, r.d1, r.d2, r.idx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int e1 = __shfl_xor_sync(0xffffffffu, d1, o);
      const int e2 = __shfl_xor_sync(0xffffffffu, d2, o);
      const long long eidx = __shfl_xor_sync(0xffffffffu, idx, o);
      fold_top2(d1, d2, idx, e1, e2, eidx);
    }
    // lanes 0 .. world-1 each serve one destination rank: one 16-byte store per peer
    if (lane < x.world) {
      int4 v;
      v.x = d1; v.y = d2; v.z = (int)(unsigned)(idx & 0xffffffffll); v.w = (int)(idx >> 32);
      *reinterpret_cast<int4 *>(window_records(x.peers[lane], x, epoch, x.rank) + qi) = v;
    }
  }
  __threadfence_system();                                  // this thread's peer stores are visible system-wide ...
  __syncthreads();                                         // ... for every thread of the CTA
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(x.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < x.world) {
    __threadfence_system();
    volatile unsigned *flag = window_flags(x.peers[threadIdx.x], x, epoch) + x.rank;
    *flag = epoch;                                         // arrival of rank x.rank in rank threadIdx.x's window
    if (threadIdx.x == 0) *x.done = 0u;                    // re-arm for the next step (stream order separates the steps)
  }
}

// Waits until every rank's records of `epoch` have arrived in the LOCAL window, then folds them in rank order.
__global__ void __launch_bounds__(256)
k_top2_wait_merge(int nq, const ExchangeView x, unsigned epoch, Top2Rec *__restrict__ result, float nnratio, int thLow,
                  long long *__restrict__ match) {
  if (threadIdx.x < x.world) {
    volatile unsigned *flag = window_flags(x.local, x, epoch) + threadIdx.x;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag != epoch) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) { atomicExch(x.error, 1u + (unsigned)threadIdx.x); break; }   // a peer never arrived: report, do not hang
      __nanosleep(200);
    }
    __threadfence_system();
  }
  __syncthreads();
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  int d1 = 256, d2 = 256;
  long long idx = -1;
  for (int r = lane; r < x.world; r += 32) {
    const int4 v = __ldcg(reinterpret_cast<const int4 *>(window_records(x.local, x, epoch, r) + qi));   // L2: where the peer stores land
    fold_top2(d1, d2, idx, v.x, v.y, ((long long)v.w << 32) | (unsigned)v.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int e1 = __shfl_xor_sync(0xffffffffu, d1, o);
    const int e2 = __shfl_xor_sync(0xffffffffu, d2, o);
    const long long eidx = __shfl_xor_sync(0xffffffffu, idx, o);
    fold_top2(d1, d2, idx, e1, e2, eidx);
  }
  if (lane == 0) {
    if (result) { Top2Rec o; o.d1 = d1; o.d2 = d2; o.idx = idx; result[qi] = o; }
    if (match) {
      const bool ok = d1 <= thLow && (float)d1 < __fmul_rn(nnratio, (float)d2);   // ORBmatcher.cc:224-226
      match[qi] = ok ? idx : -1ll;
    }
  }
}

// Candidate-list scan: one warp per query, lanes stride the list; merge keeps list order semantics.
__global__ void k_hamming_csr(const uint8_t *__restrict__ queries, int nq, const uint8_t *__restrict__ train,
                              const int *__restrict__ offsets, const int *__restrict__ cand, int *__restrict__ idx1,
                              int *__restrict__ od1, int *__restrict__ od2) {
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  const uint4 a0 = reinterpret_cast<const uint4 *>(queries)[2 * (size_t)qi];
  const uint4 a1 = reinterpret_cast<const uint4 *>(queries)[2 * (size_t)qi + 1];
  const int lo = offsets[qi], hi = offsets[qi + 1];
  int d1 = 256, d2 = 256, p1 = 0x7fffffff;
  for (int p = lo + lane; p < hi; p += 32) {
    const uint4 *t = reinterpret_cast<const uint4 *>(train) + 2 * (size_t)cand[p];
    const uint4 b0 = t[0], b1 = t[1];
    const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    if (d < d1) { d2 = d1; d1 = d; p1 = p; }
    else if (d < d2) d2 = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int e1 = __shfl_xor_sync(0xffffffffu, d1, o);
    const int e2 = __shfl_xor_sync(0xffffffffu, d2, o);
    const int q1 = __shfl_xor_sync(0xffffffffu, p1, o);
    if (e1 < d1 || (e1 == d1 && q1 < p1)) { d2 = min(d1, e2); d1 = e1; p1 = q1; }
    else d2 = min(d2, e1);
  }
  if (lane == 0) {
    idx1[qi] = d1 < 256 ? cand[p1] : -1;
    od1[qi] = d1;
    od2[qi] = d2;
  }
}

__global__ void k_pair_distance(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int n, int *__restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned *pa = reinterpret_cast<const unsigned *>(a) + 8 * (size_t)i;
  const unsigned *pb = reinterpret_cast<const unsigned *>(b) + 8 * (size_t)i;
  int d = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) d += __popc(pa[w] ^ pb[w]);
  dist[i] = d;
}

// Register-resident issue-rate probes: 8 independent dependency chains per thread.
// modes: 0 POPC, 1 LOP3, 2 VIMNMX3.S16x2, 3 VIMNMX.S16x2, 4 VIADD.16x2, 5 PRMT, 6 IMAD, 7 VIMNMX3 (s32), 8 IADD3
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe_probe(unsigned *out, int iters, unsigned seed) {
  unsigned v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = seed * (threadIdx.x + 1) + j * 0x9e3779b9u;
  const unsigned a = seed ^ 0x5bd1e995u, b = seed * 3u + 1u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (MODE == 0) asm volatile("popc.b32 %0, %0;" : "+r"(v[j]));
        else if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(a), "r"(b));
        else if (MODE == 2) v[j] = (u & 1) ? __vimin3_s16x2(v[j], v[(j + 1) & 7], v[(j + 2) & 7]) : __vimax3_s16x2(v[j], v[(j + 3) & 7], v[(j + 5) & 7]);
        else if (MODE == 3) v[j] = (u & 1) ? __vmins2(v[j], v[(j + 1) & 7]) : __vmaxs2(v[j], v[(j + 3) & 7]);
        else if (MODE == 4) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(v[j]) : "r"(a));
        else if (MODE == 5) asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(v[j]) : "r"(a));
        else if (MODE == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(a), "r"(b));
        else if (MODE == 7) v[j] = (u & 1) ? (unsigned)__vimin3_s32((int)v[j], (int)v[(j + 1) & 7], (int)v[(j + 2) & 7]) : (unsigned)__vimax3_s32((int)v[j], (int)v[(j + 3) & 7], (int)v[(j + 5) & 7]);
        else asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(v[j]) : "r"(a), "r"(b));
      }
    }
  }
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc ^= v[j];
  if (acc == 0x12345678u) out[0] = acc;
}

}  // namespace

int hamming_dense_q(int nq) {
  static const int forced = [] { const char *e = getenv("ORBGPU_HAMMING_Q"); return e ? atoi(e) : 0; }();
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  return nq > 4 * kConsumers ? 8 : nq > 2 * kConsumers ? 4 : nq > kConsumers ? 2 : 1;
}

int hamming_dense_chunks(int rows, int nq, int smCount) {
  const int Q = hamming_dense_q(nq);
  const int qTiles = (nq + kConsumers * Q - 1) / (kConsumers * Q);
  int chunks = max(1, (2 * smCount + qTiles - 1) / qTiles);
  int chunkRows = max((rows + chunks - 1) / chunks, kTileRows);
  return max(1, (rows + chunkRows - 1) / chunkRows);
}

template <int Q>
static void launch_dense_q(int variant, dim3 grid, cudaStream_t s, const uint8_t *bank, int rows, int chunkRows,
                           long long indexBase, const uint8_t *q, int nq, Top2Rec *partial) {
  switch (variant) {
    case 0: k_hamming_dense<Q, 0><<<grid, kDenseThreads, 0, s>>>(bank, rows, chunkRows, indexBase, q, nq, partial); break;
    case 1: k_hamming_dense<Q, 1><<<grid, kDenseThreads, 0, s>>>(bank, rows, chunkRows, indexBase, q, nq, partial); break;
    case 3: k_hamming_dense<Q, 3><<<grid, kDenseThreads, 0, s>>>(bank, rows, chunkRows, indexBase, q, nq, partial); break;
    default: k_hamming_dense<Q, 2><<<grid, kDenseThreads, 0, s>>>(bank, rows, chunkRows, indexBase, q, nq, partial); break;
  }
}

cudaError_t launch_hamming_scan_only(const uint8_t *bank, int rows, long long indexBase, const uint8_t *q, int nq, Top2Rec *partial,
                                     int chunks, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  static const int variant = [] { const char *e = getenv("ORBGPU_HAMMING_VARIANT"); return e ? atoi(e) : 2; }();
  if (rows > 0) {
    const int Q = hamming_dense_q(nq);
    const int chunkRows = (rows + chunks - 1) / chunks;
    dim3 grid(chunks, (nq + kConsumers * Q - 1) / (kConsumers * Q));
    switch (Q) {
      case 8: launch_dense_q<8>(variant, grid, s, bank, rows, chunkRows, indexBase, q, nq, partial); break;
      case 4: launch_dense_q<4>(variant, grid, s, bank, rows, chunkRows, indexBase, q, nq, partial); break;
      case 2: launch_dense_q<2>(variant, grid, s, bank, rows, chunkRows, indexBase, q, nq, partial); break;
      default: launch_dense_q<1>(variant, grid, s, bank, rows, chunkRows, indexBase, q, nq, partial); break;
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_hamming_dense(const uint8_t *bank, int rows, long long indexBase, const uint8_t *q,
                                 int nq, Top2Rec *partial, int chunks, Top2Rec *result, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  cudaError_t e = launch_hamming_scan_only(bank, rows, indexBase, q, nq, partial, chunks, s);
  if (e != cudaSuccess) return e;
  return launch_top2_merge(partial, rows > 0 ? chunks : 0, nq, result, 0.f, 0, nullptr, s);
}

// ---- sharded bank: window layout, push and wait + merge ------------------------------------------------------------
// window = [2 epochs][world][maxQueries] Top2Rec, then [2][world] arrival flags (u32, padded to 128 B)
size_t exchange_window_bytes(int world, int maxQueries) {
  return (size_t)2 * world * maxQueries * sizeof(Top2Rec) + 128 + (size_t)2 * world * sizeof(unsigned) + 128;
}
cudaError_t launch_top2_push(const Top2Rec *partial, int nparts, int nq, const ExchangeView &x, unsigned epoch, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  k_top2_push<<<(nq + 7) / 8, 256, 0, s>>>(partial, nparts, nq, x, epoch);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_top2_wait_merge(int nq, const ExchangeView &x, unsigned epoch, Top2Rec *result, float nnratio, int thLow,
                                   long long *match, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  k_top2_wait_merge<<<(nq + 7) / 8, 256, 0, s>>>(nq, x, epoch, result, nnratio, thLow, match);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_top2_merge(const Top2Rec *parts, int nparts, int nq, Top2Rec *result, float nnratio,
                              int thLow, long long *match, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  k_top2_merge<<<(nq + 7) / 8, 256, 0, s>>>(parts, nparts, nq, result, nnratio, thLow, match);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_hamming_csr(const uint8_t *q, int nq, const uint8_t *train, const int *offsets,
                               const int *cand, int *idx1, int *d1, int *d2, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  k_hamming_csr<<<(nq + 3) / 4, 128, 0, s>>>(q, nq, train, offsets, cand, idx1, d1, d2);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pair_distance(const uint8_t *a, const uint8_t *b, int n, int *dist, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_pair_distance<<<(n + 255) / 256, 256, 0, s>>>(a, b, n, dist);
  count_launch();
  return cudaGetLastError();
}

template <int MODE>
static void launch_probe(int blocks, unsigned *out, int iters, unsigned seed) { k_pipe_probe<MODE><<<blocks, 256>>>(out, iters, seed); }

// Lane-operations per second of one instruction kind (see k_pipe_probe for the mode numbers).
cudaError_t run_pipe_probe(int smCount, int mode, double *opsPerSec) {
  unsigned *out = nullptr;
  cudaError_t e = cudaMalloc(&out, 4);
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 4096, blocks = smCount * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a);
    switch (mode) {
      case 0: launch_probe<0>(blocks, out, iters, 12345u + rep); break;
      case 1: launch_probe<1>(blocks, out, iters, 12345u + rep); break;
      case 2: launch_probe<2>(blocks, out, iters, 12345u + rep); break;
      case 3: launch_probe<3>(blocks, out, iters, 12345u + rep); break;
      case 4: launch_probe<4>(blocks, out, iters, 12345u + rep); break;
      case 5: launch_probe<5>(blocks, out, iters, 12345u + rep); break;
      case 6: launch_probe<6>(blocks, out, iters, 12345u + rep); break;
      case 7: launch_probe<7>(blocks, out, iters, 12345u + rep); break;
      default: launch_probe<8>(blocks, out, iters, 12345u + rep); break;
    }
    cudaEventRecord(b);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  if (e != cudaSuccess) return e;
  *opsPerSec = (double)blocks * 256.0 * iters * 32.0 / (best * 1e-3);
  return cudaGetLastError();
}

cudaError_t run_popc_microbench(int smCount, double *popcPerSec, double *lop3PerSec) {
  cudaError_t e = run_pipe_probe(smCount, 0, popcPerSec);
  if (e != cudaSuccess) return e;
  return run_pipe_probe(smCount, 1, lop3PerSec);
}

}  // namespace orbgpu
