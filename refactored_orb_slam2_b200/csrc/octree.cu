// ORBextractor::DistributeOctTree + ExtractorNode::DivideNode (ORBextractor.cc:475-731) on device.
//
// The reference walks a std::list, pushing children to the front and erasing parents; the result
// (which keypoint each node keeps AND the order of the nodes) depends on that list discipline.
// Here one CTA owns one (image, level) and keeps the node list as ARRAYS IN LIST ORDER in shared
// memory; every round is re-expressed as position arithmetic with block-wide scans
// (tests/octree_model.py is the executable specification, validated against the literal
// std::list restatement in oracle/oracle_cpu.cpp):
//
//   full round    L' = reverse(children in creation order) ++ (single-key nodes, old order)
//                 creation order = parents in list order, children n1..n4 (UL, UR, BL, BR)
//   careful round (entered when |L| + 3*nToExpand > N): the multi-key children of the previous
//                 round are processed by (key count desc, list position asc) until |L| >= N;
//                 L' = reverse(children of the processed nodes) ++ (all other nodes, old order).
//                 "position asc" == "created later first", the deterministic stand-in for the
//                 reference's heap-address tie-break (ORBextractor.cc:659-661).
//   end           every node keeps its first key with maximal response, "first" in the
//                 reference's candidate order (cell-major, then row-major inside the cell), which
//                 is recomputed from (x, y) so the candidate list itself may be in any order.
//
// All splits of a round are evaluated speculatively in parallel (child key counts by shared-memory
// atomics); an inclusive scan of (children - 1) in processing order finds the reference's
// `break` position.  Candidates carry their node's list position (u16) in an HBM scratch array.
#include "internal.cuh"

namespace orbgpu {

namespace {

// This translation unit is compiled twice: as is (throughput shape) and through *_lat.cu with a second thread shape for
// frame-at-a-time calls (a stereo pair keeps only 16-32 CTAs busy, so the latency build gives each more threads);
// ORBGPU_VARIANT suffixes the exported launchers.
#ifndef ORBGPU_VARIANT
#define ORBGPU_VARIANT
#endif
#define ORBGPU_CAT2(a, b) a##b
#define ORBGPU_CAT(a, b) ORBGPU_CAT2(a, b)
#define VNAME(x) ORBGPU_CAT(x, ORBGPU_VARIANT)
#ifndef ORBGPU_OCTREE_THREADS
#define ORBGPU_OCTREE_THREADS 256      // measured on B200 (KITTI batch): 128 -> 0.47 ms, 192 -> 0.41, 256 -> 0.34, 512 -> 0.40, 1024 -> 0.57
#endif
constexpr int kThreads = ORBGPU_OCTREE_THREADS;
constexpr int kWarps = kThreads / 32;

struct Smem {
  short4 *rect[2];            // x0, x1, y0, y1  (UL.x, UR.x, UL.y, BL.y)
  int *cnt[2];
  int4 *cc;                   // child key counts per processing slot
  int *slotOf;                // node -> processing slot, -1 = not split this round
  int *nodeAt;                // processing slot -> node
  int *base;                  // slot -> creation index of its first child
  int *keepPos;               // node -> new position if it is kept
  unsigned long long *key;    // sort keys; later per-node best (response, order)
};

__device__ __forceinline__ Smem carve(uint8_t *p, int M, int M2) {
  Smem s;
  s.key = reinterpret_cast<unsigned long long *>(p); p += (size_t)M2 * 8;
  s.cc = reinterpret_cast<int4 *>(p); p += (size_t)M * 16;
  s.rect[0] = reinterpret_cast<short4 *>(p); p += (size_t)M * 8;
  s.rect[1] = reinterpret_cast<short4 *>(p); p += (size_t)M * 8;
  s.cnt[0] = reinterpret_cast<int *>(p); p += (size_t)M * 4;
  s.cnt[1] = reinterpret_cast<int *>(p); p += (size_t)M * 4;
  s.slotOf = reinterpret_cast<int *>(p); p += (size_t)M * 4;
  s.nodeAt = reinterpret_cast<int *>(p); p += (size_t)M * 4;
  s.base = reinterpret_cast<int *>(p); p += (size_t)M * 4;
  s.keepPos = reinterpret_cast<int *>(p);
  return s;
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// In-place exclusive scan of a[0..n) by the whole CTA; returns the total.  sTmp: kWarps + 1 ints.
__device__ int block_scan_excl(int *a, int n, int *sTmp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + kThreads - 1) / kThreads;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; i++) sum += a[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sTmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kWarps ? sTmp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    if (lane < kWarps) sTmp[lane] = winc - w;
    if (lane == kWarps - 1) sTmp[kWarps] = winc;
  }
  __syncthreads();
  int run = sTmp[warp] + inc - sum;
  for (int i = lo; i < hi; i++) { const int v = a[i]; a[i] = run; run += v; }
  const int total = sTmp[kWarps];
  __syncthreads();
  return total;
}

__device__ __forceinline__ int quadrant(const short4 r, int x, int y, int &mx, int &my) {
  mx = r.x + ((r.y - r.x + 1) >> 1);   // x0 + ceil((x1 - x0) / 2.f), ORBextractor.cc:477
  my = r.z + ((r.w - r.z + 1) >> 1);   // y0 + ceil((y1 - y0) / 2.f), ORBextractor.cc:478
  return x < mx ? (y < my ? 0 : 2) : (y < my ? 1 : 3);   // ORBextractor.cc:508-516
}

__device__ __forceinline__ int nonempty_before(const int4 c, int q) {
  return (q > 0 && c.x > 0) + (q > 1 && c.y > 0) + (q > 2 && c.z > 0);
}

#ifndef ORBGPU_OCTREE_MIN_CTAS
#define ORBGPU_OCTREE_MIN_CTAS 6      // 40 registers: six 256-thread CTAs per SM instead of five (0.453 -> 0.432 ms per 640 images)
#endif
__global__ void __launch_bounds__(kThreads, ORBGPU_OCTREE_MIN_CTAS)
k_octree(const LevelTable *__restrict__ T, const Cand *__restrict__ candAll, const int *__restrict__ candCount,
         unsigned short *__restrict__ nodeOfAll, int4 *__restrict__ selAll, int *__restrict__ selCount,
         unsigned *__restrict__ status, int M, int M2, int imgBase) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  __shared__ int sTmp[kWarps + 2];
  __shared__ int sRootCnt[kMaxRoots], sRootPos[kMaxRoots];
  __shared__ int sSize, sNProc, sExpand;

  // x = image, y = level: CTAs are issued x-fastest, so every level-0 CTA (the longest: most candidates, most rounds) starts
  // before any level-7 one and the short ones fill the tail of the launch
  const int level = blockIdx.y, img = imgBase + blockIdx.x;
  const LevelGeom &G = T->lv[level];
  const int tid = threadIdx.x;
  const int N = G.quota;
  const int n = min(candCount[img * kMaxLevels + level], G.candCap);
  const Cand *cand = candAll + G.candOffset + (size_t)img * G.candCap;
  unsigned short *nodeOf = nodeOfAll + G.candOffset + (size_t)img * G.candCap;
  int4 *sel = selAll + G.selOffset + (size_t)img * G.selCap;
  int *outCount = selCount + img * kMaxLevels + level;
  if (n == 0 || G.nIni <= 0) {
    if (tid == 0) *outCount = 0;
    return;
  }
  Smem S = carve(smemRaw, M, M2);

  // ---- roots (ORBextractor.cc:535-572)
  if (tid < kMaxRoots) sRootCnt[tid] = 0;
  __syncthreads();
  for (int c = tid; c < n; c += kThreads) {
    const int x = cand[c].xy & 0xffff;
    int r = (int)__fdiv_rn((float)x, G.hX);
    r = min(max(r, 0), G.nIni - 1);
    atomicAdd(&sRootCnt[r], 1);
    nodeOf[c] = (unsigned short)r;
  }
  __syncthreads();
  if (tid == 0) {
    int pos = 0;
    for (int i = 0; i < G.nIni; i++) {
      sRootPos[i] = pos;
      if (sRootCnt[i] > 0) {
        S.rect[0][pos] = make_short4((short)G.rootX0[i], (short)G.rootX1[i], 0, (short)(G.maxBY - kMinBorder));
        S.cnt[0][pos] = sRootCnt[i];
        pos++;
      }
    }
    sSize = pos;
  }
  __syncthreads();
  for (int c = tid; c < n; c += kThreads) nodeOf[c] = (unsigned short)sRootPos[nodeOf[c]];
  __syncthreads();

  int cur = 0, size = sSize, Tprev = 0;
  bool careful = false, overflow = false;
  for (;;) {
    const int prevSize = size;
    short4 *rect = S.rect[cur];
    int *cnt = S.cnt[cur];
    int nP;   // nodes split speculatively this round, in processing order
    if (!careful) {
      // every multi-key node, list order
      for (int i = tid; i < size; i += kThreads) S.base[i] = cnt[i] > 1;
      __syncthreads();
      nP = block_scan_excl(S.base, size, sTmp);
      for (int i = tid; i < size; i += kThreads) {
        const bool multi = cnt[i] > 1;
        S.slotOf[i] = multi ? S.base[i] : -1;
        if (multi) S.nodeAt[S.base[i]] = i;
      }
    } else {
      // multi-key children of the previous round, sorted by (count desc, position asc)
      for (int i = tid; i < size; i += kThreads) { S.base[i] = (i < Tprev && cnt[i] > 1); S.slotOf[i] = -1; }
      __syncthreads();
      nP = block_scan_excl(S.base, size, sTmp);
      const int P2 = next_pow2(max(nP, 1));
      for (int i = tid; i < P2; i += kThreads) S.key[i] = 0ull;
      __syncthreads();
      for (int i = tid; i < Tprev; i += kThreads)
        if (cnt[i] > 1) S.key[S.base[i]] = ((unsigned long long)(unsigned)cnt[i] << 32) | (unsigned)(0xffffffffu - (unsigned)i);
      __syncthreads();
      for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < P2; i += kThreads) {
            const int o = i ^ j;
            if (o > i) {
              const unsigned long long a = S.key[i], b = S.key[o];
              const bool desc = (i & k) == 0;
              if ((a < b) == desc) { S.key[i] = b; S.key[o] = a; }
            }
          }
          __syncthreads();
        }
      }
      for (int j = tid; j < nP; j += kThreads) {
        const int i = (int)(0xffffffffu - (unsigned)(S.key[j] & 0xffffffffu));
        S.nodeAt[j] = i;
        S.slotOf[i] = j;
      }
    }
    for (int j = tid; j < nP; j += kThreads) S.cc[j] = make_int4(0, 0, 0, 0);
    __syncthreads();

    // ---- speculative split: child key counts
    for (int c = tid; c < n; c += kThreads) {
      const int i = nodeOf[c];
      const int j = S.slotOf[i];
      if (j < 0) continue;
      const unsigned xy = cand[c].xy;
      int mx, my;
      const int q = quadrant(rect[i], xy & 0xffff, xy >> 16, mx, my);
      atomicAdd(reinterpret_cast<int *>(&S.cc[j]) + q, 1);
    }
    __syncthreads();

    // ---- how many slots are really processed (careful rounds stop at |L| >= N)
    for (int j = tid; j < nP; j += kThreads) {
      const int4 c = S.cc[j];
      S.base[j] = (c.x > 0) + (c.y > 0) + (c.z > 0) + (c.w > 0);
    }
    if (tid == 0) { sNProc = nP; sExpand = 0; }
    __syncthreads();
    if (careful) {
      // inclusive running size after processing slot j: size + sum_{i<=j} (children_i - 1)
      for (int j = tid; j < nP; j += kThreads) S.keepPos[j] = S.base[j] - 1;
      __syncthreads();
      block_scan_excl(S.keepPos, nP, sTmp);
      for (int j = tid; j < nP; j += kThreads) {
        const int before = size + S.keepPos[j];
        const int after = before + S.base[j] - 1;
        if (after >= N && before < N) sNProc = j + 1;    // gains are >= 0, so this j is unique
      }
      __syncthreads();
    }
    const int nproc = sNProc;
    // creation index of each processed slot's first child
    for (int j = tid + nproc; j < nP; j += kThreads) S.base[j] = 0;
    __syncthreads();
    const int Tnew = block_scan_excl(S.base, nproc, sTmp);
    // kept nodes keep their relative order behind the new children
    for (int i = tid; i < size; i += kThreads) {
      const int j = S.slotOf[i];
      S.keepPos[i] = (j < 0 || j >= nproc) ? 1 : 0;
    }
    __syncthreads();
    const int nKeep = block_scan_excl(S.keepPos, size, sTmp);
    const int newSize = Tnew + nKeep;
    if (newSize > M) { overflow = true; break; }

    // ---- write the new list
    short4 *nrect = S.rect[cur ^ 1];
    int *ncnt = S.cnt[cur ^ 1];
    for (int i = tid; i < size; i += kThreads) {
      const int j = S.slotOf[i];
      if (j < 0 || j >= nproc) {
        const int p = Tnew + S.keepPos[i];
        S.keepPos[i] = p;
        nrect[p] = rect[i];
        ncnt[p] = cnt[i];
      }
    }
    int expand = 0;
    for (int j = tid; j < nproc; j += kThreads) {
      const short4 r = rect[S.nodeAt[j]];
      const int4 c = S.cc[j];
      const int mx = r.x + ((r.y - r.x + 1) >> 1), my = r.z + ((r.w - r.z + 1) >> 1);
      int k = S.base[j];
      if (c.x > 0) { const int p = Tnew - 1 - k++; nrect[p] = make_short4(r.x, (short)mx, r.z, (short)my); ncnt[p] = c.x; expand += c.x > 1; }
      if (c.y > 0) { const int p = Tnew - 1 - k++; nrect[p] = make_short4((short)mx, r.y, r.z, (short)my); ncnt[p] = c.y; expand += c.y > 1; }
      if (c.z > 0) { const int p = Tnew - 1 - k++; nrect[p] = make_short4(r.x, (short)mx, (short)my, r.w); ncnt[p] = c.z; expand += c.z > 1; }
      if (c.w > 0) { const int p = Tnew - 1 - k++; nrect[p] = make_short4((short)mx, r.y, (short)my, r.w); ncnt[p] = c.w; expand += c.w > 1; }
    }
    if (expand) atomicAdd(&sExpand, expand);
    __syncthreads();
    // ---- move the candidates
    for (int c = tid; c < n; c += kThreads) {
      const int i = nodeOf[c];
      const int j = S.slotOf[i];
      int p;
      if (j >= 0 && j < nproc) {
        const unsigned xy = cand[c].xy;
        int mx, my;
        const int q = quadrant(rect[i], xy & 0xffff, xy >> 16, mx, my);
        p = Tnew - 1 - (S.base[j] + nonempty_before(S.cc[j], q));
      } else {
        p = S.keepPos[i];
      }
      nodeOf[c] = (unsigned short)p;
    }
    __syncthreads();
    size = newSize;
    cur ^= 1;
    Tprev = Tnew;
    const int nToExpand = sExpand;
    __syncthreads();
    // ---- ORBextractor.cc:647-649 / :706-707
    if (size >= N || size == prevSize) break;
    if (!careful && size + 3 * nToExpand > N) careful = true;
  }

  if (overflow) {
    if (tid == 0) { atomicOr(status, kStatusNodeOverflow); *outCount = 0; }
    return;
  }

  // ---- keep the best key of every node (ORBextractor.cc:711-728)
  for (int i = tid; i < size; i += kThreads) S.key[i] = 0ull;
  __syncthreads();
  for (int c = tid; c < n; c += kThreads) {
    const Cand k = cand[c];
    const int x = (k.xy & 0xffff) - 3, y = (k.xy >> 16) - 3;     // relative to the first detection pixel
    const int cj = x / G.wCell, ci = y / G.hCell;
    const unsigned order = (unsigned)(((ci * G.nCols + cj) * G.hCell + (y - ci * G.hCell)) * G.wCell + (x - cj * G.wCell));
    const unsigned long long packed = ((unsigned long long)(k.resp + 1u) << 32) | (0xffffffffu - order);
    atomicMax(&S.key[nodeOf[c]], packed);
  }
  __syncthreads();
  const bool fits = size <= G.selCap;
  for (int c = tid; c < n; c += kThreads) {
    const Cand k = cand[c];
    const int x = (k.xy & 0xffff) - 3, y = (k.xy >> 16) - 3;
    const int cj = x / G.wCell, ci = y / G.hCell;
    const unsigned order = (unsigned)(((ci * G.nCols + cj) * G.hCell + (y - ci * G.hCell)) * G.wCell + (x - cj * G.wCell));
    const unsigned long long packed = ((unsigned long long)(k.resp + 1u) << 32) | (0xffffffffu - order);
    const int i = nodeOf[c];
    if (fits && S.key[i] == packed)
      sel[i] = make_int4((int)(k.xy & 0xffff) + kMinBorder, (int)(k.xy >> 16) + kMinBorder, (int)k.resp, 0);
  }
  if (tid == 0) {
    *outCount = fits ? size : 0;
    if (!fits) atomicOr(status, kStatusKpOverflow);
  }
}

}  // namespace

#ifndef ORBGPU_VARIANT_IS_SECONDARY
size_t octree_smem_bytes(int maxNodes) {
  int m2 = 1;
  while (m2 < maxNodes) m2 <<= 1;
  return (size_t)m2 * 8 + (size_t)maxNodes * (16 + 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4);
}

#endif
cudaError_t VNAME(octree_configure)(int smemBytes) {
  return cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
}

cudaError_t VNAME(launch_octree)(const LevelTable &T, const ExtractBuffers &B, int imgBase, int batch, int smemBytes,
                          cudaStream_t s) {
  int M = 0;
  for (int l = 0; l < T.nlevels; l++) M = max(M, T.lv[l].maxNodes);
  int M2 = 1;
  while (M2 < M) M2 <<= 1;
  dim3 grid(batch, T.nlevels);
  k_octree<<<grid, kThreads, smemBytes, s>>>(B.dT, B.cand, B.candCount, B.nodeOf, B.sel, B.selCount, B.status, M, M2, imgBase);
  count_launch();
  return cudaGetLastError();
}

}  // namespace orbgpu
