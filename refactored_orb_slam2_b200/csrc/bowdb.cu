// Bag-of-words keyframe database scoring: the data-parallel part of KeyFrameDatabase::DetectLoopCandidates and
// DetectRelocalizationCandidates (Source/Libraries/ORB_SLAM2/src/KeyFrameDatabase.cc:72-306) — for EVERY keyframe of the
// database at once, (1) the number of vocabulary words it shares with the query (what the inverted-file walk of
// :81-99 / :195-213 accumulates into mnLoopWords / mnRelocWords), (2) the smallest shared word (which fixes the position at
// which that walk meets the keyframe first, i.e. the order of lKFsSharingWords) and (3) DBoW2's L1 score
// (ThirdParty/DBoW2/DBoW2-local/src/ScoringObject.cpp:23-67) of the two BowVectors.
//
// Layout in HBM: the database is a CSR of the keyframes' BowVectors in insertion order — start[slots + 1], word ids (u32,
// ascending inside a keyframe: a BowVector is a std::map) and tf-idf weights (f64) — plus one "alive" byte per slot; slots
// are handed out in add() order and never reused, so slot order is the push_back order of every inverted-file list.
// One warp per keyframe: 32 of its words per round are looked up in the query by binary search; the terms
// |v - w| - |v| - |w| of the shared words are then added one by one in ascending word order — the order of DBoW2's merge
// loop — so the double-precision sum is DBoW2's bit for bit, and the score converts to float like `float si = score(...)`.
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.cuh"

namespace orbgpu {
namespace {

__global__ void __launch_bounds__(128)
k_bowdb_query(const int *__restrict__ start, const unsigned *__restrict__ words, const double *__restrict__ weights,
              const uint8_t *__restrict__ alive, int slots, const unsigned *__restrict__ qWords, const double *__restrict__ qWeights,
              int nq, int *__restrict__ common, unsigned *__restrict__ firstWord, double *__restrict__ score) {
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (slot >= slots) return;
  int n = 0;
  unsigned first = 0xffffffffu;
  double sum = 0.0;
  if (alive[slot]) {
    const int s = start[slot], e = start[slot + 1];
    for (int base = s; base < e; base += 32) {
      const int i = base + lane;
      bool hit = false;
      unsigned w = 0u;
      double term = 0.0;
      if (i < e) {
        w = words[i];
        int lo = 0, hi = nq;                       // first query word >= w
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (qWords[mid] < w) lo = mid + 1; else hi = mid;
        }
        if (lo < nq && qWords[lo] == w) {
          hit = true;
          const double vi = qWeights[lo], wi = weights[i];          // score(query, keyframe): v1 = query, v2 = keyframe
          term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi));
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask && first == 0xffffffffu) first = __shfl_sync(0xffffffffu, w, __ffs((int)mask) - 1);
      n += __popc(mask);
      while (mask) {                               // ascending word order = DBoW2's accumulation order
        const int l = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, term, l));
      }
    }
  }
  if (lane == 0) {
    common[slot] = n;
    firstWord[slot] = first;
    score[slot] = -sum / 2.0;
  }
}

template <class T>
cudaError_t grow(T *&p, size_t &cap, size_t used, size_t need) {
  if (need <= cap) return cudaSuccess;
  size_t ncap = std::max(need, std::max(cap * 2, (size_t)1 << 16));
  T *q = nullptr;
  cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
  if (e != cudaSuccess) return e;
  if (used) e = cudaMemcpy(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice);
  cudaFree(p);
  p = q; cap = ncap;
  return e;
}

struct DevScope {
  int prev = -1; bool ok = true;
  explicit DevScope(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; } if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess; }
  ~DevScope() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace
}  // namespace orbgpu

using namespace orbgpu;

struct orbgpu_bowdb {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<int> hStart{0};               // host mirror of the CSR offsets
  int *dStart = nullptr; size_t startCap = 0;
  unsigned *dWords = nullptr; size_t wordsCap = 0;
  double *dWeights = nullptr; size_t weightsCap = 0;
  uint8_t *dAlive = nullptr; size_t aliveCap = 0;
  // query / result scratch
  unsigned *dQWords = nullptr; double *dQWeights = nullptr; size_t qCap = 0;
  int *dCommon = nullptr; unsigned *dFirst = nullptr; double *dScore = nullptr; size_t resCap = 0;
  int slots() const { return (int)hStart.size() - 1; }
};

extern "C" {

int orbgpu_bowdb_create(int device, orbgpu_bowdb **out) {
  if (!out) { set_error("null output"); return ORBGPU_E_INVALID; }
  *out = nullptr;
  if (orbgpu_device_count() < 1) { set_error("no CUDA device available (this library has no CPU path)"); return ORBGPU_E_CUDA; }
  DevScope guard(device);
  if (!guard.ok) { set_error("cannot select device %d", device); return ORBGPU_E_CUDA; }
  orbgpu_bowdb *db = new orbgpu_bowdb;
  cudaGetDevice(&db->device);
  if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess) { delete db; set_error("stream creation failed"); return ORBGPU_E_CUDA; }
  *out = db;
  return ORBGPU_OK;
}

void orbgpu_bowdb_destroy(orbgpu_bowdb *db) {
  if (!db) return;
  DevScope guard(db->device);
  cudaFree(db->dStart); cudaFree(db->dWords); cudaFree(db->dWeights); cudaFree(db->dAlive);
  cudaFree(db->dQWords); cudaFree(db->dQWeights); cudaFree(db->dCommon); cudaFree(db->dFirst); cudaFree(db->dScore);
  if (db->stream) cudaStreamDestroy(db->stream);
  delete db;
}

int orbgpu_bowdb_slots(const orbgpu_bowdb *db) { return db ? db->slots() : 0; }

int orbgpu_bowdb_add(orbgpu_bowdb *db, const uint32_t *word_ids, const double *weights, int n) {
  if (!db || n < 0 || (n > 0 && (!word_ids || !weights))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  for (int i = 1; i < n; i++)
    if (word_ids[i] <= word_ids[i - 1]) { set_error("word ids must be strictly ascending (a BowVector is a std::map)"); return ORBGPU_E_INVALID; }
  DevScope guard(db->device);
  const int slot = db->slots();
  const size_t used = (size_t)db->hStart.back();
  ORBGPU_CUDA_CHECK(grow(db->dWords, db->wordsCap, used, used + n));
  ORBGPU_CUDA_CHECK(grow(db->dWeights, db->weightsCap, used, used + n));
  ORBGPU_CUDA_CHECK(grow(db->dStart, db->startCap, (size_t)slot, (size_t)slot + 2));      // start[slot] is rewritten below
  ORBGPU_CUDA_CHECK(grow(db->dAlive, db->aliveCap, (size_t)slot, (size_t)slot + 1));
  const int ends[2] = {(int)used, (int)used + n};
  const uint8_t one = 1;
  if (n) {
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dWords + used, word_ids, (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice, db->stream));
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dWeights + used, weights, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, db->stream));
  }
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dStart + slot, ends, 2 * sizeof(int), cudaMemcpyHostToDevice, db->stream));
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dAlive + slot, &one, 1, cudaMemcpyHostToDevice, db->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(db->stream));        // the caller's arrays may go away
  db->hStart.push_back((int)used + n);
  return slot;
}

int orbgpu_bowdb_erase(orbgpu_bowdb *db, int slot) {
  if (!db || slot < 0 || slot >= db->slots()) { set_error("slot out of range"); return ORBGPU_E_INVALID; }
  DevScope guard(db->device);
  const uint8_t zero = 0;
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dAlive + slot, &zero, 1, cudaMemcpyHostToDevice, db->stream));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(db->stream));
  return ORBGPU_OK;
}

int orbgpu_bowdb_clear(orbgpu_bowdb *db) {
  if (!db) { set_error("null database"); return ORBGPU_E_INVALID; }
  db->hStart.assign(1, 0);
  return ORBGPU_OK;
}

int orbgpu_bowdb_query(orbgpu_bowdb *db, const uint32_t *word_ids, const double *weights, int n, int32_t *common_words,
                       uint32_t *first_common_word, double *score, int capacity) {
  if (!db || n < 0 || (n > 0 && (!word_ids || !weights))) { set_error("bad arguments"); return ORBGPU_E_INVALID; }
  const int slots = db->slots();
  if (capacity < slots || (slots > 0 && (!common_words || !first_common_word || !score))) { set_error("result arrays hold %d slots, the database has %d", capacity, slots); return ORBGPU_E_CAPACITY; }
  for (int i = 1; i < n; i++)
    if (word_ids[i] <= word_ids[i - 1]) { set_error("word ids must be strictly ascending (a BowVector is a std::map)"); return ORBGPU_E_INVALID; }
  if (slots == 0) return ORBGPU_OK;
  DevScope guard(db->device);
  if ((size_t)n > db->qCap) {
    cudaFree(db->dQWords); cudaFree(db->dQWeights);
    db->qCap = std::max((size_t)n * 2, (size_t)4096);
    ORBGPU_CUDA_CHECK(cudaMalloc(&db->dQWords, db->qCap * sizeof(unsigned)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&db->dQWeights, db->qCap * sizeof(double)));
  }
  if ((size_t)slots > db->resCap) {
    cudaFree(db->dCommon); cudaFree(db->dFirst); cudaFree(db->dScore);
    db->resCap = std::max((size_t)slots * 2, (size_t)1024);
    ORBGPU_CUDA_CHECK(cudaMalloc(&db->dCommon, db->resCap * sizeof(int)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&db->dFirst, db->resCap * sizeof(unsigned)));
    ORBGPU_CUDA_CHECK(cudaMalloc(&db->dScore, db->resCap * sizeof(double)));
  }
  cudaStream_t s = db->stream;
  if (n) {
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dQWords, word_ids, (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice, s));
    ORBGPU_CUDA_CHECK(cudaMemcpyAsync(db->dQWeights, weights, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  k_bowdb_query<<<(slots + 3) / 4, 128, 0, s>>>(db->dStart, db->dWords, db->dWeights, db->dAlive, slots, db->dQWords, db->dQWeights, n, db->dCommon,
                                                db->dFirst, db->dScore);
  count_launch();
  ORBGPU_CUDA_CHECK(cudaGetLastError());
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(common_words, db->dCommon, (size_t)slots * sizeof(int), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(first_common_word, db->dFirst, (size_t)slots * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaMemcpyAsync(score, db->dScore, (size_t)slots * sizeof(double), cudaMemcpyDeviceToHost, s));
  ORBGPU_CUDA_CHECK(cudaStreamSynchronize(s));
  return ORBGPU_OK;
}

}  // extern "C"
