// ref_api.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own ORB_SLAM2::ORBextractor
// (compiled from /root/reference, see Makefile) + a monotone bump allocator for this library's operator new.
//
// Why the allocator: DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:659-661), so ties between
// nodes with equal key counts fall to heap addresses.  While an extraction runs, every allocation of the calling
// thread is served from an ever-growing arena owned by the extractor handle, so a later-created node always has a
// higher address — exactly the "creation sequence" rule the oracle and the CUDA kernel implement.  Each handle
// alternates between two arenas (objects of call k-1, e.g. the pyramid Mats, die during call k).  Outside an
// extraction, and for every other library in the process (-Bsymbolic), allocation is plain malloc/free.
#include <sys/mman.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ORBextractor.h"

extern "C" int orb_oracle_stereo_views(int nlevels, const uint8_t *const *planesL, const uint8_t *const *planesR, const int *w,
                                       const int *h, const int *strideL, const int *strideR, const float *scale,
                                       const float *invScale, const void *kL, const uint8_t *dL, int nL, const void *kR,
                                       const uint8_t *dR, int nR, float mbf, float mb, float *uRight, float *depth);

namespace {
constexpr size_t kSlab = (size_t)2 << 30;          // one arena: 2 GiB of address space, committed on touch
constexpr size_t kReserve = (size_t)2 << 40;       // 2 TiB reservation = 1024 slabs = 512 extractor handles ALIVE at once
char *g_base = nullptr;
size_t g_slabs = 0;                                // slabs ever carved out of the reservation (guarded by g_slabMutex)
size_t g_freeSlab[kReserve / kSlab];               // slabs handed back by destroyed handles, reused before new ones are carved
size_t g_nFree = 0;                                // (a plain array: this file replaces operator new, so no std::vector here)
std::mutex g_slabMutex;
std::atomic<int> g_init{0};

void reserve_once() {
  int expected = 0;
  if (g_init.compare_exchange_strong(expected, 1)) {
    void *p = mmap(nullptr, kReserve, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { fprintf(stderr, "ref arena: mmap failed\n"); abort(); }
    g_base = (char *)p;
    g_init.store(2);
  } else {
    while (g_init.load() != 2) std::this_thread::yield();
  }
}

struct Arena {
  char *slab[2] = {nullptr, nullptr};
  size_t used = 0;
  int cur = 0;
  void begin() {
    reserve_once();
    cur ^= 1;
    if (!slab[cur]) {
      std::lock_guard<std::mutex> lock(g_slabMutex);
      size_t i;
      if (g_nFree) {
        i = g_freeSlab[--g_nFree];
      } else {
        i = g_slabs++;
        if ((i + 1) * kSlab > kReserve) { fprintf(stderr, "ref arena: out of slabs (more than %zu arenas alive)\n", kReserve / kSlab); abort(); }
      }
      slab[cur] = g_base + i * kSlab;
    }
    used = 0;
  }
  // A handle's slabs go back to the free list when it dies; the pages it touched are dropped so that a process that
  // creates and destroys handles for hours (bench.py's reference arm: two per worker thread per call) stays bounded in
  // both address space and resident memory.  Objects still pointing into the slab (the extractor's pyramid Mats, destroyed
  // after the arena) are only ever passed to release(), which ignores arena addresses.
  ~Arena() {
    for (int k = 0; k < 2; k++) {
      if (!slab[k]) continue;
      madvise(slab[k], high[k] ? high[k] : 1, MADV_DONTNEED);
      std::lock_guard<std::mutex> lock(g_slabMutex);
      g_freeSlab[g_nFree++] = (size_t)(slab[k] - g_base) / kSlab;
    }
  }
  size_t high[2] = {0, 0};                          // high-water mark per slab = what MADV_DONTNEED has to cover
  void *take(size_t n) {
    n = (n + 15) & ~(size_t)15;
    if (used + n > kSlab) { fprintf(stderr, "ref arena exhausted\n"); abort(); }
    void *p = slab[cur] + used;
    used += n;
    if (used > high[cur]) high[cur] = used;
    return p;
  }
};
thread_local Arena *t_active = nullptr;     // the arena serving this thread's allocations, if an extraction is running

inline bool in_arena(void *p) { return g_base && (char *)p >= g_base && (char *)p < g_base + kReserve; }
inline void *alloc(size_t n) {
  if (t_active) return t_active->take(n);
  void *p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
inline void release(void *p) { if (p && !in_arena(p)) free(p); }
}  // namespace

void *operator new(size_t n) { return alloc(n); }
void *operator new[](size_t n) { return alloc(n); }
void operator delete(void *p) noexcept { release(p); }
void operator delete[](void *p) noexcept { release(p); }
void operator delete(void *p, size_t) noexcept { release(p); }
void operator delete[](void *p, size_t) noexcept { release(p); }

struct RefKp { float x, y, size, angle, response; int octave, class_id; };
static_assert(sizeof(cv::KeyPoint) == sizeof(RefKp), "KeyPoint layout");

namespace {
struct RefHandle {
  ORB_SLAM2::ORBextractor ex;
  Arena arena;
  RefHandle(int nf, float sf, int nl, int ini, int mn) : ex(nf, sf, nl, ini, mn) {}
};

int extract_one(RefHandle &H, const unsigned char *img, int w, int hgt, int step, void *kps, unsigned char *desc, int cap) {
  ORB_SLAM2::ORBextractor &ex = H.ex;
  H.arena.begin();
  t_active = &H.arena;
  int n;
  {
    cv::Mat image(hgt, w, CV_8UC1, (void *)img, (size_t)step);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    ex(image, cv::Mat(), k, d);     // the reference's ORBextractor::operator() (ORBextractor.cc:978-1039)
    n = (int)k.size();
    if (n <= cap) {
      for (int i = 0; i < n; i++) {
        std::memcpy((char *)kps + (size_t)i * sizeof(RefKp), &k[i], sizeof(RefKp));
        std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
      }
    } else {
      n = -n;
    }
  }
  t_active = nullptr;
  return n;
}
}  // namespace

extern "C" {
void *ref_orb_create(int nf, float sf, int nl, int ini, int mn) { return new RefHandle(nf, sf, nl, ini, mn); }
void ref_orb_destroy(void *h) { delete (RefHandle *)h; }
int ref_orb_extract(void *h, const unsigned char *img, int w, int hgt, int step, void *kps, unsigned char *desc, int cap) {
  return extract_one(*(RefHandle *)h, img, w, hgt, step, kps, desc, cap);
}
int ref_orb_level(void *h, int level, unsigned char *dst, int dstStep, int *w, int *hgt) {
  ORB_SLAM2::ORBextractor &ex = ((RefHandle *)h)->ex;
  if (level < 0 || level >= ex.GetLevels()) return -1;
  const cv::Mat &m = ex.mvImagePyramid[level];
  *w = m.cols; *hgt = m.rows;
  if (dst) for (int y = 0; y < m.rows; y++) std::memcpy(dst + (size_t)y * dstStep, m.ptr(y), m.cols);
  return 0;
}

// Whole stereo frames, one pair per thread as in Frame.cc:87-99: the reference's own extraction for the left and right
// image, then the oracle's restatement of Frame::ComputeStereoMatches on the reference-built pyramids.
void ref_orb_stereo_batch(int nf, float sf, int nl, int ini, int mn, const unsigned char *left, const unsigned char *right,
                          int npairs, int w, int hgt, long imgStride, float mbf, float mb, int cap, int *countsL, int *countsR,
                          int *matched, int nthreads) {
  nthreads = std::max(1, std::min(nthreads, npairs));
  std::atomic<int> next(0);
  auto worker = [&] {
    RefHandle HL(nf, sf, nl, ini, mn), HR(nf, sf, nl, ini, mn);
    ORB_SLAM2::ORBextractor &exL = HL.ex, &exR = HR.ex;
    std::vector<RefKp> kL(cap), kR(cap);
    std::vector<unsigned char> dL((size_t)cap * 32), dR((size_t)cap * 32);
    std::vector<float> u(cap), z(cap);
    const std::vector<float> scale = exL.GetScaleFactors(), inv = exL.GetInverseScaleFactors();
    for (int i; (i = next.fetch_add(1)) < npairs;) {
      const int a = extract_one(HL, left + (size_t)i * imgStride, w, hgt, w, kL.data(), dL.data(), cap);
      const int b = extract_one(HR, right + (size_t)i * imgStride, w, hgt, w, kR.data(), dR.data(), cap);
      countsL[i] = a; countsR[i] = b;
      matched[i] = 0;
      if (a > 0 && b > 0) {
        const unsigned char *pl[16], *pr[16];
        int lw[16], lh[16], sl[16], sr[16];
        for (int l = 0; l < nl; l++) {
          const cv::Mat &ml = exL.mvImagePyramid[l], &mr = exR.mvImagePyramid[l];
          pl[l] = ml.data; pr[l] = mr.data; lw[l] = ml.cols; lh[l] = ml.rows; sl[l] = (int)ml.step.v; sr[l] = (int)mr.step.v;
        }
        matched[i] = orb_oracle_stereo_views(nl, pl, pr, lw, lh, sl, sr, scale.data(), inv.data(), kL.data(), dL.data(), a,
                                             kR.data(), dR.data(), b, mbf, mb, u.data(), z.data());
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
  for (auto &x : th) x.join();
}
}
