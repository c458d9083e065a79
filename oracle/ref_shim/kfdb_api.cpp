// kfdb_api.cpp — TEST INFRASTRUCTURE.  C entry points around ORB_SLAM2::KeyFrameDatabase over the reference's real
// vocabulary code (ORBVocabulary.cc + the vendored DBoW2, see voc_api.cpp) and the KeyFrame / Frame slices of
// kfdb_standins/.  Three builds (Makefile): libref_kfdb.so = the reference's OWN, unmodified src/KeyFrameDatabase.cc;
// libshim_kfdb_cpu.so / libshim_kfdb_gpu.so = shim/KeyFrameDatabase.cc on the CPU oracle backend / on liborbgpu.so.
#include <cstdint>
#include <vector>

#include "KeyFrameDatabase.h"

using namespace ORB_SLAM2;

namespace {
struct Harness {
  ORBVocabulary *voc;
  KeyFrameDatabase db;
  std::vector<KeyFrame *> kfs;
  explicit Harness(ORBVocabulary *v) : voc(v), db(*v) {}
  ~Harness() { for (KeyFrame *k : kfs) delete k; }
  int index_of(KeyFrame *k) const { for (size_t i = 0; i < kfs.size(); i++) if (kfs[i] == k) return (int)i; return -1; }
};
void fill(DBoW2::BowVector &v, const uint32_t *ids, const double *w, int n) {
  v.clear();
  for (int i = 0; i < n; i++) v.insert(v.end(), std::make_pair((DBoW2::WordId)ids[i], (DBoW2::WordValue)w[i]));
}
}  // namespace

extern "C" {

void *kfdb_create(void *voc) { return new Harness((ORBVocabulary *)voc); }
void kfdb_destroy(void *h) { delete (Harness *)h; }
// a KeyFrame with this id and BowVector (not yet in the database); returns its index
int kfdb_new_keyframe(void *h, unsigned long id, const uint32_t *ids, const double *w, int n) {
  Harness *H = (Harness *)h;
  KeyFrame *k = new KeyFrame;
  k->mnId = id;
  fill(k->mBowVec, ids, w, n);
  H->kfs.push_back(k);
  return (int)H->kfs.size() - 1;
}
void kfdb_set_graph(void *h, int kf, const int *connected, int nc, const int *covisible, int ncv) {
  Harness *H = (Harness *)h;
  H->kfs[kf]->connected.clear(); H->kfs[kf]->covisible.clear();
  for (int i = 0; i < nc; i++) H->kfs[kf]->connected.insert(H->kfs[connected[i]]);
  for (int i = 0; i < ncv; i++) H->kfs[kf]->covisible.push_back(H->kfs[covisible[i]]);
}
void kfdb_add(void *h, int kf) { Harness *H = (Harness *)h; H->db.add(H->kfs[kf]); }
void kfdb_erase(void *h, int kf) { Harness *H = (Harness *)h; H->db.erase(H->kfs[kf]); }
void kfdb_clear(void *h) { ((Harness *)h)->db.clear(); }
int kfdb_detect_loop(void *h, int kf, float minScore, int *out, int cap) {
  Harness *H = (Harness *)h;
  std::vector<KeyFrame *> r = H->db.DetectLoopCandidates(H->kfs[kf], minScore);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = H->index_of(r[i]);
  return (int)r.size();
}
int kfdb_detect_reloc(void *h, unsigned long frameId, const uint32_t *ids, const double *w, int n, int *out, int cap) {
  Harness *H = (Harness *)h;
  Frame F;
  F.mnId = frameId;
  fill(F.mBowVec, ids, w, n);
  std::vector<KeyFrame *> r = H->db.DetectRelocalizationCandidates(&F);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = H->index_of(r[i]);
  return (int)r.size();
}
// the bookkeeping both detections leave on the KeyFrame objects: [loopQuery, loopWords, relocQuery, relocWords], [loopScore, relocScore]
void kfdb_state(void *h, int kf, long *ints, float *scores) {
  KeyFrame *k = ((Harness *)h)->kfs[kf];
  ints[0] = (long)k->mnLoopQuery; ints[1] = k->mnLoopWords; ints[2] = (long)k->mnRelocQuery; ints[3] = k->mnRelocWords;
  scores[0] = k->mLoopScore; scores[1] = k->mRelocScore;
}
double kfdb_voc_score(void *h, const uint32_t *ids1, const double *w1, int n1, const uint32_t *ids2, const double *w2, int n2) {
  DBoW2::BowVector a, b;
  fill(a, ids1, w1, n1); fill(b, ids2, w2, n2);
  return ((Harness *)h)->voc->score(a, b);
}

}  // extern "C"
