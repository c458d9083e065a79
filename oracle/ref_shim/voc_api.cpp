// voc_api.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN vocabulary code: ORBVocabulary.cc
// (Source/Libraries/ORB_SLAM2/src) and the vendored DBoW2 (Source/ThirdParty/DBoW2/DBoW2-local: TemplatedVocabulary.h,
// FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp), all compiled unmodified from /root/reference against
// the OpenCV API stand-in.  Three builds (Makefile): libref_voc.so = reference only; libshim_voc_cpu.so /
// libshim_voc_gpu.so = the same plus shim/ORBVocabularyGPU.h on the CPU oracle backend / on liborbgpu.so.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define private public
#define protected public
#include "ORBVocabulary.h"
#ifdef VOC_SHIM
#include "ORBVocabularyGPU.h"
#endif
#undef private
#undef protected

#ifdef VOC_SHIM
typedef ORB_SLAM2::ORBVocabularyGPU Voc;
#else
typedef ORB_SLAM2::ORBVocabulary Voc;
#endif

namespace {
std::vector<cv::Mat> rows_of(const uint8_t *desc, int n) {
  std::vector<cv::Mat> v;
  for (int i = 0; i < n; i++) v.push_back(cv::Mat(1, 32, CV_8U, (void *)(desc + (size_t)i * 32)));
  return v;
}
}  // namespace

extern "C" {

void *voc_load_text(const char *path) {
  Voc *v = new Voc();
  if (!v->loadFromTextFile(path) || v->empty()) { delete v; return nullptr; }
  return v;
}
void voc_destroy(void *h) { delete (Voc *)h; }
int voc_nodes(void *h) { return (int)((Voc *)h)->m_nodes.size(); }
int voc_words(void *h) { return (int)((Voc *)h)->m_words.size(); }
int voc_depth(void *h) { return ((Voc *)h)->m_L; }
void voc_export(void *h, int *childStart, int *children, uint8_t *desc, int *wordId, double *weight) {
  Voc *v = (Voc *)h;
  int k = 0;
  for (size_t i = 0; i < v->m_nodes.size(); i++) {
    childStart[i] = k;
    for (DBoW2::NodeId c : v->m_nodes[i].children) children[k++] = (int)c;
    wordId[i] = v->m_nodes[i].isLeaf() ? (int)v->m_nodes[i].word_id : -1;
    weight[i] = v->m_nodes[i].weight;
    if (i > 0) std::memcpy(desc + i * 32, v->m_nodes[i].descriptor.ptr(0), 32); else std::memset(desc, 0, 32);
  }
  childStart[v->m_nodes.size()] = k;
}
// The reference's single-feature descent (TemplatedVocabulary.h:1216-1258), feature by feature; nids[i] stays 0xffffffff
// where the reference leaves *nid untouched (leaf above level L - levelsup).
void voc_transform_each(void *h, const uint8_t *desc, int n, int levelsup, int *words, double *weights, unsigned *nids) {
  Voc *v = (Voc *)h;
  std::vector<cv::Mat> f = rows_of(desc, n);
  for (int i = 0; i < n; i++) {
    DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid = 0xffffffffu;
    v->OrbVocabulary::transform(f[i], id, w, &nid, levelsup);
    words[i] = (int)id; weights[i] = w; nids[i] = nid;
  }
}
// transform(features, BowVector, FeatureVector, levelsup) through the VIRTUAL call, as Frame::ComputeBoW makes it
int voc_transform(void *h, const uint8_t *desc, int n, int levelsup, int *bowIds, double *bowVals, int *nbow, int *fvNodes, int *fvOffsets,
                  int *fvIdx, int *nfv) {
  ORB_SLAM2::ORBVocabulary *v = (Voc *)h;
  std::vector<cv::Mat> f = rows_of(desc, n);
  DBoW2::BowVector bow;
  DBoW2::FeatureVector fv;
  v->transform(f, bow, fv, levelsup);
  int k = 0;
  for (auto &e : bow) { bowIds[k] = (int)e.first; bowVals[k] = e.second; k++; }
  *nbow = k;
  int m = 0, j = 0;
  for (auto &e : fv) {
    fvNodes[m] = (int)e.first; fvOffsets[m] = j;
    for (unsigned idx : e.second) fvIdx[j++] = (int)idx;
    m++;
  }
  fvOffsets[m] = j;
  *nfv = m;
  return 0;
}

}  // extern "C"
