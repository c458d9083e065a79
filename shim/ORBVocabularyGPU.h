// shim/ORBVocabularyGPU.h — ORBVocabulary whose per-feature tree descent runs on the device.
//
// The reference creates its vocabulary in System.cc (`mpVocabulary = new ORBVocabulary();`, src/System.cc:68) and reaches
// it only through Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:412-417, src/KeyFrame.cc:63-72), which call the
// VIRTUAL TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
// (Source/ThirdParty/DBoW2/DBoW2-local/include/DBoW2/TemplatedVocabulary.h:133, body :1125-1190).  Replacing that one
// `new` by `new ORBVocabularyGPU()` is the whole integration: loading (text / binary) and scoring stay the reference's.
//
// The override keeps the reference's bookkeeping — addWeight / addFeature in feature order, the "stopped word" test,
// the normalisation through the scoring object — and replaces only the descent of every feature
// (TemplatedVocabulary.h:1216-1258: k FORB::distance evaluations per level) by one orbgpu_vocabulary_transform call on a
// copy of the tree kept in HBM (uploaded on first use and again whenever the node count changes).
#pragma once
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBVocabulary.h"
#include "orbgpu.h"

namespace ORB_SLAM2 {

class ORBVocabularyGPU : public ORBVocabulary {
 public:
  using ORBVocabulary::ORBVocabulary;
  using ORBVocabulary::transform;                         // keep the other overloads visible
  ~ORBVocabularyGPU() override { orbgpu_vocabulary_destroy(mpDevice); }

  void transform(const std::vector<cv::Mat> &features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const override {
    v.clear();
    fv.clear();
    if (empty()) return;                                  // TemplatedVocabulary.h:1131
    const int n = (int)features.size();
    std::vector<unsigned char> rows((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(&rows[(size_t)i * 32], features[i].ptr(0), 32);
    std::vector<int32_t> words(n), nodes(n);
    {
      std::lock_guard<std::mutex> lock(mMutex);
      Upload();
    }
    if (n > 0 && orbgpu_vocabulary_transform(mpDevice, rows.data(), n, levelsup, words.data(), nodes.data()) != ORBGPU_OK)
      throw std::runtime_error(std::string("orbgpu_vocabulary_transform: ") + orbgpu_last_error());

    DBoW2::LNorm norm;
    const bool must = m_scoring_object->mustNormalize(norm);
    const bool tf = m_weighting == DBoW2::TF || m_weighting == DBoW2::TF_IDF;
    for (int i = 0; i < n; i++) {
      const DBoW2::WordValue w = m_words[words[i]]->weight;     // idf (or 1): the leaf's weight, TemplatedVocabulary.h:1255
      if (!(w > 0)) continue;                                    // stopped word
      if (tf) v.addWeight((DBoW2::WordId)words[i], w); else v.addIfNotExist((DBoW2::WordId)words[i], w);
      fv.addFeature((DBoW2::NodeId)nodes[i], (unsigned int)i);
    }
    if (tf && !v.empty() && !must) {
      const double nd = v.size();
      for (DBoW2::BowVector::iterator vit = v.begin(); vit != v.end(); vit++) vit->second /= nd;
    }
    if (must) v.normalize(norm);
  }

 private:
  void Upload() const {
    if (mpDevice && mUploadedNodes == m_nodes.size()) return;
    orbgpu_vocabulary_destroy(mpDevice);
    mpDevice = nullptr;
    const int nn = (int)m_nodes.size();
    std::vector<int32_t> start(nn + 1, 0), kids, word(nn, -1);
    std::vector<unsigned char> desc((size_t)nn * 32, 0);
    for (int i = 0; i < nn; i++) {
      start[i] = (int32_t)kids.size();
      for (DBoW2::NodeId c : m_nodes[i].children) kids.push_back((int32_t)c);
      if (m_nodes[i].isLeaf()) word[i] = (int32_t)m_nodes[i].word_id;
      if (i > 0) std::memcpy(&desc[(size_t)i * 32], m_nodes[i].descriptor.ptr(0), 32);
    }
    start[nn] = (int32_t)kids.size();
    if (orbgpu_vocabulary_create(nn, start.data(), kids.data(), desc.data(), word.data(), m_L, -1, &mpDevice) != ORBGPU_OK)
      throw std::runtime_error(std::string("orbgpu_vocabulary_create: ") + orbgpu_last_error());
    mUploadedNodes = m_nodes.size();
  }
  mutable orbgpu_vocabulary *mpDevice = nullptr;
  mutable size_t mUploadedNodes = 0;
  mutable std::mutex mMutex;
};

}  // namespace ORB_SLAM2
