// orbgpu_oracle_backend.cpp — TEST INFRASTRUCTURE.  NOT part of the product and never linked into liborbgpu.so.
//
// Implements the entry points of include/orbgpu.h that shim/ORBmatcher.cc and shim/ORBVocabularyGPU.h call
// (orbgpu_features_create / _destroy, orbgpu_search, orbgpu_vocabulary_create / _destroy / _transform, orbgpu_last_error)
// on top of the CPU oracle (oracle_search.cpp), so that the HOST policy of the shim —
// query construction before the candidate loops, pointer-graph updates after them — can be compared with the
// reference's own ORBmatcher.cc (oracle/_ref/libref_slam.so) in the CPU test suite, and so that the oracle's
// restatement of the candidate loops is itself pinned to the reference.  The GPU tests link the same shim against the
// real liborbgpu.so instead.
#include <cstdint>
#include <map>
#include <mutex>
#include <vector>

#include "../include/orbgpu.h"

struct OracleFeatures;
extern "C" {
OracleFeatures *orb_oracle_features_create(const orbgpu_features_desc *desc);
void orb_oracle_features_destroy(OracleFeatures *F);
int orb_oracle_search(const OracleFeatures *F, const orbgpu_query *queries, const uint8_t *qdesc, int nq, const uint8_t *occupied,
                      const int32_t *cand, int ncand, const orbgpu_search_policy *P, orbgpu_match *result);

int orb_oracle_voc_transform(int nodes, const int32_t *childStart, const int32_t *children, const uint8_t *nodeDesc, const int32_t *wordId,
                             int depthLevels, const uint8_t *features, int n, int levelsup, int32_t *wordIds, int32_t *nodeIds);

struct OracleVoc { std::vector<int32_t> childStart, children, wordId; std::vector<uint8_t> desc; int depth; };
int orbgpu_vocabulary_create(int32_t nodes, const int32_t *child_start, const int32_t *children, const uint8_t *descriptors, const int32_t *word_id,
                             int32_t depth_levels, int, orbgpu_vocabulary **out) {
  OracleVoc *v = new OracleVoc;
  v->childStart.assign(child_start, child_start + nodes + 1);
  v->children.assign(children, children + child_start[nodes]);
  v->wordId.assign(word_id, word_id + nodes);
  v->desc.assign(descriptors, descriptors + (size_t)nodes * 32);
  v->depth = depth_levels;
  *out = reinterpret_cast<orbgpu_vocabulary *>(v);
  return ORBGPU_OK;
}
void orbgpu_vocabulary_destroy(orbgpu_vocabulary *v) { delete reinterpret_cast<OracleVoc *>(v); }
int orbgpu_vocabulary_transform(orbgpu_vocabulary *h, const uint8_t *descriptors, int n, int levelsup, int32_t *word_ids, int32_t *node_ids) {
  OracleVoc *v = reinterpret_cast<OracleVoc *>(h);
  return orb_oracle_voc_transform((int)v->wordId.size(), v->childStart.data(), v->children.data(), v->desc.data(), v->wordId.data(), v->depth, descriptors, n,
                                  levelsup, word_ids, node_ids);
}

// Lifetime audit (what a freed device buffer would punish on the GPU, made observable on the CPU): every live set has a
// count of searches running on it; destroying a set that is being searched, or searching one that is not alive, is a
// violation the concurrency test of tests/test_matcher_ref.py reads back.
static std::mutex g_liveMutex;
static std::map<orbgpu_features *, int> g_live;
static int g_violations = 0;
int orbgpu_oracle_backend_violations(void) { std::lock_guard<std::mutex> l(g_liveMutex); return g_violations; }

int orbgpu_features_create(const orbgpu_features_desc *desc, orbgpu_features **out) {
  *out = reinterpret_cast<orbgpu_features *>(orb_oracle_features_create(desc));
  std::lock_guard<std::mutex> l(g_liveMutex);
  g_live[*out] = 0;
  return ORBGPU_OK;
}
void orbgpu_features_destroy(orbgpu_features *f) {
  {
    std::lock_guard<std::mutex> l(g_liveMutex);
    auto it = g_live.find(f);
    if (it == g_live.end() || it->second != 0) g_violations++;
    if (it != g_live.end()) g_live.erase(it);
  }
  orb_oracle_features_destroy(reinterpret_cast<OracleFeatures *>(f));
}
int orbgpu_search(orbgpu_features *f, const orbgpu_query *queries, const uint8_t *query_descriptors, int nq, const uint8_t *occupied,
                  const int32_t *cand, int ncand, const orbgpu_search_policy *policy, orbgpu_match *result, int *iterations) {
  if (iterations) *iterations = 1;
  {
    std::lock_guard<std::mutex> l(g_liveMutex);
    auto it = g_live.find(f);
    if (it == g_live.end()) { g_violations++; return ORBGPU_E_INVALID; }
    it->second++;
  }
  const int rc = orb_oracle_search(reinterpret_cast<const OracleFeatures *>(f), queries, query_descriptors, nq, occupied, cand, ncand, policy, result);
  {
    std::lock_guard<std::mutex> l(g_liveMutex);
    auto it = g_live.find(f);
    if (it == g_live.end()) g_violations++; else it->second--;
  }
  return rc;
}
// ---- KeyFrameDatabase scoring (shim/KeyFrameDatabase.cc on the CPU): the database as the same CSR the device keeps
void orb_oracle_bow_query(int slots, const int *start, const uint32_t *words, const double *weights, const uint8_t *alive,
                          const uint32_t *qWords, const double *qWeights, int nq, int32_t *common, uint32_t *firstWord, double *score);
struct OracleBowDb { std::vector<int> start{0}; std::vector<uint32_t> words; std::vector<double> weights; std::vector<uint8_t> alive; };
int orbgpu_bowdb_create(int, orbgpu_bowdb **out) { *out = reinterpret_cast<orbgpu_bowdb *>(new OracleBowDb); return ORBGPU_OK; }
void orbgpu_bowdb_destroy(orbgpu_bowdb *db) { delete reinterpret_cast<OracleBowDb *>(db); }
int orbgpu_bowdb_slots(const orbgpu_bowdb *db) { return (int)reinterpret_cast<const OracleBowDb *>(db)->alive.size(); }
int orbgpu_bowdb_add(orbgpu_bowdb *h, const uint32_t *word_ids, const double *weights, int n) {
  OracleBowDb *db = reinterpret_cast<OracleBowDb *>(h);
  db->words.insert(db->words.end(), word_ids, word_ids + n);
  db->weights.insert(db->weights.end(), weights, weights + n);
  db->start.push_back((int)db->words.size());
  db->alive.push_back(1);
  return (int)db->alive.size() - 1;
}
int orbgpu_bowdb_erase(orbgpu_bowdb *h, int slot) { reinterpret_cast<OracleBowDb *>(h)->alive[slot] = 0; return ORBGPU_OK; }
int orbgpu_bowdb_clear(orbgpu_bowdb *h) { *reinterpret_cast<OracleBowDb *>(h) = OracleBowDb(); return ORBGPU_OK; }
int orbgpu_bowdb_query(orbgpu_bowdb *h, const uint32_t *word_ids, const double *weights, int n, int32_t *common_words, uint32_t *first_common_word,
                       double *score, int capacity) {
  OracleBowDb *db = reinterpret_cast<OracleBowDb *>(h);
  if (capacity < (int)db->alive.size()) return ORBGPU_E_CAPACITY;
  orb_oracle_bow_query((int)db->alive.size(), db->start.data(), db->words.data(), db->weights.data(), db->alive.data(), word_ids, weights, n,
                       common_words, first_common_word, score);
  return ORBGPU_OK;
}
const char *orbgpu_last_error(void) { return "oracle backend"; }
}
