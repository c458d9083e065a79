// orbgpu_oracle_backend.cpp — TEST INFRASTRUCTURE.  NOT part of the product and never linked into liborbgpu.so.
//
// Implements the four entry points of include/orbgpu.h that shim/ORBmatcher.cc calls (orbgpu_features_create / _destroy,
// orbgpu_search, orbgpu_last_error) on top of the CPU oracle (oracle_search.cpp), so that the HOST policy of the shim —
// query construction before the candidate loops, pointer-graph updates after them — can be compared with the
// reference's own ORBmatcher.cc (oracle/_ref/libref_slam.so) in the CPU test suite, and so that the oracle's
// restatement of the candidate loops is itself pinned to the reference.  The GPU tests link the same shim against the
// real liborbgpu.so instead.
#include <cstdint>

#include "../include/orbgpu.h"

struct OracleFeatures;
extern "C" {
OracleFeatures *orb_oracle_features_create(const orbgpu_features_desc *desc);
void orb_oracle_features_destroy(OracleFeatures *F);
int orb_oracle_search(const OracleFeatures *F, const orbgpu_query *queries, const uint8_t *qdesc, int nq, const uint8_t *occupied,
                      const int32_t *cand, int ncand, const orbgpu_search_policy *P, orbgpu_match *result);

int orbgpu_features_create(const orbgpu_features_desc *desc, orbgpu_features **out) {
  *out = reinterpret_cast<orbgpu_features *>(orb_oracle_features_create(desc));
  return ORBGPU_OK;
}
void orbgpu_features_destroy(orbgpu_features *f) { orb_oracle_features_destroy(reinterpret_cast<OracleFeatures *>(f)); }
int orbgpu_search(orbgpu_features *f, const orbgpu_query *queries, const uint8_t *query_descriptors, int nq, const uint8_t *occupied,
                  const int32_t *cand, int ncand, const orbgpu_search_policy *policy, orbgpu_match *result, int *iterations) {
  if (iterations) *iterations = 1;
  return orb_oracle_search(reinterpret_cast<const OracleFeatures *>(f), queries, query_descriptors, nq, occupied, cand, ncand, policy, result);
}
const char *orbgpu_last_error(void) { return "oracle backend"; }
}
