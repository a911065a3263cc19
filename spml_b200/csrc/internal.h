// Internal interfaces between the translation units of libspml_b200 (not part of the ABI).
#pragma once

#include "common.cuh"

namespace spml {

// topk.cu ------------------------------------------------------------------------------
// Optional restrictions of the candidate set (nearest-neighbour tag propagation,
// reference spml/models/utils.py:157-223): a (query, prototype) pair only counts when both
// carry the same group id and the prototype's label is below `plab_limit`; `sim` receives
// the similarity of every retrieved prototype (-inf where fewer than k qualify).
struct TopkExtra {
  const int64_t* qgroup;   // [nq] or nullptr
  const int64_t* pgroup;   // [m]  or nullptr
  int64_t plab_limit;      // prototypes with plab >= limit are skipped when has_limit
  int has_limit;
  float* sim;              // [nq, k] or nullptr
};

int topk_launch(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                const uint8_t* pvalid, int k, int64_t* topk_labels, int64_t* topk_index,
                int32_t* hit_count, const TopkExtra& extra, cudaStream_t st);

}  // namespace spml
