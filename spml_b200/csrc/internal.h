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

// segsort.cu --------------------------------------------------------------------------
int segsort_fwd_partial(const spml_segsort_desc* d, float* stats, float* nll, void* workspace,
                        size_t workspace_bytes, cudaStream_t st, const float** partial_out,
                        int* tiles_x_out);
// backward with d(prototypes) limited to the first `proto_rows` prototypes and, with
// `partial_out`, left as [*chunks_out][proto_rows][dim] partial sums in the workspace
int segsort_bwd_impl(const spml_segsort_desc* d, const float* stats, const float* grad_loss,
                     float beta, float* demb, int64_t ld_demb, float* dprotos, int64_t proto_rows,
                     const float** partial_out, int* chunks_out, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);
// out[i] = sum over the chunks of pa and pb (either may have 0 chunks)
int segsort_reduce_two(const float* pa, int ca, const float* pb, int cb, int64_t count, float* out,
                       cudaStream_t st);

// topk_tc.cu: top_k_ranking on the tensor cores for a large prototype bank (no masks / groups)
bool topk_tc_supported(int64_t nq, int64_t m, int dim, int k);
int topk_tc_launch(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                   const int64_t* qlab, const int64_t* plab, int k, int64_t* topk_labels,
                   int64_t* topk_index, int32_t* hit_count, cudaStream_t st);

// compact.cu: the phases of spml_unique_inverse (preparation / insertion, after which *count is
// final / ranking + inverse map)
int unique_prepare(bool has_hi, const int64_t* lo, int64_t n, const int32_t* n_dev, int64_t bound,
                   int32_t* count, void* workspace, size_t workspace_bytes, cudaStream_t st);
int unique_insert(const int64_t* hi, const int64_t* lo, int64_t n, const int32_t* n_dev,
                  int64_t bound, int32_t* count, void* workspace, cudaStream_t st);
// (the insertion for the keys (image, k-means cluster, label) of the clustering stage, built on
// the fly; its last block publishes {rows, segments, status} to device and pinned host memory)
int cluster_insert(const int32_t* km, const int64_t* batch, const int64_t* labels, int64_t n,
                   const int32_t* n_dev, int64_t num_clusters, int32_t* count, int64_t divisor,
                   int64_t* sem_out, int64_t* inst_out, int32_t* status, int32_t* counts_out,
                   int32_t* host_out, int32_t seq, void* workspace, cudaStream_t st);
int unique_finish(bool has_hi, int64_t n, const int32_t* n_dev, int64_t bound, int64_t* inverse,
                  int64_t* uniq_hi, int64_t* uniq_lo, const int32_t* count, int64_t* bound_out,
                  void* workspace, cudaStream_t st);

// normalize.cu: label packing + valid-pixel scan + normalise / pack in one launch (A8 front half)
int scan_normalize_pack(const float* emb, const float* loc, int64_t loc_batch_stride, int loc_ch,
                        const int64_t* labels, int has_ignore, int64_t ignore_index,
                        const int64_t* ignore_dev, const int64_t* sem, const int64_t* inst,
                        int64_t divisor, int64_t semantic_ignore, int64_t dropped,
                        const int64_t* seeds, int64_t seed_batch_stride, int batch, int dim, int n,
                        int64_t batch_index_offset, float eps, int32_t* dst, int32_t* img_off,
                        float* e, float* el, float* nx, float* nc, int64_t* labels_out,
                        int64_t* batch_out, int32_t* seed_out, void* workspace,
                        size_t workspace_bytes, cudaStream_t st);

#ifdef __CUDACC__
// The loss of one problem from the per-tile partial sums of its row losses (loss.py:149-190
// reduction; SPML_REDUCE_* in the header).  Executed by ONE warp; every lane gets the value.
__device__ inline float segsort_finalize_loss(const spml_segsort_desc& d,
                                              const float* __restrict__ partial, int tiles_x) {
  const int lane = threadIdx.x & 31;
  float total = 0.f;
  int nonempty = 0;
  for (int g = 0; g < d.num_groups; ++g) {
    float v = 0.f;
    for (int x = lane; x < tiles_x; x += 32) v += partial[(size_t)g * tiles_x + x];
    v = warp_sum(v);
    if (d.reduction == SPML_REDUCE_GROUP_MEAN && d.group_off) {
      const int ng = d.group_off[g + 1] - d.group_off[g];
      if (ng > 0) {
        total += v / (float)ng;
        ++nonempty;
      }
    } else {
      total += v;
    }
  }
  if (d.reduction == SPML_REDUCE_SUM) return total;
  if (d.reduction == SPML_REDUCE_GROUP_MEAN && d.group_off) return total / (float)nonempty;
  const int64_t rows =
      d.group_off ? (int64_t)d.group_off[d.num_groups] - d.group_off[0] : d.n_rows;
  return total / (float)rows;
}
#endif

}  // namespace spml
