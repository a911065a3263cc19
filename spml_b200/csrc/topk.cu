// C3: top_k_ranking (reference spml/utils/segsort/eval.py:9-52) without the full
// [Q, M] argsort: each warp owns one query row, streams the prototype bank through
// shared memory, keeps a sorted top-k list per lane in registers and merges the 32
// lists with k rounds of a warp arg-max.  Order: similarity descending, lowest
// prototype index first on ties.
#include <math.h>

#include "common.cuh"

namespace spml {

constexpr int kTopkWarps = 8;
constexpr int kTopkCols = 64;  // prototypes staged per step

template <int KMAX>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ p, int64_t m,
            int dim, const int64_t* __restrict__ qlab, const int64_t* __restrict__ plab,
            const uint8_t* __restrict__ qvalid, const uint8_t* __restrict__ pvalid, int k,
            int64_t* __restrict__ topk_labels, int64_t* __restrict__ topk_index,
            int32_t* hit_count) {
  extern __shared__ float smem[];
  const int ldq = dim, ldp = dim + 1;
  float* Qs = smem;                          // [warps][dim]
  float* Ps = Qs + kTopkWarps * ldq;         // [kTopkCols][dim + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kTopkWarps + warp;
  const bool active = row < nq && (!qvalid || qvalid[row]);
  if (!__syncthreads_or(active)) return;     // fixed-capacity buffers: a block of padding rows

  for (int d = lane; d < dim; d += 32) Qs[warp * ldq + d] = active ? q[row * dim + d] : 0.f;

  float lv[KMAX];
  int li[KMAX];
#pragma unroll
  for (int s = 0; s < KMAX; ++s) lv[s] = -INFINITY, li[s] = 0x7fffffff;

  for (int64_t c0 = 0; c0 < m; c0 += kTopkCols) {
    const int cc = (int)min((int64_t)kTopkCols, m - c0);
    if (pvalid) {   // skip tiles made of dead columns only (block-uniform)
      const bool live = threadIdx.x < cc && pvalid[c0 + threadIdx.x];
      if (!__syncthreads_or(live)) continue;
    }
    __syncthreads();
    for (int j = warp; j < kTopkCols; j += kTopkWarps)
      for (int d = lane; d < dim; d += 32)
        Ps[j * ldp + d] = j < cc ? p[(c0 + j) * dim + d] : 0.f;
    __syncthreads();
    float a0 = 0.f, a1 = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float qv = Qs[warp * ldq + d];
      a0 = fmaf(qv, Ps[lane * ldp + d], a0);
      a1 = fmaf(qv, Ps[(lane + 32) * ldp + d], a1);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      const float v = h ? a1 : a0;
      if (j < cc && (!pvalid || pvalid[c0 + j]) && v > lv[KMAX - 1]) {
        // columns arrive in increasing index, so a strict '>' keeps the lowest index on ties
        lv[KMAX - 1] = v;
        li[KMAX - 1] = (int)(c0 + j);
#pragma unroll
        for (int s = KMAX - 1; s > 0; --s) {
          if (lv[s] > lv[s - 1]) {
            const float tv = lv[s]; lv[s] = lv[s - 1]; lv[s - 1] = tv;
            const int ti = li[s]; li[s] = li[s - 1]; li[s - 1] = ti;
          }
        }
      }
    }
  }

  int hits = 0;
  const int64_t ql = active ? qlab[row] : 0;
  for (int r = 0; r < k; ++r) {
    float bv = lv[0];
    int bi = li[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    if (li[0] == bi) {  // the winner pops its head
#pragma unroll
      for (int s = 0; s + 1 < KMAX; ++s) lv[s] = lv[s + 1], li[s] = li[s + 1];
      lv[KMAX - 1] = -INFINITY;
      li[KMAX - 1] = 0x7fffffff;
    }
    if (active && lane == 0) {
      const bool found = bi != 0x7fffffff;
      const int64_t lab = found ? plab[bi] : -1;
      topk_labels[row * k + r] = lab;
      if (topk_index) topk_index[row * k + r] = found ? bi : -1;
      hits += found && lab == ql;
    }
  }
  if (lane == 0 && hits) atomicAdd(hit_count, hits);
  if (lane == 0 && active) atomicAdd(hit_count + 1, 1);   // queries that took part
}

}  // namespace spml

extern "C" {

int spml_topk_ranking(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                      const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                      const uint8_t* pvalid, int k, int64_t* topk_labels, int64_t* topk_index,
                      int32_t* hit_count, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(nq >= 0 && m >= 0 && dim > 0 && k > 0 && hit_count, "topk_ranking: bad arguments");
  SPML_CHECK_SUPPORTED(k <= SPML_MAX_TOPK, "topk_ranking: k %d exceeds %d", k, SPML_MAX_TOPK);
  SPML_CHECK_SUPPORTED(dim <= 1024 && m < (1ll << 31), "topk_ranking: problem too large");
  SPML_CHECK_ARG(m >= k, "topk_ranking: fewer prototypes (%lld) than k (%d)", (long long)m, k);
  cudaStream_t st = as_stream(stream);
  SPML_CUDA(cudaMemsetAsync(hit_count, 0, 2 * sizeof(int32_t), st));
  if (nq == 0) return SPML_OK;
  SPML_CHECK_ARG(q && p && qlab && plab && topk_labels, "topk_ranking: null pointer");
  const size_t smem = ((size_t)kTopkWarps * dim + (size_t)kTopkCols * (dim + 1)) * sizeof(float);
  const unsigned blocks = (unsigned)ceil_div(nq, kTopkWarps);
  if (k <= 8) {
    SPML_CUDA(cudaFuncSetAttribute(topk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    topk_kernel<8><<<blocks, kTopkWarps * 32, smem, st>>>(q, nq, p, m, dim, qlab, plab, qvalid,
                                                          pvalid, k, topk_labels, topk_index,
                                                          hit_count);
  } else {
    SPML_CUDA(cudaFuncSetAttribute(topk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    topk_kernel<32><<<blocks, kTopkWarps * 32, smem, st>>>(q, nq, p, m, dim, qlab, plab, qvalid,
                                                           pvalid, k, topk_labels, topk_index,
                                                           hit_count);
  }
  SPML_LAUNCH_CHECK("topk_kernel");
  return SPML_OK;
}

}  // extern "C"
