// C3: top_k_ranking (reference spml/utils/segsort/eval.py:9-52) without the full
// [Q, M] argsort: each warp owns one query row, streams prototype tiles through
// shared memory, keeps a sorted top-k list per lane in registers and merges the 32
// lists with k rounds of a warp arg-max.  Order: similarity descending, lowest
// prototype index first on ties.
//
// The prototype bank is split over a thread-block CLUSTER of 8 CTAs (CTA r takes the
// column tiles r, r + 8, ...): the problem is small (M^2 D flops) and latency-bound per
// tile, so the tiles of one query block run on 8 SMs at once.  Each CTA leaves its
// per-query top-k in its own shared memory; CTA 0 of the cluster merges the 8 lists
// through distributed shared memory.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace cg = cooperative_groups;

namespace spml {

constexpr int kTopkWarps = 8;
constexpr int kTopkCols = 64;  // prototypes staged per step
constexpr int kTopkCluster = 8;  // CTAs sharing one block of queries


template <int KMAX, int QW, bool kExtra>
__global__ void __launch_bounds__(kTopkWarps * 32, QW > 1 ? 2 : 1)
topk_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ p, int64_t m,
            int dim, const int64_t* __restrict__ qlab, const int64_t* __restrict__ plab,
            const uint8_t* __restrict__ qvalid, const uint8_t* __restrict__ pvalid, int k,
            int64_t* __restrict__ topk_labels, int64_t* __restrict__ topk_index,
            int32_t* hit_count, TopkExtra x) {
  // A warp owns QW query rows: a prototype value read from shared memory feeds QW FMAs
  // (the kernel is bound by shared-memory reads, not by the FMAs).
  constexpr int kQ = kTopkWarps * QW;        // queries per CTA (and per cluster)
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_val[kQ][KMAX];          // this CTA's top-k per query, read by cluster rank 0
  __shared__ int s_idx[kQ][KMAX];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int ldq = (dim + 3) & ~3, ldp = dim + 1;
  float* Qs = smem;                          // [kQ][ldq], zero padded to a multiple of 4
  float* Ps = Qs + kQ * ldq;                 // [kTopkCols][dim + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)(blockIdx.x / kTopkCluster) * kQ + warp * QW;
  bool active[QW];
  bool any = false;
  int qg[QW];   // group ids (image indices) fit 32 bits
#pragma unroll
  for (int i = 0; i < QW; ++i) {
    active[i] = row0 + i < nq && (!qvalid || qvalid[row0 + i]);
    any |= active[i];
    qg[i] = (kExtra && x.qgroup && active[i]) ? (int)x.qgroup[row0 + i] : 0;
  }
  // fixed-capacity buffers: a block of padding rows (the same for the whole cluster)
  if (!__syncthreads_or(any)) return;

#pragma unroll
  for (int i = 0; i < QW; ++i)
    for (int d = lane; d < ldq; d += 32)
      Qs[(warp * QW + i) * ldq + d] = (active[i] && d < dim) ? q[(row0 + i) * dim + d] : 0.f;

  float lv[QW][KMAX];
  int li[QW][KMAX];
#pragma unroll
  for (int i = 0; i < QW; ++i)
#pragma unroll
    for (int s = 0; s < KMAX; ++s) lv[i][s] = -INFINITY, li[i][s] = 0x7fffffff;

  // Which of this CTA's column tiles hold a live prototype (fixed-capacity banks are mostly
  // padding): one batch of loads up front instead of a dependent L2 round trip per tile.
  constexpr int kMaxTiles = 64;
  __shared__ int s_live[kMaxTiles];
  const int64_t stride = (int64_t)kTopkCluster * kTopkCols;
  const int64_t first = (int64_t)rank * kTopkCols;
  const int my_tiles = first < m ? (int)((m - first + stride - 1) / stride) : 0;
  const bool mask_known = pvalid && my_tiles <= kMaxTiles;
  if (mask_known) {
    if (threadIdx.x < kMaxTiles) s_live[threadIdx.x] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < my_tiles * kTopkCols; i0 += 4 * kTopkWarps * 32) {
      uint8_t f[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * kTopkWarps * 32 + threadIdx.x;
        const int64_t c = first + (int64_t)(idx >> 6) * stride + (idx & 63);
        f[u] = (idx < my_tiles * kTopkCols && c < m) ? pvalid[c] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * kTopkWarps * 32 + threadIdx.x;
        if (f[u]) s_live[idx >> 6] = 1;
      }
    }
    __syncthreads();
  }

  int tile = 0;
  for (int64_t c0 = first; c0 < m; c0 += stride, ++tile) {
    const int cc = (int)min((int64_t)kTopkCols, m - c0);
    if (mask_known) {
      if (!s_live[tile]) continue;   // block-uniform
    } else if (pvalid) {
      const bool live = threadIdx.x < cc && pvalid[c0 + threadIdx.x];
      if (!__syncthreads_or(live)) continue;
    }
    __syncthreads();
    if (dim <= 128) {
      // eight prototype rows per warp, all their loads in flight before the first store
      float v[kTopkCols / kTopkWarps][4];
#pragma unroll
      for (int i = 0; i < kTopkCols / kTopkWarps; ++i) {
        const int j = warp + i * kTopkWarps;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int d = lane + 32 * u;
          v[i][u] = (j < cc && d < dim) ? p[(c0 + j) * dim + d] : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < kTopkCols / kTopkWarps; ++i) {
        const int j = warp + i * kTopkWarps;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int d = lane + 32 * u;
          if (d < dim) Ps[j * ldp + d] = v[i][u];
        }
      }
    } else {
      for (int j = warp; j < kTopkCols; j += kTopkWarps)
        for (int d = lane; d < dim; d += 32)
          Ps[j * ldp + d] = j < cc ? p[(c0 + j) * dim + d] : 0.f;
    }
    __syncthreads();
    float acc[QW][2];
#pragma unroll
    for (int i = 0; i < QW; ++i) acc[i][0] = acc[i][1] = 0.f;
    const float* p0 = Ps + lane * ldp;
    const float* p1 = Ps + (lane + 32) * ldp;
    const float* qw = Qs + warp * QW * ldq;
    int d = 0;
    for (; d + 4 <= dim; d += 4) {   // d ascending: the same fmaf chain per (query, column) as ever
      const float b0[4] = {p0[d], p0[d + 1], p0[d + 2], p0[d + 3]};
      const float b1[4] = {p1[d], p1[d + 1], p1[d + 2], p1[d + 3]};
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const float4 q4 = *reinterpret_cast<const float4*>(qw + i * ldq + d);
        const float qa[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[i][0] = fmaf(qa[u], b0[u], acc[i][0]);
          acc[i][1] = fmaf(qa[u], b1[u], acc[i][1]);
        }
      }
    }
    for (; d < dim; ++d) {
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const float qv = qw[i * ldq + d];
        acc[i][0] = fmaf(qv, p0[d], acc[i][0]);
        acc[i][1] = fmaf(qv, p1[d], acc[i][1]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      bool col_ok = j < cc && (!pvalid || pvalid[c0 + j]);
      if (kExtra && col_ok && x.has_limit) col_ok = plab[c0 + j] < x.plab_limit;
      const int pg = (kExtra && x.pgroup && col_ok) ? (int)x.pgroup[c0 + j] : 0;
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const float v = acc[i][h];
        if (col_ok && (!kExtra || !x.pgroup || pg == qg[i]) && v > lv[i][KMAX - 1]) {
          // columns arrive in increasing index, so a strict '>' keeps the lowest index on ties
          lv[i][KMAX - 1] = v;
          li[i][KMAX - 1] = (int)(c0 + j);
#pragma unroll
          for (int s = KMAX - 1; s > 0; --s) {
            if (lv[i][s] > lv[i][s - 1]) {
              const float tv = lv[i][s]; lv[i][s] = lv[i][s - 1]; lv[i][s - 1] = tv;
              const int ti = li[i][s]; li[i][s] = li[i][s - 1]; li[i][s - 1] = ti;
            }
          }
        }
      }
    }
  }

  // ---- this CTA's top-k of each query (k rounds of a warp arg-max over the lane lists)
#pragma unroll
  for (int i = 0; i < QW; ++i) {
    for (int r = 0; r < k; ++r) {
      float bv = lv[i][0];
      int bi = li[i][0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        {   // bitwise predicates + selects: the short-circuit form compiles to divergent branches
          const bool take = (ov > bv) | ((ov == bv) & (oi < bi));
          bv = take ? ov : bv;
          bi = take ? oi : bi;
        }
      }
      if (li[i][0] == bi) {  // the winner pops its head
#pragma unroll
        for (int s = 0; s + 1 < KMAX; ++s) lv[i][s] = lv[i][s + 1], li[i][s] = li[i][s + 1];
        lv[i][KMAX - 1] = -INFINITY;
        li[i][KMAX - 1] = 0x7fffffff;
      }
      if (lane == 0) s_val[warp * QW + i][r] = bv, s_idx[warp * QW + i][r] = bi;
    }
  }
  cluster.sync();

  // ---- cluster rank 0 merges the 8 sorted lists of each query: lane l holds the list of CTA l
  if (rank == 0) {
    int hits = 0, took_part = 0;
#pragma unroll
    for (int i = 0; i < QW; ++i) {
      float mv[KMAX];
      int mi[KMAX];
#pragma unroll
      for (int s = 0; s < KMAX; ++s) mv[s] = -INFINITY, mi[s] = 0x7fffffff;
      if (lane < kTopkCluster) {
        const float* rv = cluster.map_shared_rank(&s_val[warp * QW + i][0], lane);
        const int* ri = cluster.map_shared_rank(&s_idx[warp * QW + i][0], lane);
#pragma unroll
        for (int s = 0; s < KMAX; ++s)
          if (s < k) mv[s] = rv[s], mi[s] = ri[s];
      }
      const int64_t row = row0 + i;
      const int64_t ql = active[i] ? qlab[row] : 0;
      for (int r = 0; r < k; ++r) {
        float bv = mv[0];
        int bi = mi[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          {   // bitwise predicates + selects: the short-circuit form compiles to divergent branches
          const bool take = (ov > bv) | ((ov == bv) & (oi < bi));
          bv = take ? ov : bv;
          bi = take ? oi : bi;
        }
        }
        if (mi[0] == bi) {
#pragma unroll
          for (int s = 0; s + 1 < KMAX; ++s) mv[s] = mv[s + 1], mi[s] = mi[s + 1];
          mv[KMAX - 1] = -INFINITY;
          mi[KMAX - 1] = 0x7fffffff;
        }
        if (active[i] && lane == 0) {
          const bool found = bi != 0x7fffffff;
          const int64_t lab = found ? plab[bi] : -1;
          topk_labels[row * k + r] = lab;
          if (topk_index) topk_index[row * k + r] = found ? bi : -1;
          if (kExtra && x.sim) x.sim[row * k + r] = found ? bv : -INFINITY;
          hits += found && lab == ql;
        }
      }
      took_part += active[i] ? 1 : 0;
    }
    if (lane == 0 && hits) atomicAdd(hit_count, hits);
    if (lane == 0 && took_part) atomicAdd(hit_count + 1, took_part);   // queries that took part
  }
  cluster.sync();   // the other CTAs' shared memory stays alive until rank 0 has read it
}

}  // namespace spml

namespace spml {

int topk_launch(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                const uint8_t* pvalid, int k, int64_t* topk_labels, int64_t* topk_index,
                int32_t* hit_count, const TopkExtra& extra, cudaStream_t st) {
  SPML_CHECK_ARG(nq >= 0 && m >= 0 && dim > 0 && k > 0 && hit_count, "topk_ranking: bad arguments");
  SPML_CHECK_SUPPORTED(k <= SPML_MAX_TOPK, "topk_ranking: k %d exceeds %d", k, SPML_MAX_TOPK);
  SPML_CHECK_SUPPORTED(dim <= 1024 && m < (1ll << 31), "topk_ranking: problem too large");
  SPML_CHECK_ARG(m >= k, "topk_ranking: fewer prototypes (%lld) than k (%d)", (long long)m, k);
  SPML_CUDA(cudaMemsetAsync(hit_count, 0, 2 * sizeof(int32_t), st));
  if (nq == 0) return SPML_OK;
  SPML_CHECK_ARG(q && p && qlab && plab && topk_labels, "topk_ranking: null pointer");
  const int qw = k <= 8 ? 4 : 1;   // query rows per warp (the long lists of k > 8 leave no registers)
  const size_t smem = ((size_t)kTopkWarps * qw * ((dim + 3) & ~3) + (size_t)kTopkCols * (dim + 1)) *
                      sizeof(float);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ceil_div(nq, kTopkWarps * qw) * kTopkCluster));
  cfg.blockDim = dim3(kTopkWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kTopkCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const bool extra_on = extra.qgroup || extra.pgroup || extra.has_limit || extra.sim;
  // a large bank without masks (retrieval inference): tcgen05 candidates + exact re-scoring
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  SPML_CUDA(cudaStreamIsCapturing(st, &capturing));   // (its scratch may have to be allocated)
  if (!extra_on && !qvalid && !pvalid && capturing == cudaStreamCaptureStatusNone &&
      topk_tc_supported(nq, m, dim, k))
    return topk_tc_launch(q, nq, p, m, dim, qlab, plab, k, topk_labels, topk_index, hit_count, st);
#define SPML_TOPK_LAUNCH(KMAXV, QWV, EXTRAV)                                                    \
  do {                                                                                          \
    SPML_CUDA(cudaFuncSetAttribute(topk_kernel<KMAXV, QWV, EXTRAV>,                             \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    SPML_CUDA(cudaLaunchKernelEx(&cfg, topk_kernel<KMAXV, QWV, EXTRAV>, q, nq, p, m, dim, qlab, \
                                 plab, qvalid, pvalid, k, topk_labels, topk_index, hit_count,   \
                                 extra));                                                       \
  } while (0)
  if (k <= 8) {
    if (extra_on) SPML_TOPK_LAUNCH(8, 4, true); else SPML_TOPK_LAUNCH(8, 4, false);
  } else {
    if (extra_on) SPML_TOPK_LAUNCH(32, 1, true); else SPML_TOPK_LAUNCH(32, 1, false);
  }
#undef SPML_TOPK_LAUNCH
  SPML_LAUNCH_CHECK("topk_kernel");
  return SPML_OK;
}

// f3: tags[q, c] = 1 iff one of the retrieved prototypes of q has label c and is at least
// `threshold` similar (models/utils.py:207-221).
__global__ void nn_tags_kernel(const int64_t* __restrict__ labels, const float* __restrict__ sim,
                               int64_t nq, int k, int num_classes, float threshold,
                               int64_t* __restrict__ tags, int64_t* __restrict__ masks) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nq) return;
  unsigned long long bits = 0;
  for (int i = 0; i < k; ++i) {
    const int64_t lab = labels[r * k + i];
    if (lab >= 0 && lab < num_classes && lab < 64 && !(sim[r * k + i] < threshold))
      bits |= 1ull << lab;
  }
  if (tags)
    for (int c = 0; c < num_classes; ++c) tags[r * num_classes + c] = c < 64 ? (bits >> c) & 1ull : 0;
  if (masks) masks[r] = (int64_t)bits;
}

}  // namespace spml

extern "C" {

int spml_topk_ranking(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                      const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                      const uint8_t* pvalid, int k, int64_t* topk_labels, int64_t* topk_index,
                      int32_t* hit_count, void* stream) {
  spml::TopkExtra none{};
  return spml::topk_launch(q, nq, p, m, dim, qlab, plab, qvalid, pvalid, k, topk_labels,
                           topk_index, hit_count, none, spml::as_stream(stream));
}

size_t spml_nn_multiset_labels_workspace_bytes(int64_t nq, int top_k) {
  if (nq <= 0 || top_k <= 0) return 16;
  return 16 + (size_t)nq * top_k * (sizeof(int64_t) + sizeof(float));
}

int spml_nn_multiset_labels(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                            const int64_t* plab, const int64_t* qgroup, const int64_t* pgroup,
                            int num_classes, int top_k, float threshold, int64_t* tags,
                            int64_t* masks, void* workspace, size_t workspace_bytes,
                            void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(nq >= 0 && m >= 0 && dim > 0 && top_k > 0 && num_classes > 0 && (tags || masks),
                 "nn_multiset_labels: bad arguments");
  SPML_CHECK_SUPPORTED(!masks || num_classes <= 64,
                       "nn_multiset_labels: bit masks hold at most 64 classes");
  if (nq == 0) return SPML_OK;
  SPML_CHECK_ARG(q && p && plab && workspace, "nn_multiset_labels: null pointer");
  const size_t need = spml_nn_multiset_labels_workspace_bytes(nq, top_k);
  if (workspace_bytes < need) {
    set_error("nn_multiset_labels: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  char* base = reinterpret_cast<char*>(workspace);
  int32_t* hits = reinterpret_cast<int32_t*>(base);
  int64_t* labels = reinterpret_cast<int64_t*>(base + 16);
  float* sim = reinterpret_cast<float*>(base + 16 + (size_t)nq * top_k * sizeof(int64_t));
  TopkExtra extra{};
  extra.qgroup = qgroup;
  extra.pgroup = pgroup;
  extra.plab_limit = num_classes;
  extra.has_limit = 1;
  extra.sim = sim;
  // the reference's torch.topk raises when there are fewer prototypes than top_k
  int rc = topk_launch(q, nq, p, m, dim, plab /* query labels are unused */, plab, nullptr,
                       nullptr, top_k, labels, nullptr, hits, extra, st);
  if (rc != SPML_OK) return rc;
  nn_tags_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(labels, sim, nq, top_k, num_classes,
                                                              threshold, tags, masks);
  SPML_LAUNCH_CHECK("nn_tags_kernel");
  return SPML_OK;
}

}  // extern "C"
