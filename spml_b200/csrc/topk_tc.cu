// C3 / f2 at retrieval scale: top_k_ranking (reference spml/utils/segsort/eval.py:9-52, called
// by predictions/segsort.py:68-125 with the segment prototypes of an image against a memory
// bank of 10^4 - 10^5 prototypes, k = 20) on the tensor cores.
//
// topk.cu streams the bank through shared memory on the FMA pipe: fine for the few hundred
// prototypes of a training step, ~1.1 ms for 576 queries against 50 000 bank rows (twice what
// cuBLAS + torch.topk take).  Here
//   1. topk_tc_candidates_kernel: CTA (query tile of 128 rows, bank slice s) walks the bank tiles
//      s, s + S, ... : the tile is split into bf16 hi / lo in shared memory (SWIZZLE_128B K-major,
//      the layout of kmeans_small.cu), three tcgen05 products (hi.hi + lo.hi + hi.lo, error of a
//      score ~1.5e-5) into TMEM; the 128 x 128 scores go through shared memory to the warps, which
//      take the rows one at a time: the 32 lanes compare the row's 128 scores with the C-th best
//      seen so far (C = k + 8; almost always nothing passes: one ballot per 32 scores) and insert
//      what passes into the row's sorted list, which sits one entry per lane in registers during
//      the insertion (shuffles, no divergence: per-thread lists cost 30x more in divergent
//      branches);
//   2. topk_tc_merge_kernel: one warp per query merges the S sorted lists of its row down to
//      the best C, scores those C prototypes EXACTLY (the fmaf chain of topk.cu, d ascending),
//      orders them by (score descending, index ascending) and writes the first k.
// The exact top-k is among the C candidates unless more than 8 other prototypes lie within 3e-5
// of the k-th best score; ties are kept in index order at every stage.  Same outputs as topk.cu
// (labels, indices, hit counters); no masks / groups (those callers have small problems).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "internal.h"
#include "tc_common.cuh"

namespace spml {

constexpr int kTtRows = 128;                 // queries per CTA (UMMA M)
constexpr int kTtCols = 128;                 // bank rows per tile (UMMA N; N = 256 with the lo part)
constexpr int kTtThreads = 512;             // 16 warps: the row phase is a latency-bound shuffle chain per row
constexpr int kTtMaxC = 32;                  // candidates per list: k + 8 <= 32
constexpr int kTtBlockBytes = 128 * 128;     // one 64-wide K block of a 128-row bf16 tile
constexpr int kTtLd = kTtCols + 1;           // row stride of the score tile in shared memory

struct TopkTcArgs {
  const float* q;
  int64_t nq;
  const float* p;
  int64_t m;
  int dim, nkb, ksteps;
  int slices;        // bank slices (gridDim.y)
  int cand;          // C = k + 8
  float* cand_val;   // [nq][slices][cand]
  int32_t* cand_idx;
};

// byte offset of element (row, d) of a 128-row bf16 K-major tile in the layout a SWIZZLE_128B
// TMA box {64, 128} would write (tc_common.cuh), K blocks `block_stride` bytes apart
__device__ __forceinline__ uint32_t tt_offset(int row, int d, uint32_t block_stride) {
  const uint32_t sw = row & 7;
  return (uint32_t)(d >> 6) * block_stride + (uint32_t)(row >> 3) * 1024 + sw * 128 +
         (((((uint32_t)d & 63) >> 3) ^ sw) << 4) + ((uint32_t)d & 7) * 2;
}

// rows [row0, row0 + 128) of x[rows_total, dim] -> bf16 hi / lo tiles (zero rows / columns
// beyond the data).  hi and lo are `part` bytes apart inside a K block.
__device__ __forceinline__ void tt_split_tile(const float* __restrict__ x, int64_t row0,
                                              int64_t rows_total, int dim, int nkb, uint8_t* hi,
                                              uint32_t lo_delta, uint32_t block_stride) {
  const int nch = nkb * 8;   // 8-element chunks per row
  for (int idx = threadIdx.x; idx < nch * kTtRows; idx += kTtThreads) {
    const int c = idx >> 7, row = idx & (kTtRows - 1);
    const int64_t r = row0 + row;
    uint32_t h[4], l[4];
    float xv[8];
    if ((dim & 3) == 0 && r < rows_total && c * 8 + 8 <= dim &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {   // two 128-bit loads
      const float4 u0 = *reinterpret_cast<const float4*>(x + r * dim + c * 8);
      const float4 u1 = *reinterpret_cast<const float4*>(x + r * dim + c * 8 + 4);
      xv[0] = u0.x, xv[1] = u0.y, xv[2] = u0.z, xv[3] = u0.w;
      xv[4] = u1.x, xv[5] = u1.y, xv[6] = u1.z, xv[7] = u1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        xv[e] = (r < rows_total && c * 8 + e < dim) ? x[r * dim + c * 8 + e] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = xv[2 * e];
      const float x1 = xv[2 * e + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      const __nv_bfloat162 hv = __halves2bfloat162(h0, h1);
      const __nv_bfloat162 lv =
          __floats2bfloat162_rn(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
      h[e] = *reinterpret_cast<const uint32_t*>(&hv);
      l[e] = *reinterpret_cast<const uint32_t*>(&lv);
    }
    const uint32_t off = tt_offset(row, c * 8, block_stride);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(hi + off + lo_delta) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// (score, index) order of the lists: higher score first, lower index first on equal scores
__device__ __forceinline__ bool tt_better(float va, int ia, float vb, int ib) {
  return (va > vb) | ((va == vb) & (ia < ib));   // (bitwise: no short-circuit branches)
}
// N independent sequences of one element per lane -> each sorted best-first across the warp
// (bitonic network, 15 exchanges; the N chains are interleaved step by step, because a single
// chain is nothing but shuffle latency: 4 sorts one after the other took ~12k cycles per row)
template <int N>
__device__ __forceinline__ void tt_sort32(float (&v)[N], int (&i)[N], int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool want_better = ((lane & j) == 0) == ((lane & k) == 0);
      float ov[N];
      int oi[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        ov[n] = __shfl_xor_sync(0xffffffffu, v[n], j);
        oi[n] = __shfl_xor_sync(0xffffffffu, i[n], j);
      }
#pragma unroll
      for (int n = 0; n < N; ++n) {
        // (keys are distinct: "the other one is better" decides both directions; selects, no branches)
        const bool other_better = tt_better(ov[n], oi[n], v[n], i[n]);
        const bool take = other_better == want_better;
        v[n] = take ? ov[n] : v[n];
        i[n] = take ? oi[n] : i[n];
      }
    }
  }
}
// N pairs of sequences sorted best-first across the warp -> the best 32 of each pair's 64,
// sorted, in (v[n], i[n]); interleaved like tt_sort32
template <int N>
__device__ __forceinline__ void tt_merge32(float (&v)[N], int (&i)[N], const float (&bv)[N],
                                           const int (&bi)[N], int lane) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float rv = __shfl_sync(0xffffffffu, bv[n], 31 - lane);
    const int ri = __shfl_sync(0xffffffffu, bi[n], 31 - lane);
    const bool take = tt_better(rv, ri, v[n], i[n]);                  // bitonic, holds the best 32
    v[n] = take ? rv : v[n];
    i[n] = take ? ri : i[n];
  }
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    float ov[N];
    int oi[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      ov[n] = __shfl_xor_sync(0xffffffffu, v[n], j);
      oi[n] = __shfl_xor_sync(0xffffffffu, i[n], j);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const bool other_better = tt_better(ov[n], oi[n], v[n], i[n]);
      const bool take = other_better == ((lane & j) == 0);
      v[n] = take ? ov[n] : v[n];
      i[n] = take ? oi[n] : i[n];
    }
  }
}

__global__ void __launch_bounds__(kTtThreads, 1) topk_tc_candidates_kernel(const TopkTcArgs a) {
  extern __shared__ uint8_t tt_smem_raw[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t q0 = (int64_t)blockIdx.x * kTtRows;
  const int slice = blockIdx.y;
  const int C = a.cand;

  uint8_t* smem = tt_smem_raw + ((1024u - (tc::smem_u32(tt_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;                                          // [nkb][128 x 128 B]
  uint8_t* a_lo = a_hi + (size_t)a.nkb * kTtBlockBytes;
  uint8_t* b_tile = a_lo + (size_t)a.nkb * kTtBlockBytes;        // [nkb][hi | lo]
  // the scores of the current tile, row-major with an odd stride (the threads of a warp write a
  // column, the lanes of a warp read a row), and one sorted candidate list of 32 slots per row
  float* s_sc = reinterpret_cast<float*>(b_tile + (size_t)a.nkb * 2 * kTtBlockBytes);   // [128][129]
  float* l_val = s_sc + kTtRows * kTtLd;                                                 // [128][32]
  int* l_idx = reinterpret_cast<int*>(l_val + kTtRows * 32);
  float* s_thr = reinterpret_cast<float*>(l_idx + kTtRows * 32);                         // [128]

  if (tid == 0) {
    tc::mbar_init(&bar_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, 2 * kTtCols);
  for (int i = tid; i < 32 * kTtRows; i += kTtThreads) l_val[i] = -INFINITY, l_idx[i] = 0x7fffffff;
  if (tid < kTtRows) s_thr[tid] = -INFINITY;
  tt_split_tile(a.q, q0, a.nq, a.dim, a.nkb, a_hi, (uint32_t)a.nkb * kTtBlockBytes, kTtBlockBytes);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  constexpr uint32_t idesc = tc::umma_idesc_bf16(kTtRows, kTtCols, 0, 0);
  constexpr uint32_t idesc2x = tc::umma_idesc_bf16(kTtRows, 2 * kTtCols, 0, 0);
  const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
  const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
  const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
  const uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(b_tile), 16);

  const int sp = warp & 3, quarter = warp >> 2;   // TMEM sub-partition, 32-column quarter of the tile
  const int row = sp * 32 + lane;
  const int64_t ntiles = (a.m + kTtCols - 1) / kTtCols;
  uint32_t phase = 0;

#ifdef SPML_TT_TRACE
  long long tt_t[6] = {0, 0, 0, 0, 0, 0}, tt_c = clock64();
  int tt_events = 0, tt_visits = 0;
#define TT_MARK(i) do { const long long n__ = clock64(); tt_t[i] += n__ - tt_c; tt_c = n__; } while (0)
#else
#define TT_MARK(i) do { } while (0)
#endif
  for (int64_t t = slice; t < ntiles; t += a.slices) {
    const int64_t c0 = t * kTtCols;
    TT_MARK(0);
    // ---- bank tile -> bf16 hi | lo operand (hi part then lo part inside every K block)
    tt_split_tile(a.p, c0, a.m, a.dim, a.nkb, b_tile, kTtBlockBytes, 2 * kTtBlockBytes);
    tc::fence_proxy_async();
    __syncthreads();
    TT_MARK(1);
    if (warp == 0) {
      tc::tcgen05_fence_after();
      if (tc::elect_one()) {
        uint32_t accumulate = 0;
        for (int kb = 0; kb < a.nkb; ++kb) {
          const int steps = min(4, a.ksteps - kb * 4);
          uint32_t ah = ah_lo + kb * (kTtBlockBytes >> 4), al = al_lo + kb * (kTtBlockBytes >> 4);
          uint32_t bp = b_lo + kb * (2 * kTtBlockBytes >> 4);
          for (int ks = 0; ks < steps; ++ks) {   // 16 bf16 = 32 bytes inside the swizzle atom
            tc::umma_bf16_words(tmem_base, ah, hi_k, bp, hi_k, idesc2x, accumulate);   // hh | hl
            tc::umma_bf16_words(tmem_base, al, hi_k, bp, hi_k, idesc, 1);              // += lh
            accumulate = 1;
            ah += 2, al += 2, bp += 2;
          }
        }
        tc::umma_commit(&bar_done);
      }
      __syncwarp();
    }
    tc::mbar_wait(&bar_done, phase);
    phase ^= 1u;
    tc::tcgen05_fence_after();
    TT_MARK(2);
    // ---- TMEM -> shared memory: this thread's row, its 32 columns; columns past the bank = -inf
    {
      const int cb = quarter * 32;
      uint32_t v[32], w[32];
      const uint32_t taddr = tmem_base + cb + (static_cast<uint32_t>(sp * 32) << 16);
      tc::tmem_ld_32x32(taddr, v);              // hi.hi + lo.hi
      tc::tmem_ld_32x32(taddr + kTtCols, w);    // hi.lo
      tc::tmem_ld_wait();
      const int64_t live = a.m - c0 - cb;       // columns of this chunk that exist
#pragma unroll
      for (int u = 0; u < 32; ++u)
        s_sc[row * kTtLd + cb + u] = u < live ? __uint_as_float(v[u]) + __uint_as_float(w[u]) : -INFINITY;
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    TT_MARK(3);
    // ---- the warps take the rows one at a time
    constexpr int kRowsPerWarp = kTtRows / (kTtThreads / 32);
    for (int r = warp * kRowsPerWarp; r < (warp + 1) * kRowsPerWarp; ++r) {
      if (q0 + r >= a.nq) break;   // warp-uniform
      if (t == slice) {
        // the first tile of this CTA: the list is empty and every score would be inserted one
        // at a time (128 dependent insertions per row: ~100 us per CTA); sort the four blocks
        // of 32 scores and merge them instead
        float sv[4];
        int si[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sv[j] = s_sc[r * kTtLd + 32 * j + lane];
          si[j] = sv[j] == -INFINITY ? 0x7fffffff : (int)(c0 + 32 * j + lane);
        }
        tt_sort32<4>(sv, si, lane);
        {
          float av[2] = {sv[0], sv[2]}, bv[2] = {sv[1], sv[3]};
          int ai[2] = {si[0], si[2]}, bi[2] = {si[1], si[3]};
          tt_merge32<2>(av, ai, bv, bi, lane);          // blocks 0 + 1 and 2 + 3
          float cv[1] = {av[0]}, dv[1] = {av[1]};
          int ci[1] = {ai[0]}, di[1] = {ai[1]};
          tt_merge32<1>(cv, ci, dv, di, lane);
          sv[0] = cv[0], si[0] = ci[0];
        }
        if (lane >= C) sv[0] = -INFINITY, si[0] = 0x7fffffff;
        l_val[r * 32 + lane] = sv[0];
        l_idx[r * 32 + lane] = si[0];
        const float t0 = __shfl_sync(0xffffffffu, sv[0], C - 1);
        if (lane == 0) s_thr[r] = t0;
        continue;
      }
      float thr = s_thr[r];
      float xs[4];
      unsigned pass[4];
      unsigned any = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xs[j] = s_sc[r * kTtLd + 32 * j + lane];
        pass[j] = __ballot_sync(0xffffffffu, xs[j] > thr);
        any |= pass[j];
      }
      if (!any) continue;
#ifdef SPML_TT_TRACE
      ++tt_visits;
#endif
      float lv = l_val[r * 32 + lane];     // the row's sorted list, one entry per lane
      int li = l_idx[r * 32 + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned todo = pass[j];
        while (todo) {                     // ascending column index: an equal score stays behind
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
#ifdef SPML_TT_TRACE
          ++tt_events;
#endif
          // (the bar may have risen since the ballot: such a score finds pos == C and changes
          // nothing; not re-checking it keeps the list's last entry off the dependent chain)
          const float sc = __shfl_sync(0xffffffffu, xs[j], src);
          const int pos = __popc(__ballot_sync(0xffffffffu, lane < C && lv >= sc));
          const float nv = __shfl_up_sync(0xffffffffu, lv, 1);
          const int ni = __shfl_up_sync(0xffffffffu, li, 1);
          if (lane > pos && lane < C) lv = nv, li = ni;
          if (lane == pos && lane < C) lv = sc, li = (int)(c0 + 32 * j + src);
        }
      }
      thr = __shfl_sync(0xffffffffu, lv, C - 1);
      l_val[r * 32 + lane] = lv;
      l_idx[r * 32 + lane] = li;
      if (lane == 0) s_thr[r] = thr;
    }
    TT_MARK(5);
    __syncthreads();   // the score tile and the operand tile are free again
    TT_MARK(4);
  }
#ifdef SPML_TT_TRACE
  if (tid == 0 && blockIdx.x == 0 && (blockIdx.y == 0 || blockIdx.y == 7))
    printf("topk_tc CTA(0,%d): loop top %lld split %lld mma %lld tmem->smem %lld rows (warp 0) %lld wait for the other warps %lld cycles; warp 0: %d row visits with work, %d insertions\n",
           blockIdx.y, tt_t[0], tt_t[1], tt_t[2], tt_t[3], tt_t[5], tt_t[4], tt_visits, tt_events);
#endif

  // ---- lists out: [row][slices][C]
  for (int i = tid; i < kTtRows * 32; i += kTtThreads) {
    const int r = i >> 5, e = i & 31;
    if (q0 + r < a.nq && e < C) {
      const size_t at = ((size_t)(q0 + r) * a.slices + slice) * C + e;
      a.cand_val[at] = l_val[i];
      a.cand_idx[at] = l_idx[i];
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 2 * kTtCols);
  }
}

// One warp per query: merge the sorted lists, score the best C exactly, write the top k.
__global__ void __launch_bounds__(256) topk_tc_merge_kernel(
    const TopkTcArgs a, const int64_t* __restrict__ qlab, const int64_t* __restrict__ plab, int k,
    int64_t* __restrict__ topk_labels, int64_t* __restrict__ topk_index, int32_t* hit_count) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.nq) return;
  const int C = a.cand, L = a.slices;
  const float* lv = a.cand_val + (size_t)row * L * C;
  const int32_t* li = a.cand_idx + (size_t)row * L * C;
  // lane l walks the lists l, l + 32, ... with one cursor each (at most 4 lists per lane)
  int cur[4] = {0, 0, 0, 0};
  int my_cand = 0x7fffffff;   // lane r keeps the r-th best candidate
  for (int r = 0; r < C; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff, bj = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int list = lane + 32 * j;
      if (list < L && cur[j] < C) {
        const float v = lv[(size_t)list * C + cur[j]];
        const int i = li[(size_t)list * C + cur[j]];
        if (i != 0x7fffffff && (v > bv || (v == bv && i < bi))) bv = v, bi = i, bj = j;
      }
    }
    float wv = bv;
    int wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      const bool take = (ov > wv) | ((ov == wv) & (oi < wi));
      wv = take ? ov : wv;
      wi = take ? oi : wi;
    }
    if (wi == 0x7fffffff) break;          // fewer than C prototypes in total
    if (bi == wi && bj >= 0) {            // the lane that owns the winner moves on
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == bj) ++cur[j];
    }
    if (lane == r) my_cand = wi;
  }
  // exact scores: the fmaf chain of topk.cu (d ascending from 0)
  float ev = -INFINITY;
  if (my_cand != 0x7fffffff) {
    const float* qr = a.q + row * a.dim;
    const float* pr = a.p + (int64_t)my_cand * a.dim;
    float acc = 0.f;
    for (int d = 0; d < a.dim; ++d) acc = fmaf(qr[d], pr[d], acc);
    ev = acc;
  }
  // rank among the candidates: (score descending, index ascending)
  int rank = 0;
  for (int o = 0; o < 32; ++o) {
    const float ov = __shfl_sync(0xffffffffu, ev, o);
    const int oi = __shfl_sync(0xffffffffu, my_cand, o);
    if (oi != 0x7fffffff && (ov > ev || (ov == ev && oi < my_cand))) ++rank;
  }
  const int64_t ql = qlab ? qlab[row] : -1;
  bool hit = false;
  if (my_cand != 0x7fffffff && rank < k) {
    const int64_t lab = plab[my_cand];
    topk_labels[row * k + rank] = lab;
    if (topk_index) topk_index[row * k + rank] = my_cand;
    hit = lab == ql;
  }
  const unsigned hits = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) {
    if (hits) atomicAdd(hit_count, __popc(hits));
    atomicAdd(hit_count + 1, 1);          // queries that took part
  }
}

// ------------------------------------------------------------------------- host side

// Candidate lists between the two kernels: one buffer per host thread and device, grown on demand
// and handed from one call to the next in stream order (an event), because the C ABI of
// top_k_ranking has no workspace argument and a fresh stream-ordered allocation per call costs
// milliseconds whenever the pool has been trimmed at a synchronisation.
struct TopkScratch {
  void* ptr;
  size_t bytes;
  cudaEvent_t last_use;
  bool ready;
};
static thread_local TopkScratch g_topk_scratch[16];

bool topk_tc_supported(int64_t nq, int64_t m, int dim, int k) {
  const char* e = getenv("SPML_B200_TOPK");
  if (e && !strcmp(e, "fma")) return false;
  if (dim < 1 || dim > 128 || k + 8 > kTtMaxC || m < k) return false;
  if (nq >= (1ll << 31) || m >= (1ll << 31)) return false;
  if (e && !strcmp(e, "tc")) return true;
  // worth two launches and a scratch allocation only for a large bank
  return m >= 4096 && nq * m >= (1ll << 22);
}

int topk_tc_launch(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                   const int64_t* qlab, const int64_t* plab, int k, int64_t* topk_labels,
                   int64_t* topk_index, int32_t* hit_count, cudaStream_t st) {
  TopkTcArgs a{};
  a.q = q, a.nq = nq, a.p = p, a.m = m, a.dim = dim;
  a.nkb = (dim + 63) / 64;
  a.ksteps = (dim + 15) / 16;
  a.cand = k + 8;
  int device = 0, sms = 0;
  SPML_CUDA(cudaGetDevice(&device));
  SPML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int64_t qtiles = ceil_div(nq, kTtRows), ntiles = ceil_div(m, kTtCols);
  // about one wave of CTAs; a lane of the merge kernel walks at most 4 of the S lists
  a.slices = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ntiles, 128), sms / qtiles));
  const size_t entries = (size_t)nq * a.slices * a.cand;
  SPML_CHECK_SUPPORTED(device >= 0 && device < 16, "device index %d not supported", device);
  TopkScratch& sc = g_topk_scratch[device];
  if (!sc.ready) {
    SPML_CUDA(cudaEventCreateWithFlags(&sc.last_use, cudaEventDisableTiming));
    sc.ready = true;
  }
  if (sc.bytes < entries * 8) {
    if (sc.ptr) SPML_CUDA(cudaFree(sc.ptr));   // (synchronises: nobody uses the old buffer any more)
    sc.ptr = nullptr, sc.bytes = 0;
    const size_t want = align_up(entries * 8 + entries * 2, 1 << 20);
    SPML_CUDA(cudaMalloc(&sc.ptr, want));
    sc.bytes = want;
  } else {
    SPML_CUDA(cudaStreamWaitEvent(st, sc.last_use, 0));   // the previous call may have run on another stream
  }
  a.cand_val = reinterpret_cast<float*>(sc.ptr);
  a.cand_idx = reinterpret_cast<int32_t*>(a.cand_val + entries);
  const size_t smem = 1024 + (size_t)a.nkb * 4 * kTtBlockBytes +
                      (size_t)kTtRows * (kTtLd + 32 + 32 + 1) * 4;
  cudaError_t err = cudaFuncSetAttribute(topk_tc_candidates_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess) {
    topk_tc_candidates_kernel<<<dim3((unsigned)qtiles, (unsigned)a.slices), kTtThreads, smem, st>>>(a);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) {
    count_launch();
    topk_tc_merge_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(a, qlab, plab, k, topk_labels,
                                                                    topk_index, hit_count);
    err = cudaGetLastError();
    if (err == cudaSuccess) count_launch();
  }
  if (err != cudaSuccess) return cuda_fail(err, "topk_tc");
  SPML_CUDA(cudaEventRecord(sc.last_use, st));
  return SPML_OK;
}

}  // namespace spml
