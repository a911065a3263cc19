// A4-A6 for K <= 128 clusters per image (every shipped configuration: 6x6, 8x8, 8x16 seeds):
// the tcgen05 E-step of kmeans_tc.cu without its per-pass publication chain.
//
// kmeans_tc.cu separates two passes by: segment sums -> count -> ONE finalising CTA folds,
// normalises and publishes the image's prototypes -> flag -> every CTA pulls them in by TMA.
// That is ~8 dependent L2 round trips per pass and 14.5 us per pass at batch 1, 25x what the
// arithmetic needs (profiles/r1c_kmeans_tc_timeline_b1.txt).  With K <= 128 the whole
// prototype set of an image is one MMA operand tile, so here
//   * every CTA folds the image's K x dim fixed-point sums ITSELF as soon as the image's tile
//     counter is complete (one batch of L2 loads, all in flight together), normalises them
//     and writes the bf16 hi / lo operand tile and the fp32 rows straight into its own shared
//     memory: no finalising CTA, no flag, no TMA, no prototype traffic through global memory;
//   * the M-step is INCREMENTAL: pass 0 sums every row once; afterwards only the rows whose
//     label changed move (- x from the old cluster, + x to the new one; 39 % of the rows in
//     pass 1 of the VOC workload, under 1 % from pass 6 on) and the image's previous totals
//     are carried over by the CTAs in equal slices.  Sums are exact fixed-point integers, so
//     this is the same arithmetic as re-summing everything.  A tile's contribution is gathered
//     with native 32-bit shared-memory atomics on the two halves of round(x 2^32) and only its
//     non-zero entries go to the global sums (64-bit reductions).
// Sums are integers, so labels are bit-identical to the fp32 kernel and to kmeans_tc.cu
// whatever the tiling (tests/test_gpu_ops.py::test_kmeans_tensor_core_equals_fp32).
#include <math.h>

#include <algorithm>

#include "kmeans.cuh"
#include "tc_common.cuh"

namespace spml {

constexpr int kSmBN = 128;                 // most prototype rows of the operand tile (64 when K <= 64)
constexpr int kSmBlockBytes = 128 * 128;   // one 64-wide K block of a 128-row bf16 tile
constexpr int kSmWarps = kGemmThreads / 32;
constexpr int kSmMaxBatch = 255;           // images per call (offsets cached in shared memory)
constexpr int kSmMaxTilesPerFlush = 24;    // 2^19 * 128 * 24 < 2^31: the hi halves cannot overflow

#ifdef SPML_KM_TRACE
__device__ long long g_kms_trace[16 * 16];
// every CTA's stamps of its last tile of each pass: [CTA][pass][slot] (SM-local clocks: only
// differences inside one CTA mean something; "done seen" is the common reference of a pass)
__device__ long long g_kms_trace_all[160 * 16 * 16];
#define KMS(slot)                                                                          \
  do {                                                                                     \
    if (threadIdx.x == 0 && it < 16) {                                                     \
      const long long c__ = clock64();                                                     \
      if (lt0 == 0 && lt1 > 0) g_kms_trace[it * 16 + (slot)] = c__;                        \
      if (blockIdx.x < 160) g_kms_trace_all[(blockIdx.x * 16 + it) * 16 + (slot)] = c__;   \
    }                                                                                      \
  } while (0)
#define KMS_VAL(slot, v)                                                                   \
  do {                                                                                     \
    if (threadIdx.x == 0 && it < 16 && blockIdx.x < 160)                                   \
      g_kms_trace_all[(blockIdx.x * 16 + it) * 16 + (slot)] = (v);                         \
  } while (0)
#else
#define KMS_VAL(slot, v) do { } while (0)
#define KMS(slot) do { } while (0)
#endif

struct KmeansSmallArgs {
  KmeansArgs k;
  int nkb;        // 64-wide K blocks
  int ksteps;     // 16-wide K steps that hold data
  int prefetch;   // a second fp32 tile buffer fits
  int bn;         // prototype rows of the operand tile: 64 or 128 (UMMA N = bn and 2 bn)
  float tau;
};

__device__ __forceinline__ void kms_st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void kms_st_shared_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void kms_fence_acq_rel_gpu() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ unsigned kms_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void kms_cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void kms_cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src)
               : "memory");
}

// byte offset of element (row, d) inside a [nkb][hi | lo] SWIZZLE_128B operand whose hi and lo
// parts are `part_bytes` (= rows * 128) each: what a TMA box {64 bf16, rows} would write
// (tc_common.cuh)
__device__ __forceinline__ uint32_t operand_offset(int row, int d, uint32_t part_bytes) {
  const uint32_t sw = row & 7;
  return (uint32_t)(d >> 6) * (2 * part_bytes) + (uint32_t)(row >> 3) * 1024 + sw * 128 +
         (((((uint32_t)d & 63) >> 3) ^ sw) << 4) + ((uint32_t)d & 7) * 2;
}

// live tile `lt` (image-major order) -> image, first row, row count; also the image's first
// live tile index and its number of tiles.  `off` = the image offsets (shared-memory copy).
__device__ __forceinline__ void kms_locate_tile(const int* off, int lt, Tile& tile,
                                                int* img_base = nullptr, int* img_tiles = nullptr) {
  int img = 0, base = 0;
  for (;;) {
    const int first = off[img], last = off[img + 1];
    const int tiles_b = (last - first + BM - 1) >> 7;
    if (lt < base + tiles_b) {
      tile.b = img;
      tile.row0 = first + (int64_t)(lt - base) * BM;
      tile.rows = min(BM, last - (int)tile.row0);
      if (img_base) *img_base = base;
      if (img_tiles) *img_tiles = tiles_b;
      return;
    }
    base += tiles_b;
    ++img;
  }
}

// Starts the copy of a tile's rows * dim floats into `buf` (kept at the source's 16-byte
// phase so that the body moves in 16-byte pieces); returns that phase in floats.
__device__ __forceinline__ int kms_prefetch_tile(const KmeansArgs& p, const Tile& tile, float* buf) {
  const int tid = threadIdx.x;
  const float* src = p.x + tile.row0 * p.dim;
  const int total = tile.rows * p.dim;
  const int lead = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
  float* dst = buf + lead;
  const int head = min(total, (4 - lead) & 3);
  const int body = (total - head) >> 2;
  if (tid < head) kms_cp_async_4(dst + tid, src + tid);
  for (int i = tid; i < body; i += kGemmThreads) kms_cp_async_16(dst + head + 4 * i, src + head + 4 * i);
  const int done4 = head + 4 * body;
  if (tid < total - done4) kms_cp_async_4(dst + done4 + tid, src + done4 + tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  return lead;
}

// The image's unit prototypes from its fixed-point totals (common.py:39: sum / max(||sum||,
// eps); an empty cluster is the zero vector), staged in shared memory, written to THIS CTA's
// shared memory: fp32 rows for the exact re-check and the bf16 hi / lo operand tile.  One warp
// per prototype, lanes across the channels, kC prototypes per warp and round.
//
// The rebuild sits on the critical path of every pass (nothing else can run until the
// prototypes exist), so its dependent chains are kept short: the square root and the
// reciprocal of the kC norms are computed once, by lanes 0..kC-1 side by side, and the
// quotients use the branch-free shared-reciprocal division of common.cuh (sqrtf and `/` each
// carry a range check and a slow-path call per use, which serialised the 3 kC quotient chains of
// a warp: 3.9k cycles for 36 x 66 before, see profiles/).
template <int kSlots>
__device__ __forceinline__ void build_prototypes(const KmeansArgs& p,
                                                 const long long* __restrict__ totals, int kb,
                                                 float* __restrict__ pf,
                                                 uint8_t* __restrict__ b_tile,
                                                 uint32_t b_part) {
  constexpr int kC = 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dim = p.dim;
  const uint32_t pf_s = tc::smem_u32(pf), bt_s = tc::smem_u32(b_tile);   // 32-bit shared addresses
  for (int kbase = warp; kbase < kb; kbase += kC * kSmWarps) {
    float v[kC][kSlots], ss[kC];
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int k = kbase + i * kSmWarps;
      ss[i] = 0.f;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int d = lane + 32 * s;
        v[i][s] = (k < kb && d < dim) ? fixed_to_float(totals[k * dim + d]) : 0.f;
        ss[i] += v[i][s] * v[i][s];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < kC; ++i) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
    // lane i finishes prototype i
    float mine = ss[0];
#pragma unroll
    for (int i = 1; i < kC; ++i) mine = lane == i ? ss[i] : mine;
    const float nrm = sqrtf(mine);
    const float div_l = nrm >= p.eps ? nrm : p.eps;     // eps = 1e-12 >= kDivMinDivisor
    const float rcp_l = div_reciprocal(div_l);
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int k = kbase + i * kSmWarps;
      const float div = __shfl_sync(0xffffffffu, div_l, i);
      const float rcp = __shfl_sync(0xffffffffu, rcp_l, i);
      if (k < kb) {                                  // warp-uniform
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const int d = lane + 32 * s;
          // computed for every lane (the padding lanes hold zeros), only the stores are
          // predicated: no branch inside the chain
          const float u = div_by(v[i][s], div, rcp);
          const __nv_bfloat16 h = __float2bfloat16_rn(u);
          const __nv_bfloat16 l = __float2bfloat16_rn(u - __bfloat162float(h));
          const uint32_t off = bt_s + operand_offset(k, d, b_part);
          if (d < dim) {
            kms_st_shared_f32(pf_s + (uint32_t)(k * dim + d) * 4u, u);
            kms_st_shared_u16(off, __bfloat16_as_ushort(h));
            kms_st_shared_u16(off + b_part, __bfloat16_as_ushort(l));
          }
        }
      }
    }
  }
}

// s_order[i] = i-th row of the tile in label order (counting sort over K <= 128 labels; the
// order inside a label is whatever the atomics give: the sums are integers, so it is free).
// Rows with a label outside [0, K) are left out (reported through the return value).
__device__ __forceinline__ bool sort_rows_by_label(const int* s_lab, int rows, int num_clusters,
                                                   int* s_cnt, unsigned char* s_order,
                                                   int* s_sorted) {
  const int tid = threadIdx.x;
  if (tid < kSmBN) s_cnt[tid] = 0;
  __syncthreads();
  int lab = -1, pos = 0;
  bool bad = false;
  if (tid < rows) {
    lab = s_lab[tid];
    if ((unsigned)lab < (unsigned)num_clusters) pos = atomicAdd(&s_cnt[lab], 1);
    else lab = -1, bad = true;
  }
  __syncthreads();
  if (tid < 32) {   // exclusive prefix over the 128 counters: four per lane + a warp scan
    int c[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = s_cnt[tid * 4 + j], sum += c[j];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) s_cnt[tid * 4 + j] = run, run += c[j];
    if (tid == 31) *s_sorted = incl;   // rows that take part
  }
  __syncthreads();
  if (lab >= 0) s_order[s_cnt[lab] + pos] = (unsigned char)tid;
  __syncthreads();
  return bad;
}

// M-step of one tile into the CTA's shared sums.  Warp w takes 16 consecutive entries of the
// label order, lanes across the channels, sums a run of equal labels in registers as the two
// exact halves of round(x 2^32) (hi * 2^16 + lo; at most 16 rows per run, neither half can
// overflow) and adds a run to the shared sums with native 32-bit atomics when the label
// changes: about (labels in the tile + 8) flushes per tile.
template <int kSlots>
__device__ __forceinline__ bool accumulate_tile(int sorted, int dim, const float* __restrict__ xs,
                                                const int* __restrict__ s_lab,
                                                const unsigned char* __restrict__ s_order,
                                                int* s_hi, int* s_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kPer = BM / kSmWarps;   // 16
  const int e0 = warp * kPer;
  int run_hi[kSlots], run_lo[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) run_hi[s] = run_lo[s] = 0;
  int run_lab = -1;
  bool bad = false;
  auto flush = [&]() {
    if (run_lab < 0) return;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int d = lane + 32 * s;
      if (d < dim) {
        atomicAdd(&s_hi[run_lab * dim + d], run_hi[s]);
        atomicAdd(&s_lo[run_lab * dim + d], run_lo[s]);
      }
      run_hi[s] = run_lo[s] = 0;
    }
  };
#pragma unroll
  for (int g = 0; g < kPer / 4; ++g) {
    // four rows in flight: all their shared-memory loads are issued before the first add
    float v[4][kSlots];
    int lab4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + g * 4 + u;
      const bool ok = e < sorted;
      const int row = ok ? s_order[e] : 0;
      lab4[u] = ok ? s_lab[row] : -1;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int d = lane + 32 * s;
        v[u][s] = (ok && d < dim) ? xs[row * dim + d] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (lab4[u] < 0) continue;        // past the end of the tile (warp-uniform)
      if (lab4[u] != run_lab) {
        flush();
        run_lab = lab4[u];
      }
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        bad |= !(fabsf(v[u][s]) <= 8.f);
        int hi, lo;
        split_fixed(v[u][s], hi, lo);
        run_hi[s] += hi;
        run_lo[s] += lo;
      }
    }
  }
  flush();
  return bad;
}

// Exact scores of one row against the prototypes lane, lane + 32, ... (kChains of them per
// lane): the very fmaf chain of the fp32 kernel (d ascending from 0), so the label of an
// ambiguous row is the one the fp32 path (and the reference) gives.  The chains of a lane are
// independent and the loads of eight steps are issued ahead of the arithmetic: a straggler row
// costs ~0.5k cycles instead of 3.3k (one such row held up the whole pass, profiles/).
// Returns the lane's best (score, index), first index on ties.
template <int kChains>
__device__ __forceinline__ void exact_best(const float* __restrict__ xr,
                                           const float* __restrict__ pf, int dim, int kb, int lane,
                                           float& bv, int& bk) {
  float acc[kChains];
  const float* pk[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    acc[c] = 0.f;
    pk[c] = pf + min(lane + 32 * c, kb - 1) * dim;     // out-of-range lanes redo the last row
  }
  int d = 0;
  // a single warp is bound by the number of shared-memory instructions it can issue (scalar
  // loads: 2.1k cycles for 36 x 66), so the rows are read as float2 where they are 8-byte
  // aligned (even dim: every row of both arrays is)
  if ((dim & 1) == 0 && ((tc::smem_u32(xr) | tc::smem_u32(pf)) & 7u) == 0) {
#pragma unroll 2
    for (; d + 8 <= dim; d += 8) {
      float2 xv[4], pv[kChains][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float2*>(xr + d + 2 * j);
#pragma unroll
      for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) pv[c][j] = *reinterpret_cast<const float2*>(pk[c] + d + 2 * j);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = fmaf(xv[j].x, pv[c][j].x, acc[c]);
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = fmaf(xv[j].y, pv[c][j].y, acc[c]);
      }
    }
  } else {
#pragma unroll 2
    for (; d + 8 <= dim; d += 8) {
      float xv[8], pv[kChains][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = xr[d + j];
#pragma unroll
      for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[c][j] = pk[c][d + j];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = fmaf(xv[j], pv[c][j], acc[c]);
    }
  }
  for (; d < dim; ++d) {
    const float x = xr[d];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = fmaf(x, pk[c][d], acc[c]);
  }
  bv = -INFINITY;
  bk = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    const int k = lane + 32 * c;
    if (k < kb && acc[c] > bv) bv = acc[c], bk = k;
  }
}

// The exact label of one row (one warp).  NOT inlined: the recheck is rare (a few rows of the
// whole grid per pass), so on most SMs its instructions are cold when it finally runs, and
// every instruction-cache miss is an L2 round trip (~1k cycles, 2-4 of them measured per row
// on top of the 1.3k cycles of the chain itself, and the whole pass waits for that one row).
// The kernel runs it once on scratch data at its start, through this same copy of the code,
// which takes the misses off the passes (a row costs 1.8-2.2k cycles afterwards).
__device__ __noinline__ int exact_row_label(const float* xr, const float* pf, int dim, int kb,
                                            int lane) {
  float bv;
  int bk;
  switch ((kb + 31) >> 5) {
    case 1: exact_best<1>(xr, pf, dim, kb, lane, bv, bk); break;
    case 2: exact_best<2>(xr, pf, dim, kb, lane, bv, bk); break;
    case 3: exact_best<3>(xr, pf, dim, kb, lane, bv, bk); break;
    default: exact_best<4>(xr, pf, dim, kb, lane, bv, bk); break;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
    if (ov > bv || (ov == bv && ok < bk)) bv = ov, bk = ok;
  }
  return bk;
}

#define KMS_SLOT_SWITCH(dim, CALL)           \
  switch (((dim) + 31) >> 5) {               \
    case 1: { constexpr int kS = 1; CALL; } break; \
    case 2: { constexpr int kS = 2; CALL; } break; \
    case 3: { constexpr int kS = 3; CALL; } break; \
    case 4: { constexpr int kS = 4; CALL; } break; \
    default: { constexpr int kS = 5; CALL; } break; \
  }

__global__ void __launch_bounds__(kGemmThreads, 1)
kmeans_small_kernel(const KmeansSmallArgs a) {
  extern __shared__ uint8_t kms_smem_raw[];
  __shared__ __align__(8) uint64_t bar_t_full, bar_stage;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_lab[BM], s_old[BM];
  __shared__ float s_b1[BM], s_b2[BM];
  __shared__ int s_k1[BM];
  __shared__ int s_amb[BM], s_chg[BM];
  __shared__ int s_namb, s_nchg, s_sorted, s_warm;
  __shared__ int s_cnt[kSmBN];
  __shared__ unsigned char s_order[BM];

  const KmeansArgs& p = a.k;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int dim = p.dim;
  const int K = p.num_clusters;
  // image offsets in shared memory: every pass ends in a fence.acq_rel.gpu, which drops the L1
  // lines, so reading them from global memory costs an L2 round trip per tile and pass
  __shared__ int s_off[kSmMaxBatch + 1];
  for (int i = tid; i <= p.batch; i += kGemmThreads)
    s_off[i] = p.img_off ? p.img_off[i] : (i == 0 ? 0 : (int)p.rows_total);
  __syncthreads();
  // every CTA takes a CONTIGUOUS range of the live tiles (image-major order)
  int live_total = 0;
  for (int bb = 0; bb < p.batch; ++bb) live_total += (s_off[bb + 1] - s_off[bb] + BM - 1) >> 7;
  const int lt0 = (int)((int64_t)live_total * blockIdx.x / gridDim.x);
  const int lt1 = (int)((int64_t)live_total * (blockIdx.x + 1) / gridDim.x);
  const bool resident = live_total <= (int)gridDim.x;       // at most one tile per CTA: load it once
  const int per_img = K * dim;
  const size_t per_iter = (size_t)p.batch * per_img;

  // 1024-byte aligned carve-up by OFFSET, so that the pointers stay shared-space pointers
  uint8_t* smem = kms_smem_raw + ((1024u - (tc::smem_u32(kms_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;                                        // [nkb][128 x 128 B]
  uint8_t* a_lo = a_hi + (size_t)a.nkb * kSmBlockBytes;
  uint8_t* b_tile = a_lo + (size_t)a.nkb * kSmBlockBytes;      // [nkb][hi | lo]
  // the tile's contribution to the sums as two int32 halves per entry; the same bytes stage
  // the image's int64 totals while the prototypes are rebuilt (the halves are zero then)
  const uint32_t b_part = (uint32_t)a.bn * 128;                // bytes of the hi (or lo) rows of a K block
  int* s_hi = reinterpret_cast<int*>(b_tile + (size_t)a.nkb * 2 * b_part);   // [K][dim]
  int* s_lo = s_hi + per_img;
  long long* s_tot = reinterpret_cast<long long*>(s_hi);
  float* pf = reinterpret_cast<float*>(s_lo + per_img);        // [K][dim] fp32 prototypes
  float* xf = reinterpret_cast<float*>(
      reinterpret_cast<uint8_t*>(pf + per_img) +
      ((16u - (tc::smem_u32(pf + per_img) & 15u)) & 15u));     // 16-byte aligned (128-bit copies)
  const size_t xf_stride = (size_t)BM * dim + 4;               // a second buffer only with a.prefetch
  const float* xs = xf;   // xs[r * dim + d]: the fp32 tile, at the 16-byte phase of its source

  if (tid == 0) {
    tc::mbar_init(&bar_t_full, 1);
    tc::mbar_init(&bar_stage, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, 2 * kSmBN);   // one accumulator [hh + lh | hl]
  // operand tile: rows >= K and the K padding [dim, 64 nkb) must be zeros (0 x garbage = NaN)
  for (int i = tid; i < a.nkb * 2 * (int)b_part / 16; i += kGemmThreads)
    reinterpret_cast<uint4*>(b_tile)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < per_img; i += kGemmThreads) s_hi[i] = 0, s_lo[i] = 0;
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // instruction-cache warm-up of the exact recheck (see exact_row_label) on whatever bytes
  // the prototype rows hold right now; the result goes nowhere
  if (warp == kSmWarps - 1) {
    const int warm = exact_row_label(pf, pf, dim, K, lane);
    if (lane == 0) s_warm = warm;
  }

  const uint32_t idesc = a.bn == 64 ? tc::umma_idesc_bf16(BM, 64, 0, 0)
                                    : tc::umma_idesc_bf16(BM, 128, 0, 0);
  const uint32_t idesc2x = a.bn == 64 ? tc::umma_idesc_bf16(BM, 128, 0, 0)
                                      : tc::umma_idesc_bf16(BM, 256, 0, 0);
  const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
  const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
  const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
  const uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(b_tile), 16);

  uint32_t q = 0;   // tiles this CTA has pushed through the accumulator so far (barrier parity)
  uint32_t stage_phase = 0;   // parity of bar_stage (one bulk copy of the totals per rebuild)
  const bool pf_on = a.prefetch && !resident && lt0 < lt1;
  int cur = 0, lead_cur = 0, lead_nxt = 0;
  if (pf_on) {
    Tile first_tile;
    kms_locate_tile(s_off, lt0, first_tile);
    lead_cur = kms_prefetch_tile(p, first_tile, xf);
  }
  bool bad = false;

  for (int it = 0; it <= p.iterations; ++it) {
    int proto_img = -1;    // image whose prototypes of pass it - 1 sit in shared memory
    int pending = 0;       // tiles of the current image gathered in s_hi / s_lo, not flushed
    for (int lt = lt0; lt < lt1; ++lt) {
      Tile tile;
      int img_base, tiles_b;
      kms_locate_tile(s_off, lt, tile, &img_base, &tiles_b);
      const int b = tile.b;
      const int kb = p.k_per_image ? p.k_per_image[b] : K;
      KMS(0);

      if (!resident || it == 0) {
        float* dst;
        if (pf_on) {
          // ---- the tile was requested one tile ago (cp.async); ask for the next one now
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncthreads();   // landed for everybody; nobody still reads the previous tile
          dst = xf + cur * xf_stride + lead_cur;
          const int nlt = lt + 1 < lt1 ? lt + 1 : (it < p.iterations ? lt0 : -1);
          if (nlt >= 0) {
            Tile next;
            kms_locate_tile(s_off, nlt, next);
            lead_nxt = kms_prefetch_tile(p, next, xf + (cur ^ 1) * xf_stride);
          }
        } else {
          // ---- fp32 tile: one contiguous chunk of rows * dim floats, 128-bit copies where the
          // source allows (the shared copy keeps the source's 16-byte phase)
          const float* __restrict__ src = p.x + tile.row0 * dim;
          const int total = tile.rows * dim;
          const int lead = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
          dst = xf + lead;
          const int head = min(total, (4 - lead) & 3);
          const int body = (total - head) >> 2;
          __syncthreads();   // nobody still reads the previous tile
          if (tid < head) dst[tid] = src[tid];
          const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src + head);
          float4* d4 = reinterpret_cast<float4*>(dst + head);
          for (int i0 = 0; i0 < body; i0 += 4 * kGemmThreads) {   // four 128-bit loads in flight
            float4 r4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i0 + j * kGemmThreads + tid;
              if (i < body) r4[j] = s4[i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i0 + j * kGemmThreads + tid;
              if (i < body) d4[i] = r4[j];
            }
          }
          const int done4 = head + 4 * body;
          if (tid < total - done4) dst[done4 + tid] = src[done4 + tid];
          __syncthreads();
        }
        xs = dst;
        // ---- bf16 hi / lo operand tiles in the layout a SWIZZLE_128B TMA box would write
        const int nch = 2 * a.ksteps;   // 8-element chunks that the MMAs read
        for (int idx = tid; idx < nch * BM; idx += kGemmThreads) {
          const int c = idx >> 7, row = idx & (BM - 1);
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int d0 = c * 8 + 2 * h;
            const float x0 = (row < tile.rows && d0 < dim) ? dst[row * dim + d0] : 0.f;
            const float x1 = (row < tile.rows && d0 + 1 < dim) ? dst[row * dim + d0 + 1] : 0.f;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
            const __nv_bfloat162 hv = __halves2bfloat162(h0, h1);
            const __nv_bfloat162 lv = __floats2bfloat162_rn(x0 - __bfloat162float(h0),
                                                             x1 - __bfloat162float(h1));
            hi[h] = *reinterpret_cast<const uint32_t*>(&hv);
            lo[h] = *reinterpret_cast<const uint32_t*>(&lv);
          }
          const uint32_t sw = row & 7;
          const uint32_t off = (uint32_t)(c >> 3) * kSmBlockBytes + (uint32_t)(row >> 3) * 1024 +
                               sw * 128 + ((((uint32_t)c & 7) ^ sw) << 4);
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      KMS(1);

      if (it == 0) {
        // ================================================================== pass 0
        __syncthreads();
        if (tid < tile.rows) s_lab[tid] = p.labels_in[tile.row0 + tid];
        // M-step over every row (the totals start here)
        bad |= sort_rows_by_label(s_lab, tile.rows, kb, s_cnt, s_order, &s_sorted);
        KMS(10);
        KMS_SLOT_SWITCH(dim, (bad |= accumulate_tile<kS>(s_sorted, dim, xs, s_lab, s_order, s_hi,
                                                         s_lo)));
        if (!resident && tid < tile.rows) {
          // the rows' labels travel through the output buffer between passes (same CTA)
          if (p.labels_out) p.labels_out[tile.row0 + tid] = s_lab[tid];
          else p.labels_out64[tile.row0 + tid] = s_lab[tid];
        }
      } else {
        // ================================================================== E-step
        if (proto_img != b) {
          // ---- pass it - 1 of this image is complete once every one of its tiles is counted
          if (tid == 0) {
            const unsigned* done = p.done + (size_t)(it - 1) * p.batch + b;
            unsigned spins = 0;
            while (kms_ld_acquire_gpu(done) < (unsigned)tiles_b) {
              __nanosleep(40);   // ~100 pollers share this line with the CTAs still counting
              if (++spins > tc::kSpinLimit) {   // trap instead of hanging the GPU
                printf("spml_b200: k-means pass %d of image %d never completed (block %d)\n",
                       it - 1, b, blockIdx.x);
                __trap();
              }
            }
            kms_fence_acq_rel_gpu();
          }
          tc::fence_proxy_async();   // this thread's zeroing of s_hi / s_lo -> the bulk copy below
          __syncthreads();
          KMS(2);
          // ---- the image's totals after pass it - 1 into the (currently all-zero) s_hi / s_lo
          // bytes.  Every CTA of the image reads the SAME K x dim words at the same moment.
          const long long* tot_prev = p.sums + (size_t)(it - 1) * per_iter + (size_t)b * per_img;
          if ((per_img & 1) == 0) {
            // one bulk copy (16-byte granular: per_img even makes source and size aligned)
            if (tid == 0) {
              tc::fence_proxy_async_all();
              tc::mbar_expect_tx(&bar_stage, (uint32_t)per_img * 8u);
              tc::bulk_load(s_tot, tot_prev, (uint32_t)per_img * 8u, &bar_stage);
            }
            tc::mbar_wait(&bar_stage, stage_phase);
            stage_phase ^= 1u;
          } else {
            // walking the words in the same order serialises ~100 SMs on a handful of L2 lines
            // at a time (measured: 5-9k cycles for 19 KB): each CTA starts at its own rotation
            const int rot = (int)(((unsigned)blockIdx.x * 2654435761u >> 8) % (unsigned)per_img) & ~31;
            for (int i0 = 0; i0 < per_img; i0 += 12 * kGemmThreads) {
              long long r12[12];
#pragma unroll
              for (int j = 0; j < 12; ++j) {
                int i = i0 + j * kGemmThreads + tid;
                const bool in = i < per_img;
                i += rot;
                i -= i >= per_img ? per_img : 0;
                r12[j] = in ? __ldcg(tot_prev + i) : 0;
              }
#pragma unroll
              for (int j = 0; j < 12; ++j) {
                int i = i0 + j * kGemmThreads + tid;
                const bool in = i < per_img;
                i += rot;
                i -= i >= per_img ? per_img : 0;
                if (in) s_tot[i] = r12[j];
              }
            }
          }
          __syncthreads();
          KMS(11);
          if (it < p.iterations) {
            // ---- carry: the totals of pass it start from those of pass it - 1; every tile of
            // the image brings over one slice (this CTA: the slices of its tiles of the image)
            const int j0 = max(lt0, img_base) - img_base;
            const int j1 = min(lt1, img_base + tiles_b) - img_base;
            // (per_img <= 2^15 and tiles_b < 2^16 on every supported shape: 32-bit products)
            const int c0 = (int)((unsigned)per_img * (unsigned)j0 / (unsigned)tiles_b);
            const int c1 = (int)((unsigned)per_img * (unsigned)j1 / (unsigned)tiles_b);
            long long* tot_it = p.sums + (size_t)it * per_iter + (size_t)b * per_img;
            for (int i = c0 + tid; i < c1; i += kGemmThreads) {
              const long long v = s_tot[i];
              if (v != 0) atomic_add_i64(&tot_it[i], v);
            }
          }
          KMS(12);
          KMS_SLOT_SWITCH(dim, (build_prototypes<kS>(p, s_tot, kb, pf, b_tile, b_part)));
          KMS(13);
          __syncthreads();
          for (int i = tid; i < per_img; i += kGemmThreads) s_hi[i] = 0, s_lo[i] = 0;
          proto_img = b;
        }
        // the labels of the previous pass (for the incremental M-step)
        if (tid < BM) {
          if (resident) s_old[tid] = s_lab[tid];
          else if (tid < tile.rows)
            s_old[tid] = p.labels_out ? p.labels_out[tile.row0 + tid]
                                      : (int)p.labels_out64[tile.row0 + tid];
        }
        if (tid == 0) s_namb = 0, s_nchg = 0;
        tc::fence_proxy_async();   // generic-proxy stores (A and B operands) -> tensor core
        __syncthreads();
        KMS(3);
        if (warp == 0) {
          tc::tcgen05_fence_after();
          if (tc::elect_one()) {
            uint32_t accumulate = 0;
            for (int kblk = 0; kblk < a.nkb; ++kblk) {
              const int steps = min(4, a.ksteps - kblk * 4);
              uint32_t ah = ah_lo + kblk * (kSmBlockBytes >> 4);
              uint32_t al = al_lo + kblk * (kSmBlockBytes >> 4);
              uint32_t bp = b_lo + kblk * (2 * b_part >> 4);
              for (int ks = 0; ks < steps; ++ks) {   // 16 bf16 = 32 bytes inside the swizzle atom
                tc::umma_bf16_words(tmem_base, ah, hi_k, bp, hi_k, idesc2x, accumulate);   // hh | hl
                tc::umma_bf16_words(tmem_base, al, hi_k, bp, hi_k, idesc, 1);              // += lh
                accumulate = 1;
                ah += 2, al += 2, bp += 2;
              }
            }
            tc::umma_commit(&bar_t_full);
          }
          __syncwarp();
        }
        const int sp = warp & 3;            // TMEM sub-partition of this warp
        const int half = warp >> 2;         // which 32-column chunks of the accumulator
        const int row = sp * 32 + lane;
        float b1 = -INFINITY, b2 = -INFINITY;
        int k1 = 0;
        tc::mbar_wait(&bar_t_full, q & 1);
        ++q;
        tc::tcgen05_fence_after();
        KMS(4);
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
          // the halves take alternate 32-column chunks, so both work when K <= 64
          const int cb = (2 * chunk + half) * 32;
          if (cb < kb) {             // warp-uniform
            uint32_t v[32], w[32];
            const uint32_t taddr = tmem_base + cb + (static_cast<uint32_t>(sp * 32) << 16);
            tc::tmem_ld_32x32(taddr, v);            // hi.hi + lo.hi
            tc::tmem_ld_32x32(taddr + a.bn, w);     // hi.lo
            tc::tmem_ld_wait();
            const int live = kb - cb;               // columns of this chunk that exist
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              float sc = __uint_as_float(v[u]) + __uint_as_float(w[u]);
              sc = u < live ? sc : -INFINITY;
              b2 = fmaxf(b2, fminf(sc, b1));        // second best so far (a tie counts)
              k1 = sc > b1 ? cb + u : k1;
              b1 = fmaxf(b1, sc);
            }
          }
        }
        tc::tcgen05_fence_before();
        // ---- the two column halves of a row meet; ambiguous rows go to the exact path
        if (half == 1) s_b1[row] = b1, s_b2[row] = b2, s_k1[row] = k1;
        __syncthreads();
        if (half == 0 && row < tile.rows) {
          const float o1 = s_b1[row], o2 = s_b2[row];
          if (o1 > b1) {
            b2 = fmaxf(b1, o2), b1 = o1, k1 = s_k1[row];
          } else {
            b2 = fmaxf(o1, b2);
          }
          s_lab[row] = k1;
          if (!(b1 - b2 >= a.tau)) s_amb[atomicAdd(&s_namb, 1)] = row;   // also catches NaN
        }
        __syncthreads();
        const int namb = s_namb;
        KMS(5);
        KMS_VAL(14, namb);
        // exact re-check: the very fmaf chain of the fp32 kernel (d ascending from 0), first
        // index on ties, against the fp32 prototypes in shared memory
        for (int i = warp; i < namb; i += kSmWarps) {
          const int r = s_amb[i];
#ifdef SPML_KM_TRACE
          const long long t0__ = clock64();
#endif
          const int bk = exact_row_label(xs + r * dim, pf, dim, kb, lane);
          if (lane == 0) s_lab[r] = bk;
#ifdef SPML_KM_TRACE
          KMS_VAL(15, clock64() - t0__);
#endif
        }
        __syncthreads();
        KMS(6);
        if (tid < tile.rows) {
          int lab = s_lab[tid];
          // an input outside the fixed-point range (|x| > 8, NaN) or an undeclared initial label
          // was seen in an M-step: the sums are meaningless, say so instead of returning ids
          if (it == p.iterations && *reinterpret_cast<volatile int*>(p.poison)) lab = -1;
          if (!resident || it == p.iterations) {
            if (p.labels_out) p.labels_out[tile.row0 + tid] = lab;
            if (p.labels_out64 && (it == p.iterations || !p.labels_out))
              p.labels_out64[tile.row0 + tid] = lab;
          }
          if (it < p.iterations && lab != s_old[tid]) s_chg[atomicAdd(&s_nchg, 1)] = tid;
        }
        if (it < p.iterations) {
          // ================================================================ incremental M-step
          __syncthreads();
          const int nchg = s_nchg;
          for (int i = warp; i < nchg; i += kSmWarps) {
            const int r = s_chg[i];
            const int to = s_lab[r] * dim, from = s_old[r];
            for (int d = lane; d < dim; d += 32) {
              const float v = xs[r * dim + d];
              bad |= !(fabsf(v) <= 8.f);
              int hi, lo;
              split_fixed(v, hi, lo);
              atomicAdd(&s_hi[to + d], hi);
              atomicAdd(&s_lo[to + d], lo);
              if ((unsigned)from < (unsigned)kb) {   // (a row whose initial label was invalid has
                atomicAdd(&s_hi[from * dim + d], -hi);   //  nothing to take back)
                atomicAdd(&s_lo[from * dim + d], -lo);
              }
            }
          }
        }
      }

      if (it < p.iterations) {
        ++pending;
        bool flush = lt + 1 >= lt1 || pending >= kSmMaxTilesPerFlush;
        if (!flush) {
          Tile next;
          kms_locate_tile(s_off, lt + 1, next);
          flush = next.b != b;
        }
        const bool any_bad = __syncthreads_or(bad);
        KMS(7);
        if (any_bad && tid == 0) *p.poison = 1;   // published before this pass is counted
        if (flush) {
          // ---- non-zero entries -> the image's global sums (64-bit reductions), zero for reuse
          long long* tot_it = p.sums + (size_t)it * per_iter + (size_t)b * per_img;
          const int rot = (int)(((unsigned)blockIdx.x * 2654435761u >> 8) % (unsigned)per_img) & ~31;
          for (int i1 = tid; i1 < per_img; i1 += kGemmThreads) {
            int i = i1 + rot;                 // staggered like the reads: spread the L2 lines
            i -= i >= per_img ? per_img : 0;
            const int hi = s_hi[i], lo = s_lo[i];
            if ((hi | lo) != 0) {
              atomic_add_i64(&tot_it[i], (long long)hi * 65536ll + lo);
              s_hi[i] = 0;
              s_lo[i] = 0;
            }
          }
          __syncthreads();
          KMS(8);
          if (tid == 0) {
            kms_fence_acq_rel_gpu();   // cumulative: the CTA's reductions (ordered by the barrier) first
            atomicAdd(p.done + (size_t)it * p.batch + b, (unsigned)pending);
          }
          pending = 0;
        }
        KMS(9);
      }
      if (pf_on) cur ^= 1, lead_cur = lead_nxt;
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 2 * kSmBN);
  }
}

// ------------------------------------------------------------------------- host side

static bool kmeans_small_geometry(int dim, int num_clusters, int* nkb, int* bn, int* prefetch,
                                  size_t* smem) {
  const int blocks = (dim + 63) / 64;
  if (dim < 1 || blocks > 2 || num_clusters < 1 || num_clusters > kSmBN) return false;
  if (num_clusters * dim > (1 << 15)) return false;
  const size_t tile = ((size_t)BM * dim + 4) * sizeof(float);
  *bn = num_clusters <= 64 ? 64 : kSmBN;
  const size_t fixed = 1024 + (size_t)2 * blocks * kSmBlockBytes +            // A hi / lo
                       (size_t)2 * blocks * *bn * 128 +                       // B hi / lo
                       (size_t)num_clusters * dim * (sizeof(float) + 2 * sizeof(int)) + 64;   // sums, fp32 rows
  if (fixed + tile > 220 * 1024) return false;
  *nkb = blocks;
  *prefetch = fixed + 2 * tile <= 220 * 1024;
  *smem = fixed + (*prefetch ? 2 : 1) * tile;
  return true;
}

bool kmeans_small_supported(int dim, int num_clusters, int batch, int64_t rows) {
  int nkb, bn, prefetch;
  size_t smem;
  // (offsets are cached as int32 in shared memory; an image has fewer than 2^16 tiles)
  if (batch > kSmMaxBatch || rows >= (1ll << 22) * batch || rows >= (1ll << 31)) return false;
  return kmeans_small_geometry(dim, num_clusters, &nkb, &bn, &prefetch, &smem);
}

int kmeans_small_launch(const KmeansArgs& p, int sms, cudaStream_t st) {
  KmeansSmallArgs a{};
  size_t smem = 0;
  if (!kmeans_small_geometry(p.dim, p.num_clusters, &a.nkb, &a.bn, &a.prefetch, &smem)) {
    set_error("kmeans(small): dim %d, %d clusters are not supported", p.dim, p.num_clusters);
    return SPML_E_UNSUPPORTED;
  }
  a.k = p;
  a.ksteps = (p.dim + 15) / 16;
  a.tau = 1e-4f;
  SPML_CUDA(cudaFuncSetAttribute(kmeans_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_img;
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sms);
  void* args[] = {&a};
  SPML_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kmeans_small_kernel), dim3(grid),
                                        dim3(kGemmThreads), args, smem, st));
  SPML_LAUNCH_CHECK("kmeans_small_kernel");
  return SPML_OK;
}

}  // namespace spml

#ifdef SPML_KM_TRACE
extern "C" int spml_debug_kms_trace_all(long long* trace) {
  return cudaMemcpyFromSymbol(trace, spml::g_kms_trace_all, sizeof(long long) * 160 * 16 * 16) ==
                 cudaSuccess
             ? 0 : -2;
}
extern "C" int spml_debug_kms_trace(long long* trace) {
  return cudaMemcpyFromSymbol(trace, spml::g_kms_trace, sizeof(long long) * 16 * 16) == cudaSuccess
             ? 0 : -2;
}
#endif
