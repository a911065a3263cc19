// A4-A6 when one image fits in the shared memory of ONE thread-block cluster (K <= 128
// clusters, the 512 x 512 configurations): spherical k-means with no grid-wide synchronisation
// and (almost) no global-memory traffic between the passes.
//
// kmeans_small.cu spreads an image over ~116 CTAs (one 128-pixel tile each) and separates two
// passes by an all-to-all through L2: 64-bit reductions into the image's sums, a fence, a tile
// counter, a poll, and K x dim sums pulled back into every CTA.  Here
//   * an image is owned by a cluster of up to 16 CTAs (non-portable size); a CTA keeps its
//     <= 8 pixel tiles resident as ONE fp16 MMA operand (no-swizzle K-major core matrices, so the
//     K extent is dim rounded up to 16, not to 64) for all passes: 160 bytes per pixel at
//     dim = 66;
//   * the E-step is one fp16 product per score (tcgen05, kind::f16, fp32 accumulation in TMEM,
//     all tiles of the CTA in flight in separate accumulators, issued by a warp that does
//     nothing else).  fp16 inputs bound the error of a score by 2^-10 |x| (Cauchy-Schwarz), so
//     a row whose best two scores are further apart than tau = 2^-9 |x| (+ slack) has its exact
//     label; the others (about 1 %) are settled by the very fmaf chain of the fp32 kernel,
//     evaluated only for the candidates within tau of the best, one thread per (row,
//     candidate), on fp32 rows staged in shared memory with cp.async while the E-step runs;
//   * the M-step is incremental as in kmeans_small.cu (only rows whose label changed move),
//     into per-CTA 32-bit shared sums of the two halves of round(x 2^32);
//   * between two passes: cluster barrier -> CTA r adds, through distributed shared memory and
//     only from the CTAs whose bit mask says they touched them, the contributions to ITS
//     prototypes (k = r mod cluster size) to running 64-bit totals, normalises them, stores the
//     fp16 operand row into every CTA of the cluster (16-byte DSMEM stores) and the fp32 row into
//     a global scratch (every CTA copies the scratch into its shared memory behind the MMAs)
//     -> cluster barrier.
// Sums are exact integers and the recheck is the fp32 chain, so the labels are those of the
// other three kernels bit for bit (tests/test_gpu_ops.py::test_kmeans_tensor_core_equals_fp32).
// Measured against kmeans_small.cu in kmeans.cu::kmeans_path (default from batch 4 on) and
// DESIGN.md section 4; per-phase timeline: scripts/trace_kmeans_cluster.py.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>

#include "kmeans.cuh"
#include "tc_common.cuh"

namespace spml {
namespace cg = cooperative_groups;

constexpr int kKcThreads = 512;
constexpr int kKcWarps = kKcThreads / 32;
constexpr int kKcMaxCluster = 16;
constexpr int kKcMaxTiles = 8;        // pixel tiles per CTA: 8 accumulators of 64 columns = TMEM
constexpr int kKcAmbCap = 192;        // ambiguous rows per CTA and pass settled by candidate pairs
constexpr int kKcStage = 40;          // ... of which the first ones have their fp32 row staged in shared memory
constexpr int kKcCand = 4;            // candidates per ambiguous row (more: warp-per-row path)
constexpr int kKcSmemLimit = 227 * 1024;

#ifdef SPML_KM_TRACE
// [CTA][pass][slot] clock64 stamps (SM-local clocks: differences inside one CTA only)
__device__ long long g_kc_trace[256 * 16 * 16];
#define KCT(slot)                                                                      \
  do {                                                                                 \
    if (threadIdx.x == 0 && it < 16 && blockIdx.x < 256)                               \
      g_kc_trace[(blockIdx.x * 16 + it) * 16 + (slot)] = clock64();                    \
  } while (0)
#define KCT_VAL(slot, v)                                                               \
  do {                                                                                 \
    if (threadIdx.x == 0 && it < 16 && blockIdx.x < 256)                               \
      g_kc_trace[(blockIdx.x * 16 + it) * 16 + (slot)] = (v);                          \
  } while (0)
#else
#define KCT(slot) do { } while (0)
#define KCT_VAL(slot, v) do { } while (0)
#endif

struct KmeansClusterArgs {
  KmeansArgs k;
  int kp;          // K extent of the operands: dim rounded up to 16
  int bn;          // prototype rows of the operand tile (UMMA N): 64 or 128
  int max_tiles;   // pixel tiles a CTA can hold
  int own;         // prototypes a CTA owns at most: ceil(K / cluster size)
  int tmem_cols;
};

// byte offsets of the dynamic shared memory (from a 128-byte aligned base)
struct KcLayout {
  uint32_t a, b, hi, lo, tot, hst, pf, xst, amb, pk, sc, lab, old, total;
};
__host__ __device__ inline uint32_t kc_up(uint32_t x, uint32_t al) { return (x + al - 1) / al * al; }
__host__ __device__ inline KcLayout kc_layout(int K, int dim, int kp, int bn, int max_tiles, int own) {
  KcLayout L;
  uint32_t o = 0;
  L.a = o, o += (uint32_t)max_tiles * BM * kp * 2;        // [tile][128 x kp] fp16, core matrices
  L.b = o, o += (uint32_t)bn * kp * 2;                     // [bn x kp] fp16
  L.hi = o, o += (uint32_t)K * dim * 4;                    // this CTA's contribution, high halves
  L.lo = o, o = kc_up(o + (uint32_t)K * dim * 4, 16);
  L.tot = o, o = kc_up(o + (uint32_t)own * dim * 8, 16);   // running totals of the owned prototypes
  L.hst = o, o = kc_up(o + (uint32_t)own * kp * 2, 16);    // the owned prototypes as fp16 operand rows (staging)
  L.pf = o, o += kc_up((uint32_t)K * dim, 4) * 4;          // fp32 prototypes of the pass (copy of the scratch)
  L.xst = o, o = kc_up(o + (uint32_t)kKcStage * dim * 4, 16);    // fp32 rows of the first ambiguous rows
  const uint32_t need_a = kKcAmbCap * (1 + 2 * kKcCand), need_c = (uint32_t)max_tiles * BM;
  const uint32_t lists = need_a > need_c ? need_a : need_c;
  L.amb = o;                                               // ambiguous rows | changed rows (later)
  L.pk = o + kKcAmbCap * 4;
  L.sc = L.pk + kKcAmbCap * kKcCand * 4;
  o += lists * 4;
  L.lab = o, o += (uint32_t)max_tiles * BM;                // uint8 labels of this pass
  L.old = o, o = kc_up(o + (uint32_t)max_tiles * BM, 16);  // ... of the previous pass
  L.total = o;
  return L;
}

// Operand layout (fp16, K-major, no swizzle): 8 x 8 core matrices of 128 contiguous bytes; the
// matrices of one 8-element K chunk follow each other down the rows, i.e. element (row, d) sits
// at (d / 8) * rows * 16 + row * 16 + (d % 8) * 2 (UMMA descriptor: SBO = 128, LBO = 16 rows).

// kind::f16 instruction descriptor, fp16 x fp16 -> fp32, both operands K-major
// (tc_common.cuh::umma_idesc_bf16 with a_format = b_format = 0)
__host__ __device__ constexpr uint32_t kc_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void kc_cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void kc_cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void kc_cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
// one fp32 row global -> shared, asynchronously, by ONE thread (a row that has just turned out
// ambiguous: its values are needed by the exact re-check after the E-step)
__device__ __forceinline__ void kc_stage_row(float* dst, const float* src, int dim) {
  if ((dim & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
    for (int d = 0; d < dim; d += 2) kc_cp_async_8(dst + d, src + d);
  } else {
    for (int d = 0; d < dim; ++d) kc_cp_async_4(dst + d, src + d);
  }
}
// the fmaf chain of the fp32 kernel (d ascending from 0) on two rows in shared memory
__device__ __forceinline__ float kc_chain_shared(const float* __restrict__ xr,
                                                 const float* __restrict__ pr, int dim) {
  float acc = 0.f;
  int d = 0;
  if ((dim & 1) == 0 && ((tc::smem_u32(xr) | tc::smem_u32(pr)) & 7u) == 0) {
#pragma unroll 2
    for (; d + 8 <= dim; d += 8) {
      float2 xv[4], pv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float2*>(xr + d + 2 * j);
#pragma unroll
      for (int j = 0; j < 4; ++j) pv[j] = *reinterpret_cast<const float2*>(pr + d + 2 * j);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = fmaf(xv[j].x, pv[j].x, acc);
        acc = fmaf(xv[j].y, pv[j].y, acc);
      }
    }
  }
  for (; d < dim; ++d) acc = fmaf(xr[d], pr[d], acc);
  return acc;
}
__device__ __forceinline__ void kc_prefetch_row(const float* row, int dim) {
  const char* p0 = reinterpret_cast<const char*>(row);
  for (int o = 0; o < dim * 4 + 127; o += 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + min(o, dim * 4 - 4)) : "memory");
}

// the fp32 score of the fp32 kernel: acc = fmaf(x[d], p[d], acc), d ascending from 0, for the
// rare warp-per-row path: xr is a row in global memory, pr a prototype row in shared memory.
template <int kSlots>
__device__ __forceinline__ float kc_exact_score(const float* __restrict__ xr,
                                                const float* __restrict__ pr, int dim) {
  float acc = 0.f;
  if ((dim & 1) == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(pr)) & 7) == 0) {   // 8-byte aligned rows
    constexpr int kH = 24;   // float2 of the prototype row in flight together
    const int n2 = dim >> 1;
    for (int h0 = 0; h0 < n2; h0 += kH) {
      float2 pv[kH];
#pragma unroll
      for (int c = 0; c < kH; ++c) pv[c] = reinterpret_cast<const float2*>(pr)[min(h0 + c, n2 - 1)];
#pragma unroll
      for (int c0 = 0; c0 < kH; c0 += 8) {
        if (h0 + c0 < n2) {
          float2 xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            xv[j] = __ldca(reinterpret_cast<const float2*>(xr) + min(h0 + c0 + j, n2 - 1));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (h0 + c0 + j < n2) {
              acc = fmaf(xv[j].x, pv[c0 + j].x, acc);
              acc = fmaf(xv[j].y, pv[c0 + j].y, acc);
            }
        }
      }
    }
    return acc;
  }
  for (int d = 0; d < dim; ++d) acc = fmaf(__ldca(xr + d), pr[d], acc);
  return acc;
}

template <int kSlots>
__global__ void __launch_bounds__(kKcThreads, 1) kmeans_cluster_kernel(const KmeansClusterArgs a) {
  extern __shared__ uint8_t kc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_tile[kKcMaxTiles];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_namb, s_ndefer, s_nchg, s_maxn;
  __shared__ uint32_t s_touch[4];                     // prototypes this CTA's contribution touches
  __shared__ uint32_t s_rmask[kKcMaxCluster * 4];     // ... of every CTA of the cluster

  cg::cluster_group cluster = cg::this_cluster();
  const KmeansArgs& p = a.k;
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = (int)blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int dim = p.dim, K = p.num_clusters, kp = a.kp, bn = a.bn;
  const int per_img = K * dim;
  const int T = p.iterations;

  const int first = p.img_off ? p.img_off[b] : 0;
  const int last = p.img_off ? p.img_off[b + 1] : (int)p.rows_total;
  const int tiles_b = (last - first + BM - 1) >> 7;
  if ((tiles_b + cs - 1) / cs > a.max_tiles) {
    // more rows than the caller declared (max_rows_per_image): the whole cluster leaves
    if (tid == 0) *p.poison = 1;
    return;
  }
  const int t_lo = (int)((int64_t)tiles_b * rank / cs), t_hi = (int)((int64_t)tiles_b * (rank + 1) / cs);
  const int nt = t_hi - t_lo;
  const int64_t row_lo = (int64_t)first + (int64_t)t_lo * BM;
  const int nrows = nt > 0 ? min(last, first + t_hi * BM) - (first + t_lo * BM) : 0;
  const int kb = p.k_per_image ? p.k_per_image[b] : K;
  const float* __restrict__ x = p.x + row_lo * dim;
  // fp32 prototypes of the image (workspace): written by their owners, read by the re-checks
  // (16-byte aligned blocks: they are copied into shared memory with 16-byte cp.async)
  const int pf_stride = (K * dim + 3) & ~3;
  float* gpf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.protos) + 15) & ~uintptr_t(15)) +
               (size_t)b * pf_stride;

  uint8_t* smem = kc_smem_raw + ((128u - (tc::smem_u32(kc_smem_raw) & 127u)) & 127u);
  const KcLayout L = kc_layout(K, dim, kp, bn, a.max_tiles, a.own);
  uint8_t* a_op = smem + L.a;
  uint8_t* b_op = smem + L.b;
  int* s_hi = reinterpret_cast<int*>(smem + L.hi);
  int* s_lo = reinterpret_cast<int*>(smem + L.lo);
  long long* s_tot = reinterpret_cast<long long*>(smem + L.tot);
  __half* s_hst = reinterpret_cast<__half*>(smem + L.hst);
  float* pf = reinterpret_cast<float*>(smem + L.pf);          // [K][dim]
  float* s_xst = reinterpret_cast<float*>(smem + L.xst);      // [kKcStage][dim]
  uint32_t* s_amb = reinterpret_cast<uint32_t*>(smem + L.amb);   // row | candidates << 16
  int* s_pk = reinterpret_cast<int*>(smem + L.pk);               // [slot][kKcCand] prototype index
  float* s_sc = reinterpret_cast<float*>(smem + L.sc);           // [slot][kKcCand] exact score
  uint32_t* s_chg = s_amb;                                       // row | old << 12 | new << 20
  uint8_t* s_lab = smem + L.lab;
  uint8_t* s_old = smem + L.old;
  const uint32_t tile_bytes = (uint32_t)BM * kp * 2;

  // ---- everything starts as zeros: operand padding (0 x garbage = NaN), sums, totals
  for (uint32_t i = tid; i < L.total / 16; i += kKcThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < kKcMaxTiles; ++i) tc::mbar_init(&bar_tile[i], 1);
    tc::fence_barrier_init();
    s_maxn = 0;
  }
  if (tid < 4) s_touch[tid] = 0;
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, (uint32_t)a.tmem_cols);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  int it = 0;
  KCT(0);

  // ================================================================== pass 0
  // rows -> fp16 operand, largest row norm (for tau), initial M-step.  A warp takes a
  // contiguous range of rows and sums runs of equal labels in registers (the seeds of
  // common.py:145-153 are blocks of the raster: ~20 rows per run); the next four rows are
  // requested before the current four are worked on.
  bool bad = false;
  {
    const int rpw = (nrows + kKcWarps - 1) / kKcWarps;
    const int r0 = warp * rpw, r1 = min(nrows, r0 + rpw);
    int run_hi[kSlots], run_lo[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) run_hi[s] = run_lo[s] = 0;
    int run_lab = -1, in_run = 0;
    float max_ss = 0.f;
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
      if (lane == 0) atomicOr(&s_touch[run_lab >> 5], 1u << (run_lab & 31));
    };
    uint32_t col_off[kSlots];   // byte offset of this lane's channels inside an operand row group
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int d = lane + 32 * s;
      col_off[s] = (uint32_t)(d >> 3) * (uint32_t)(BM * 16) + (uint32_t)(d & 7) * 2u;
    }
    auto load4 = [&](int r, float (&v)[4][kSlots], int (&lab4)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u;
        const bool ok = rr < r1;
        lab4[u] = ok ? p.labels_in[row_lo + rr] : -1;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const int d = lane + 32 * s;
          v[u][s] = (ok && d < dim) ? __ldcg(x + (int64_t)rr * dim + d) : 0.f;
        }
      }
    };
    auto work4 = [&](int r, const float (&v)[4][kSlots], const int (&lab4)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u;
        if (rr >= r1) continue;   // warp-uniform
        float ss = 0.f;
        uint8_t* row_ptr = a_op + (uint32_t)(rr >> 7) * tile_bytes + (uint32_t)(rr & (BM - 1)) * 16u;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const int d = lane + 32 * s;
          ss = fmaf(v[u][s], v[u][s], ss);
          bad |= !(fabsf(v[u][s]) <= 8.f);
          if (d < dim) *reinterpret_cast<__half*>(row_ptr + col_off[s]) = __float2half_rn(v[u][s]);
        }
        max_ss = fmaxf(max_ss, ss);   // per lane; summed over the lanes below
        int lab = lab4[u];
        if ((unsigned)lab >= (unsigned)kb) lab = -1, bad = true;
        if (lane == 0) s_lab[rr] = (uint8_t)(lab < 0 ? 255 : lab);
        if (lab < 0) continue;
        if (lab != run_lab || in_run >= 2048) {   // (2^19 x 2048 < 2^31: the halves cannot overflow)
          flush();
          run_lab = lab;
          in_run = 0;
        }
        ++in_run;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          int hi, lo;
          split_fixed(v[u][s], hi, lo);
          run_hi[s] += hi;
          run_lo[s] += lo;
        }
      }
    };
    float va[4][kSlots], vb[4][kSlots];
    int la[4], lb[4];
    load4(r0, va, la);
    for (int r = r0; r < r1; r += 8) {
      load4(r + 4, vb, lb);
      work4(r, va, la);
      load4(r + 8, va, la);
      work4(r + 4, vb, lb);
    }
    flush();
    // largest squared norm of the CTA's rows (the sum of the per-lane maxima bounds every row)
    max_ss = warp_sum(max_ss);
    if (lane == 0) atomicMax(&s_maxn, __float_as_int(max_ss));
  }
  // the fp32 rows of the owned prototypes start as the zero vector (an empty cluster is never
  // rebuilt, and the scratch holds whatever the last call left there)
  for (int j = warp; rank + j * cs < kb; j += kKcWarps)
    for (int d = lane; d < dim; d += 32) gpf[(rank + j * cs) * dim + d] = 0.f;
  if (__syncthreads_or(bad) && tid == 0) *p.poison = 1;
  // |score error| <= (2 u + u^2) |x| |p| + subnormal terms with u = 2^-11 (fp16), |p| <= 1, plus
  // the accumulation in the tensor core; a gap of twice that decides the label
  const float nrm = sqrtf(__int_as_float(s_maxn));
  const float tau = fmaf(nrm, 0.001953125f * 1.002f + 6e-5f, 7e-7f * (1.f + nrm));
  KCT(1);

  const uint32_t idesc = kc_idesc_f16(BM, bn);
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);   // SBO = 128, version 1, no swizzle
  const uint32_t a_lo0 = tc::umma_desc_lo(tc::smem_u32(a_op), BM * 16);
  const uint32_t b_lo0 = tc::umma_desc_lo(tc::smem_u32(b_op), (uint32_t)bn * 16);
  const int ksteps = kp >> 4;
  const int tpr = min(a.max_tiles, a.tmem_cols / bn);   // tiles per round (accumulators in TMEM)
  uint32_t phase = 0;                                   // parity bit per bar_tile
  const int sp = warp & 3, grp = warp >> 2;             // TMEM sub-partition, tile group
  const int nchunk = kp >> 3;                           // 16-byte pieces of an operand row

  for (it = 1; it <= T; ++it) {
    // ================================================================ prototypes of pass it - 1
    KCT(0);
    cluster.sync();   // every CTA's contribution to the sums is complete
    KCT(2);
    if (tid < cs * 4) s_rmask[tid] = cluster.map_shared_rank(s_touch, tid >> 2)[tid & 3];
    __syncthreads();
    for (int j = warp; rank + j * cs < kb; j += kKcWarps) {
      const int k = rank + j * cs;
      // the CTAs whose rows moved into or out of prototype k in the last M-step; nobody: the
      // prototype (and its operand row everywhere) stays what it is
      const bool mine = lane < cs && ((s_rmask[lane * 4 + (k >> 5)] >> (k & 31)) & 1u);
      unsigned srcs = __ballot_sync(0xffffffffu, mine);
      if (srcs == 0) continue;
      long long delta[kSlots];
#pragma unroll
      for (int s = 0; s < kSlots; ++s) delta[s] = 0;
      while (srcs) {   // four CTAs' words in flight together
        int rh[4][kSlots], rl[4][kSlots];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool any = srcs != 0;
          const int src = any ? __ffs(srcs) - 1 : 0;
          srcs &= srcs - 1;
          const int* rhi = cluster.map_shared_rank(s_hi, src);
          const int* rlo = cluster.map_shared_rank(s_lo, src);
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            const int d = lane + 32 * s;
            const bool in = any && d < dim;
            rh[q][s] = in ? rhi[k * dim + d] : 0;
            rl[q][s] = in ? rlo[k * dim + d] : 0;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int s = 0; s < kSlots; ++s) delta[s] += (long long)rh[q][s] * 65536ll + rl[q][s];
      }
      KCT(9);
      float v[kSlots];
      float ss = 0.f;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int d = lane + 32 * s;
        v[s] = 0.f;
        if (d < dim) {
          const long long tot = s_tot[j * dim + d] + delta[s];
          s_tot[j * dim + d] = tot;
          v[s] = fixed_to_float(tot);
        }
        ss = fmaf(v[s], v[s], ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      // common.py:39: sum / max(||sum||, eps); an empty cluster is the zero vector
      const float n2 = sqrtf(ss);
      const float div = n2 >= p.eps ? n2 : p.eps;
      const float rcp = div_reciprocal(div);
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int d = lane + 32 * s;
        if (d < dim) {
          const float u = div_by(v[s], div, rcp);
          gpf[k * dim + d] = u;                         // fp32 row for the re-checks: through L2
          s_hst[j * kp + d] = __float2half_rn(u);       // (the padding [dim, kp) stays zero)
        }
      }
      __syncwarp();
      KCT(10);
      // the fp16 operand row into every CTA of the cluster, 16 bytes (one core-matrix row) a store
      for (int idx = lane; idx < cs * nchunk; idx += 32) {
        const int dst = idx / nchunk, c = idx - dst * nchunk;
        const uint4 piece = *reinterpret_cast<const uint4*>(s_hst + j * kp + c * 8);
        *reinterpret_cast<uint4*>(cluster.map_shared_rank(b_op, dst) + (uint32_t)c * (uint32_t)(bn * 16) +
                                  (uint32_t)k * 16u) = piece;
      }
    }
    tc::fence_proxy_async_all();   // the operand rows written above -> the tensor core of their CTA
    KCT(3);
    cluster.sync();   // the operand rows of every owner have arrived
    KCT(4);
    // the fp32 prototypes of this pass (written by their owners in front of the barrier) into
    // shared memory, behind the MMAs and the epilogue: nobody needs them before the re-check
    for (int i = tid; i < pf_stride / 4; i += kKcThreads) kc_cp_async_16(pf + 4 * i, gpf + 4 * i);
    for (int i = tid; i < per_img; i += kKcThreads) s_hi[i] = 0, s_lo[i] = 0;
    for (int i = tid; i < (nrows + 3) / 4; i += kKcThreads)
      reinterpret_cast<uint32_t*>(s_old)[i] = reinterpret_cast<const uint32_t*>(s_lab)[i];
    if (tid == 0) s_namb = 0, s_ndefer = 0, s_nchg = 0;
    if (tid < 4) s_touch[tid] = 0;
    tc::fence_proxy_async();
    __syncthreads();

    // ================================================================ E-step
    for (int t0 = 0; t0 < nt; t0 += tpr) {
      const int ntr = min(tpr, nt - t0);
      if (warp == 0) {
        tc::tcgen05_fence_after();
        if (tc::elect_one()) {
          tc::fence_proxy_async();
          for (int tt = 0; tt < ntr; ++tt) {
            uint32_t al = a_lo0 + (uint32_t)(t0 + tt) * (tile_bytes >> 4);
            uint32_t bl = b_lo0;
            const uint32_t acc = tmem_base + (uint32_t)(tt * bn);
            for (int ks = 0; ks < ksteps; ++ks) {   // 16 fp16 = two K chunks
              tc::umma_bf16_words(acc, al, desc_hi, bl, desc_hi, idesc, ks > 0);
              al += (2 * BM * 16) >> 4, bl += (uint32_t)(2 * bn * 16) >> 4;
            }
            tc::umma_commit(&bar_tile[tt]);
          }
        }
        __syncwarp();
      }
      KCT(8);
      // warp 0 has just spent the time of all the MMAs issuing them: the tiles of its TMEM
      // sub-partition are shared by warps 4, 8 and 12, those of the others by four warps each
      const int nshare = sp == 0 ? kKcWarps / 4 - 1 : kKcWarps / 4;
      const int myshare = sp == 0 ? grp - 1 : grp;
      for (int tt = myshare; tt < ntr && myshare >= 0; tt += nshare) {
        tc::mbar_wait(&bar_tile[tt], (phase >> tt) & 1u);
        phase ^= 1u << tt;
        tc::tcgen05_fence_after();
        const int rloc = (t0 + tt) * BM + sp * 32 + lane;
        const bool valid = rloc < nrows;
        const uint32_t taddr = tmem_base + (uint32_t)(tt * bn) + (static_cast<uint32_t>(sp * 32) << 16);
        float b1 = -INFINITY, b2 = -INFINITY;
        int k1 = 0;
        for (int cb = 0; cb < kb; cb += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(taddr + cb, v);
          tc::tmem_ld_wait();
          const int live = kb - cb;   // columns of this chunk that exist (warp-uniform)
#pragma unroll
          for (int u0 = 0; u0 < 32; u0 += 4) {
            if (u0 < live) {
#pragma unroll
              for (int u = u0; u < u0 + 4; ++u) {
                float sc = __uint_as_float(v[u]);
                if (u0 + 4 > live) sc = u < live ? sc : -INFINITY;
                b2 = fmaxf(b2, fminf(sc, b1));   // second best so far (a tie counts)
                k1 = sc > b1 ? cb + u : k1;
                b1 = fmaxf(b1, sc);
              }
            }
          }
        }
        const bool amb = valid && !(b1 - b2 >= tau);   // also catches NaN
        if (valid) s_lab[rloc] = (uint8_t)k1;
        // the fp32 row of a row that is ambiguous or moves is needed in a moment: into L1
        if (valid && !amb && it < T && k1 != (int)s_old[rloc]) kc_prefetch_row(x + (int64_t)rloc * dim, dim);
        if (__any_sync(0xffffffffu, amb)) {
          // ---- the candidates of an ambiguous row: every prototype within tau of the best
          int slot = -1, cnt = 0;
          if (amb) {
            slot = atomicAdd(&s_namb, 1);
            if (slot >= kKcAmbCap) slot = -1;
            if (slot >= 0 && slot < kKcStage) kc_stage_row(s_xst + slot * dim, x + (int64_t)rloc * dim, dim);
            else kc_prefetch_row(x + (int64_t)rloc * dim, dim);
          }
          const float thr = b1 - tau;
          for (int cb = 0; cb < kb; cb += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(taddr + cb, v);
            tc::tmem_ld_wait();
            const int live = kb - cb;
            // candidates of this chunk as a bit mask (bitwise predicates and selects: branches
            // inside the unrolled loop cost a convergence barrier each), then the few set bits
            unsigned mc = 0;
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              const float sc = __uint_as_float(v[u]);
              mc |= ((sc >= thr) | (sc != sc)) ? (1u << u) : 0u;
            }
            if (live < 32) mc &= (1u << live) - 1u;
            if (slot < 0) mc = 0;
            while (mc) {
              const int u = __ffs(mc) - 1;
              mc &= mc - 1;
              if (cnt < kKcCand) s_pk[slot * kKcCand + cnt] = cb + u;
              ++cnt;
            }
          }
          if (amb) {
            if (slot >= 0) s_amb[slot] = (uint32_t)rloc | ((uint32_t)min(cnt, 255) << 16);
            if (slot < 0 || cnt > kKcCand || cnt == 0) {   // -> the warp-per-row path below
              s_lab[rloc] = 255;
              atomicAdd(&s_ndefer, 1);
            }
          }
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");   // fp32 prototypes, rows of the ambiguous rows
      tc::tcgen05_fence_before();
      __syncthreads();   // the accumulators are free for the next round
    }
    KCT(5);

    // ================================================================ exact re-check
    {
      const int namb = min(s_namb, kKcAmbCap);
      KCT_VAL(14, namb);
      for (int i = tid; i < namb * kKcCand; i += kKcThreads) {
        const uint32_t rec = s_amb[i / kKcCand];
        const int cnt = (int)(rec >> 16), c = i % kKcCand;
        if (cnt <= kKcCand && c < cnt) {
          const int k = s_pk[i];
          s_sc[i] = i / kKcCand < kKcStage
                        ? kc_chain_shared(s_xst + (i / kKcCand) * dim, pf + k * dim, dim)
                        : kc_exact_score<kSlots>(x + (int64_t)(rec & 0xffffu) * dim, pf + k * dim, dim);
        }
      }
      __syncthreads();
      for (int sl = tid; sl < namb; sl += kKcThreads) {
        const uint32_t rec = s_amb[sl];
        const int cnt = (int)(rec >> 16);
        if (cnt < 1 || cnt > kKcCand) continue;
        float bv = s_sc[sl * kKcCand];
        int bk = s_pk[sl * kKcCand];
        for (int c = 1; c < cnt; ++c) {   // candidates in ascending index order: first index on ties
          const float v = s_sc[sl * kKcCand + c];
          if (v > bv) bv = v, bk = s_pk[sl * kKcCand + c];
        }
        s_lab[rec & 0xffffu] = (uint8_t)bk;
      }
      if (s_ndefer > 0) {
        // rows with many candidates (or more ambiguous rows than slots): every prototype, one
        // warp per row, lanes across the prototypes.  (The barrier keeps the label stores just
        // above apart from the scan for the 255 markers below: they never touch a marked row,
        // but they are stores to the array this loop reads.)
        __syncthreads();
        for (int r = warp; r < nrows; r += kKcWarps) {
          if (s_lab[r] != 255) continue;   // warp-uniform
          float bv = -INFINITY;
          int bk = 0;
          for (int k = lane; k < kb; k += 32) {
            const float v = kc_exact_score<kSlots>(x + (int64_t)r * dim, pf + k * dim, dim);
            if (v > bv) bv = v, bk = k;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ov > bv || (ov == bv && ok < bk)) bv = ov, bk = ok;
          }
          __syncwarp();
          if (lane == 0) s_lab[r] = (uint8_t)bk;
        }
      }
      __syncthreads();
    }
    KCT(6);

    if (it < T) {
      // ================================================================ incremental M-step
      for (int r = tid; r < nrows; r += kKcThreads) {
        const uint32_t nl = s_lab[r], ol = s_old[r];
        if (nl != ol) s_chg[atomicAdd(&s_nchg, 1)] = (uint32_t)r | (ol << 12) | (nl << 20);
      }
      __syncthreads();
      const int nchg = s_nchg;
      KCT_VAL(15, nchg);
      constexpr int kFly = 4;   // rows in flight per warp (prefetched into L1 by the E-step)
      for (int i0 = warp; i0 < nchg; i0 += kFly * kKcWarps) {
        float v[kFly][kSlots];
        uint32_t rec[kFly];
#pragma unroll
        for (int u = 0; u < kFly; ++u) {
          const int i = i0 + u * kKcWarps;
          rec[u] = i < nchg ? s_chg[i] : 0xffffffffu;
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            const int d = lane + 32 * s;
            v[u][s] = (i < nchg && d < dim) ? __ldca(x + (int64_t)(rec[u] & 0xfffu) * dim + d) : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < kFly; ++u) {
          if (rec[u] == 0xffffffffu) continue;   // warp-uniform
          const int from = (int)((rec[u] >> 12) & 0xffu), to = (int)(rec[u] >> 20);
          if (lane == 0) {
            atomicOr(&s_touch[to >> 5], 1u << (to & 31));
            if (from < kb) atomicOr(&s_touch[from >> 5], 1u << (from & 31));
          }
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            const int d = lane + 32 * s;
            if (d < dim) {
              int hi, lo;
              split_fixed(v[u][s], hi, lo);
              atomicAdd(&s_hi[to * dim + d], hi);
              atomicAdd(&s_lo[to * dim + d], lo);
              if (from < kb) {   // (a row whose initial label was invalid has nothing to take back)
                atomicAdd(&s_hi[from * dim + d], -hi);
                atomicAdd(&s_lo[from * dim + d], -lo);
              }
            }
          }
        }
      }
      KCT(7);
    }
  }

  // ================================================================ labels out
  __syncthreads();
  const bool poisoned = *reinterpret_cast<volatile int*>(p.poison) != 0;
  for (int r = tid; r < nrows; r += kKcThreads) {
    // an input outside the fixed-point range (|x| > 8, NaN) or an undeclared initial label:
    // the sums are meaningless, say so instead of returning ids
    const int lab = poisoned ? -1 : (int)s_lab[r];
    if (p.labels_out) p.labels_out[row_lo + r] = lab;
    if (p.labels_out64) p.labels_out64[row_lo + r] = lab;
  }
  tc::tcgen05_fence_before();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------- host side

struct KcGeometry {
  int cs, kp, bn, max_tiles, own, tmem_cols, slots;
  size_t smem;
};

static bool kmeans_cluster_geometry(int dim, int num_clusters, int max_rows, KcGeometry* g) {
  if (dim < 1 || dim > SPML_MAX_DIM || num_clusters < 1 || num_clusters > 128 || max_rows < 1)
    return false;
  const int tiles = (max_rows + BM - 1) / BM;
  int cs = 1;
  while (cs < kKcMaxCluster && (tiles + cs - 1) / cs > 4) cs *= 2;
  g->cs = cs;
  g->max_tiles = (tiles + cs - 1) / cs;
  if (g->max_tiles > kKcMaxTiles || g->max_tiles * BM > 4096) return false;
  g->kp = (dim + 15) / 16 * 16;
  g->bn = num_clusters <= 64 ? 64 : 128;
  g->own = (num_clusters + cs - 1) / cs;
  int cols = 32;
  while (cols < std::min(g->max_tiles * g->bn, 512)) cols *= 2;
  g->tmem_cols = cols;
  g->slots = (dim + 31) / 32;
  g->smem = kc_layout(num_clusters, dim, g->kp, g->bn, g->max_tiles, g->own).total + 128;
  return g->smem + 1024 <= (size_t)kKcSmemLimit;   // + the static shared memory of the kernel
}

template <int kSlots>
static int kmeans_cluster_launch_t(const KmeansClusterArgs& a, const KcGeometry& g, int batch,
                                   cudaStream_t st, bool probe) {
  auto kernel = kmeans_cluster_kernel<kSlots>;
  SPML_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  if (g.cs > 8) SPML_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(batch * g.cs));
  cfg.blockDim = dim3(kKcThreads);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)g.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (probe) {
    int clusters = 0;
    SPML_CUDA(cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg));
    return clusters >= 1 ? SPML_OK : SPML_E_UNSUPPORTED;
  }
  SPML_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  SPML_LAUNCH_CHECK("kmeans_cluster_kernel");
  return SPML_OK;
}

static int kmeans_cluster_dispatch(const KmeansClusterArgs& a, const KcGeometry& g, int batch,
                                   cudaStream_t st, bool probe) {
  switch (g.slots) {
    case 1: return kmeans_cluster_launch_t<1>(a, g, batch, st, probe);
    case 2: return kmeans_cluster_launch_t<2>(a, g, batch, st, probe);
    case 3: return kmeans_cluster_launch_t<3>(a, g, batch, st, probe);
    case 4: return kmeans_cluster_launch_t<4>(a, g, batch, st, probe);
    default: return kmeans_cluster_launch_t<5>(a, g, batch, st, probe);
  }
}

// Can a cluster of this shape be resident at all on this device?  (Asked once per shape: the
// answer depends on the GPC layout of the part, not on the call.)
bool kmeans_cluster_supported(int dim, int num_clusters, int batch, int max_rows) {
  KcGeometry g;
  if (batch < 1 || !kmeans_cluster_geometry(dim, num_clusters, max_rows, &g)) return false;
  struct Seen { size_t smem; int cs, slots, ok; };
  static thread_local Seen seen[8];
  static thread_local int nseen = 0;
  for (int i = 0; i < nseen; ++i)
    if (seen[i].smem == g.smem && seen[i].cs == g.cs && seen[i].slots == g.slots) return seen[i].ok != 0;
  KmeansClusterArgs a{};
  const int ok = kmeans_cluster_dispatch(a, g, 1, nullptr, true) == SPML_OK;
  if (!ok) {
    (void)cudaGetLastError();
    clear_error();
  }
  if (nseen < 8) seen[nseen++] = Seen{g.smem, g.cs, g.slots, ok};
  return ok != 0;
}

int kmeans_cluster_launch(const KmeansArgs& p, int max_rows, cudaStream_t st) {
  KcGeometry g;
  if (!kmeans_cluster_geometry(p.dim, p.num_clusters, max_rows, &g)) {
    set_error("kmeans(cluster): dim %d, %d clusters, %d rows per image are not supported", p.dim,
              p.num_clusters, max_rows);
    return SPML_E_UNSUPPORTED;
  }
  KmeansClusterArgs a{};
  a.k = p;
  a.kp = g.kp;
  a.bn = g.bn;
  a.max_tiles = g.max_tiles;
  a.own = g.own;
  a.tmem_cols = g.tmem_cols;
  return kmeans_cluster_dispatch(a, g, p.batch, st, false);
}

}  // namespace spml

#ifdef SPML_KM_TRACE
extern "C" int spml_debug_kc_trace(long long* trace) {
  return cudaMemcpyFromSymbol(trace, spml::g_kc_trace, sizeof(long long) * 256 * 16 * 16) ==
                 cudaSuccess
             ? 0 : -2;
}
#endif
