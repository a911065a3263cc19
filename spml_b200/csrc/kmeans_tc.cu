// A4-A6 with the E-step on the 5th-generation tensor cores
// (reference spml/utils/segsort/common.py:11-97; fp32 CUDA-core twin in kmeans.cu).
//
// Same persistent structure as kmeans.cu: ONE cooperative launch runs every iteration of
// every image; an iteration is separated from the next by per-image flags, not by kernel
// boundaries.  Per 128-pixel tile and iteration:
//
//   E-step   The tile is loaded once as fp32 (row-major in shared memory: the M-step and
//            the exact re-check read it) and split in place into bf16 hi + lo SWIZZLE_128B
//            operand tiles.  The image's prototypes, published by the finalising CTA as bf16
//            hi + lo rows, arrive by TMA.  Per 16-wide K step: A_hi x [P_hi ; P_lo] (N = 256)
//            and A_lo x P_hi (N = 128), fp32 accumulation in TMEM: a cosine carries ~16
//            mantissa bits (|err| < 2e-5).  Epilogue: tcgen05.ld, running best / second best.
//   exact    Segment ids must equal the fp32 result bit for bit, so the tensor-core scores
//            only DECIDE pixels whose best and second-best score are further apart than
//            tau = 1e-4 (5x the error bound).  The others (a fraction of a percent) are
//            re-scored with the very fmaf chain of the fp32 kernel (gemm_nt_tile: d ascending
//            from 0) and the same first-index tie-break: both kernels return identical labels.
//   M-step   Rows are ranked by label inside the tile (so that a warp sees long runs of one
//            label whatever the spatial layout), summed per run in registers as exact
//            fixed point and added to the image's K x dim sums with 64-bit reductions; the
//            sums are replicated kKmReplicas times to spread the hot L2 lines.  The CTA that
//            adds an image's last tile folds the replicas, L2-normalises and publishes.
#include <math.h>

#include <algorithm>

#include "kmeans.cuh"
#include "tc_common.cuh"

namespace spml {

constexpr int kKmBN = 128;                 // prototype columns per accumulator
constexpr int kKmBlockBytes = 128 * 128;   // one 64-wide K block of a 128-row bf16 tile
constexpr int kKmWarps = kGemmThreads / 32;

#ifdef SPML_KM_TRACE
__device__ long long g_kmt_trace[16 * 16];        // CTA 0: phases of its last tile of each pass
__device__ long long g_kmt_fin[16 * 16];          // the finalising CTA's sub-phases per pass
__device__ long long g_kmt_cta[160 * 12 * 4];     // [CTA][pass][start, flag seen, counted, published]
#define KMT(slot)                                                                          \
  do {                                                                                     \
    if (blockIdx.x == 0 && threadIdx.x == 0 && it < 16) g_kmt_trace[it * 16 + (slot)] = clock64(); \
  } while (0)
#define KMT_FIN(slot)                                                                      \
  do {                                                                                     \
    if (threadIdx.x == 0 && it < 16) g_kmt_fin[it * 16 + (slot)] = clock64();              \
  } while (0)
#define KMT_CTA(slot)                                                                      \
  do {                                                                                     \
    if (threadIdx.x == 0 && blockIdx.x < 160 && it < 12)                                   \
      g_kmt_cta[(blockIdx.x * 12 + it) * 4 + (slot)] = clock64();                          \
  } while (0)
#else
#define KMT(slot) do { } while (0)
#define KMT_FIN(slot) do { } while (0)
#define KMT_CTA(slot) do { } while (0)
#endif

struct KmeansTcArgs {
  KmeansArgs k;
  __nv_bfloat16* ph;   // [iterations][batch][K][dp] prototypes, bf16 hi part
  __nv_bfloat16* pl;   //                                      bf16 lo part
  int nkb;             // 64-wide K blocks (dp = 64 nkb)
  int ksteps;          // 16-wide K steps that hold data
  int stages;          // prototype ring depth (1 or 2)
  int box_rows;        // rows of a TMA box: min(128, K rounded up to 8)
  int prefetch;        // a second fp32 tile buffer fits: the next tile is fetched with cp.async
  float tau;
};

__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
// __threadfence() compiles to MEMBAR.SC.GPU + CCTL.IVALL here; acquire / release is enough
__device__ __forceinline__ void fence_acq_rel_gpu() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------- M-step pieces

// s_order[i] = i-th row of the tile in (label, row) order.  All-pairs ranking: 128 x 128
// compares over 256 threads.
__device__ __forceinline__ void rank_rows_by_label(const int* s_lab, int rows, int* s_rank,
                                                   unsigned char* s_order) {
  const int tid = threadIdx.x;
  if (tid < BM) s_rank[tid] = 0;
  __syncthreads();
  const int r = tid & (BM - 1), j0 = (tid >> 7) * (BM / 2);
  const int mine = r < rows ? s_lab[r] : 0x7fffffff;
  int cnt = 0;
#pragma unroll 8
  for (int j = j0; j < j0 + BM / 2; ++j) {
    const int lj = j < rows ? s_lab[j] : 0x7fffffff;
    cnt += (lj < mine) || (lj == mine && j < r);
  }
  atomicAdd(&s_rank[r], cnt);
  __syncthreads();
  if (tid < BM) s_order[s_rank[tid]] = (unsigned char)tid;
  __syncthreads();
}

// Each warp sums 16 consecutive entries of the ranked order, lanes across the channels, and
// adds a run to the image's sums whenever the label changes: about (labels in the tile + 8)
// flushes per tile.  A value is round(x * 2^32) kept as hi * 2^16 + lo in two exact 32-bit
// integers (at most 16 rows per run: neither half can overflow).
template <int kSlots>
__device__ __forceinline__ void accumulate_ranked(const KmeansArgs& p, int rows, const float* xs,
                                                  const int* s_lab, const unsigned char* s_order,
                                                  long long* __restrict__ sums_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dim = p.dim;
  constexpr int kPer = BM / kKmWarps;   // 16
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
      if (d < dim && (run_hi[s] | run_lo[s]) != 0)
        atomic_add_i64(&sums_b[(int64_t)run_lab * dim + d],
                       (long long)run_hi[s] * 65536ll + run_lo[s]);
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
      const bool ok = e < rows;
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
  if (bad) *p.poison = 1;
}

// The image's last tile is in: unit prototypes from the sums (common.py:39:
// sum / max(||sum||, eps); an empty cluster is the zero vector).  One warp per prototype,
// lanes across the channels; the loads of five prototypes x R replicas are all in flight
// together, so the K x dim words of the shipped configurations cost ONE round trip to L2.
// Writes the fp32 rows (exact re-check) and the bf16 hi / lo rows (TMA-fed operand; their
// padding columns [dim, dp) were zeroed by the host).
template <int kSlots, int R>
__device__ __forceinline__ void finalize_image(const KmeansArgs& p,
                                               const long long* __restrict__ sums_it,
                                               size_t per_iter, int kb, float* __restrict__ out,
                                               __nv_bfloat16* __restrict__ out_hi,
                                               __nv_bfloat16* __restrict__ out_lo, int dp) {
  constexpr int kC = 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dim = p.dim;
  for (int kbase = warp; kbase < kb; kbase += kC * kKmWarps) {
    long long raw[kC][kSlots][R];
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int k = kbase + i * kKmWarps;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int d = lane + 32 * s;
#pragma unroll
        for (int r = 0; r < R; ++r)
          raw[i][s][r] = (k < kb && d < dim)
                             ? __ldcg(sums_it + (size_t)r * per_iter + (size_t)k * dim + d) : 0;
      }
    }
    float v[kC][kSlots], ss[kC];
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      ss[i] = 0.f;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        long long t = raw[i][s][0];
#pragma unroll
        for (int r = 1; r < R; ++r) t += raw[i][s][r];
        v[i][s] = fixed_to_float(t);
        ss[i] += v[i][s] * v[i][s];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < kC; ++i) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int k = kbase + i * kKmWarps;
      const float nrm = sqrtf(ss[i]);
      const float div = nrm >= p.eps ? nrm : p.eps;
      if (k < kb) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const int d = lane + 32 * s;
          if (d < dim) {
            const float u = v[i][s] / div;
            out[(int64_t)k * dim + d] = u;
            const __nv_bfloat16 h = __float2bfloat16_rn(u);
            out_hi[(int64_t)k * dp + d] = h;
            out_lo[(int64_t)k * dp + d] = __float2bfloat16_rn(u - __bfloat162float(h));
          }
        }
      }
    }
  }
}

// the channel loops are unrolled over ceil(dim / 32) slots per lane
#define KM_SLOT_SWITCH(dim, CALL)            \
  switch (((dim) + 31) >> 5) {               \
    case 1: { constexpr int kS = 1; CALL; } break; \
    case 2: { constexpr int kS = 2; CALL; } break; \
    case 3: { constexpr int kS = 3; CALL; } break; \
    case 4: { constexpr int kS = 4; CALL; } break; \
    default: { constexpr int kS = 5; CALL; } break; \
  }

// live tile `lt` (image-major order) -> image, first row, row count
__device__ __forceinline__ void locate_tile(const KmeansArgs& p, int lt, Tile& tile) {
  int img = 0, base = 0;
  for (;;) {
    const int64_t first = p.img_off ? (int64_t)p.img_off[img] : 0;
    const int64_t last = p.img_off ? (int64_t)p.img_off[img + 1] : p.rows_total;
    const int tiles_b = (int)((last - first + BM - 1) / BM);
    if (lt < base + tiles_b) {
      tile.b = img;
      tile.row0 = first + (int64_t)(lt - base) * BM;
      tile.rows = (int)min((int64_t)BM, last - tile.row0);
      return;
    }
    base += tiles_b;
    ++img;
  }
}

__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Starts the copy of a tile's rows * dim floats into `buf` (kept at the source's 16-byte
// phase so that the body moves in 16-byte pieces); returns that phase in floats.
__device__ __forceinline__ int prefetch_tile(const KmeansArgs& p, const Tile& tile, float* buf) {
  const int tid = threadIdx.x;
  const float* src = p.x + tile.row0 * p.dim;
  const int total = tile.rows * p.dim;
  const int lead = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
  float* dst = buf + lead;
  const int head = min(total, (4 - lead) & 3);
  const int body = (total - head) >> 2;
  if (tid < head) cp_async_4(dst + tid, src + tid);
  for (int i = tid; i < body; i += kGemmThreads) cp_async_16(dst + head + 4 * i, src + head + 4 * i);
  const int done4 = head + 4 * body;
  if (tid < total - done4) cp_async_4(dst + done4 + tid, src + done4 + tid);
  cp_async_commit();
  return lead;
}

// ------------------------------------------------------------------------- the kernel

__global__ void __launch_bounds__(kGemmThreads, 1)
kmeans_tc_kernel(const __grid_constant__ CUtensorMap map_ph,
                 const __grid_constant__ CUtensorMap map_pl, const KmeansTcArgs a) {
  extern __shared__ uint8_t km_smem_raw[];
  __shared__ __align__(8) uint64_t bar_b_full[2], bar_t_full[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_lab[BM];
  __shared__ float s_b1[BM], s_b2[BM];
  __shared__ int s_k1[BM];
  __shared__ int s_amb[BM];
  __shared__ int s_rank[BM];
  __shared__ unsigned char s_order[BM];
  __shared__ int s_namb, s_last;

  const KmeansArgs& p = a.k;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int dim = p.dim;
  // Every CTA takes a CONTIGUOUS range of the live tiles (image-major order): its tiles then
  // belong to one image (two at a boundary), so it waits for ONE image's prototypes per pass.
  // With round-robin tiles every CTA visited every image and the finalising CTA of each image
  // sat on the critical path of all of them, once per image and pass.
  int live_total = 0;
  for (int bb = 0; bb < p.batch; ++bb) {
    const int64_t rows_b = p.img_off ? (int64_t)p.img_off[bb + 1] - p.img_off[bb] : p.rows_total;
    live_total += (int)((rows_b + BM - 1) / BM);
  }
  const int lt0 = (int)((int64_t)live_total * blockIdx.x / gridDim.x);
  const int lt1 = (int)((int64_t)live_total * (blockIdx.x + 1) / gridDim.x);
  const bool resident = live_total <= (int)gridDim.x;       // at most one tile per CTA: load it once
  const size_t per_img = (size_t)p.num_clusters * dim;
  const size_t per_iter = (size_t)p.batch * per_img;
  const int dp = a.nkb * 64;

  // 1024-byte aligned carve-up by OFFSET, so that the pointers stay shared-space pointers
  uint8_t* smem = km_smem_raw + ((1024u - (tc::smem_u32(km_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;                                        // [nkb][128 x 128 B]
  uint8_t* a_lo = a_hi + (size_t)a.nkb * kKmBlockBytes;
  uint8_t* b_ring = a_lo + (size_t)a.nkb * kKmBlockBytes;      // [stages][nkb][hi | lo]
  const uint32_t stage_bytes = 2u * a.nkb * kKmBlockBytes;
  float* xf = reinterpret_cast<float*>(b_ring + (size_t)a.stages * stage_bytes);
  const size_t xf_stride = (size_t)BM * dim + 4;          // a second buffer only with a.prefetch
  const float* xs = xf;   // xs[r * dim + d]: the fp32 tile, at the 16-byte phase of its source

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_b_full[s], 1);
      tc::mbar_init(&bar_t_full[s], 1);
    }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&map_ph);
    tc::prefetch_tensormap(&map_pl);
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, 4 * kKmBN);   // two accumulators [hh + lh | hl]
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, kKmBN, 0, 0);
  constexpr uint32_t idesc2x = tc::umma_idesc_bf16(BM, 2 * kKmBN, 0, 0);
  const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
  const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
  const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
  const uint32_t ring_lo = tc::umma_desc_lo(tc::smem_u32(b_ring), 16);

  uint32_t q = 0;   // column tiles this CTA has pushed through the ring / accumulators so far
  bool tma_ahead = false;   // the next tile's first prototype tile is already in flight
  int pending = 0;          // tiles of the current image accumulated but not yet counted
  const bool pf = a.prefetch && !resident && lt0 < lt1;
  int cur = 0, lead_cur = 0, lead_nxt = 0;
  if (pf) {
    Tile first_tile;
    locate_tile(p, lt0, first_tile);
    lead_cur = prefetch_tile(p, first_tile, xf);
  }

  for (int it = 0; it <= p.iterations; ++it) {
    for (int lt = lt0; lt < lt1; ++lt) {
      Tile tile;
      locate_tile(p, lt, tile);
      const int b = tile.b;
      const int kb = p.k_per_image ? p.k_per_image[b] : p.num_clusters;
      KMT(0);
      KMT_CTA(0);

      if (!resident || it == 0) {
        float* dst;
        if (pf) {
          // ---- the tile was requested one tile ago (cp.async); ask for the next one now
          cp_async_wait_all();
          __syncthreads();   // landed for everybody; nobody still reads the previous tile
          dst = xf + cur * xf_stride + lead_cur;
          const int nlt = lt + 1 < lt1 ? lt + 1 : (it < p.iterations ? lt0 : -1);
          if (nlt >= 0) {
            Tile next;
            locate_tile(p, nlt, next);
            lead_nxt = prefetch_tile(p, next, xf + (cur ^ 1) * xf_stride);
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
          const uint32_t off = (uint32_t)(c >> 3) * kKmBlockBytes + (uint32_t)(row >> 3) * 1024 +
                               sw * 128 + ((((uint32_t)c & 7) ^ sw) << 4);
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
      }
      KMT(1);

      if (it == 0) {
        if (tid < tile.rows) s_lab[tid] = p.labels_in[tile.row0 + tid];
        __syncthreads();
      } else {
        // ================================================================== E-step
        if (tid == 0) {
          // (with the prototype tile already requested, this image's flag was seen last tile)
          const unsigned* flag = p.ready + (size_t)(it - 1) * p.batch + b;
          while (!tma_ahead && ld_acquire_gpu(flag) == 0) {
          }
          s_namb = 0;
        }
        __syncthreads();
        KMT(2);
        KMT_CTA(1);
        const int ntiles = (kb + kKmBN - 1) / kKmBN;
        const int32_t prow0 = (int32_t)(((size_t)(it - 1) * p.batch + b) * p.num_clusters);

        auto issue_tma = [&](int j, uint32_t qq) {   // warp 0, converged
          const uint32_t s = qq % a.stages;
          if (tc::elect_one()) {
            // a box holds only the rows that exist (K <= 128: fewer bytes in flight); the rows
            // of the slot it leaves alone feed accumulator columns that nobody reads
            tc::mbar_expect_tx(&bar_b_full[s], 2u * a.nkb * a.box_rows * 128u);
            uint8_t* bh = b_ring + (size_t)s * stage_bytes;
            for (int kblk = 0; kblk < a.nkb; ++kblk) {
              // per K block the hi tile is followed by the lo tile: ONE 256-row B operand
              tc::tma_load_2d(&map_ph, &bar_b_full[s], bh + (size_t)(2 * kblk) * kKmBlockBytes,
                              kblk * 64, prow0 + j * kKmBN);
              tc::tma_load_2d(&map_pl, &bar_b_full[s], bh + (size_t)(2 * kblk + 1) * kKmBlockBytes,
                              kblk * 64, prow0 + j * kKmBN);
            }
          }
          __syncwarp();
        };
        auto issue_mma = [&](uint32_t qq) {          // warp 0, converged
          const uint32_t s = qq % a.stages, use = qq / a.stages, acc = qq & 1;
          tc::mbar_wait(&bar_b_full[s], use & 1);
          tc::tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 2 * kKmBN;
          const uint32_t b_lo = ring_lo + s * (stage_bytes >> 4);
          if (tc::elect_one()) {
            uint32_t accumulate = 0;
            for (int kblk = 0; kblk < a.nkb; ++kblk) {
              const int steps = min(4, a.ksteps - kblk * 4);
              uint32_t ah = ah_lo + kblk * (kKmBlockBytes >> 4);
              uint32_t al = al_lo + kblk * (kKmBlockBytes >> 4);
              uint32_t bp = b_lo + kblk * (2 * kKmBlockBytes >> 4);
              for (int ks = 0; ks < steps; ++ks) {   // 16 bf16 = 32 bytes inside the swizzle atom
                tc::umma_bf16_words(d_tmem, ah, hi_k, bp, hi_k, idesc2x, accumulate);   // hh | hl
                tc::umma_bf16_words(d_tmem, al, hi_k, bp, hi_k, idesc, 1);              // += lh
                accumulate = 1;
                ah += 2, al += 2, bp += 2;
              }
            }
            tc::umma_commit(&bar_t_full[acc]);
          }
          __syncwarp();
        };

        if (warp == 0 && ntiles > 0) {
          if (!tma_ahead) {
            fence_proxy_async_all();   // prototypes were written through the generic proxy
            issue_tma(0, q);
          }
          if (a.stages == 2 && ntiles > 1) issue_tma(1, q + 1);
          issue_mma(q);
        }
        tma_ahead = false;
        KMT(3);
        const int sp = warp & 3;            // TMEM sub-partition of this warp
        const int half = warp >> 2;         // which 32-column chunks of an accumulator
        const int row = sp * 32 + lane;
        float b1 = -INFINITY, b2 = -INFINITY;
        int k1 = 0;
        for (int j = 0; j < ntiles; ++j) {
          const uint32_t qq = q + j;
          const uint32_t acc = qq & 1, tpar = (qq >> 1) & 1;
          if (warp == 0 && j + 1 < ntiles) {
            if (a.stages == 1) {            // the single slot is free once MMA j has read it
              tc::mbar_wait(&bar_t_full[acc], tpar);
              issue_tma(j + 1, qq + 1);
            }
            issue_mma(qq + 1);              // accumulator (qq + 1) & 1 was drained before the last barrier
            if (a.stages == 2 && j + 2 < ntiles) {
              tc::mbar_wait(&bar_t_full[acc], tpar);
              issue_tma(j + 2, qq + 2);
            }
          }
          tc::mbar_wait(&bar_t_full[acc], tpar);
          tc::tcgen05_fence_after();
          if (j == 0) KMT(4);
          const int c0 = j * kKmBN;
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            // the halves take alternate 32-column chunks, so both work when K <= 64
            const int cb = (2 * chunk + half) * 32;
            if (c0 + cb < kb) {             // warp-uniform
              uint32_t v[32], w[32];
              const uint32_t taddr = tmem_base + acc * 2 * kKmBN + cb +
                                     (static_cast<uint32_t>(sp * 32) << 16);
              tc::tmem_ld_32x32(taddr, v);            // hi.hi + lo.hi
              tc::tmem_ld_32x32(taddr + kKmBN, w);    // hi.lo
              tc::tmem_ld_wait();
              const int live = kb - c0 - cb;          // columns of this chunk that exist
#pragma unroll
              for (int u = 0; u < 32; ++u) {
                float s = __uint_as_float(v[u]) + __uint_as_float(w[u]);
                s = u < live ? s : -INFINITY;
                b2 = fmaxf(b2, fminf(s, b1));         // second best so far (a tie counts)
                k1 = s > b1 ? c0 + cb + u : k1;
                b1 = fmaxf(b1, s);
              }
            }
          }
          tc::tcgen05_fence_before();
          __syncthreads();
        }
        q += ntiles;
        KMT(5);

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
        KMT(6);
        const float* __restrict__ protos_b = p.protos + (size_t)(it - 1) * per_iter + b * per_img;
        if (namb > 0) {   // block-uniform
          // The MMAs of this tile are complete, so the prototype ring is idle: stage the
          // image's fp32 prototypes in it (one coalesced pass) instead of chasing them through
          // L2 element by element (measured: tiles with near-ties took 2x as long).
          const bool in_smem = (size_t)kb * dim * sizeof(float) <= (size_t)a.stages * stage_bytes;
          float* pf = reinterpret_cast<float*>(b_ring);
          if (in_smem) {
            const int n = kb * dim;
            for (int i0 = 0; i0 < n; i0 += 8 * kGemmThreads) {
              float r8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int i = i0 + j * kGemmThreads + tid;
                r8[j] = i < n ? __ldcg(protos_b + i) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int i = i0 + j * kGemmThreads + tid;
                if (i < n) pf[i] = r8[j];
              }
            }
            __syncthreads();
          }
          for (int i = warp; i < namb; i += kKmWarps) {
            const int r = s_amb[i];
            const float* xr = xs + r * dim;
            float bv = -INFINITY;
            int bk = 0;
            for (int k = lane; k < kb; k += 32) {
              float accv = 0.f;
              if (in_smem) {
                const float* pk = pf + k * dim;
#pragma unroll 4
                for (int d = 0; d < dim; ++d) accv = fmaf(xr[d], pk[d], accv);
              } else {
                const float* pk = protos_b + (size_t)k * dim;
                int d = 0;
                for (; d + 8 <= dim; d += 8) {   // eight loads in flight, the fmaf chain in order
                  float pv[8];
#pragma unroll
                  for (int u = 0; u < 8; ++u) pv[u] = __ldcg(pk + d + u);
#pragma unroll
                  for (int u = 0; u < 8; ++u) accv = fmaf(xr[d + u], pv[u], accv);
                }
                for (; d < dim; ++d) accv = fmaf(xr[d], __ldcg(pk + d), accv);
              }
              if (accv > bv) bv = accv, bk = k;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
              if (ov > bv || (ov == bv && ok < bk)) bv = ov, bk = ok;
            }
            if (lane == 0) s_lab[r] = bk;
          }
        }
        __syncthreads();
        KMT(7);
        // The next tile of this CTA is usually in the same image and pass: it needs the same
        // prototypes, and the ring slot is free (this tile's MMAs are complete), so its
        // first prototype tile is requested now and lands during the M-step.
        if (!resident && lt + 1 < lt1 && ntiles > 0) {
          Tile next;
          locate_tile(p, lt + 1, next);
          if (next.b == b) {
            tma_ahead = true;
            if (warp == 0) {
              fence_proxy_async_all();   // the exact re-check may have used the ring as scratch
              issue_tma(0, q);
            }
          }
        }
#ifdef SPML_KM_TRACE
        if (blockIdx.x == 0 && tid == 0 && it < 16) g_kmt_trace[it * 16 + 15] = namb;
#endif
        if (it == p.iterations && tid < tile.rows) {
          if (p.labels_out) p.labels_out[tile.row0 + tid] = s_lab[tid];
          if (p.labels_out64) p.labels_out64[tile.row0 + tid] = s_lab[tid];
        }
      }

      if (it < p.iterations) {
        // ================================================================== M-step
        long long* sums_it = p.sums + (size_t)it * p.replicas * per_iter + b * per_img;
        long long* sums_b = sums_it + (size_t)(blockIdx.x % p.replicas) * per_iter;
        rank_rows_by_label(s_lab, tile.rows, s_rank, s_order);
        KMT(11);
        KM_SLOT_SWITCH(dim, (accumulate_ranked<kS>(p, tile.rows, xs, s_lab, s_order, sums_b)));
        __syncthreads();
        KMT(8);
        // The CTA's consecutive tiles of one image are counted together: one fence (which has to
        // wait for the reductions to be performed) and one counter round trip per image and
        // pass instead of one per tile.
        ++pending;
        bool flush_count = lt + 1 >= lt1;
        if (!flush_count) {
          Tile next;
          locate_tile(p, lt + 1, next);
          flush_count = next.b != b;
        }
        if (flush_count) {
          if (tid == 0) {
            const int64_t rows_b = (int64_t)(p.img_off ? p.img_off[b + 1] - p.img_off[b]
                                                       : p.rows_total);
            const unsigned tiles_b = (unsigned)((rows_b + BM - 1) / BM);
            fence_acq_rel_gpu();   // cumulative: the CTA's reductions (ordered by the barriers) first
            s_last = atomicAdd(p.done + (size_t)it * p.batch + b, (unsigned)pending) + pending == tiles_b;
            fence_acq_rel_gpu();
          }
          pending = 0;
          __syncthreads();
        }
        KMT(9);
        KMT_CTA(2);
        if (flush_count && s_last) {
          // ---- the image's last tile is in: fold the replicas, normalise, publish
          KMT_FIN(10);
          KMT_FIN(11);
          const size_t slab = ((size_t)it * p.batch + b) * p.num_clusters * dp;
          float* out = p.protos + (size_t)it * per_iter + b * per_img;
          KM_SLOT_SWITCH(dim, (finalize_image<kS, kKmReplicas>(p, sums_it, per_iter, kb, out,
                                                               a.ph + slab, a.pl + slab, dp)));
          KMT_FIN(13);
          fence_proxy_async_all();   // the consumers read the bf16 rows through TMA
          __syncthreads();
          KMT_FIN(14);
          if (tid == 0) {
            fence_acq_rel_gpu();     // cumulative over the CTA's stores (ordered by the barrier)
            atomicExch(p.ready + (size_t)it * p.batch + b, 1u);
          }
          KMT_CTA(3);
          KMT_FIN(15);
        }
        KMT(10);
      }
      if (pf) cur ^= 1, lead_cur = lead_nxt;
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 4 * kKmBN);
  }
}

// ------------------------------------------------------------------------- host side

// K padded to 64-wide blocks; the fp32 tile must fit next to the operand tiles
static bool kmeans_tc_geometry(int dim, int* nkb, int* stages, size_t* smem) {
  const int blocks = (dim + 63) / 64;
  if (dim < 1 || blocks > 2) return false;
  const int st = blocks == 1 ? 2 : 1;
  const size_t bytes = 1024 + (size_t)2 * blocks * kKmBlockBytes +
                       (size_t)st * 2 * blocks * kKmBlockBytes +
                       ((size_t)BM * dim + 4) * sizeof(float);
  if (bytes > 220 * 1024) return false;
  *nkb = blocks, *stages = st, *smem = bytes;
  return true;
}

bool kmeans_tc_supported(int dim) {
  int nkb, stages;
  size_t smem;
  return kmeans_tc_geometry(dim, &nkb, &stages, &smem);
}

// bf16 hi + lo prototype rows of every iteration and image
size_t kmeans_tc_split_bytes(int batch, int num_clusters, int dim, int iterations) {
  int nkb, stages;
  size_t smem;
  if (!kmeans_tc_geometry(dim, &nkb, &stages, &smem)) return 0;
  return 2 * align_up((size_t)iterations * batch * num_clusters * nkb * 64 * 2, 256);
}

int kmeans_tc_launch(const KmeansArgs& p, void* split_protos, int sms, cudaStream_t st) {
  KmeansTcArgs a{};
  size_t smem = 0;
  if (!kmeans_tc_geometry(p.dim, &a.nkb, &a.stages, &smem)) {
    set_error("kmeans(tc): dim %d is not supported", p.dim);
    return SPML_E_UNSUPPORTED;
  }
  a.k = p;
  a.ksteps = (p.dim + 15) / 16;
  a.tau = 1e-4f;
  a.box_rows = std::min(kKmBN, (p.num_clusters + 7) & ~7);
  const size_t second = ((size_t)BM * p.dim + 4) * sizeof(float);
  if (smem + second <= 220 * 1024) {
    a.prefetch = 1;
    smem += second;
  }
  const size_t split_rows = (size_t)p.iterations * p.batch * p.num_clusters;
  const size_t split_bytes = align_up(split_rows * a.nkb * 64 * 2, 256);
  char* split = reinterpret_cast<char*>(split_protos);
  SPML_CHECK_ARG((reinterpret_cast<uintptr_t>(split) & 15) == 0,
                 "kmeans: workspace must be 16-byte aligned");
  SPML_CHECK_SUPPORTED(split_rows < (1ull << 31), "kmeans: too many prototype rows");
  a.ph = reinterpret_cast<__nv_bfloat16*>(split);
  a.pl = reinterpret_cast<__nv_bfloat16*>(split + split_bytes);
  // K padding of the bf16 rows (columns [dim, 64 nkb)): zero x garbage could be NaN
  if (p.dim < a.nkb * 64) SPML_CUDA(cudaMemsetAsync(split, 0, 2 * split_bytes, st));
  CUtensorMap map_ph, map_pl;
  const uint64_t pitch = (uint64_t)a.nkb * 64 * 2;
  int rc;
  if ((rc = make_tensor_map_bf16_2d(&map_ph, a.ph, (uint64_t)a.nkb * 64, split_rows, pitch, 64,
                                    a.box_rows)))
    return rc;
  if ((rc = make_tensor_map_bf16_2d(&map_pl, a.pl, (uint64_t)a.nkb * 64, split_rows, pitch, 64,
                                    a.box_rows)))
    return rc;
  SPML_CUDA(cudaFuncSetAttribute(kmeans_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  int per_sm = 0;
  SPML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmeans_tc_kernel, kGemmThreads,
                                                          smem));
  SPML_CHECK_SUPPORTED(per_sm >= 1, "kmeans(tc): kernel does not fit on an SM (dim %d)", p.dim);
  // one CTA per SM: each allocates all 512 TMEM columns
  const int64_t tiles = (int64_t)p.batch * p.tiles_per_img;
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sms);
  void* args[] = {&map_ph, &map_pl, &a};
  SPML_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kmeans_tc_kernel), dim3(grid),
                                        dim3(kGemmThreads), args, smem, st));
  SPML_LAUNCH_CHECK("kmeans_tc_kernel");
  return SPML_OK;
}

}  // namespace spml

#ifdef SPML_KM_TRACE
extern "C" int spml_debug_kmtc_trace(long long* trace, long long* fin, long long* cta) {
  if (cudaMemcpyFromSymbol(trace, spml::g_kmt_trace, sizeof(long long) * 16 * 16) != cudaSuccess)
    return -2;
  if (cudaMemcpyFromSymbol(fin, spml::g_kmt_fin, sizeof(long long) * 16 * 16) != cudaSuccess)
    return -2;
  if (cudaMemcpyFromSymbol(cta, spml::g_kmt_cta, sizeof(long long) * 160 * 12 * 4) != cudaSuccess)
    return -2;
  return 0;
}
#endif
