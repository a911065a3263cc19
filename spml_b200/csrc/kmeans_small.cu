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
//   * the M-step accumulates a tile with native 32-bit shared-memory atomics on the two
//     halves of the exact fixed-point value (hi * 2^16 + lo = round(x 2^32)), then adds only
//     the non-zero entries to the global sums with 64-bit reductions: no ranking pass, and
//     several consecutive tiles of a CTA share one flush.
// Sums are integers, so labels are bit-identical to the fp32 kernel and to kmeans_tc.cu
// whatever the tiling (tests/test_gpu_ops.py::test_kmeans_tensor_core_equals_fp32).
#include <math.h>

#include <algorithm>

#include "kmeans.cuh"
#include "tc_common.cuh"

namespace spml {

constexpr int kSmBN = 128;                 // prototype rows of the operand tile
constexpr int kSmBlockBytes = 128 * 128;   // one 64-wide K block of a 128-row bf16 tile
constexpr int kSmWarps = kGemmThreads / 32;
constexpr int kSmMaxTilesPerFlush = 24;    // 2^19 * 128 * 24 < 2^31: the hi halves cannot overflow

#ifdef SPML_KM_TRACE
__device__ long long g_kms_trace[16 * 16];
#define KMS(slot)                                                                          \
  do {                                                                                     \
    if (blockIdx.x == 0 && threadIdx.x == 0 && it < 16) g_kms_trace[it * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define KMS(slot) do { } while (0)
#endif

struct KmeansSmallArgs {
  KmeansArgs k;
  int nkb;        // 64-wide K blocks
  int ksteps;     // 16-wide K steps that hold data
  int prefetch;   // a second fp32 tile buffer fits
  float tau;
};

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

// byte offset of element (row, d) inside a [nkb][hi | lo] SWIZZLE_128B operand: what a TMA
// box {64 bf16, 128 rows} would write (tc_common.cuh)
__device__ __forceinline__ uint32_t operand_offset(int row, int d) {
  const uint32_t sw = row & 7;
  return (uint32_t)(d >> 6) * (2 * kSmBlockBytes) + (uint32_t)(row >> 3) * 1024 + sw * 128 +
         (((((uint32_t)d & 63) >> 3) ^ sw) << 4) + ((uint32_t)d & 7) * 2;
}

// live tile `lt` (image-major order) -> image, first row, row count
__device__ __forceinline__ void kms_locate_tile(const KmeansArgs& p, int lt, Tile& tile) {
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

// The image's unit prototypes from its fixed-point sums (common.py:39: sum / max(||sum||, eps);
// an empty cluster is the zero vector), written to THIS CTA's shared memory: fp32 rows for the
// exact re-check and the bf16 hi / lo operand tile.  One warp per prototype, lanes across the
// channels, the loads of kC prototypes x R replicas in flight together.
template <int kSlots, int R>
__device__ __forceinline__ void build_prototypes(const KmeansArgs& p,
                                                 const long long* __restrict__ sums_it,
                                                 size_t per_iter, int kb, float* __restrict__ pf,
                                                 uint8_t* __restrict__ b_tile) {
  constexpr int kC = 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dim = p.dim;
  for (int kbase = warp; kbase < kb; kbase += kC * kSmWarps) {
    long long raw[kC][kSlots][R];
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      const int k = kbase + i * kSmWarps;
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
      const int k = kbase + i * kSmWarps;
      const float nrm = sqrtf(ss[i]);
      const float div = nrm >= p.eps ? nrm : p.eps;
      if (k < kb) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const int d = lane + 32 * s;
          if (d < dim) {
            const float u = v[i][s] / div;
            pf[k * dim + d] = u;
            const __nv_bfloat16 h = __float2bfloat16_rn(u);
            const uint32_t off = operand_offset(k, d);
            *reinterpret_cast<__nv_bfloat16*>(b_tile + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(b_tile + off + kSmBlockBytes) =
                __float2bfloat16_rn(u - __bfloat162float(h));
          }
        }
      }
    }
  }
}

// M-step of one tile into the CTA's shared sums: the two exact halves of round(x 2^32) with
// native 32-bit shared atomics (a 64-bit shared atomicAdd compiles to a CAS loop).  Warp w
// takes rows 16 w .. 16 w + 15, lanes across the channels.
template <int kSlots>
__device__ __forceinline__ bool accumulate_tile(int rows, int dim, int num_clusters,
                                                const float* __restrict__ xs,
                                                const int* __restrict__ s_lab, int* s_hi,
                                                int* s_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kPer = BM / kSmWarps;
  bool bad = false;
#pragma unroll 4
  for (int i = 0; i < kPer; ++i) {
    const int r = warp * kPer + i;
    if (r >= rows) break;
    const int lab = s_lab[r];
    if ((unsigned)lab >= (unsigned)num_clusters) {   // a label the caller never declared
      bad = true;
      continue;
    }
    const int base = lab * dim;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int d = lane + 32 * s;
      if (d < dim) {
        const float v = xs[r * dim + d];
        bad |= !(fabsf(v) <= 8.f);
        int hi, lo;
        split_fixed(v, hi, lo);
        atomicAdd(&s_hi[base + d], hi);
        atomicAdd(&s_lo[base + d], lo);
      }
    }
  }
  return bad;
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
  __shared__ __align__(8) uint64_t bar_t_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_lab[BM];
  __shared__ float s_b1[BM], s_b2[BM];
  __shared__ int s_k1[BM];
  __shared__ int s_amb[BM];
  __shared__ int s_namb;

  const KmeansArgs& p = a.k;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int dim = p.dim;
  const int K = p.num_clusters;
  // every CTA takes a CONTIGUOUS range of the live tiles (image-major order)
  int live_total = 0;
  for (int bb = 0; bb < p.batch; ++bb) {
    const int64_t rows_b = p.img_off ? (int64_t)p.img_off[bb + 1] - p.img_off[bb] : p.rows_total;
    live_total += (int)((rows_b + BM - 1) / BM);
  }
  const int lt0 = (int)((int64_t)live_total * blockIdx.x / gridDim.x);
  const int lt1 = (int)((int64_t)live_total * (blockIdx.x + 1) / gridDim.x);
  const bool resident = live_total <= (int)gridDim.x;       // at most one tile per CTA: load it once
  const size_t per_img = (size_t)K * dim;
  const size_t per_iter = (size_t)p.batch * per_img;

  // 1024-byte aligned carve-up by OFFSET, so that the pointers stay shared-space pointers
  uint8_t* smem = kms_smem_raw + ((1024u - (tc::smem_u32(kms_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_hi = smem;                                        // [nkb][128 x 128 B]
  uint8_t* a_lo = a_hi + (size_t)a.nkb * kSmBlockBytes;
  uint8_t* b_tile = a_lo + (size_t)a.nkb * kSmBlockBytes;      // [nkb][hi | lo]
  float* pf = reinterpret_cast<float*>(b_tile + (size_t)a.nkb * 2 * kSmBlockBytes);   // [K][dim]
  int* s_hi = reinterpret_cast<int*>(pf + per_img);            // [K][dim]
  int* s_lo = s_hi + per_img;
  float* xf = reinterpret_cast<float*>(
      reinterpret_cast<uint8_t*>(s_lo + per_img) +
      ((16u - (tc::smem_u32(s_lo + per_img) & 15u)) & 15u));   // 16-byte aligned (128-bit copies)
  const size_t xf_stride = (size_t)BM * dim + 4;               // a second buffer only with a.prefetch
  const float* xs = xf;   // xs[r * dim + d]: the fp32 tile, at the 16-byte phase of its source

  if (tid == 0) {
    tc::mbar_init(&bar_t_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, 2 * kSmBN);   // one accumulator [hh + lh | hl]
  // operand tile: rows >= K and the K padding [dim, 64 nkb) must be zeros (0 x garbage = NaN)
  for (int i = tid; i < a.nkb * 2 * kSmBlockBytes / 16; i += kGemmThreads)
    reinterpret_cast<uint4*>(b_tile)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (int)per_img; i += kGemmThreads) s_hi[i] = 0, s_lo[i] = 0;
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, kSmBN, 0, 0);
  constexpr uint32_t idesc2x = tc::umma_idesc_bf16(BM, 2 * kSmBN, 0, 0);
  const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
  const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
  const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
  const uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(b_tile), 16);

  uint32_t q = 0;   // tiles this CTA has pushed through the accumulator so far (barrier parity)
  const bool pf_on = a.prefetch && !resident && lt0 < lt1;
  int cur = 0, lead_cur = 0, lead_nxt = 0;
  if (pf_on) {
    Tile first_tile;
    kms_locate_tile(p, lt0, first_tile);
    lead_cur = kms_prefetch_tile(p, first_tile, xf);
  }

  for (int it = 0; it <= p.iterations; ++it) {
    int proto_img = -1;    // image whose prototypes of pass it - 1 sit in shared memory
    int pending = 0;       // tiles of the current image accumulated in s_hi / s_lo, not flushed
    bool bad = false;
    for (int lt = lt0; lt < lt1; ++lt) {
      Tile tile;
      kms_locate_tile(p, lt, tile);
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
            kms_locate_tile(p, nlt, next);
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
        __syncthreads();
        if (tid < tile.rows) s_lab[tid] = p.labels_in[tile.row0 + tid];
        __syncthreads();
      } else {
        // ================================================================== E-step
        if (proto_img != b) {
          // ---- pass it - 1 of this image is complete once every one of its tiles is counted
          if (tid == 0) {
            const int64_t rows_b = (int64_t)(p.img_off ? p.img_off[b + 1] - p.img_off[b]
                                                       : p.rows_total);
            const unsigned tiles_b = (unsigned)((rows_b + BM - 1) / BM);
            const unsigned* done = p.done + (size_t)(it - 1) * p.batch + b;
            unsigned spins = 0;
            while (kms_ld_acquire_gpu(done) < tiles_b) {
              if (++spins > tc::kSpinLimit) {   // trap instead of hanging the GPU
                printf("spml_b200: k-means pass %d of image %d never completed (block %d)\n",
                       it - 1, b, blockIdx.x);
                __trap();
              }
            }
            kms_fence_acq_rel_gpu();
          }
          __syncthreads();
          KMS(2);
          const long long* sums_prev = p.sums + (size_t)(it - 1) * p.replicas * per_iter + b * per_img;
          KMS_SLOT_SWITCH(dim, (build_prototypes<kS, kKmReplicas>(p, sums_prev, per_iter, kb, pf,
                                                                  b_tile)));
          proto_img = b;
        }
        if (tid == 0) s_namb = 0;
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
              uint32_t bp = b_lo + kblk * (2 * kSmBlockBytes >> 4);
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
            tc::tmem_ld_32x32(taddr + kSmBN, w);    // hi.lo
            tc::tmem_ld_wait();
            const int live = kb - cb;               // columns of this chunk that exist
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              float s = __uint_as_float(v[u]) + __uint_as_float(w[u]);
              s = u < live ? s : -INFINITY;
              b2 = fmaxf(b2, fminf(s, b1));         // second best so far (a tie counts)
              k1 = s > b1 ? cb + u : k1;
              b1 = fmaxf(b1, s);
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
        // exact re-check: the very fmaf chain of the fp32 kernel (d ascending from 0), first
        // index on ties, against the fp32 prototypes in shared memory
        for (int i = warp; i < namb; i += kSmWarps) {
          const int r = s_amb[i];
          const float* xr = xs + r * dim;
          float bv = -INFINITY;
          int bk = 0;
          for (int k = lane; k < kb; k += 32) {
            float accv = 0.f;
            const float* pk = pf + k * dim;
#pragma unroll 4
            for (int d = 0; d < dim; ++d) accv = fmaf(xr[d], pk[d], accv);
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
        __syncthreads();
        KMS(6);
        if (it == p.iterations && tid < tile.rows) {
          if (p.labels_out) p.labels_out[tile.row0 + tid] = s_lab[tid];
          if (p.labels_out64) p.labels_out64[tile.row0 + tid] = s_lab[tid];
        }
      }

      if (it < p.iterations) {
        // ================================================================== M-step
        KMS_SLOT_SWITCH(dim, (bad |= accumulate_tile<kS>(tile.rows, dim, kb, xs, s_lab, s_hi, s_lo)));
        ++pending;
        bool flush = lt + 1 >= lt1 || pending >= kSmMaxTilesPerFlush;
        if (!flush) {
          Tile next;
          kms_locate_tile(p, lt + 1, next);
          flush = next.b != b;
        }
        __syncthreads();
        KMS(7);
        if (flush) {
          // ---- non-zero entries -> the image's global sums (64-bit reductions), zero for reuse
          long long* sums_b = p.sums + (size_t)it * p.replicas * per_iter + b * per_img +
                              (size_t)(blockIdx.x % p.replicas) * per_iter;
          for (int i = tid; i < (int)per_img; i += kGemmThreads) {
            const int hi = s_hi[i], lo = s_lo[i];
            if ((hi | lo) != 0) {
              atomic_add_i64(&sums_b[i], (long long)hi * 65536ll + lo);
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
    if (bad) *p.poison = 1;
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 2 * kSmBN);
  }
}

// ------------------------------------------------------------------------- host side

static bool kmeans_small_geometry(int dim, int num_clusters, int* nkb, int* prefetch, size_t* smem) {
  const int blocks = (dim + 63) / 64;
  if (dim < 1 || blocks > 2 || num_clusters < 1 || num_clusters > kSmBN) return false;
  const size_t tile = ((size_t)BM * dim + 4) * sizeof(float);
  const size_t fixed = 1024 + (size_t)4 * blocks * kSmBlockBytes +            // A hi/lo, B hi/lo
                       (size_t)num_clusters * dim * (sizeof(float) + 2 * sizeof(int)) + 64;
  if (fixed + tile > 220 * 1024) return false;
  *nkb = blocks;
  *prefetch = fixed + 2 * tile <= 220 * 1024;
  *smem = fixed + (*prefetch ? 2 : 1) * tile;
  return true;
}

bool kmeans_small_supported(int dim, int num_clusters) {
  int nkb, prefetch;
  size_t smem;
  return kmeans_small_geometry(dim, num_clusters, &nkb, &prefetch, &smem);
}

int kmeans_small_launch(const KmeansArgs& p, int sms, cudaStream_t st) {
  KmeansSmallArgs a{};
  size_t smem = 0;
  if (!kmeans_small_geometry(p.dim, p.num_clusters, &a.nkb, &a.prefetch, &smem)) {
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
extern "C" int spml_debug_kms_trace(long long* trace) {
  return cudaMemcpyFromSymbol(trace, spml::g_kms_trace, sizeof(long long) * 16 * 16) == cudaSuccess
             ? 0 : -2;
}
#endif
