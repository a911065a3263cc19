// C1 / C2 on the 5th-generation tensor cores: SegSort / SetSegSort forward as a
// TMA-fed tcgen05 GEMM with the exp / label-mask / row-sum epilogue read straight out
// of TMEM (reference spml/utils/segsort/loss.py:15-130; math in SURVEY.md 7.3).
//
// Precision: the fp32 unit vectors are split into bf16 hi + lo parts by a pre-pass and
// every 16-wide K step issues three MMAs (hi.hi + lo.hi + hi.lo, fp32 accumulation in
// TMEM), which keeps ~16 mantissa bits per product: cosines are exact to ~1e-5, inside
// the 1e-3 parity bar with two orders of margin, at 3/1 of the bf16 MMA cost.
//
// Kernel layout (one CTA per 128-row tile, 18 warps):
//   warp 0      TMA producer: A tile (all K blocks, hi + lo) once, then the prototype
//               bank streamed in 128-column tiles through a 4-stage shared-memory ring
//   warp 1      TMEM allocator and MMA issuer (elect.sync); accumulators double-buffered
//               in TMEM so the epilogue of tile j overlaps the MMAs of j+1
//   warps 2-17  epilogue: tcgen05.ld 32 lanes x 32 columns, ex2, code compare, running
//               same / diff sums per row; four warps share each TMEM sub-partition, one per
//               32-column quarter of the tile
#include <math.h>

#include <stdlib.h>

#include <algorithm>

#include "tc_common.cuh"
#include "segsort_tc.h"

#include <chrono>

namespace spml {

// ------------------------------------------------------------------------- host: tensor map

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                         &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                            uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SPML_E_CUDA;
  }
  const cuuint64_t gdim[2] = {inner, rows};
  const cuuint64_t gstride[1] = {row_pitch_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner %llu rows %llu pitch %llu", (int)r,
              (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)row_pitch_bytes);
    return SPML_E_CUDA;
  }
  return SPML_OK;
}

// ------------------------------------------------------------------------- pre-pass kernels

// compact list of the valid prototype columns (sem_ann's labelled-prototype filter):
// dst[c] = compact index or -1, src[k] = original column, *count = number kept.
__global__ void compact_cols_kernel(const uint8_t* __restrict__ valid, int m,
                                    int32_t* __restrict__ dst, int32_t* __restrict__ src,
                                    int32_t* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < m; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const bool keep = c < m && valid[c] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
      if (w < warp) before += s_warp[w];
      total += s_warp[w];
    }
    const int base = s_base;
    if (c < m) {
      const int k = keep ? base + before + __popc(ballot & ((1u << lane) - 1)) : -1;
      dst[c] = k;
      if (keep) src[k] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_base;
}

constexpr int kPadRows = 128;   // rows past the live count that are kept finite (zero)

__device__ __forceinline__ void split_store(float x, float y, __nv_bfloat16* hi,
                                            __nv_bfloat16* lo, int64_t at) {
  const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
  const __nv_bfloat16 xl = __float2bfloat16_rn(x - __bfloat162float(xh));
  const __nv_bfloat16 yl = __float2bfloat16_rn(y - __bfloat162float(yh));
  *reinterpret_cast<__nv_bfloat162*>(hi + at) = __halves2bfloat162(xh, yh);
  *reinterpret_cast<__nv_bfloat162*>(lo + at) = __halves2bfloat162(xl, yl);
}

// one warp per problem row: gather, scale, split into bf16 hi / lo (zero padded to dp
// columns), narrow the labels to int32 and translate the segment id to a compact column.
// `scale` = kappa * log2(e): the similarity GEMM then delivers the exponent of exp2
// directly (one multiply less per similarity in every epilogue); the kernel that
// re-uses the pixel rows as the second GEMM's operand folds 1 / scale into its weights.
__global__ void split_rows_kernel(const float* __restrict__ x, int64_t ld, int dim, int dp,
                                  float scale,
                                  const int32_t* __restrict__ row_index,
                                  const int32_t* __restrict__ group_off, int num_groups,
                                  int64_t n_rows, const int64_t* __restrict__ pix_code,
                                  const int64_t* __restrict__ seg,
                                  const int32_t* __restrict__ col_dst, int64_t m,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  int32_t* __restrict__ rcode, int32_t* __restrict__ rseg) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t live = group_off ? (int64_t)group_off[num_groups] : n_rows;
  if (r >= n_rows) return;
  if (r >= live) {
    // tiles that straddle the live count feed these rows to the K dimension of the second
    // GEMM, where even a zero weight cannot cancel a NaN: keep one tile of zeros
    if (r < live + kPadRows)
      for (int q = lane * 2; q < dp; q += 64) split_store(0.f, 0.f, hi, lo, r * dp + q);
    return;
  }
  const int64_t orig = row_index ? (int64_t)row_index[r] : r;
  const float* xr = x + orig * ld;
  for (int q = lane * 2; q < dp; q += 64) {
    const float a = q < dim ? xr[q] * scale : 0.f;
    const float b = q + 1 < dim ? xr[q + 1] * scale : 0.f;
    split_store(a, b, hi, lo, r * dp + q);
  }
  if (lane == 0) {
    rcode[r] = (int32_t)pix_code[orig];
    const int64_t s = seg[orig];
    int32_t col = -1;
    if (s >= 0 && s < m) col = col_dst ? col_dst[s] : (int32_t)s;
    rseg[r] = col;
  }
}

__global__ void split_protos_kernel(const float* __restrict__ p, int64_t ld, int dim, int dp,
                                    const int32_t* __restrict__ col_src,
                                    const int32_t* __restrict__ col_count, int64_t m,
                                    const int64_t* __restrict__ proto_code,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    int32_t* __restrict__ ccode) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t live = col_count ? (int64_t)*col_count : m;
  if (k >= m) return;
  if (k >= live) {
    if (k < live + kPadRows)
      for (int q = lane * 2; q < dp; q += 64) split_store(0.f, 0.f, hi, lo, k * dp + q);
    return;
  }
  const int64_t orig = col_src ? (int64_t)col_src[k] : k;
  const float* pr = p + orig * ld;
  for (int q = lane * 2; q < dp; q += 64) {
    const float a = q < dim ? pr[q] : 0.f;
    const float b = q + 1 < dim ? pr[q + 1] : 0.f;
    split_store(a, b, hi, lo, k * dp + q);
  }
  if (lane == 0) ccode[k] = (int32_t)proto_code[orig];
}

// ------------------------------------------------------------------------- forward kernel

constexpr int kTcBM = 128;          // rows per CTA = UMMA M
constexpr int kTcBN = 128;          // prototype columns per tile = UMMA N
// backward: producer + MMA + 16 epilogue warps (a warp owns 32 owner rows x 16 of the 64
// streamed entities of a tile)
// Warp 1 issues GEMM 1 (S), the LAST warp issues GEMM 2 (d(owner) += G . streamed): one
// issuing thread spent ~1300 cycles per tile on the 16 MMAs plus ~900 on its four barrier
// waits, more than the tensor pipe (~900) or the epilogue needed (SPML_TC_TRACE timeline).
constexpr int kBwdEpiWarps = 16;
constexpr int kBwdGemm2Warp = 2 + kBwdEpiWarps;
constexpr int kTcThreads = (3 + kBwdEpiWarps) * 32;
constexpr int kTcEpiThreads = kBwdEpiWarps * 32;
constexpr int kBwdEpiCols = 64 / (kBwdEpiWarps / 4);   // S columns per epilogue warp
// forward: producer + MMA + 16 epilogue warps.  The epilogue (exp, masks, row sums: ~10
// instructions per similarity, MUFU-co-bound) is what bounds the kernel, and with 8 warps
// there are 2 per scheduler: too few to cover the tcgen05.ld -> ex2 -> add chains.
constexpr int kFwdEpiWarps = 16;
constexpr int kFwdThreads = (2 + kFwdEpiWarps) * 32;
constexpr int kFwdEpiThreads = kFwdEpiWarps * 32;
constexpr int kKBlockBytesA = kTcBM * 128;  // one 64-wide K block of the A tile (bf16)
constexpr int kKBlockBytesB = kTcBN * 128;

struct TcFwdArgs {
  const int32_t* group_off;   // [num_groups + 1] or nullptr
  const int32_t* col_off;     // [num_groups + 1] or nullptr
  const int32_t* col_count;   // compact column count (proto_valid) or nullptr
  int64_t n_rows;
  int64_t m;
  const int32_t* rcode;
  const int32_t* rseg;
  const int32_t* ccode;
  float* stats;
  float* nll;
  float* partial;
  int mode;
  int nkb;       // 64-wide K blocks
  int ksteps;    // 16-wide K steps
  int stages;    // B ring depth (1 or 2)
};

template <int kMode>
__global__ void __launch_bounds__(kFwdThreads, 1)
segsort_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_eh,
                      const __grid_constant__ CUtensorMap map_el,
                      const __grid_constant__ CUtensorMap map_ph,
                      const __grid_constant__ CUtensorMap map_pl, const TcFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full, bar_b_full[4], bar_b_empty[4], bar_t_full[2],
      bar_t_empty[2];
  __shared__ uint32_t s_tmem_base;
  // prototype codes of the tiles in flight, written by warp 0.  The producer runs at most
  // `stages` tiles ahead of the MMA and an MMA only starts once every epilogue warp is
  // past the tile three before it, so a ring of stages + 3 <= 8 slots is never overrun.
  __shared__ __align__(16) int32_t s_ccode[8][kTcBN];
  __shared__ float s_part[3][kTcBM][3];   // partial sums of the column quarters 1..3
  __shared__ float s_nll[kTcBM];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform role id
  const int g = blockIdx.y;
  const int64_t r_begin = a.group_off ? a.group_off[g] : 0;
  const int64_t r_end = a.group_off ? a.group_off[g + 1] : a.n_rows;
  const int64_t row0 = r_begin + (int64_t)blockIdx.x * kTcBM;
  float* my_partial = a.partial + (size_t)g * gridDim.x + blockIdx.x;
  if (row0 >= r_end) {
    if (tid == 0) *my_partial = 0.f;
    return;
  }
  const int rows = (int)min((int64_t)kTcBM, r_end - row0);
  const int c_begin = a.col_off ? a.col_off[g] : 0;
  const int c_end = a.col_count ? *a.col_count : (a.col_off ? a.col_off[g + 1] : (int)a.m);
  const int ntiles = c_end > c_begin ? (c_end - c_begin + kTcBN - 1) / kTcBN : 0;

  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_hi = smem;                                   // [nkb][128 x 128 B]
  uint8_t* a_lo = smem + (size_t)a.nkb * kKBlockBytesA;
  uint8_t* b_ring = smem + (size_t)2 * a.nkb * kKBlockBytesA;
  const uint32_t stage_bytes = 2u * a.nkb * kKBlockBytesB;  // hi blocks then lo blocks

  if (warp == 0 && lane == 0) {
    tc::mbar_init(&bar_a_full, 1);
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&bar_b_full[s], 1);
      tc::mbar_init(&bar_b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_t_full[s], 1);
      tc::mbar_init(&bar_t_empty[s], kFwdEpiWarps);
    }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&map_eh);
    tc::prefetch_tensormap(&map_el);
    tc::prefetch_tensormap(&map_ph);
    tc::prefetch_tensormap(&map_pl);
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, 4 * kTcBN);   // two accumulators of [hh+lh | hl]
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_a_full, 2u * a.nkb * kKBlockBytesA);
      for (int kb = 0; kb < a.nkb; ++kb) {
        tc::tma_load_2d(&map_eh, &bar_a_full, a_hi + (size_t)kb * kKBlockBytesA, kb * 64,
                        (int32_t)row0);
        tc::tma_load_2d(&map_el, &bar_a_full, a_lo + (size_t)kb * kKBlockBytesA, kb * 64,
                        (int32_t)row0);
      }
    }
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % a.stages, use = j / a.stages;
      const int32_t c0 = c_begin + j * kTcBN;
      // the tile's prototype codes do not depend on the ring slot: fetch them before the wait
      int code4[kTcBN / 32];
#pragma unroll
      for (int h = 0; h < kTcBN / 32; ++h) {
        const int k = lane + 32 * h;
        code4[h] = c0 + k < c_end ? a.ccode[c0 + k] : 0;
      }
      tc::mbar_wait(&bar_b_empty[s], (use & 1) ^ 1);      // every lane: the stage is free
      // they ride along with the stage (ordered before lane 0's arrive by __syncwarp, seen by
      // the epilogue through b_full -> MMA -> t_full)
#pragma unroll
      for (int h = 0; h < kTcBN / 32; ++h) s_ccode[j & 7][lane + 32 * h] = code4[h];
      __syncwarp();
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&bar_b_full[s], stage_bytes);
        // per K block the hi tile is followed by the lo tile: read as ONE 256-row B operand
        uint8_t* bh = b_ring + (size_t)s * stage_bytes;
        for (int kb = 0; kb < a.nkb; ++kb) {
          tc::tma_load_2d(&map_ph, &bar_b_full[s], bh + (size_t)(2 * kb) * kKBlockBytesB, kb * 64, c0);
          tc::tma_load_2d(&map_pl, &bar_b_full[s], bh + (size_t)(2 * kb + 1) * kKBlockBytesB, kb * 64,
                          c0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // the whole warp runs the loop (uniform control flow and operands); one elected lane
    // issues the MMAs and commits
    {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kTcBM, kTcBN, 0, 0);
      constexpr uint32_t idesc2x = tc::umma_idesc_bf16(kTcBM, 2 * kTcBN, 0, 0);
      const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
      const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
      const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
      const uint32_t ring_lo = tc::umma_desc_lo(tc::smem_u32(b_ring), 16);
      tc::mbar_wait(&bar_a_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % a.stages, use = j / a.stages;
        const int acc = j & 1, ause = j >> 1;
        tc::mbar_wait(&bar_t_empty[acc], (ause & 1) ^ 1);
        tc::mbar_wait(&bar_b_full[s], use & 1);
        tc::tcgen05_fence_after();
        // the three split products run as two instructions per K step (an SS MMA with
        // M = 128 costs max(~48, N/2) cycles, scripts/micro/mma_issue.cu):
        //   A_hi x [B_hi ; B_lo]  (N = 256)  ->  columns [0,128) = hi.hi, [128,256) = hi.lo
        //   A_lo x  B_hi          (N = 128)  ->  added to columns [0,128)
        const uint32_t d_tmem = tmem_base + acc * 2 * kTcBN;
        const uint32_t b_lo = ring_lo + s * (stage_bytes >> 4);   // K-major: LBO unused (1)
        if (tc::elect_one()) {
          uint32_t accumulate = 0;
          for (int kb = 0; kb < a.nkb; ++kb) {
            const int steps = min(4, a.ksteps - kb * 4);
            uint32_t ah = ah_lo + kb * (kKBlockBytesA >> 4), al = al_lo + kb * (kKBlockBytesA >> 4);
            uint32_t bp = b_lo + kb * (2 * kKBlockBytesB >> 4);
            for (int ks = 0; ks < steps; ++ks) {   // 16 bf16 = 32 bytes inside the swizzle atom
              tc::umma_bf16_words(d_tmem, ah, hi_k, bp, hi_k, idesc2x, accumulate);
              tc::umma_bf16_words(d_tmem, al, hi_k, bp, hi_k, idesc, 1);
              accumulate = 1;
              ah += 2, al += 2, bp += 2;
            }
          }
          tc::umma_commit(&bar_b_empty[s]);    // the ring slot can be refilled
          tc::umma_commit(&bar_t_full[acc]);   // the accumulator is complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue
    const int sp = warp & 3;                  // TMEM sub-partition of this warp
    const int quarter = (warp - 2) >> 2;      // which 32-column quarter of the tile
    const int row = sp * 32 + lane;
    const bool row_ok = row < rows;
    const int code_i = row_ok ? a.rcode[row0 + row] : 0;
    const int seg_i = row_ok ? a.rseg[row0 + row] : -1;
    float same = 0.f, diff = 0.f, self = 0.f;
    const int cb = quarter * 32;

    for (int j = 0; j < ntiles; ++j) {
      const int acc = j & 1, ause = j >> 1, st = j & 7;
      const int c0 = c_begin + j * kTcBN;
      tc::mbar_wait(&bar_t_full[acc], ause & 1);
      tc::tcgen05_fence_after();
      const bool tail = c0 + kTcBN > c_end;          // only the last tile has dead columns
      const int seg_rel = seg_i - c0 - cb;           // own segment relative to this quarter
      uint32_t v[32], w[32];
      const uint32_t taddr = tmem_base + acc * 2 * kTcBN + cb + (static_cast<uint32_t>(sp * 32) << 16);
      tc::tmem_ld_32x32(taddr, v);             // hi.hi + lo.hi
      tc::tmem_ld_32x32(taddr + kTcBN, w);     // hi.lo
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();              // the loads are done: hand the accumulator back
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_t_empty[acc]);
      const int4* codes = reinterpret_cast<const int4*>(&s_ccode[st][cb]);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const int4 c4 = codes[q4];
        const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q4 * 4 + u;
          float s = tc::fast_exp2(__uint_as_float(v[q]) + __uint_as_float(w[q]));
          if (tail) s = c0 + cb + q < c_end ? s : 0.f;
          const bool match = kMode == SPML_MODE_TAGS ? (code_i & cc[u]) != 0 : code_i == cc[u];
          if (match) same += s; else diff += s;
          v[q] = __float_as_uint(s);
        }
      }
      // a pixel's own segment is ONE column of the bank: only the warps that hold it look
      if (__any_sync(0xffffffffu, static_cast<unsigned>(seg_rel) < 32u)) {
#pragma unroll
        for (int q = 0; q < 32; ++q)
          if (q == seg_rel) self += __uint_as_float(v[q]);
      }
    }
    // the four column quarters of a row live in different warps: combine in a fixed order
    if (quarter > 0) {
      s_part[quarter - 1][row][0] = same;
      s_part[quarter - 1][row][1] = diff;
      s_part[quarter - 1][row][2] = self;
    }
    tc::named_bar_sync(1, kFwdEpiThreads);
    if (quarter == 0) {
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) {
        same += s_part[qq][row][0];
        diff += s_part[qq][row][1];
        self += s_part[qq][row][2];
      }
      float nll = 0.f;
      if (row_ok) {
        const float others = same - self;   // loss.py:64-70
        const bool pos = others > 0.f;
        const float num = pos ? others : self;
        const float den = diff + num;
        nll = -logf(num / den);
        float* st = a.stats + (row0 + row) * 3;
        st[0] = num;
        st[1] = den;
        st[2] = pos ? 1.f : 0.f;
        if (a.nll) a.nll[row0 + row] = nll;
      }
      s_nll[row] = nll;
    }
    tc::named_bar_sync(1, kFwdEpiThreads);
    if (warp == 2) {
      float v = s_nll[lane] + s_nll[lane + 32] + s_nll[lane + 64] + s_nll[lane + 96];
      v = warp_sum(v);
      if (lane == 0) *my_partial = v;
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 4 * kTcBN);
  }
}

// ------------------------------------------------------------------------- backward kernel
//
// One kernel, two roles.  The OWNER side is a resident tile of 128 entities whose
// gradient accumulates in TMEM; the STREAMED side passes by in tiles of 64:
//   kProtoOwner = false : owner = 128 pixel rows, streamed = prototypes, output dE
//   kProtoOwner = true  : owner = 128 prototypes, streamed = pixel rows, output dP
// Per streamed tile:   GEMM 1  S^(T) = owner . streamed^T          (K = D, both K-major)
//                      epilogue G = coef S ((diff + numset) / den - numset / num)
//                               -> bf16 hi / lo, written as a K-major SW128 operand
//                      GEMM 2  d(owner) += G . streamed            (K = 64; the streamed
//                               tile is re-used as an MN-major operand: same bytes, the
//                               128-byte rows now index K)
// The similarity is recomputed instead of stored, as in the fp32 path.

constexpr int kBwdBM = 128;
constexpr int kBwdBN = 64;
constexpr int kBwdTileBytesA = kBwdBM * 128;   // one 64-wide K block of the owner tile
constexpr int kBwdTileBytesB = kBwdBN * 128;   // one 64-wide block of the streamed tile
constexpr int kBwdGBytes = kBwdBM * 128;       // G tile, 128 x 64 bf16
constexpr int kBwdMaxStages = 4;

#ifdef SPML_TC_TRACE
#define SPML_TRACE(slot)                                                              \
  do {                                                                                \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 64 && \
        (threadIdx.x == 32 || threadIdx.x == 64))                                     \
      a.trace[j * 16 + (slot)] = clock64();                                           \
  } while (0)
#else
#define SPML_TRACE(slot) do { } while (0)
#endif

struct TcBwdArgs {
  long long* trace;           // SPML_TC_TRACE builds: per-tile timestamps of CTA 0
  spml_segsort_desc d;        // for group ranges / reduction weights
  const int32_t* col_count;   // compact column count (proto_valid) or nullptr
  const int32_t* col_src;     // compact column -> original column, or nullptr
  const int32_t* rcode;
  const int32_t* rseg;
  const int32_t* ccode;
  const float* stats;
  const float* grad_loss;
  float4* pm;                 // [n_rows] per-pixel gradient weights: written by the pixel-owner
                              // kernel (or pix_meta_kernel), read by the prototype-owner kernel
  float* out;                 // demb, or the [chunks][m][dim] prototype partials
  int64_t ld_out;
  float beta;
  float inv_scale;            // 1 / (kappa log2 e): the pixel operand rows are pre-scaled
  int nkb, ksteps, stages;
  int n2;                     // GEMM 2 N: dim rounded up to 16
  int tmem_cols;
  int stacked;                // nkb == 1: hi/lo streamed tiles read as one operand (2 MMAs per K step)
  int64_t proto_rows;         // prototype-owner kernel: only prototypes [0, proto_rows) get a
                              // gradient; the partials are [chunks][proto_rows][dim]
};

// Per-pixel gradient weights.  G_ij = S_ij * w(match_ij, own_ij) with
//   w = coef ((diff + numset) / den - numset / num),  diff = 1 - match,
//   numset = pos ? match - own : own        (loss.py:64-80)
// which only takes four values per pixel.
struct PixMeta {
  float w00, w10, w01, w11;   // w[match][own]: w00 = (0,0), w10 = (1,0), w01 = (0,1), w11 = (1,1)
};

__device__ __forceinline__ PixMeta load_pix_meta(const TcBwdArgs& a, int64_t r, float weight) {
  const float* st = a.stats + r * 3;
  const float inv_num = 1.f / st[0], inv_den = 1.f / st[1];
  const bool pos = st[2] != 0.f;
  const float coef = a.d.kappa * (*a.grad_loss) * weight;
  PixMeta pm;
  pm.w00 = coef * inv_den;
  pm.w10 = pos ? coef * (inv_den - inv_num) : 0.f;
  pm.w01 = pos ? coef * inv_num : coef * (2.f * inv_den - inv_num);
  pm.w11 = pos ? 0.f : coef * (inv_den - inv_num);
  return pm;
}

// the weights of every live row, for a backward call that asks for d(prototypes) only
__global__ void pix_meta_kernel(const TcBwdArgs a) {
  const int g = blockIdx.y;
  const spml_segsort_desc& d = a.d;
  const int64_t r_begin = d.group_off ? d.group_off[g] : 0;
  const int64_t r_end = d.group_off ? d.group_off[g + 1] : d.n_rows;
  const int64_t r = r_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r_end) return;
  const PixMeta pm = load_pix_meta(a, r, reduction_weight(d, g));
  const float u = a.inv_scale;
  a.pm[r] = make_float4(pm.w00 * u, pm.w10 * u, pm.w01 * u, pm.w11 * u);
}

template <int kMode>
__device__ __forceinline__ float grad_elem(float z, int code_pix, int code_pro,
                                           bool own, const PixMeta& pm) {
  const float s = tc::fast_exp2(z);
  const bool match = kMode == SPML_MODE_TAGS ? (code_pix & code_pro) != 0 : code_pix == code_pro;
  const float w_own = match ? pm.w11 : pm.w01;
  const float w_oth = match ? pm.w10 : pm.w00;
  return s * (own ? w_own : w_oth);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float x, float y) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <bool kProtoOwner, int kMode>
__global__ void __launch_bounds__(kTcThreads, 1)
segsort_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_ah,
                      const __grid_constant__ CUtensorMap map_al,
                      const __grid_constant__ CUtensorMap map_bh,
                      const __grid_constant__ CUtensorMap map_bl, const TcBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full, bar_b_full[kBwdMaxStages],
      bar_b_empty[kBwdMaxStages], bar_t_full[2], bar_t_empty[2], bar_g_full[2], bar_g_empty[2],
      bar_d_full;
  __shared__ uint32_t s_tmem_base;
  // metadata of the streamed tile in each ring stage, written by warp 0 before the stage's
  // TMA is armed; a stage is only recycled after GEMM 2 of its tile, i.e. after the
  // epilogue that reads these arrays
  __shared__ __align__(16) int32_t s_code[kBwdMaxStages][kBwdBN];
  __shared__ __align__(16) int32_t s_seg[kBwdMaxStages][kBwdBN];
  __shared__ __align__(16) PixMeta s_pm[kBwdMaxStages][kBwdBN];
  __shared__ int32_t s_segmin[kBwdMaxStages], s_segmax[kBwdMaxStages];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform role id
  const int g = blockIdx.y;
  const spml_segsort_desc& d = a.d;
  const int64_t r_begin = d.group_off ? d.group_off[g] : 0;
  const int64_t r_end = d.group_off ? d.group_off[g + 1] : d.n_rows;
  const int c_begin = d.col_off ? d.col_off[g] : 0;
  const int c_end = a.col_count ? *a.col_count : (d.col_off ? d.col_off[g + 1] : (int)d.m);

  // owner tile [o0, o0 + 128) and streamed range [s_lo, s_hi) in steps of 64
  int64_t o0, o_end, s_lo, s_hi;
  if (!kProtoOwner) {
    o0 = r_begin + (int64_t)blockIdx.x * kBwdBM;
    o_end = r_end;
    s_lo = c_begin;
    s_hi = c_end;
  } else {
    o0 = c_begin + (int64_t)blockIdx.x * kBwdBM;
    o_end = a.col_src ? (int64_t)c_end : min((int64_t)c_end, a.proto_rows);
    const int64_t steps = (r_end - r_begin + kBwdBN - 1) / kBwdBN;
    const int64_t per = (steps + gridDim.z - 1) / gridDim.z;
    s_lo = r_begin + (int64_t)blockIdx.z * per * kBwdBN;
    s_hi = min(r_end, s_lo + per * kBwdBN);
  }
  if (o0 >= o_end) return;
  // compacted columns keep their order: a tile that starts past the limit has nothing to do
  if (kProtoOwner && a.col_src && a.col_src[o0] >= a.proto_rows) return;
  const int owned = (int)min((int64_t)kBwdBM, o_end - o0);
  if (s_lo >= s_hi) {
    // nothing streams past this tile: dP partials are pre-zeroed, dE rows are written here
    if (!kProtoOwner && warp >= 2 && (warp - 2) < 4) {   // one warp per 32 owner rows
      const int row = (warp - 2) * 32 + lane;
      if (row < owned) {
        const int64_t orow = o0 + row;
        float* out_row = a.out + (d.row_index ? (int64_t)d.row_index[orow] : orow) * a.ld_out;
        for (int col = 0; col < d.dim; ++col)
          out_row[col] = a.beta != 0.f ? a.beta * out_row[col] : 0.f;
      }
    }
    return;
  }
  const int ntiles = (int)((s_hi - s_lo + kBwdBN - 1) / kBwdBN);
  const float weight = reduction_weight(d, g);

  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + (size_t)a.nkb * kBwdTileBytesA;
  uint8_t* b_ring = a_lo + (size_t)a.nkb * kBwdTileBytesA;
  const uint32_t stage_bytes = 2u * a.nkb * kBwdTileBytesB;
  uint8_t* g_ring = b_ring + (size_t)a.stages * stage_bytes;   // [2][hi | lo][16 KB]

  if (warp == 0 && lane == 0) {
    tc::mbar_init(&bar_a_full, 1);
    tc::mbar_init(&bar_d_full, 1);
    for (int s = 0; s < kBwdMaxStages; ++s) {
      tc::mbar_init(&bar_b_full[s], 1);
      tc::mbar_init(&bar_b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_t_full[s], 1);
      tc::mbar_init(&bar_t_empty[s], kTcEpiThreads / 32);
      tc::mbar_init(&bar_g_full[s], kTcEpiThreads / 32);
      tc::mbar_init(&bar_g_empty[s], 1);
    }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&map_ah);
    tc::prefetch_tensormap(&map_al);
    tc::prefetch_tensormap(&map_bh);
    tc::prefetch_tensormap(&map_bl);
  }
  if (warp == 1) tc::tmem_alloc(&s_tmem_base, a.tmem_cols);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // stacked: an S buffer is [hi.hi + lo.hi | hi.lo] = 128 columns, d(owner) likewise
  const uint32_t s_stride = a.stacked ? 2 * kBwdBN : kBwdBN;
  const uint32_t tmem_acc = tmem_base + 2 * s_stride;   // d(owner) accumulator columns

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_a_full, 2u * a.nkb * kBwdTileBytesA);
      for (int kb = 0; kb < a.nkb; ++kb) {
        tc::tma_load_2d(&map_ah, &bar_a_full, a_hi + (size_t)kb * kBwdTileBytesA, kb * 64,
                        (int32_t)o0);
        tc::tma_load_2d(&map_al, &bar_a_full, a_lo + (size_t)kb * kBwdTileBytesA, kb * 64,
                        (int32_t)o0);
      }
    }
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % a.stages, use = j / a.stages;
      const int64_t s0 = s_lo + (int64_t)j * kBwdBN;
      int seg_min = 0x7fffffff, seg_max = -1;
      // the tile's metadata does not depend on the ring slot: fetch it (both halves of the 64
      // entries at once) while the slot is still busy
      int code2[2], seg2[2];
      float4 pm2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t e = s0 + lane + 32 * h;
        const bool in = e < s_hi;
        if (!kProtoOwner) {
          code2[h] = in ? a.ccode[e] : 0;
        } else {
          code2[h] = in ? a.rcode[e] : 0;
          seg2[h] = in ? a.rseg[e] : -1;
          pm2[h] = in ? a.pm[e] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc::mbar_wait(&bar_b_empty[s], (use & 1) ^ 1);      // every lane: the stage is free
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        s_code[s][k] = code2[h];
        if (kProtoOwner) {
          s_seg[s][k] = seg2[h];
          *reinterpret_cast<float4*>(&s_pm[s][k]) = pm2[h];
          if (seg2[h] >= 0) seg_min = min(seg_min, seg2[h]), seg_max = max(seg_max, seg2[h]);
        }
      }
      if (kProtoOwner) {
        seg_min = __reduce_min_sync(0xffffffffu, seg_min);
        seg_max = __reduce_max_sync(0xffffffffu, seg_max);
        if (lane == 0) s_segmin[s] = seg_min, s_segmax[s] = seg_max;
      }
      __syncwarp();
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&bar_b_full[s], stage_bytes);
        uint8_t* bh = b_ring + (size_t)s * stage_bytes;
        uint8_t* bl = bh + (size_t)a.nkb * kBwdTileBytesB;
        for (int kb = 0; kb < a.nkb; ++kb) {
          tc::tma_load_2d(&map_bh, &bar_b_full[s], bh + (size_t)kb * kBwdTileBytesB, kb * 64,
                          (int32_t)s0);
          tc::tma_load_2d(&map_bl, &bar_b_full[s], bl + (size_t)kb * kBwdTileBytesB, kb * 64,
                          (int32_t)s0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== GEMM 1 issuer
    // the whole warp runs the loop (uniform control flow and operands); one elected lane
    // issues the MMAs and commits
    {
      constexpr uint32_t idesc1 = tc::umma_idesc_bf16(kBwdBM, kBwdBN, 0, 0);
      constexpr uint32_t idesc1s = tc::umma_idesc_bf16(kBwdBM, 2 * kBwdBN, 0, 0);
      const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
      const uint32_t ah_lo = tc::umma_desc_lo(tc::smem_u32(a_hi), 16);
      const uint32_t al_lo = tc::umma_desc_lo(tc::smem_u32(a_lo), 16);
      const uint32_t ring_lo = tc::umma_desc_lo(tc::smem_u32(b_ring), 16);
      auto gemm1 = [&](int j) {
        const int s = j % a.stages, use = j / a.stages;
        const int acc = j & 1, ause = j >> 1;
        SPML_TRACE(8);
        tc::mbar_wait(&bar_t_empty[acc], (ause & 1) ^ 1);
        SPML_TRACE(9);
        tc::mbar_wait(&bar_b_full[s], use & 1);
        tc::tcgen05_fence_after();
        SPML_TRACE(10);
        const uint32_t d_tmem = tmem_base + acc * s_stride;
        const uint32_t bh_lo = ring_lo + s * (stage_bytes >> 4);
        const uint32_t bl_lo = bh_lo + a.nkb * (kBwdTileBytesB >> 4);
        uint32_t accumulate = 0;
        if (tc::elect_one()) {
        if (a.stacked) {
          // the lo tile follows the hi tile: [B_hi ; B_lo] is one 128-row K-major operand:
          // 2 MMAs (N = 128, 64) instead of 3 (N = 64) per K step; an SS MMA never costs
          // less than ~48 cycles
          uint32_t ah = ah_lo, al = al_lo, bp = bh_lo;
          for (int ks = 0; ks < a.ksteps; ++ks) {
            tc::umma_bf16_words(d_tmem, ah, hi_k, bp, hi_k, idesc1s, accumulate);   // hh | hl
            tc::umma_bf16_words(d_tmem, al, hi_k, bp, hi_k, idesc1, 1);             // += lh
            accumulate = 1;
            ah += 2, al += 2, bp += 2;
          }
        } else
        for (int kb = 0; kb < a.nkb; ++kb) {
          const int steps = min(4, a.ksteps - kb * 4);
          uint32_t ah = ah_lo + kb * (kBwdTileBytesA >> 4), al = al_lo + kb * (kBwdTileBytesA >> 4);
          uint32_t bh = bh_lo + kb * (kBwdTileBytesB >> 4), bl = bl_lo + kb * (kBwdTileBytesB >> 4);
          for (int ks = 0; ks < steps; ++ks) {
            tc::umma_bf16_words(d_tmem, ah, hi_k, bh, hi_k, idesc1, accumulate);
            tc::umma_bf16_words(d_tmem, al, hi_k, bh, hi_k, idesc1, 1);
            tc::umma_bf16_words(d_tmem, ah, hi_k, bl, hi_k, idesc1, 1);
            accumulate = 1;
            ah += 2, al += 2, bh += 2, bl += 2;
          }
        }
        tc::umma_commit(&bar_t_full[acc]);
        }
        __syncwarp();
        SPML_TRACE(11);
      };
      tc::mbar_wait(&bar_a_full, 0);
      for (int j = 0; j < ntiles; ++j) gemm1(j);   // the ring and t_empty barriers pace it
    }
  } else if (warp == kBwdGemm2Warp) {
    // ===================================================================== GEMM 2 issuer
    {
      const uint32_t idesc2 = tc::umma_idesc_bf16(kBwdBM, a.n2, 0, 1);
      const uint32_t idesc2s = tc::umma_idesc_bf16(kBwdBM, 64 + a.n2, 0, 1);
      const uint32_t hi_k = tc::umma_desc_hi_sw128(1024);
      // G tile K-major; streamed tile MN-major with LBO = next 64-wide N atom
      const uint32_t g_lo = tc::umma_desc_lo(tc::smem_u32(g_ring), 16);
      const uint32_t ring_mn_lo = tc::umma_desc_lo(tc::smem_u32(b_ring), kBwdTileBytesB);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % a.stages, use = j / a.stages, gb = j & 1, guse = j >> 1;
        tc::mbar_wait(&bar_b_full[s], use & 1);     // this thread reads the streamed tile too
        tc::mbar_wait(&bar_g_full[gb], guse & 1);
        tc::tcgen05_fence_after();
        uint32_t bh = ring_mn_lo + s * (stage_bytes >> 4);
        uint32_t bl = bh + a.nkb * (kBwdTileBytesB >> 4);
        uint32_t accumulate = j > 0 ? 1u : 0u;
        if (tc::elect_one()) {
        if (a.stacked) {
          // MN-major B: the next 64-wide N atom (LBO) of the hi tile IS the lo tile, so
          // G_hi x [P_hi | P_lo] is one MMA of N = 64 + n2; G_lo x P_hi adds to block 0
          uint32_t gh = g_lo + gb * (2 * kBwdGBytes >> 4), gl = gh + (kBwdGBytes >> 4);
          for (int ks = 0; ks < kBwdBN / 16; ++ks) {
            tc::umma_bf16_words(tmem_acc, gh, hi_k, bh, hi_k, idesc2s, accumulate);
            tc::umma_bf16_words(tmem_acc, gl, hi_k, bh, hi_k, idesc2, 1);
            accumulate = 1;
            gh += 2, gl += 2, bh += 128;
          }
        } else {
          uint32_t gh = g_lo + gb * (2 * kBwdGBytes >> 4), gl = gh + (kBwdGBytes >> 4);
          for (int ks = 0; ks < kBwdBN / 16; ++ks) {
            // A: 16 columns of G = 32 bytes inside the swizzle atom; B: 16 K rows = 2048 bytes
            tc::umma_bf16_words(tmem_acc, gh, hi_k, bh, hi_k, idesc2, accumulate);
            tc::umma_bf16_words(tmem_acc, gl, hi_k, bh, hi_k, idesc2, 1);
            tc::umma_bf16_words(tmem_acc, gh, hi_k, bl, hi_k, idesc2, 1);
            accumulate = 1;
            gh += 2, gl += 2, bh += 128, bl += 128;
          }
        }
        // GEMM 1 of this tile completed before its G existed: the slot is free after GEMM 2
        tc::umma_commit(&bar_b_empty[s]);
        tc::umma_commit(&bar_g_empty[gb]);
        if (j + 1 == ntiles) tc::umma_commit(&bar_d_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue
    const int sp = warp & 3;
    const int quarter = (warp - 2) >> 2;       // 16-column quarter of the 64-wide S tile
    const int row = sp * 32 + lane;            // owner entity of this thread
    const bool row_ok = row < owned;
    const int64_t orow = o0 + row;
    // owner-side constants
    int code_o = 0, seg_o = -1;
    PixMeta pm_o = {0.f, 0.f, 0.f, 0.f};   // padding rows: all-zero weights
    if (row_ok) {
      if (!kProtoOwner) {
        code_o = a.rcode[orow];
        seg_o = a.rseg[orow];
        pm_o = load_pix_meta(a, orow, weight);
        if (quarter == 0 && a.pm) {   // for the prototype-owner kernel: its second GEMM multiplies
          const float u = a.inv_scale;   // G with the SCALED pixel rows
          a.pm[orow] = make_float4(pm_o.w00 * u, pm_o.w10 * u, pm_o.w01 * u, pm_o.w11 * u);
        }
      } else {
        code_o = a.ccode[orow];
      }
    }
    const uint32_t sw = (row & 7);
    const uint32_t g_row_off = (row >> 3) * 1024 + sw * 128;

    for (int j = 0; j < ntiles; ++j) {
      const int acc = j & 1, ause = j >> 1, gb = j & 1, guse = j >> 1, st = j % a.stages;
      const int64_t s0 = s_lo + (int64_t)j * kBwdBN;
      SPML_TRACE(0);
      tc::mbar_wait(&bar_t_full[acc], ause & 1);
      tc::tcgen05_fence_after();
      SPML_TRACE(1);
      uint32_t v[kBwdEpiCols];
      const int cb = quarter * kBwdEpiCols;
      const uint32_t s_addr = tmem_base + acc * s_stride + cb + (static_cast<uint32_t>(sp * 32) << 16);
      tc::tmem_ld_32x16(s_addr, v);
      if (a.stacked) {
        uint32_t w[kBwdEpiCols];
        tc::tmem_ld_32x16(s_addr + kBwdBN, w);   // the hi.lo block
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < kBwdEpiCols; ++q)
          v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w[q]));
      } else {
        tc::tmem_ld_wait();
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_t_empty[acc]);

      // G = exp(kappa S) * w(match, own), computed in place.  `own` (the pixel's own segment)
      // holds for one column per pixel, so almost every (warp, tile) pair takes the lean loop
      // that only applies the match weights: the epilogue is instruction-issue-bound (the
      // tensor pipe needs ~900 cycles per step, this loop was ~2400 with the own selects in).
      bool any_own;
      int own_rel = 0;
      if (!kProtoOwner) {
        own_rel = seg_o - (int)s0 - cb;   // column of this pixel's own prototype
        any_own = __any_sync(0xffffffffu, static_cast<unsigned>(own_rel) < (unsigned)kBwdEpiCols);
      } else {
        // pixels of the streamed tile own prototypes in [segmin, segmax] only
        const int r_lo = (int)o0 + sp * 32;
        any_own = s_segmax[st] >= r_lo && s_segmin[st] <= r_lo + 31;
      }
      if (!any_own) {
#pragma unroll
        for (int q4 = 0; q4 < kBwdEpiCols / 4; ++q4) {
          const int4 cc = *reinterpret_cast<const int4*>(&s_code[st][cb + q4 * 4]);
          const int c4[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int q = q4 * 4 + r;
            const float e = tc::fast_exp2(__uint_as_float(v[q]));
            const bool match = kMode == SPML_MODE_TAGS ? (code_o & c4[r]) != 0 : code_o == c4[r];
            float w;
            if (!kProtoOwner) {
              w = match ? pm_o.w10 : pm_o.w00;
            } else {
              const float2 wq = *reinterpret_cast<const float2*>(&s_pm[st][cb + q]);   // w00, w10
              w = match ? wq.y : wq.x;
            }
            v[q] = __float_as_uint(e * w);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < kBwdEpiCols; ++q) {
          const int k = cb + q;
          const float z = __uint_as_float(v[q]);
          float gq;
          if (!kProtoOwner) {
            gq = grad_elem<kMode>(z, code_o, s_code[st][k], q == own_rel, pm_o);
          } else {
            gq = grad_elem<kMode>(z, s_code[st][k], code_o, s_seg[st][k] == (int)orow,
                                  s_pm[st][k]);
          }
          v[q] = __float_as_uint(gq);
        }
      }
      if (kProtoOwner && !row_ok) {   // padding prototypes of the last owner tile
#pragma unroll                        // (padding PIXEL rows have all-zero weights instead)
        for (int q = 0; q < kBwdEpiCols; ++q) v[q] = 0u;
      }
      if (s0 + kBwdBN > s_hi) {       // last streamed tile: columns past the end
#pragma unroll
        for (int q = 0; q < kBwdEpiCols; ++q) v[q] = s0 + cb + q < s_hi ? v[q] : 0u;
      }
      SPML_TRACE(2);
      tc::mbar_wait(&bar_g_empty[gb], (guse & 1) ^ 1);
      SPML_TRACE(3);
      {
        uint8_t* gh = g_ring + (size_t)gb * 2 * kBwdGBytes;
        uint8_t* gl = gh + kBwdGBytes;
#pragma unroll
        for (int c4 = 0; c4 < kBwdEpiCols / 8; ++c4) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2) {
            // hi = the top 16 bits (truncation: one byte-permute packs both), lo = the exact
            // remainder rounded to bf16: the pair still carries ~16 mantissa bits
            const uint32_t xb = v[c4 * 8 + p2 * 2], yb = v[c4 * 8 + p2 * 2 + 1];
            hi[p2] = __byte_perm(xb, yb, 0x7632);
            lo[p2] = pack_bf16x2(__uint_as_float(xb) - __uint_as_float(xb & 0xffff0000u),
                                 __uint_as_float(yb) - __uint_as_float(yb & 0xffff0000u));
          }
          const uint32_t chunk = static_cast<uint32_t>(quarter * (kBwdEpiCols / 8) + c4);
          const uint32_t off = g_row_off + ((chunk ^ sw) << 4);
          *reinterpret_cast<uint4*>(gh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(gl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc::fence_proxy_async();      // this thread's G stores -> visible to the tensor core
      }
      SPML_TRACE(4);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_g_full[gb]);
      SPML_TRACE(5);
    }

    // ---- d(owner) out of TMEM
    tc::mbar_wait(&bar_d_full, 0);
    tc::tcgen05_fence_after();
    float* out_row = nullptr;
    if (row_ok) {
      if (!kProtoOwner) {
        const int64_t orig = d.row_index ? (int64_t)d.row_index[orow] : orow;
        out_row = a.out + orig * a.ld_out;
      } else {
        const int64_t orig = a.col_src ? (int64_t)a.col_src[orow] : orow;
        out_row = a.out + ((size_t)blockIdx.z * a.proto_rows + orig) * d.dim;
        if (orig >= a.proto_rows) out_row = nullptr;
      }
    }
    const int nchunks = (a.n2 + 31) / 32;
    for (int ch = quarter; ch < nchunks; ch += kBwdEpiWarps / 4) {
      uint32_t v[32];
      const uint32_t d_addr = tmem_acc + ch * 32 + (static_cast<uint32_t>(sp * 32) << 16);
      tc::tmem_ld_32x32(d_addr, v);
      if (a.stacked) {
        uint32_t w[32];
        tc::tmem_ld_32x32(d_addr + 64, w);       // G_hi x P_lo block
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w[q]));
      } else {
        tc::tmem_ld_wait();
      }
      if (out_row) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int col = ch * 32 + q;
          if (col < d.dim) {
            const float val = __uint_as_float(v[q]);
            out_row[col] = (!kProtoOwner && a.beta != 0.f) ? a.beta * out_row[col] + val : val;
          }
        }
      }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------------- host side

static int tc_tiles_x(const spml_segsort_desc& d) {
  return (int)std::max<int64_t>(1, ceil_div(d.max_rows_per_group, kTcBM));
}

bool segsort_tc_supported(const spml_segsort_desc& d) {
  if (d.dim > 192 || d.dim < 1) return false;
  if (d.n_rows <= 0 || d.m <= 0) return false;
  if (d.proto_valid && (d.col_off || d.num_groups > 1)) return false;
  if (d.n_rows >= (1ll << 31) - kTcBM || d.m >= (1ll << 31) - kTcBN) return false;
  return true;
}

TcPlan segsort_tc_plan(const spml_segsort_desc& d, void* base) {
  TcPlan p{};
  p.dp = (d.dim + 7) & ~7;
  p.nkb = (d.dim + 63) / 64;
  p.ksteps = (d.dim + 15) / 16;
  p.stages = p.nkb == 1 ? 4 : (p.nkb == 2 ? 2 : 1);   // forward ring: what fits next to the A tile
  char* ptr = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = ptr + off;
    off += align_up(bytes, 256);
    return r;
  };
  p.partial = reinterpret_cast<float*>(take((size_t)d.num_groups * tc_tiles_x(d) * sizeof(float)));
  p.eh = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.n_rows * p.dp * 2));
  p.el = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.n_rows * p.dp * 2));
  p.ph = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.m * p.dp * 2));
  p.pl = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.m * p.dp * 2));
  p.rcode = reinterpret_cast<int32_t*>(take((size_t)d.n_rows * 4));
  p.rseg = reinterpret_cast<int32_t*>(take((size_t)d.n_rows * 4));
  p.ccode = reinterpret_cast<int32_t*>(take((size_t)d.m * 4));
  p.col_dst = reinterpret_cast<int32_t*>(take((size_t)d.m * 4));
  p.col_src = reinterpret_cast<int32_t*>(take((size_t)d.m * 4));
  p.col_count = reinterpret_cast<int32_t*>(take(16));
  p.pm = reinterpret_cast<float4*>(take((size_t)d.n_rows * sizeof(float4)));
  p.bytes = off;
  return p;
}

// pre-pass shared by forward and backward: compact columns, split both operands
int segsort_tc_prepare(const spml_segsort_desc& d, const TcPlan& p, cudaStream_t st) {
  const bool compact = d.proto_valid != nullptr;
  if (compact) {
    compact_cols_kernel<<<1, 1024, 0, st>>>(d.proto_valid, (int)d.m, p.col_dst, p.col_src,
                                            p.col_count);
    SPML_LAUNCH_CHECK("compact_cols_kernel");
  }
  split_rows_kernel<<<(unsigned)ceil_div(d.n_rows, 8), 256, 0, st>>>(
      d.emb, d.ld_emb, d.dim, p.dp, (float)((double)d.kappa * 1.4426950408889634), d.row_index, d.group_off, d.num_groups, d.n_rows, d.pix_code,
      d.seg, compact ? p.col_dst : nullptr, d.m, p.eh, p.el, p.rcode, p.rseg);
  SPML_LAUNCH_CHECK("split_rows_kernel");
  split_protos_kernel<<<(unsigned)ceil_div(d.m, 8), 256, 0, st>>>(
      d.protos, d.ld_protos, d.dim, p.dp, compact ? p.col_src : nullptr,
      compact ? p.col_count : nullptr, d.m, d.proto_code, p.ph, p.pl, p.ccode);
  SPML_LAUNCH_CHECK("split_protos_kernel");
  return SPML_OK;
}

int segsort_fwd_tc(const spml_segsort_desc& d, const TcPlan& p, float* stats, float* nll,
                   cudaStream_t st) {
  int rc = segsort_tc_prepare(d, p, st);
  if (rc != SPML_OK) return rc;
  CUtensorMap map_eh, map_el, map_ph, map_pl;
  const uint64_t pitch = (uint64_t)p.dp * 2;
  if ((rc = make_tensor_map_bf16_2d(&map_eh, p.eh, p.dp, d.n_rows, pitch, 64, kTcBM))) return rc;
  if ((rc = make_tensor_map_bf16_2d(&map_el, p.el, p.dp, d.n_rows, pitch, 64, kTcBM))) return rc;
  if ((rc = make_tensor_map_bf16_2d(&map_ph, p.ph, p.dp, d.m, pitch, 64, kTcBN))) return rc;
  if ((rc = make_tensor_map_bf16_2d(&map_pl, p.pl, p.dp, d.m, pitch, 64, kTcBN))) return rc;

  TcFwdArgs a{};
  a.group_off = d.group_off;
  a.col_off = d.col_off;
  a.col_count = d.proto_valid ? p.col_count : nullptr;
  a.n_rows = d.n_rows;
  a.m = d.m;
  a.rcode = p.rcode;
  a.rseg = p.rseg;
  a.ccode = p.ccode;
  a.stats = stats;
  a.nll = nll;
  a.partial = p.partial;
  a.mode = d.mode;
  a.nkb = p.nkb;
  a.ksteps = p.ksteps;
  a.stages = p.stages;
  const size_t smem = 1024 + (size_t)2 * p.nkb * kKBlockBytesA +
                      (size_t)p.stages * 2 * p.nkb * kKBlockBytesB;
  if (smem > 227 * 1024) {
    set_error("segsort_fwd(tc): needs %zu bytes of shared memory", smem);
    return SPML_E_UNSUPPORTED;
  }
  dim3 grid((unsigned)tc_tiles_x(d), (unsigned)d.num_groups);
  if (d.mode == SPML_MODE_TAGS) {
    SPML_CUDA(cudaFuncSetAttribute(segsort_fwd_tc_kernel<SPML_MODE_TAGS>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segsort_fwd_tc_kernel<SPML_MODE_TAGS>
        <<<grid, kFwdThreads, smem, st>>>(map_eh, map_el, map_ph, map_pl, a);
  } else {
    SPML_CUDA(cudaFuncSetAttribute(segsort_fwd_tc_kernel<SPML_MODE_CLASS>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segsort_fwd_tc_kernel<SPML_MODE_CLASS>
        <<<grid, kFwdThreads, smem, st>>>(map_eh, map_el, map_ph, map_pl, a);
  }
  SPML_LAUNCH_CHECK("segsort_fwd_tc_kernel");
  return SPML_OK;
}


#ifdef SPML_TC_TRACE
__device__ long long g_tc_trace[64 * 16];
#endif

static int proto_chunks_for(const spml_segsort_desc& d, int64_t m) {
  const int64_t col_tiles = std::max<int64_t>(1, ceil_div(m, kBwdBM));
  const int64_t steps = std::max<int64_t>(1, ceil_div(d.max_rows_per_group, kBwdBN));
  // with per-group column ranges (img_sim) a group only owns ~1 / num_groups of the column
  // tiles: the others exit at once, so they must not count as work when the rows are split
  const int64_t live_tiles = d.col_off ? ceil_div(col_tiles, d.num_groups) * d.num_groups
                                       : col_tiles * d.num_groups;
  int64_t chunks = ceil_div(2 * 148, live_tiles);
  // with a column mask the buffer is capacity-sized and most column tiles are dead:
  // keep enough row chunks for the live ones to fill the GPU
  if (d.proto_valid) chunks = std::max<int64_t>(chunks, 16);
  return (int)std::max<int64_t>(1, std::min<int64_t>(chunks, steps));
}

// Row chunks of the prototype-owner kernel.  The workspace is sized for proto_rows = m; with
// fewer gradient rows there are fewer column tiles, so the rows are split finer as long as
// the [chunks][proto_rows][dim] partials still fit into that space.
int segsort_tc_proto_chunks(const spml_segsort_desc& d, int64_t proto_rows) {
  const int full = proto_chunks_for(d, d.m);
  if (proto_rows >= d.m || proto_rows <= 0) return full;
  const int64_t fits = (int64_t)full * d.m / proto_rows;
  return (int)std::max<int64_t>(1, std::min<int64_t>(proto_chunks_for(d, proto_rows), fits));
}

int segsort_bwd_tc(const spml_segsort_desc& d, const TcPlan& p, const float* stats,
                   const float* grad_loss, float beta, float* demb, int64_t ld_demb,
                   float* proto_partial, int chunks, int64_t proto_rows, bool prepared,
                   cudaStream_t st) {
  int rc = prepared ? SPML_OK : segsort_tc_prepare(d, p, st);
  if (rc != SPML_OK) return rc;
  const uint64_t pitch = (uint64_t)p.dp * 2;
  TcBwdArgs a{};
  a.d = d;
  a.col_count = d.proto_valid ? p.col_count : nullptr;
  a.col_src = d.proto_valid ? p.col_src : nullptr;
  a.rcode = p.rcode;
  a.rseg = p.rseg;
  a.ccode = p.ccode;
  a.stats = stats;
  a.grad_loss = grad_loss;
  // the weights are published by the pixel-owner kernel only when the prototype-owner kernel of
  // THIS call reads them (a d(embedding)-only call may run next to a d(prototypes)-only call)
  a.pm = proto_partial ? p.pm : nullptr;
  a.inv_scale = (float)(1.0 / ((double)d.kappa * 1.4426950408889634));
  a.nkb = p.nkb;
  a.ksteps = p.ksteps;
  a.stages = p.nkb == 1 ? 4 : (p.nkb == 2 ? 2 : 1);   // what fits next to the A and G tiles
  a.n2 = (d.dim + 15) & ~15;
  static int stack_mode = -1;
  if (stack_mode < 0) {
    const char* e = getenv("SPML_B200_STACK");
    stack_mode = e ? atoi(e) : 1;
  }
  a.stacked = stack_mode && p.nkb == 1;
  a.proto_rows = proto_rows;
  // TMEM columns: S buffers | d(owner)
  const int s_cols = a.stacked ? 4 * kBwdBN : 2 * kBwdBN;
  const int d_cols = a.stacked ? 64 + a.n2 : a.n2;
  a.tmem_cols = s_cols + d_cols <= 256 ? 256 : 512;
  const size_t smem = 1024 + (size_t)2 * p.nkb * kBwdTileBytesA +
                      (size_t)a.stages * 2 * p.nkb * kBwdTileBytesB + (size_t)4 * kBwdGBytes;
  if (smem > 227 * 1024) {
    set_error("segsort_bwd(tc): needs %zu bytes of shared memory", smem);
    return SPML_E_UNSUPPORTED;
  }
  CUtensorMap e128h, e128l, e64h, e64l, p128h, p128l, p64h, p64l;
  if (demb) {
    if ((rc = make_tensor_map_bf16_2d(&e128h, p.eh, p.dp, d.n_rows, pitch, 64, kBwdBM))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&e128l, p.el, p.dp, d.n_rows, pitch, 64, kBwdBM))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&p64h, p.ph, p.dp, d.m, pitch, 64, kBwdBN))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&p64l, p.pl, p.dp, d.m, pitch, 64, kBwdBN))) return rc;
    a.out = demb;
    a.ld_out = ld_demb;
    a.beta = beta;
#ifdef SPML_TC_TRACE
    if (!getenv("SPML_B200_TRACE_PROTO"))
      SPML_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&a.trace), g_tc_trace));
#endif
    dim3 grid((unsigned)tc_tiles_x(d), (unsigned)d.num_groups, 1);
    if (d.mode == SPML_MODE_TAGS) {
      SPML_CUDA(cudaFuncSetAttribute(segsort_bwd_tc_kernel<false, SPML_MODE_TAGS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      segsort_bwd_tc_kernel<false, SPML_MODE_TAGS>
          <<<grid, kTcThreads, smem, st>>>(e128h, e128l, p64h, p64l, a);
    } else {
      SPML_CUDA(cudaFuncSetAttribute(segsort_bwd_tc_kernel<false, SPML_MODE_CLASS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      segsort_bwd_tc_kernel<false, SPML_MODE_CLASS>
          <<<grid, kTcThreads, smem, st>>>(e128h, e128l, p64h, p64l, a);
    }
    SPML_LAUNCH_CHECK("segsort_bwd_tc_kernel<emb>");
  }
  if (proto_partial) {
    a.pm = p.pm;
    if (!demb) {   // nobody has written the per-pixel weights yet
      dim3 pgrid((unsigned)std::max<int64_t>(1, ceil_div(d.max_rows_per_group, 256)),
                 (unsigned)d.num_groups);
      pix_meta_kernel<<<pgrid, 256, 0, st>>>(a);
      SPML_LAUNCH_CHECK("pix_meta_kernel");
    }
    if ((rc = make_tensor_map_bf16_2d(&p128h, p.ph, p.dp, d.m, pitch, 64, kBwdBM))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&p128l, p.pl, p.dp, d.m, pitch, 64, kBwdBM))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&e64h, p.eh, p.dp, d.n_rows, pitch, 64, kBwdBN))) return rc;
    if ((rc = make_tensor_map_bf16_2d(&e64l, p.el, p.dp, d.n_rows, pitch, 64, kBwdBN))) return rc;
    a.out = proto_partial;
    a.ld_out = d.dim;
    a.beta = 0.f;
    a.trace = nullptr;
#ifdef SPML_TC_TRACE
    if (getenv("SPML_B200_TRACE_PROTO"))
      SPML_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&a.trace), g_tc_trace));
#endif
    // compacted columns: the live tiles are only known on the device (dead ones exit at once)
    const int64_t owner_cols = d.proto_valid ? d.m : std::min<int64_t>(d.m, proto_rows);
    dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(owner_cols, kBwdBM)),
              (unsigned)d.num_groups, (unsigned)chunks);
    if (d.mode == SPML_MODE_TAGS) {
      SPML_CUDA(cudaFuncSetAttribute(segsort_bwd_tc_kernel<true, SPML_MODE_TAGS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      segsort_bwd_tc_kernel<true, SPML_MODE_TAGS>
          <<<grid, kTcThreads, smem, st>>>(p128h, p128l, e64h, e64l, a);
    } else {
      SPML_CUDA(cudaFuncSetAttribute(segsort_bwd_tc_kernel<true, SPML_MODE_CLASS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      segsort_bwd_tc_kernel<true, SPML_MODE_CLASS>
          <<<grid, kTcThreads, smem, st>>>(p128h, p128l, e64h, e64l, a);
    }
    SPML_LAUNCH_CHECK("segsort_bwd_tc_kernel<proto>");
  }
  return SPML_OK;
}

}  // namespace spml

// Host-side cost of the driver calls a stage-group entry makes, in nanoseconds per call
// (scripts/host_costs.py): out = {tensor-map encode, cudaFuncSetAttribute, empty kernel
// launch, 16-byte cudaMemsetAsync, event record + stream wait}.
namespace spml {
__global__ void empty_kernel() {}
}
extern "C" int spml_debug_host_costs(double* out, void* stream) {
  using namespace spml;
  using clk = std::chrono::steady_clock;
  cudaStream_t st = as_stream(stream);
  const int n = 2000;
  void* buf = nullptr;
  SPML_CUDA(cudaMalloc(&buf, 1 << 20));
  cudaEvent_t ev;
  SPML_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaStream_t side;
  SPML_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  auto ns = [&](clk::time_point t0) {
    return std::chrono::duration<double, std::nano>(clk::now() - t0).count() / n;
  };
  CUtensorMap map;
  auto t0 = clk::now();
  for (int i = 0; i < n; ++i) make_tensor_map_bf16_2d(&map, buf, 64, 1024 + i % 7, 128, 64, 128);
  out[0] = ns(t0);
  t0 = clk::now();
  for (int i = 0; i < n; ++i)
    cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + (i & 1));
  out[1] = ns(t0);
  SPML_CUDA(cudaStreamSynchronize(st));
  t0 = clk::now();
  for (int i = 0; i < n; ++i) empty_kernel<<<1, 32, 0, st>>>();
  out[2] = ns(t0);
  SPML_CUDA(cudaStreamSynchronize(st));
  t0 = clk::now();
  for (int i = 0; i < n; ++i) cudaMemsetAsync(buf, 0, 16, st);
  out[3] = ns(t0);
  SPML_CUDA(cudaStreamSynchronize(st));
  t0 = clk::now();
  for (int i = 0; i < n; ++i) {
    cudaEventRecord(ev, st);
    cudaStreamWaitEvent(side, ev, 0);
  }
  out[4] = ns(t0);
  SPML_CUDA(cudaStreamSynchronize(st));
  SPML_CUDA(cudaStreamSynchronize(side));
  cudaStreamDestroy(side);
  cudaEventDestroy(ev);
  cudaFree(buf);
  return SPML_OK;
}

#ifdef SPML_TC_TRACE
extern "C" int spml_debug_tc_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, spml::g_tc_trace, sizeof(long long) * 64 * 16) == cudaSuccess
             ? 0 : -2;
}
#endif
