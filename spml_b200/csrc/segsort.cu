// C1 / C2: SegSort and SetSegSort pixel-to-segment contrastive losses
// (reference spml/utils/segsort/loss.py:15-251) and their backward, fp32 CUDA-core
// path.  The [N, M] similarity matrix is never written: each CTA owns a 128-row
// tile, streams the prototype bank through shared memory in 64-column tiles,
// and folds exp / label masks / row sums into the GEMM epilogue.  The backward
// recomputes the similarities (flash-attention style) instead of saving them:
//   fwd     S tile -> masked row sums -> {num, den, branch} per row, nll
//   bwd dE  S tile -> G tile (shared memory) -> dE += G . P          (row owner)
//   bwd dP  S tile -> G tile (shared memory) -> dP += G^T . E        (column owner,
//           rows split into chunks, partials reduced in a fixed order)
// Math: SURVEY.md section 7.3 (checked against the reference's autograd).
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "internal.h"
#include "segsort_tc.h"
#include "tile_gemm.cuh"

namespace spml {

struct TileInfo {
  int64_t r_begin, r_end, row0;
  int rows;
  int c_begin, c_end;
};

__device__ __forceinline__ void group_range(const spml_segsort_desc& d, int g, TileInfo& t) {
  t.r_begin = d.group_off ? d.group_off[g] : 0;
  t.r_end = d.group_off ? d.group_off[g + 1] : d.n_rows;
  t.c_begin = d.col_off ? d.col_off[g] : 0;
  t.c_end = d.col_off ? d.col_off[g + 1] : (int)d.m;
}

__device__ __forceinline__ bool codes_match(int mode, int64_t a, int64_t b) {
  return mode == SPML_MODE_TAGS ? (a & b) != 0 : a == b;
}

// stage a tile of prototype rows: Bt transposed, optionally Pf row-major, plus codes.
template <bool kRowMajorToo>
__device__ __forceinline__ void load_proto_tile(const spml_segsort_desc& d, int c0, int c_end,
                                                int dpad, int ldp, float* Bt, float* Pf,
                                                int64_t* s_pcode, unsigned char* s_pvalid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < BN; k += kGemmThreads / 32) {
    const int c = c0 + k;
    const bool in = c < c_end;
    const float* rp = d.protos + (int64_t)c * d.ld_protos;
    for (int q = lane; q < dpad; q += 32) {
      const float v = (in && q < d.dim) ? rp[q] : 0.f;
      Bt[q * LDB + k] = v;
      if (kRowMajorToo) Pf[k * ldp + q] = v;
    }
  }
  if (threadIdx.x < BN) {
    const int c = c0 + threadIdx.x;
    const bool in = c < c_end;
    s_pcode[threadIdx.x] = in ? d.proto_code[c] : 0;
    s_pvalid[threadIdx.x] = in && (!d.proto_valid || d.proto_valid[c]);
  }
}

// stage a tile of embedding rows: At transposed, optionally Ef row-major, plus labels.
template <bool kRowMajorToo>
__device__ __forceinline__ void load_emb_tile(const spml_segsort_desc& d, int64_t row0, int rows,
                                              int dpad, int ldp, float* At, float* Ef,
                                              int64_t* s_code, int* s_seg, int64_t* s_orig) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < BM; r += kGemmThreads / 32) {
    const float* rp = nullptr;
    if (r < rows)
      rp = d.emb + (d.row_index ? (int64_t)d.row_index[row0 + r] : row0 + r) * d.ld_emb;
    for (int q = lane; q < dpad; q += 32) {
      const float v = (rp && q < d.dim) ? rp[q] : 0.f;
      At[q * LDA + r] = v;
      if (kRowMajorToo) Ef[r * ldp + q] = v;
    }
  }
  if (threadIdx.x < BM) {
    const int r = threadIdx.x;
    int64_t orig = -1;
    if (r < rows) orig = d.row_index ? (int64_t)d.row_index[row0 + r] : row0 + r;
    s_orig[r] = orig;
    s_code[r] = orig >= 0 ? d.pix_code[orig] : 0;
    s_seg[r] = orig >= 0 ? (int)d.seg[orig] : -1;
  }
}

// ------------------------------------------------------------------------- forward

__global__ void __launch_bounds__(kGemmThreads)
segsort_fwd_kernel(spml_segsort_desc d, int dpad, float* __restrict__ stats,
                   float* __restrict__ nll_out, float* __restrict__ partial) {
  extern __shared__ __align__(16) float smem[];
  float* At = smem;
  float* Bt = At + (size_t)dpad * LDA;
  __shared__ int64_t s_code[BM], s_orig[BM], s_pcode[BN];
  __shared__ int s_seg[BM];
  __shared__ unsigned char s_pvalid[BN];
  __shared__ float s_nll[BM];

  const int g = blockIdx.y;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  TileInfo t;
  group_range(d, g, t);
  t.row0 = t.r_begin + (int64_t)blockIdx.x * BM;
  float* my_partial = partial + (size_t)g * gridDim.x + blockIdx.x;
  if (t.row0 >= t.r_end) {
    if (tid == 0) *my_partial = 0.f;
    return;
  }
  t.rows = (int)min((int64_t)BM, t.r_end - t.row0);
  load_emb_tile<false>(d, t.row0, t.rows, dpad, 0, At, nullptr, s_code, s_seg, s_orig);

  float same[TM], diff[TM], self[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) same[i] = diff[i] = self[i] = 0.f;

  for (int c0 = t.c_begin; c0 < t.c_end; c0 += BN) {
    __syncthreads();
    load_proto_tile<false>(d, c0, t.c_end, dpad, 0, Bt, nullptr, s_pcode, s_pvalid);
    __syncthreads();
    float acc[TM][TN];
    gemm_nt_tile(At, Bt, dpad, ty, tx, acc);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int k = tx * TN + j;
      if (!s_pvalid[k]) continue;
      const int64_t pc = s_pcode[k];
      const int c = c0 + k;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
        const float s = expf(d.kappa * acc[i][j]);
        if (codes_match(d.mode, s_code[r], pc)) same[i] += s; else diff[i] += s;
        if (c == s_seg[r]) self[i] += s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      same[i] += __shfl_xor_sync(0xffffffffu, same[i], o);
      diff[i] += __shfl_xor_sync(0xffffffffu, diff[i], o);
      self[i] += __shfl_xor_sync(0xffffffffu, self[i], o);
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = ty * TM + i;
      float nll = 0.f;
      if (r < t.rows) {
        // loss.py:64-80: the own segment is subtracted from the same-class sum, and
        // the pixel falls back to its own segment when nothing else is left
        const float others = same[i] - self[i];
        const bool pos = others > 0.f;
        const float num = pos ? others : self[i];
        const float den = diff[i] + num;
        nll = -logf(num / den);
        float* st = stats + (t.row0 + r) * 3;
        st[0] = num;
        st[1] = den;
        st[2] = pos ? 1.f : 0.f;
        if (nll_out) nll_out[t.row0 + r] = nll;
      }
      s_nll[r] = nll;
    }
  }
  __syncthreads();
  if (tid < 32) {  // fixed-order block sum
    float v = s_nll[tid] + s_nll[tid + 32] + s_nll[tid + 64] + s_nll[tid + 96];
    v = warp_sum(v);
    if (tid == 0) *my_partial = v;
  }
}

// loss[0] from the per-CTA partial sums, in a fixed order.
__global__ void segsort_loss_finalize_kernel(spml_segsort_desc d, const float* __restrict__ partial,
                                             int tiles_x, float* __restrict__ loss) {
  const float v = segsort_finalize_loss(d, partial, tiles_x);   // one warp
  if (threadIdx.x == 0) *loss = v;
}

// ------------------------------------------------------------------------- backward

// G_ij = coef_i S_ij ((diff_ij + numset_ij) / den_i - numset_ij / num_i)
struct RowGrad {
  float inv_num, inv_den, pos, coef;
};

__device__ __forceinline__ void load_row_grads(const spml_segsort_desc& d,
                                               const float* __restrict__ stats,
                                               const float* __restrict__ grad_loss,
                                               const float* __restrict__ grad_rows, int64_t row0,
                                               int rows, float weight, RowGrad* s_rg) {
  if (threadIdx.x < BM) {
    const int r = threadIdx.x;
    RowGrad rg{0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
      const float* st = stats + (row0 + r) * 3;
      rg.inv_num = 1.f / st[0];
      rg.inv_den = 1.f / st[1];
      rg.pos = st[2];
      rg.coef = d.kappa * (grad_rows ? grad_rows[row0 + r] : *grad_loss * weight);
    }
    s_rg[r] = rg;
  }
}

__device__ __forceinline__ float grad_entry(const spml_segsort_desc& d, const RowGrad& rg,
                                            float z, bool match, bool own) {
  const float s = expf(d.kappa * z);
  const float same = match ? 1.f : 0.f, self = own ? 1.f : 0.f;
  const float numset = rg.pos != 0.f ? same - self : self;
  const float diffv = 1.f - same;
  return rg.coef * s * ((diffv + numset) * rg.inv_den - numset * rg.inv_num);
}

template <int ND>
__global__ void __launch_bounds__(kGemmThreads)
segsort_bwd_emb_kernel(spml_segsort_desc d, int dpad, const float* __restrict__ stats,
                       const float* __restrict__ grad_loss, const float* __restrict__ grad_rows,
                       float beta, float* __restrict__ demb, int64_t ld_demb) {
  extern __shared__ __align__(16) float smem[];
  const int ldp = dpad + 4;
  float* At = smem;                        // [dpad][LDA]
  float* Bt = At + (size_t)dpad * LDA;     // [dpad][LDB]
  float* Pf = Bt + (size_t)dpad * LDB;     // [BN][ldp]
  float* Gs = Pf + (size_t)BN * ldp;       // [BN][LDA]   G^T tile: Gs[col][row]
  __shared__ int64_t s_code[BM], s_orig[BM], s_pcode[BN];
  __shared__ int s_seg[BM];
  __shared__ unsigned char s_pvalid[BN];
  __shared__ RowGrad s_rg[BM];

  const int g = blockIdx.y;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  TileInfo t;
  group_range(d, g, t);
  t.row0 = t.r_begin + (int64_t)blockIdx.x * BM;
  if (t.row0 >= t.r_end) return;
  t.rows = (int)min((int64_t)BM, t.r_end - t.row0);
  load_emb_tile<false>(d, t.row0, t.rows, dpad, 0, At, nullptr, s_code, s_seg, s_orig);
  load_row_grads(d, stats, grad_loss, grad_rows, t.row0, t.rows, reduction_weight(d, g), s_rg);

  float dacc[TM][4 * ND];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int q = 0; q < 4 * ND; ++q) dacc[i][q] = 0.f;

  for (int c0 = t.c_begin; c0 < t.c_end; c0 += BN) {
    __syncthreads();  // previous Bt / Pf / Gs fully consumed
    load_proto_tile<true>(d, c0, t.c_end, dpad, ldp, Bt, Pf, s_pcode, s_pvalid);
    __syncthreads();
    float acc[TM][TN];
    gemm_nt_tile(At, Bt, dpad, ty, tx, acc);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int k = tx * TN + j;
      const bool valid = s_pvalid[k];
      const int64_t pc = s_pcode[k];
      const int c = c0 + k;
      float gv[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
        gv[i] = valid ? grad_entry(d, s_rg[r], acc[i][j], codes_match(d.mode, s_code[r], pc),
                                   c == s_seg[r])
                      : 0.f;
      }
      float4* gp = reinterpret_cast<float4*>(Gs + k * LDA + ty * TM);
      gp[0] = make_float4(gv[0], gv[1], gv[2], gv[3]);
      gp[1] = make_float4(gv[4], gv[5], gv[6], gv[7]);
    }
    __syncthreads();
    // dE[row][q] += sum_k G[row][k] P[k][q]
#pragma unroll 2
    for (int k = 0; k < BN; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(Gs + k * LDA + ty * TM);
      const float4 a1 = *reinterpret_cast<const float4*>(Gs + k * LDA + ty * TM + 4);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        const int q0 = n * 64 + tx * 4;
        if (q0 < dpad) {
          const float4 b = *reinterpret_cast<const float4*>(Pf + k * ldp + q0);
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            dacc[i][n * 4 + 0] = fmaf(a[i], b.x, dacc[i][n * 4 + 0]);
            dacc[i][n * 4 + 1] = fmaf(a[i], b.y, dacc[i][n * 4 + 1]);
            dacc[i][n * 4 + 2] = fmaf(a[i], b.z, dacc[i][n * 4 + 2]);
            dacc[i][n * 4 + 3] = fmaf(a[i], b.w, dacc[i][n * 4 + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty * TM + i;
    if (r >= t.rows) continue;
    float* out = demb + s_orig[r] * ld_demb;
#pragma unroll
    for (int n = 0; n < ND; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = n * 64 + tx * 4 + j;
        if (q < d.dim) out[q] = beta != 0.f ? beta * out[q] + dacc[i][n * 4 + j] : dacc[i][n * 4 + j];
      }
  }
}

// the A tile of the prototype-gradient kernel doubles as its G tile [BM][LDB]
__host__ __device__ inline size_t proto_at_floats(int dpad) {
  const size_t a = (size_t)dpad * LDA, g = (size_t)BM * LDB;
  return a > g ? a : g;
}

template <int ND>
__global__ void __launch_bounds__(kGemmThreads)
segsort_bwd_proto_kernel(spml_segsort_desc d, int dpad, const float* __restrict__ stats,
                         const float* __restrict__ grad_loss, const float* __restrict__ grad_rows,
                         int64_t proto_rows, float* __restrict__ partial /* [chunks][proto_rows][dim] */) {
  extern __shared__ __align__(16) float smem[];
  const int ldp = dpad + 4;
  float* At = smem;                        // [dpad][LDA]; reused as Gs[BM][LDB] after GEMM 1
  float* Bt = At + proto_at_floats(dpad);  // [dpad][LDB]
  float* Ef = Bt + (size_t)dpad * LDB;     // [BM][ldp]
  float* Gs = At;
  __shared__ int64_t s_code[BM], s_orig[BM], s_pcode[BN];
  __shared__ int s_seg[BM];
  __shared__ unsigned char s_pvalid[BN];
  __shared__ RowGrad s_rg[BM];

  const int g = blockIdx.y, chunk = blockIdx.z, chunks = gridDim.z;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  TileInfo t;
  group_range(d, g, t);
  const int c0 = t.c_begin + blockIdx.x * BN;
  if (c0 >= t.c_end || c0 >= proto_rows) return;
  const int64_t tiles = (t.r_end - t.r_begin + BM - 1) / BM;
  const int64_t per = (tiles + chunks - 1) / chunks;
  const int64_t tile_lo = chunk * per, tile_hi = min(tiles, tile_lo + per);
  if (tile_lo >= tile_hi) return;  // partial is pre-zeroed
  const float weight = reduction_weight(d, g);

  load_proto_tile<false>(d, c0, t.c_end, dpad, 0, Bt, nullptr, s_pcode, s_pvalid);

  float pacc[4][4 * ND];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4 * ND; ++q) pacc[i][q] = 0.f;

  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    const int64_t row0 = t.r_begin + tile * BM;
    const int rows = (int)min((int64_t)BM, t.r_end - row0);
    __syncthreads();  // Gs (aliases At) and Ef of the previous tile fully consumed
    load_emb_tile<true>(d, row0, rows, dpad, ldp, At, Ef, s_code, s_seg, s_orig);
    load_row_grads(d, stats, grad_loss, grad_rows, row0, rows, weight, s_rg);
    __syncthreads();
    float acc[TM][TN];
    gemm_nt_tile(At, Bt, dpad, ty, tx, acc);
    __syncthreads();  // every thread is done reading At before it becomes Gs
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = ty * TM + i;
      float gv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int k = tx * TN + j;
        gv[j] = s_pvalid[k] ? grad_entry(d, s_rg[r], acc[i][j],
                                         codes_match(d.mode, s_code[r], s_pcode[k]),
                                         c0 + k == s_seg[r])
                            : 0.f;
      }
      *reinterpret_cast<float4*>(Gs + r * LDB + tx * TN) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
    __syncthreads();
    // dP[col][q] += sum_r G[r][col] E[r][q];  thread: cols ty*4..+3, q = n*64 + tx*4..+3
#pragma unroll 2
    for (int r = 0; r < BM; ++r) {
      const float4 a4 = *reinterpret_cast<const float4*>(Gs + r * LDB + ty * 4);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        const int q0 = n * 64 + tx * 4;
        if (q0 < dpad) {
          const float4 b = *reinterpret_cast<const float4*>(Ef + r * ldp + q0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pacc[i][n * 4 + 0] = fmaf(a[i], b.x, pacc[i][n * 4 + 0]);
            pacc[i][n * 4 + 1] = fmaf(a[i], b.y, pacc[i][n * 4 + 1]);
            pacc[i][n * 4 + 2] = fmaf(a[i], b.z, pacc[i][n * 4 + 2]);
            pacc[i][n * 4 + 3] = fmaf(a[i], b.w, pacc[i][n * 4 + 3]);
          }
        }
      }
    }
  }
  float* out = partial + (size_t)chunk * proto_rows * d.dim;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= t.c_end || c >= proto_rows) continue;
#pragma unroll
    for (int n = 0; n < ND; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = n * 64 + tx * 4 + j;
        if (q < d.dim) out[(size_t)c * d.dim + q] = pacc[i][n * 4 + j];
      }
  }
}

// ------------------------------------------------------------------------- host side

static int tiles_x_of(const spml_segsort_desc& d) {
  return (int)std::max<int64_t>(1, ceil_div(d.max_rows_per_group, BM));
}

static int proto_chunks_of(const spml_segsort_desc& d) {
  const int64_t col_tiles = std::max<int64_t>(1, ceil_div(d.m, BN));
  int64_t chunks = ceil_div(2 * 148, col_tiles * d.num_groups);
  chunks = std::min<int64_t>(chunks, tiles_x_of(d));
  return (int)std::max<int64_t>(1, chunks);
}

static int check_desc(const spml_segsort_desc* d, const char* who) {
  SPML_CHECK_ARG(d, "%s: null descriptor", who);
  SPML_CHECK_ARG(d->dim > 0 && d->num_groups >= 1 && d->n_rows >= 0 && d->m >= 0 &&
                     d->max_rows_per_group >= 0 && d->ld_emb >= d->dim && d->ld_protos >= d->dim,
                 "%s: bad sizes", who);
  SPML_CHECK_ARG(d->n_rows == 0 || d->m == 0 || (d->emb && d->pix_code && d->seg && d->protos &&
                                                d->proto_code),
                 "%s: null pointer", who);
  SPML_CHECK_ARG((d->group_off != nullptr) || d->num_groups == 1,
                 "%s: num_groups > 1 needs group_off", who);
  SPML_CHECK_SUPPORTED(d->dim <= SPML_MAX_DIM, "%s: dim %d exceeds %d", who, d->dim,
                       SPML_MAX_DIM);
  SPML_CHECK_SUPPORTED(d->num_groups <= 65535 && d->m < (1ll << 31) && d->n_rows < (1ll << 31),
                       "%s: problem too large", who);
  SPML_CHECK_SUPPORTED(d->mode == SPML_MODE_CLASS || d->mode == SPML_MODE_TAGS, "%s: bad mode",
                       who);
  SPML_CHECK_SUPPORTED(d->reduction >= SPML_REDUCE_MEAN && d->reduction <= SPML_REDUCE_SUM,
                       "%s: bad reduction", who);
  return SPML_OK;
}

// desc.reserved: bit 0 forces the fp32 CUDA-core path, bit 1 forces the tensor-core path
// (tests compare the two), bit 2 (backward only) says the workspace still holds the operands
// prepared by the forward call of the same problem, so the pre-pass is skipped; otherwise the tensor-core path runs whenever it supports the
// problem.  SPML_B200_SEGSORT=fp32|tc overrides the default.
static bool use_tc_path(const spml_segsort_desc& d) {
  if (!segsort_tc_supported(d)) return false;
  if (d.reserved & 1) return false;
  if (d.reserved & 2) return true;
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("SPML_B200_SEGSORT");
    env = !e ? 0 : (!strcmp(e, "fp32") ? 1 : (!strcmp(e, "tc") ? 2 : 0));
  }
  return env != 1;
}

template <typename Kernel>
static int set_smem(Kernel k, size_t bytes, const char* who) {
  if (bytes > 227 * 1024) {
    set_error("%s: needs %zu bytes of shared memory", who, bytes);
    return SPML_E_UNSUPPORTED;
  }
  SPML_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SPML_OK;
}

}  // namespace spml

extern "C" size_t spml_segsort_workspace_bytes(const spml_segsort_desc* d);

namespace spml {

// The forward without its last step: per-tile partial sums of the row losses are left in the
// workspace (`*partial_out`, [num_groups][*tiles_x_out]); segsort_finalize_loss turns them
// into the loss (internal.h).  The stage-group entry finishes its three losses in one kernel.
int segsort_fwd_partial(const spml_segsort_desc* d, float* stats, float* nll, void* workspace,
                        size_t workspace_bytes, cudaStream_t st, const float** partial_out,
                        int* tiles_x_out) {
  int rc = check_desc(d, "segsort_fwd");
  if (rc != SPML_OK) return rc;
  SPML_CHECK_ARG(workspace && (stats || d->n_rows == 0), "segsort_fwd: null pointer");
  if (workspace_bytes < spml_segsort_workspace_bytes(d)) {
    set_error("segsort_fwd: workspace %zu < %zu bytes", workspace_bytes,
              spml_segsort_workspace_bytes(d));
    return SPML_E_WORKSPACE;
  }
  const int dpad = pad4(d->dim);
  const int tiles_x = tiles_x_of(*d);
  *tiles_x_out = tiles_x;
  if (use_tc_path(*d)) {
    const TcPlan plan = segsort_tc_plan(*d, workspace);
    *partial_out = plan.partial;
    return segsort_fwd_tc(*d, plan, stats, nll, st);
  }
  float* partial = reinterpret_cast<float*>(workspace);
  *partial_out = partial;
  const size_t smem = (size_t)dpad * (LDA + LDB) * sizeof(float);
  rc = set_smem(segsort_fwd_kernel, smem, "segsort_fwd");
  if (rc != SPML_OK) return rc;
  dim3 grid((unsigned)tiles_x, (unsigned)d->num_groups);
  segsort_fwd_kernel<<<grid, kGemmThreads, smem, st>>>(*d, dpad, stats, nll, partial);
  SPML_LAUNCH_CHECK("segsort_fwd_kernel");
  return SPML_OK;
}

}  // namespace spml

namespace spml {

__global__ void reduce_chunks_kernel(const float* __restrict__ partial, int chunks, int64_t count,
                                     float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = 0.f;
  for (int c = 0; c < chunks; ++c) v += partial[(size_t)c * count + i];
  out[i] = v;
}

// out = sum of the chunks of up to two partial buffers (the stage-group backward adds the
// prototype gradients of sem_occ and sem_ann in one pass)
__global__ void reduce_two_kernel(const float* __restrict__ pa, int ca,
                                  const float* __restrict__ pb, int cb, int64_t count,
                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = 0.f;
  for (int c = 0; c < ca; ++c) v += pa[(size_t)c * count + i];
  for (int c = 0; c < cb; ++c) v += pb[(size_t)c * count + i];
  out[i] = v;
}

int segsort_reduce_two(const float* pa, int ca, const float* pb, int cb, int64_t count, float* out,
                       cudaStream_t st) {
  if (count <= 0) return SPML_OK;
  reduce_two_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(pa, ca, pb, cb, count, out);
  SPML_LAUNCH_CHECK("reduce_two_kernel");
  return SPML_OK;
}

// The backward behind spml_segsort_bwd.  `proto_rows` <= m limits d(prototypes) to the first
// proto_rows prototypes (the memory bank behind them is detached): dprotos is
// [proto_rows, dim].  With `partial_out` the per-chunk partial sums
// [*chunks_out][proto_rows][dim] are left in the workspace instead of being reduced into
// `dprotos` (then unused); *chunks_out = 0 when there is nothing to add.
int segsort_bwd_impl(const spml_segsort_desc* d, const float* stats, const float* grad_loss,
                     float beta, float* demb, int64_t ld_demb, float* dprotos, int64_t proto_rows,
                     const float** partial_out, int* chunks_out, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  int rc = check_desc(d, "segsort_bwd");
  if (rc != SPML_OK) return rc;
  const bool want_protos = dprotos || partial_out;
  SPML_CHECK_ARG(grad_loss && (stats || d->n_rows == 0) && (demb || want_protos),
                 "segsort_bwd: null pointer");
  SPML_CHECK_ARG(!demb || ld_demb >= d->dim, "segsort_bwd: bad ld_demb");
  SPML_CHECK_ARG(!partial_out || chunks_out, "segsort_bwd: null pointer");
  proto_rows = std::max<int64_t>(0, std::min<int64_t>(proto_rows, d->m));
  const int dpad = pad4(d->dim);
  const int ldp = dpad + 4;
  const int nd = (dpad + 63) / 64;
  const int tiles_x = tiles_x_of(*d);
  const int64_t count = proto_rows * d->dim;
  if (partial_out) {
    *partial_out = nullptr;
    *chunks_out = 0;
  }
  if (d->n_rows == 0 || d->m == 0 || d->max_rows_per_group == 0) {
    if (dprotos && !partial_out && count > 0)
      SPML_CUDA(cudaMemsetAsync(dprotos, 0, (size_t)count * sizeof(float), st));
    return SPML_OK;
  }
  const bool protos_now = want_protos && proto_rows > 0;

  if (use_tc_path(*d)) {
    if (!workspace || workspace_bytes < spml_segsort_workspace_bytes(d)) {
      set_error("segsort_bwd: workspace %zu < %zu bytes", workspace_bytes,
                spml_segsort_workspace_bytes(d));
      return SPML_E_WORKSPACE;
    }
    const TcPlan plan = segsort_tc_plan(*d, workspace);
    const int chunks = segsort_tc_proto_chunks(*d, proto_rows);
    float* partial = nullptr;
    if (protos_now) {
      partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + plan.bytes);
      SPML_CUDA(cudaMemsetAsync(partial, 0, (size_t)chunks * count * sizeof(float), st));
    }
    rc = segsort_bwd_tc(*d, plan, stats, grad_loss, beta, demb, ld_demb, partial, chunks,
                        proto_rows, (d->reserved & 4) != 0, st);
    if (rc != SPML_OK) return rc;
    if (protos_now) {
      if (partial_out) {
        *partial_out = partial;
        *chunks_out = chunks;
      } else {
        reduce_chunks_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(partial, chunks,
                                                                             count, dprotos);
        SPML_LAUNCH_CHECK("reduce_chunks_kernel");
      }
    }
    return SPML_OK;
  }

  if (demb) {
    const size_t smem = ((size_t)dpad * (LDA + LDB) + (size_t)BN * ldp + (size_t)BN * LDA) * sizeof(float);
    dim3 grid((unsigned)tiles_x, (unsigned)d->num_groups);
#define SPML_LAUNCH_EMB(NDV)                                                                   \
  do {                                                                                         \
    rc = set_smem(segsort_bwd_emb_kernel<NDV>, smem, "segsort_bwd(emb)");                       \
    if (rc != SPML_OK) return rc;                                                              \
    segsort_bwd_emb_kernel<NDV><<<grid, kGemmThreads, smem, st>>>(*d, dpad, stats, grad_loss,   \
                                                                  nullptr, beta, demb, ld_demb); \
  } while (0)
    if (nd == 1) SPML_LAUNCH_EMB(1); else if (nd == 2) SPML_LAUNCH_EMB(2); else SPML_LAUNCH_EMB(3);
#undef SPML_LAUNCH_EMB
    SPML_LAUNCH_CHECK("segsort_bwd_emb_kernel");
  }
  if (protos_now) {
    // the workspace holds proto_chunks_of(d) x m x dim floats: the rows are split finer when
    // fewer prototypes need a gradient, as long as the partials still fit
    int chunks = proto_chunks_of(*d);
    if (proto_rows < d->m) {
      const int64_t col_tiles = std::max<int64_t>(1, ceil_div(proto_rows, BN));
      const int64_t want =
          std::min<int64_t>(ceil_div(2 * 148, col_tiles * d->num_groups), tiles_x);
      chunks = (int)std::max<int64_t>(
          1, std::min<int64_t>(want, (int64_t)chunks * d->m / proto_rows));
    }
    const size_t need = (size_t)chunks * count * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("segsort_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
      return SPML_E_WORKSPACE;
    }
    float* partial = reinterpret_cast<float*>(workspace);
    SPML_CUDA(cudaMemsetAsync(partial, 0, need, st));
    const size_t smem =
        (proto_at_floats(dpad) + (size_t)dpad * LDB + (size_t)BM * ldp) * sizeof(float);
    dim3 grid((unsigned)ceil_div(proto_rows, BN), (unsigned)d->num_groups, (unsigned)chunks);
#define SPML_LAUNCH_PROTO(NDV)                                                                  \
  do {                                                                                          \
    rc = set_smem(segsort_bwd_proto_kernel<NDV>, smem, "segsort_bwd(protos)");                  \
    if (rc != SPML_OK) return rc;                                                               \
    segsort_bwd_proto_kernel<NDV><<<grid, kGemmThreads, smem, st>>>(                            \
        *d, dpad, stats, grad_loss, nullptr, proto_rows, partial);                              \
  } while (0)
    if (nd == 1) SPML_LAUNCH_PROTO(1); else if (nd == 2) SPML_LAUNCH_PROTO(2); else SPML_LAUNCH_PROTO(3);
#undef SPML_LAUNCH_PROTO
    SPML_LAUNCH_CHECK("segsort_bwd_proto_kernel");
    if (partial_out) {
      *partial_out = partial;
      *chunks_out = chunks;
    } else {
      reduce_chunks_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(partial, chunks, count,
                                                                           dprotos);
      SPML_LAUNCH_CHECK("reduce_chunks_kernel");
    }
  }
  return SPML_OK;
}

}  // namespace spml

extern "C" {

size_t spml_segsort_workspace_bytes(const spml_segsort_desc* d) {
  if (!d) return 0;
  const size_t fwd = (size_t)d->num_groups * spml::tiles_x_of(*d) * sizeof(float);
  const size_t bwd = (size_t)spml::proto_chunks_of(*d) * d->m * d->dim * sizeof(float);
  size_t need = 16 + (fwd > bwd ? fwd : bwd);
  if (spml::segsort_tc_supported(*d)) {
    const size_t tc = spml::segsort_tc_plan(*d, nullptr).bytes +
                      (size_t)spml::segsort_tc_proto_chunks(*d, d->m) * d->m * d->dim * sizeof(float);
    need = std::max(need, tc);
  }
  return need;
}

int spml_segsort_fwd(const spml_segsort_desc* d, float* stats, float* nll, float* loss,
                     void* workspace, size_t workspace_bytes, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(loss, "segsort_fwd: null pointer");
  const float* partial = nullptr;
  int tiles_x = 0;
  int rc = segsort_fwd_partial(d, stats, nll, workspace, workspace_bytes, as_stream(stream),
                               &partial, &tiles_x);
  if (rc != SPML_OK) return rc;
  segsort_loss_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(*d, partial, tiles_x, loss);
  SPML_LAUNCH_CHECK("segsort_loss_finalize_kernel");
  return SPML_OK;
}

int spml_segsort_bwd(const spml_segsort_desc* d, const float* stats, const float* grad_loss,
                     float beta, float* demb, int64_t ld_demb, float* dprotos, void* workspace,
                     size_t workspace_bytes, void* stream) {
  return spml::segsort_bwd_impl(d, stats, grad_loss, beta, demb, ld_demb, dprotos,
                                d ? d->m : 0, nullptr, nullptr, workspace, workspace_bytes,
                                spml::as_stream(stream));
}

int spml_segsort_bwd_rows(const spml_segsort_desc* d, const float* stats, const float* grad_loss,
                          float beta, float* demb, int64_t ld_demb, float* dprotos,
                          int64_t dprotos_rows, void* workspace, size_t workspace_bytes,
                          void* stream) {
  SPML_CHECK_ARG(!d || (dprotos_rows >= 0 && dprotos_rows <= d->m), "segsort_bwd: bad dprotos_rows");
  return spml::segsort_bwd_impl(d, stats, grad_loss, beta, demb, ld_demb, dprotos, dprotos_rows,
                                nullptr, nullptr, workspace, workspace_bytes,
                                spml::as_stream(stream));
}

}  // extern "C"
