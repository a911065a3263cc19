// fp32 CUDA-core tile GEMM shared by the k-means E-step and the SegSort kernels.
//
// A CTA of 256 threads owns a BM x BN = 128 x 64 tile of  Z = A . B^T  where A is a
// tile of embedding rows and B a tile of prototype rows, both staged in shared
// memory TRANSPOSED (k-major: At[d][row], Bt[d][col]) so that a thread reads its
// 8 rows / 4 columns with 128-bit shared loads.  Thread (ty, tx) = (tid / 16,
// tid % 16) accumulates rows ty*8..+7 x cols tx*4..+3 in registers.
#pragma once

#include "common.cuh"

namespace spml {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int kGemmThreads = 256;
constexpr int TM = 8;
constexpr int TN = 4;
constexpr int LDA = BM + 4;  // floats; rows stay 16-byte aligned
constexpr int LDB = BN + 4;

__host__ __device__ inline int pad4(int d) { return (d + 3) & ~3; }

constexpr int kMaxSlots = (SPML_MAX_DIM + 31) / 32;   // 32-channel slots per lane

// T[d * LD + r] = src[row(r)][d] for r < RMAX, d < dpad; zero for r >= rows or d >= dim.
// row(r) = row_index ? row_index[row0 + r] : row0 + r.  One warp per row, lanes across d
// (coalesced global reads); eight rows are fetched into registers before anything is
// stored so that the loads of a warp are all in flight together.
template <int RMAX, int LD>
__device__ __forceinline__ void load_rows_transposed(float* T, const float* __restrict__ src,
                                                     int64_t ld_src,
                                                     const int32_t* __restrict__ row_index,
                                                     int64_t row0, int rows, int dim, int dpad) {
  constexpr int kWarps = kGemmThreads / 32;
  constexpr int kBatch = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rb = warp; rb < RMAX; rb += kWarps * kBatch) {
    float v[kBatch][kMaxSlots];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int r = rb + i * kWarps;
      const float* rp = nullptr;
      if (r < rows) rp = src + (row_index ? (int64_t)row_index[row0 + r] : row0 + r) * ld_src;
#pragma unroll
      for (int s = 0; s < kMaxSlots; ++s) {
        const int d = lane + 32 * s;
        v[i][s] = (rp && d < dim) ? rp[d] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int r = rb + i * kWarps;
      if (r < RMAX) {
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
          const int d = lane + 32 * s;
          if (d < dpad) T[d * LD + r] = v[i][s];
        }
      }
    }
  }
}

// acc[i][j] = sum_d At[d][ty*8+i] * Bt[d][tx*4+j]
__device__ __forceinline__ void gemm_nt_tile(const float* At, const float* Bt, int dpad, int ty,
                                             int tx, float (&acc)[TM][TN]) {
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const float* ap = At + ty * TM;
  const float* bp = Bt + tx * TN;
#pragma unroll 4
  for (int d = 0; d < dpad; ++d) {
    const float4 a0 = *reinterpret_cast<const float4*>(ap + d * LDA);
    const float4 a1 = *reinterpret_cast<const float4*>(ap + d * LDA + 4);
    const float4 b = *reinterpret_cast<const float4*>(bp + d * LDB);
    const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bb[TN] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
  }
}

}  // namespace spml
