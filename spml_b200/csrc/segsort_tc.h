// Internal interface between segsort.cu (dispatch, C ABI) and segsort_tc.cu (tcgen05 path).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace spml {

// workspace carve-up of the tensor-core path
struct TcPlan {
  int dp;       // embedding dim padded to a multiple of 8 (row pitch of the bf16 operands)
  int nkb;      // 64-wide K blocks
  int ksteps;   // 16-wide K steps
  int stages;   // depth of the prototype-tile ring
  float* partial;
  __nv_bfloat16 *eh, *el, *ph, *pl;
  int32_t *rcode, *rseg, *ccode;
  int32_t *col_dst, *col_src, *col_count;
  float4* pm;   // [n_rows] per-pixel gradient weights of the backward pass
  size_t bytes;
};

#ifdef __CUDACC__
// 1 / (what a row's nll is divided by); see SPML_REDUCE_* in the header.
__device__ inline float reduction_weight(const spml_segsort_desc& d, int g) {
  if (d.reduction == SPML_REDUCE_SUM) return 1.f;
  if (d.reduction == SPML_REDUCE_MEAN || !d.group_off) {
    const int64_t total =
        d.group_off ? (int64_t)d.group_off[d.num_groups] - d.group_off[0] : d.n_rows;
    return 1.f / (float)total;
  }
  int nonempty = 0;
  for (int q = 0; q < d.num_groups; ++q) nonempty += d.group_off[q + 1] > d.group_off[q];
  return 1.f / ((float)(d.group_off[g + 1] - d.group_off[g]) * (float)nonempty);
}
#endif

bool segsort_tc_supported(const spml_segsort_desc& d);
TcPlan segsort_tc_plan(const spml_segsort_desc& d, void* base);
int segsort_tc_prepare(const spml_segsort_desc& d, const TcPlan& p, cudaStream_t st);
int segsort_fwd_tc(const spml_segsort_desc& d, const TcPlan& p, float* stats, float* nll,
                   cudaStream_t st);
// backward: demb (nullable) and dprotos (nullable); `proto_partial` is the zeroed
// [chunks][proto_rows][dim] buffer the prototype-gradient CTAs write (only prototypes
// [0, proto_rows) get a gradient), reduced by the caller.
int segsort_tc_proto_chunks(const spml_segsort_desc& d, int64_t proto_rows);
int segsort_bwd_tc(const spml_segsort_desc& d, const TcPlan& p, const float* stats,
                   const float* grad_loss, float beta, float* demb, int64_t ld_demb,
                   float* proto_partial, int chunks, int64_t proto_rows, bool prepared,
                   cudaStream_t st);

}  // namespace spml
