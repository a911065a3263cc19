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
  size_t bytes;
};

bool segsort_tc_supported(const spml_segsort_desc& d);
TcPlan segsort_tc_plan(const spml_segsort_desc& d, void* base);
int segsort_tc_prepare(const spml_segsort_desc& d, const TcPlan& p, cudaStream_t st);
int segsort_fwd_tc(const spml_segsort_desc& d, const TcPlan& p, float* stats, float* nll,
                   cudaStream_t st);

}  // namespace spml
