// A4 / B1: segment prototypes (reference spml/utils/segsort/common.py:11-41,
// spml/models/utils.py:113-116) and their backward.
//
// Forward: order-independent 64-bit fixed-point segment sums (run-length compressed
// atomics: consecutive rows of one segment are folded in registers before a single
// atomic per (run, channel)), then one warp per prototype normalises.
// Backward: one warp per pixel row gathers its segment's d(prototype) and applies the
// normalisation Jacobian on the fly (no [M, D] intermediate).
#include "common.cuh"

namespace spml {

constexpr int kSumThreads = 256;
constexpr int kSumRows = 64;  // rows per CTA

__global__ void __launch_bounds__(kSumThreads)
segment_sum_kernel(const float* __restrict__ x, int64_t rows_cap, const int32_t* rows_dev, int dim,
                   const int64_t* __restrict__ seg, int64_t m, long long* sums, int* poison) {
  const int64_t rows = rows_dev ? min(rows_cap, (int64_t)*rows_dev) : rows_cap;
  if ((int64_t)blockIdx.x * kSumRows >= rows) return;
  const int groups = max(1, kSumThreads / dim);
  const int g = threadIdx.x / dim, d = threadIdx.x % dim;
  if (g >= groups) return;
  const int64_t base = (int64_t)blockIdx.x * kSumRows;
  const int count = (int)min((int64_t)kSumRows, rows - base);
  const int per = (count + groups - 1) / groups;
  const int r0 = g * per, r1 = min(count, r0 + per);
  int64_t run_seg = -1;
  long long run = 0;
  for (int r = r0; r < r1; r += 4) {
    float v[4];
    int64_t s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = r + u < r1;
      v[u] = in ? x[(base + r + u) * dim + d] : 0.f;
      s[u] = in ? seg[base + r + u] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + u >= r1) break;
      if (s[u] != run_seg) {
        if (run_seg >= 0 && run_seg < m) atomic_add_i64(&sums[run_seg * dim + d], run);
        run = 0;
        run_seg = s[u];
      }
      run += to_fixed(v[u], poison);
    }
  }
  if (run_seg >= 0 && run_seg < m) atomic_add_i64(&sums[run_seg * dim + d], run);
}

__global__ void prototype_finalize_kernel(const long long* __restrict__ sums,
                                          const int* __restrict__ poison, int64_t m, int dim,
                                          float eps, float* __restrict__ protos,
                                          float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= m) return;
  const bool bad = *poison != 0;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const float v = from_fixed(sums[row * dim + d]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const bool ok = nrm >= eps;
  const float div = ok ? nrm : eps;
  const float nan = __int_as_float(0x7fc00000);
  if (eps >= kDivMinDivisor) {             // (kernel-uniform) the shared-reciprocal division
    const float r = div_reciprocal(div);
    for (int d = lane; d < dim; d += 32)
      protos[row * dim + d] = bad ? nan : div_by(from_fixed(sums[row * dim + d]), div, r);
  } else {
    for (int d = lane; d < dim; d += 32)
      protos[row * dim + d] = bad ? nan : from_fixed(sums[row * dim + d]) / div;
  }
  if (lane == 0 && norms) norms[row] = bad ? nan : (ok ? nrm : -eps);
}

__global__ void prototype_bwd_kernel(const float* __restrict__ dprotos,
                                     const float* __restrict__ protos,
                                     const float* __restrict__ norms,
                                     const int64_t* __restrict__ seg, int64_t rows_cap,
                                     const int32_t* rows_dev, int dim, int64_t m, float eps,
                                     float beta, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = rows_dev ? min(rows_cap, (int64_t)*rows_dev) : rows_cap;
  if (row >= rows) return;
  const int64_t s = seg[row];
  const bool in = s >= 0 && s < m;
  float t = 0.f;
  if (in)
    for (int d = lane; d < dim; d += 32) t += protos[s * dim + d] * dprotos[s * dim + d];
  t = warp_sum(t);
  const float nrm = in ? norms[s] : 1.f;
  for (int d = lane; d < dim; d += 32) {
    float g = 0.f;
    if (in) {
      const float dp = dprotos[s * dim + d];
      g = nrm > 0.f ? (dp - protos[s * dim + d] * t) / nrm : dp / eps;
    }
    float* o = dx + row * dim + d;
    *o = beta != 0.f ? beta * *o + g : g;
  }
}

}  // namespace spml

extern "C" {

size_t spml_segment_prototypes_workspace_bytes(int64_t m, int dim) {
  if (m <= 0 || dim <= 0) return 16;
  return 16 + (size_t)m * dim * sizeof(long long);
}

int spml_segment_prototypes_fwd(const float* x, int64_t rows, const int32_t* rows_dev, int dim,
                                const int64_t* seg, int64_t m, float eps, float* protos,
                                float* norms,
                                void* workspace, size_t workspace_bytes, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(rows >= 0 && m >= 0 && dim > 0, "segment_prototypes_fwd: bad sizes");
  if (m == 0) return SPML_OK;
  SPML_CHECK_ARG(protos && workspace && (rows == 0 || (x && seg)),
                 "segment_prototypes_fwd: null pointer");
  SPML_CHECK_SUPPORTED(dim <= kSumThreads, "segment_prototypes_fwd: dim %d exceeds %d", dim,
                       kSumThreads);
  const size_t need = spml_segment_prototypes_workspace_bytes(m, dim);
  if (workspace_bytes < need) {
    set_error("segment_prototypes_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  SPML_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  int* poison = reinterpret_cast<int*>(workspace);
  long long* sums = reinterpret_cast<long long*>(reinterpret_cast<char*>(workspace) + 16);
  if (rows > 0) {
    segment_sum_kernel<<<(unsigned)ceil_div(rows, kSumRows), kSumThreads, 0, st>>>(
        x, rows, rows_dev, dim, seg, m, sums, poison);
    SPML_LAUNCH_CHECK("segment_sum_kernel");
  }
  prototype_finalize_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, st>>>(sums, poison, m, dim, eps,
                                                                      protos, norms);
  SPML_LAUNCH_CHECK("prototype_finalize_kernel");
  return SPML_OK;
}

int spml_segment_prototypes_bwd(const float* dprotos, const float* protos, const float* norms,
                                const int64_t* seg, int64_t rows, const int32_t* rows_dev, int dim,
                                int64_t m, float eps, float beta, float* dx, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(rows >= 0 && m >= 0 && dim > 0, "segment_prototypes_bwd: bad sizes");
  if (rows == 0) return SPML_OK;
  SPML_CHECK_ARG(dprotos && protos && norms && seg && dx, "segment_prototypes_bwd: null pointer");
  prototype_bwd_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(
      dprotos, protos, norms, seg, rows, rows_dev, dim, m, eps, beta, dx);
  SPML_LAUNCH_CHECK("prototype_bwd_kernel");
  return SPML_OK;
}

}  // extern "C"
