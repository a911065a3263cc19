// Error reporting and small row-wise utilities of libspml_b200.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace spml {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void clear_error() { g_error[0] = 0; }

static std::atomic<unsigned long long> g_launches{0};   // all threads (autograd runs backward on its own)
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t err, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(err));
  return SPML_E_CUDA;
}

// ---------------------------------------------------------------------------
// A1: one warp per row, lanes across the embedding dimension.
__global__ void normalize_rows_fwd_kernel(const float* __restrict__ x, int64_t rows, int dim,
                                          float eps, float* __restrict__ y,
                                          float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    float v = xr[d];
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const bool ok = nrm >= eps;
  const float div = ok ? nrm : eps;
  for (int d = lane; d < dim; d += 32) y[row * dim + d] = xr[d] / div;
  if (lane == 0 && norms) norms[row] = ok ? nrm : -eps;
}

__global__ void normalize_rows_bwd_kernel(const float* __restrict__ dy,
                                          const float* __restrict__ y,
                                          const float* __restrict__ norms, int64_t rows,
                                          int dim, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float nrm = norms[row];
  float t = 0.f;
  for (int d = lane; d < dim; d += 32) t += dy[row * dim + d] * y[row * dim + d];
  t = warp_sum(t);
  for (int d = lane; d < dim; d += 32) {
    const float g = dy[row * dim + d];
    dx[row * dim + d] = nrm > 0.f ? (g - y[row * dim + d] * t) / nrm : g / (-nrm);
  }
}

__global__ void pack_tags_kernel(const int64_t* __restrict__ tags, int64_t rows, int cols,
                                 int64_t ld, int64_t* __restrict__ masks) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  unsigned long long m = 0;
  for (int c = 0; c < cols; ++c)
    if (tags[r * ld + c] != 0) m |= (1ull << c);
  masks[r] = (int64_t)m;
}

__global__ void fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Per-pixel label decode and per-segment labels in one pass (fast path of
// spml/models/utils.py:100-111 for ids fresh from segment_by_kmeans).
__global__ void segment_labels_kernel(const int64_t* __restrict__ labels,
                                      const int64_t* __restrict__ batch,
                                      const int64_t* __restrict__ seg, int64_t cap,
                                      const int32_t* __restrict__ rows_dev, int64_t divisor,
                                      int64_t num_classes, int64_t m_cap,
                                      int64_t* __restrict__ sem, int64_t* __restrict__ inst,
                                      int64_t* __restrict__ keep, int64_t* __restrict__ p_sem,
                                      int64_t* __restrict__ p_inst, int64_t* __restrict__ p_batch,
                                      uint8_t* __restrict__ p_live, int32_t* overflow) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cap) return;
  const bool live = r < (int64_t)*rows_dev;
  const int64_t lab = live ? labels[r] : 0;
  const int64_t s = lab / divisor, i = lab % divisor;
  sem[r] = s;
  inst[r] = i;
  if (keep) keep[r] = live && s < num_classes;
  if (!live) return;
  const int64_t c = seg[r];
  if (c < 0 || c >= m_cap) {
    *overflow = 1;
    return;
  }
  // every pixel of a segment carries the same (image, label): plain stores of equal values
  p_sem[c] = s;
  p_inst[c] = i;
  p_batch[c] = batch[r];
  p_live[c] = 1;
}

}  // namespace spml

extern "C" {

const char* spml_last_error(void) { return spml::g_error; }

int spml_abi_version(void) { return 4; }

uint64_t spml_debug_launch_count(void) { return spml::g_launches.load(); }

int spml_normalize_rows_fwd(const float* x, int64_t rows, int dim, float eps, float* y,
                            float* norms_out, void* stream) {
  SPML_CHECK_ARG(x && y && rows >= 0 && dim > 0, "normalize_rows_fwd: bad arguments");
  if (rows == 0) return SPML_OK;
  const int warps = 8;
  spml::normalize_rows_fwd_kernel<<<(unsigned)spml::ceil_div(rows, warps), warps * 32, 0,
                                    spml::as_stream(stream)>>>(x, rows, dim, eps, y, norms_out);
  SPML_LAUNCH_CHECK("normalize_rows_fwd_kernel");
  return SPML_OK;
}

int spml_normalize_rows_bwd(const float* dy, const float* y, const float* norms, int64_t rows,
                            int dim, float* dx, void* stream) {
  SPML_CHECK_ARG(dy && y && norms && dx && rows >= 0 && dim > 0,
                 "normalize_rows_bwd: bad arguments");
  if (rows == 0) return SPML_OK;
  const int warps = 8;
  spml::normalize_rows_bwd_kernel<<<(unsigned)spml::ceil_div(rows, warps), warps * 32, 0,
                                    spml::as_stream(stream)>>>(dy, y, norms, rows, dim, dx);
  SPML_LAUNCH_CHECK("normalize_rows_bwd_kernel");
  return SPML_OK;
}

int spml_segment_labels(const int64_t* labels, const int64_t* batch, const int64_t* seg,
                        int64_t cap, const int32_t* rows_dev, int64_t divisor, int64_t num_classes,
                        int64_t m_cap, int64_t dead_label, int64_t* sem, int64_t* inst,
                        int64_t* keep, int64_t* p_sem, int64_t* p_inst, int64_t* p_batch,
                        uint8_t* p_live, int32_t* overflow, void* stream) {
  SPML_CHECK_ARG(labels && batch && seg && rows_dev && sem && inst && p_sem && p_inst && p_batch &&
                     p_live && overflow && cap >= 0 && m_cap >= 0 && divisor > 0,
                 "segment_labels: bad arguments");
  cudaStream_t st = spml::as_stream(stream);
  // dead segments: label `dead_label` (the caller's "unlabelled" id), image -1, not live
  SPML_CUDA(cudaMemsetAsync(p_live, 0, (size_t)m_cap, st));
  SPML_CUDA(cudaMemsetAsync(p_batch, 0xff, (size_t)m_cap * 8, st));
  SPML_CUDA(cudaMemsetAsync(p_inst, 0, (size_t)m_cap * 8, st));
  if (cap == 0) return SPML_OK;
  spml::fill_i64_kernel<<<(unsigned)spml::ceil_div(std::max<int64_t>(m_cap, 1), 256), 256, 0, st>>>(
      p_sem, m_cap, dead_label);
  SPML_LAUNCH_CHECK("fill_i64_kernel");
  spml::segment_labels_kernel<<<(unsigned)spml::ceil_div(cap, 256), 256, 0, st>>>(
      labels, batch, seg, cap, rows_dev, divisor, num_classes, m_cap, sem, inst, keep, p_sem,
      p_inst, p_batch, p_live, overflow);
  SPML_LAUNCH_CHECK("segment_labels_kernel");
  return SPML_OK;
}

int spml_pack_tags(const int64_t* tags, int64_t rows, int cols, int64_t ld, int64_t* masks,
                   void* stream) {
  SPML_CHECK_ARG(tags && masks && rows >= 0 && ld >= cols, "pack_tags: bad arguments");
  SPML_CHECK_SUPPORTED(cols >= 0 && cols <= 64, "pack_tags: at most 64 tag columns (got %d)",
                       cols);
  if (rows == 0) return SPML_OK;
  spml::pack_tags_kernel<<<(unsigned)spml::ceil_div(rows, 256), 256, 0,
                           spml::as_stream(stream)>>>(tags, rows, cols, ld, masks);
  SPML_LAUNCH_CHECK("pack_tags_kernel");
  return SPML_OK;
}

}  // extern "C"
