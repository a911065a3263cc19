// A4-A6: batched spherical k-means with initial labels
// (reference spml/utils/segsort/common.py:11-97).
//
// One launch = E-step against the prototypes of the previous launch fused with the
// accumulation of the next M-step, so every iteration reads the embeddings once:
// T iterations take T + 1 launches.  Segment sums are 64-bit fixed point
// (common.cuh) so the clustering is bit-reproducible whatever the tiling; the
// consumer normalises the sums when it stages the prototype tile.
#include <math.h>

#include "tile_gemm.cuh"

namespace spml {

struct KmeansStep {
  const float* x;            // [rows, dim]
  const int32_t* img_off;    // [batch + 1] or nullptr (single image of `rows_total` rows)
  int64_t rows_total;
  int dim, dpad;
  int num_clusters;          // stride of the per-image prototype arrays
  const int32_t* k_per_image;
  const long long* sums_in;  // [batch, K, dim] fixed point, or nullptr
  const float* protos_in;    // [K, dim] ready prototypes (A5), or nullptr
  long long* sums_out;       // [batch, K, dim] or nullptr
  int* poison;
  const int32_t* labels_in;  // used when there is no E-step (first launch)
  int32_t* labels_out;       // nullable
  int64_t* labels_out64;     // nullable
  float eps;
};

__global__ void __launch_bounds__(kGemmThreads) kmeans_step_kernel(KmeansStep p) {
  extern __shared__ __align__(16) float smem[];
  float* At = smem;                       // [dpad][LDA]
  float* Bt = At + (size_t)p.dpad * LDA;  // [dpad][LDB]
  __shared__ int s_lab[BM];

  const int b = blockIdx.y;
  const int64_t first = p.img_off ? p.img_off[b] : 0;
  const int64_t last = p.img_off ? p.img_off[b + 1] : p.rows_total;
  const int64_t row0 = first + (int64_t)blockIdx.x * BM;
  if (row0 >= last) return;
  const int rows = (int)min((int64_t)BM, last - row0);
  const int kb = p.k_per_image ? p.k_per_image[b] : p.num_clusters;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid / 16, tx = tid % 16;

  load_rows_transposed<BM, LDA>(At, p.x, p.dim, nullptr, row0, rows, p.dim, p.dpad);

  if (p.sums_in || p.protos_in) {
    float best_v[TM];
    int best_k[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) best_v[i] = -INFINITY, best_k[i] = 0;
    for (int k0 = 0; k0 < kb; k0 += BN) {
      const int kc = min(BN, kb - k0);
      __syncthreads();  // the previous prototype tile is no longer read
      for (int k = warp; k < BN; k += kGemmThreads / 32) {
        for (int d = lane; d < p.dpad; d += 32) {
          float v = 0.f;
          if (k < kc && d < p.dim) {
            const int64_t at = ((int64_t)b * p.num_clusters + k0 + k) * p.dim + d;
            v = p.sums_in ? from_fixed(p.sums_in[at]) : p.protos_in[(int64_t)(k0 + k) * p.dim + d];
          }
          Bt[d * LDB + k] = v;
        }
      }
      __syncthreads();
      if (p.sums_in) {  // prototype = sum / max(||sum||, eps)  (common.py:39)
        if (tid < BN) {
          float ss = 0.f;
          for (int d = 0; d < p.dpad; ++d) {
            const float v = Bt[d * LDB + tid];
            ss += v * v;
          }
          const float nrm = sqrtf(ss);
          const float div = nrm >= p.eps ? nrm : p.eps;
          for (int d = 0; d < p.dpad; ++d) Bt[d * LDB + tid] /= div;
        }
        __syncthreads();
      }
      float acc[TM][TN];
      gemm_nt_tile(At, Bt, p.dpad, ty, tx, acc);
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int k = k0 + tx * TN + j;
        if (k < kb) {
#pragma unroll
          for (int i = 0; i < TM; ++i)
            if (acc[i][j] > best_v[i]) best_v[i] = acc[i][j], best_k[i] = k;
        }
      }
    }
    // argmax across the 16 column lanes; ties keep the lowest index (torch.argmax)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_v[i], o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k[i], o);
        if (ov > best_v[i] || (ov == best_v[i] && ok < best_k[i])) best_v[i] = ov, best_k[i] = ok;
      }
      if (tx == 0) s_lab[ty * TM + i] = best_k[i];
    }
    __syncthreads();
    if (tid < rows) {
      if (p.labels_out) p.labels_out[row0 + tid] = s_lab[tid];
      if (p.labels_out64) p.labels_out64[row0 + tid] = s_lab[tid];
    }
  } else {
    if (tid < rows) s_lab[tid] = p.labels_in[row0 + tid];
    __syncthreads();
  }

  if (p.sums_out) {
    // Run-length segment sum: thread (group, d) walks its rows in order and flushes a
    // fixed-point run total whenever the label changes.
    const int groups = max(1, kGemmThreads / p.dim);
    const int g = tid / p.dim, d = tid % p.dim;
    if (g < groups) {
      const int per = (rows + groups - 1) / groups;
      const int r0 = g * per, r1 = min(rows, r0 + per);
      int run_lab = -1;
      long long run = 0;
      for (int r = r0; r < r1; ++r) {
        const int lab = s_lab[r];
        if (lab != run_lab) {
          if (run_lab >= 0)
            atomic_add_i64(&p.sums_out[((int64_t)b * p.num_clusters + run_lab) * p.dim + d], run);
          run = 0;
          run_lab = lab;
        }
        run += to_fixed(At[d * LDA + r], p.poison);
      }
      if (run_lab >= 0)
        atomic_add_i64(&p.sums_out[((int64_t)b * p.num_clusters + run_lab) * p.dim + d], run);
    }
  }
}

__global__ void copy_labels_kernel(const int32_t* in, int64_t n, int32_t* out, int64_t* out64) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (out) out[i] = in[i];
  if (out64) out64[i] = in[i];
}

static size_t kmeans_smem_bytes(int dpad) {
  return (size_t)dpad * (LDA + LDB) * sizeof(float);
}

static int launch_step(const KmeansStep& s, int batch, int max_rows_per_image,
                       cudaStream_t st) {
  const size_t smem = kmeans_smem_bytes(s.dpad);
  SPML_CUDA(cudaFuncSetAttribute(kmeans_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  dim3 grid((unsigned)ceil_div(max_rows_per_image, BM), (unsigned)batch);
  kmeans_step_kernel<<<grid, kGemmThreads, smem, st>>>(s);
  SPML_LAUNCH_CHECK("kmeans_step_kernel");
  return SPML_OK;
}

}  // namespace spml

extern "C" {

size_t spml_kmeans_workspace_bytes(int batch, int num_clusters, int dim, int iterations) {
  if (batch <= 0 || num_clusters <= 0 || dim <= 0 || iterations <= 0) return 16;
  return 16 + (size_t)iterations * batch * num_clusters * dim * sizeof(long long);
}

int spml_kmeans(const float* x, const int32_t* img_off, int batch, int max_rows_per_image,
                int dim, int num_clusters, const int32_t* k_per_image, int iterations,
                const int32_t* init_labels, int32_t* labels_out, int64_t* labels_out_i64,
                void* workspace, size_t workspace_bytes, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(x && img_off && init_labels && (labels_out || labels_out_i64) && batch > 0 &&
                     dim > 0 && num_clusters > 0 && iterations >= 0 && max_rows_per_image >= 0,
                 "kmeans: bad arguments");
  SPML_CHECK_SUPPORTED(dim <= SPML_MAX_DIM, "kmeans: dim %d exceeds %d", dim, SPML_MAX_DIM);
  SPML_CHECK_SUPPORTED(batch <= 65535, "kmeans: batch %d exceeds 65535", batch);
  cudaStream_t st = as_stream(stream);
  if (max_rows_per_image == 0) return SPML_OK;
  if (iterations == 0) {
    // without img_off[batch] on the host, copy the upper bound; rows beyond the
    // packed count are padding owned by the caller
    const int64_t n = (int64_t)batch * max_rows_per_image;
    copy_labels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(init_labels, n, labels_out,
                                                                    labels_out_i64);
    SPML_LAUNCH_CHECK("copy_labels_kernel");
    return SPML_OK;
  }
  const size_t need = spml_kmeans_workspace_bytes(batch, num_clusters, dim, iterations);
  if (!workspace || workspace_bytes < need) {
    set_error("kmeans: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  SPML_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  int* poison = reinterpret_cast<int*>(workspace);
  long long* sums = reinterpret_cast<long long*>(reinterpret_cast<char*>(workspace) + 16);
  const size_t per_iter = (size_t)batch * num_clusters * dim;

  KmeansStep s{};
  s.x = x;
  s.img_off = img_off;
  s.dim = dim;
  s.dpad = pad4(dim);
  s.num_clusters = num_clusters;
  s.k_per_image = k_per_image;
  s.poison = poison;
  s.eps = 1e-12f;
  for (int t = 0; t <= iterations; ++t) {
    s.sums_in = t > 0 ? sums + (size_t)(t - 1) * per_iter : nullptr;
    s.sums_out = t < iterations ? sums + (size_t)t * per_iter : nullptr;
    s.labels_in = t == 0 ? init_labels : nullptr;
    s.labels_out = t == iterations ? labels_out : nullptr;
    s.labels_out64 = t == iterations ? labels_out_i64 : nullptr;
    const int rc = launch_step(s, batch, max_rows_per_image, st);
    if (rc != SPML_OK) return rc;
  }
  return SPML_OK;
}

int spml_nearest_prototype(const float* x, int64_t rows, int dim, const float* protos,
                           int num_protos, int64_t* out, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(x && protos && out && rows >= 0 && dim > 0 && num_protos > 0,
                 "nearest_prototype: bad arguments");
  SPML_CHECK_SUPPORTED(dim <= SPML_MAX_DIM, "nearest_prototype: dim %d exceeds %d", dim,
                       SPML_MAX_DIM);
  SPML_CHECK_SUPPORTED(rows < (1ll << 31), "nearest_prototype: too many rows");
  if (rows == 0) return SPML_OK;
  KmeansStep s{};
  s.x = x;
  s.rows_total = rows;
  s.dim = dim;
  s.dpad = pad4(dim);
  s.num_clusters = num_protos;
  s.protos_in = protos;
  s.labels_out64 = out;
  s.eps = 1e-12f;
  return launch_step(s, 1, (int)rows, as_stream(stream));
}

}  // extern "C"
