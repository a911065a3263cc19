// A4-A6: batched spherical k-means with initial labels
// (reference spml/utils/segsort/common.py:11-97).
//
// The whole clustering (T iterations, every image of the batch) is ONE persistent
// cooperative launch: each CTA owns a fixed set of 128-pixel tiles and keeps its tile
// resident in shared memory when it has only one.  Iterations are not separated by
// kernel boundaries or a grid barrier but by per-image flags: the CTA that finishes the
// last tile of an image normalises that image's prototypes once and publishes them; the
// CTAs of the image pick them up for the next E-step.  An iteration is the
// reference's M-step (segment sums, L2-normalise) + E-step (argmax of x . P^T); the
// E-step of iteration t is fused with the accumulation of the M-step of t + 1, so
// the pixels are touched once per iteration.
//
// Segment sums are 64-bit fixed point (common.cuh): integer atomics are associative,
// so the clustering is bit-reproducible whatever the tiling or scheduling.  The
// consumer turns the sums into unit prototypes when it stages them.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kmeans.cuh"

namespace spml {

#ifdef SPML_KM_TRACE
__device__ long long g_km_trace[16 * 16];
#define KM_TRACE(slot)                                                        \
  do {                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0 && it < 16) g_km_trace[it * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define KM_TRACE(slot) do { } while (0)
#endif



// Bt[d][k] = unit prototype k0 + k (k < 64) from a [K, dim] array of ready prototypes.
// One warp per prototype, lanes across d; the eight prototypes of a warp are fetched
// before anything is stored.
__device__ __forceinline__ void stage_prototypes(const float* __restrict__ protos, int dim,
                                                 int dpad, int k0, int kc, float* Bt) {
  constexpr int kWarps = kGemmThreads / 32;
  constexpr int kCols = BN / kWarps;            // 8 prototypes per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v[kCols][kMaxSlots];
#pragma unroll
  for (int i = 0; i < kCols; ++i) {
    const int k = warp + i * kWarps;
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      const int d = lane + 32 * s;
      v[i][s] = (k < kc && d < dim) ? __ldcg(protos + (int64_t)(k0 + k) * dim + d) : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < kCols; ++i) {
    const int k = warp + i * kWarps;
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      const int d = lane + 32 * s;
      if (d < dpad) Bt[d * LDB + k] = v[i][s];
    }
  }
}

// The CTA that completed the last tile of image b turns the image's fixed-point sums into
// unit prototypes (common.py:39: sum / max(||sum||, eps); an empty cluster is the zero
// vector) and writes them to `out` [K, dim].  One warp per prototype, loads batched.
__device__ __forceinline__ void finalize_prototypes(const KmeansArgs& p, const long long* sums_b,
                                                    int kb, float* out) {
  constexpr int kWarps = kGemmThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int kbase = 0; kbase < kb; kbase += 4 * kWarps) {
    long long raw[4][kMaxSlots];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = kbase + warp + i * kWarps;
#pragma unroll
      for (int s = 0; s < kMaxSlots; ++s) {
        const int d = lane + 32 * s;
        raw[i][s] = (k < kb && d < p.dim) ? __ldcg(sums_b + (int64_t)k * p.dim + d) : 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = kbase + warp + i * kWarps;
      float v[kMaxSlots], ss = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxSlots; ++s) {
        v[s] = fixed_to_float(raw[i][s]);
        ss += v[s] * v[s];
      }
      const float nrm = sqrtf(warp_sum(ss));
      const float div = nrm >= p.eps ? nrm : p.eps;
      if (k < kb) {
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
          const int d = lane + 32 * s;
          if (d < p.dim) out[(int64_t)k * p.dim + d] = v[s] / div;
        }
      }
    }
  }
}

// E-step for the tile in At: s_lab[r] = argmax_k x_r . p_k, first index on ties.
__device__ __forceinline__ void assign_tile(const KmeansArgs& p, const float* protos_b, int kb,
                                            const float* At, float* Bt, int* s_lab) {
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float best_v[TM];
  int best_k[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) best_v[i] = -INFINITY, best_k[i] = 0;
  for (int k0 = 0; k0 < kb; k0 += BN) {
    const int kc = min(BN, kb - k0);
    __syncthreads();  // the previous prototype tile is no longer read
    stage_prototypes(protos_b, p.dim, p.dpad, k0, kc, Bt);
    __syncthreads();
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
}


// M-step accumulation for the tile: each warp walks 16 consecutive rows with its lanes
// across the channels and flushes a fixed-point run total whenever the label changes
// (neighbouring pixels mostly share a cluster, so there are few atomics).
__device__ __forceinline__ void accumulate_tile(const KmeansArgs& p, const Tile& tile,
                                                const float* At, const int* s_lab,
                                                long long* sums_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = BM / (kGemmThreads / 32);
  const int r0 = warp * per, r1 = min(tile.rows, r0 + per);
  // a value is round(x * 2^32) kept as hi * 2^16 + lo with two exact 32-bit integers
  // (runs are at most 16 rows long, so neither half can overflow)
  int run_hi[kAccSlots], run_lo[kAccSlots];
#pragma unroll
  for (int s = 0; s < kAccSlots; ++s) run_hi[s] = run_lo[s] = 0;
  int run_lab = -1;
  for (int r = r0; r < r1; ++r) {
    const int lab = s_lab[r];
    if (lab != run_lab) {
      if (run_lab >= 0) {
#pragma unroll
        for (int s = 0; s < kAccSlots; ++s) {
          const int d = lane + 32 * s;
          if (d < p.dim && (run_hi[s] | run_lo[s]) != 0)
            atomic_add_i64(&sums_b[(int64_t)run_lab * p.dim + d],
                           (long long)run_hi[s] * 65536ll + run_lo[s]);
          run_hi[s] = run_lo[s] = 0;
        }
      }
      run_lab = lab;
    }
    // the tile is already in shared memory (transposed): no second trip to L2
#pragma unroll
    for (int s = 0; s < kAccSlots; ++s) {
      const int d = lane + 32 * s;
      if (d < p.dim) {
        const float v = At[d * LDA + r];
        if (!(fabsf(v) <= 8.f)) *p.poison = 1;
        int hi, lo;
        split_fixed(v, hi, lo);
        run_hi[s] += hi;
        run_lo[s] += lo;
      }
    }
  }
  if (run_lab >= 0) {
#pragma unroll
    for (int s = 0; s < kAccSlots; ++s) {
      const int d = lane + 32 * s;
      if (d < p.dim && (run_hi[s] | run_lo[s]) != 0)
        atomic_add_i64(&sums_b[(int64_t)run_lab * p.dim + d],
                       (long long)run_hi[s] * 65536ll + run_lo[s]);
    }
  }
}

// Launched cooperatively (every CTA resident), grid <= number of tiles.
__global__ void __launch_bounds__(kGemmThreads) kmeans_persistent_kernel(KmeansArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* At = smem;                       // [dpad][LDA]
  float* Bt = At + (size_t)p.dpad * LDA;  // [dpad][LDB]
  __shared__ int s_lab[BM];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int total_tiles = p.batch * p.tiles_per_img;
  const bool resident = (int)gridDim.x >= total_tiles;   // one tile per CTA: load it once
  const size_t per_img = (size_t)p.num_clusters * p.dim;
  const size_t per_iter = (size_t)p.batch * per_img;

  for (int it = 0; it <= p.iterations; ++it) {
    KM_TRACE(0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      Tile tile;
      if (!tile_of(p, t, tile)) continue;
      const int b = tile.b;
      const int kb = p.k_per_image ? p.k_per_image[b] : p.num_clusters;
      if (!resident || it == 0) {
        __syncthreads();
        load_rows_transposed<BM, LDA>(At, p.x, p.dim, nullptr, tile.row0, tile.rows, p.dim,
                                      p.dpad);
      }
      if (it == 0) {
        if (tid < tile.rows) s_lab[tid] = p.labels_in[tile.row0 + tid];
        __syncthreads();
      } else {
        // the prototypes of iteration it - 1 of THIS image (no grid-wide barrier)
        if (tid == 0) {
          volatile unsigned* flag = p.ready + (size_t)(it - 1) * p.batch + b;
          while (*flag == 0) {
          }
          __threadfence();
        }
        __syncthreads();
        KM_TRACE(1);
        assign_tile(p, p.protos + (size_t)(it - 1) * per_iter + b * per_img, kb, At, Bt, s_lab);
        KM_TRACE(2);
        if (it == p.iterations && tid < tile.rows) {
          if (p.labels_out) p.labels_out[tile.row0 + tid] = s_lab[tid];
          if (p.labels_out64) p.labels_out64[tile.row0 + tid] = s_lab[tid];
        }
      }
      if (it < p.iterations) {
        long long* sums_b = p.sums + (size_t)it * per_iter + b * per_img;
        accumulate_tile(p, tile, At, s_lab, sums_b);
        KM_TRACE(3);
        // publish: the CTA that adds the image's last tile normalises its prototypes
        __syncthreads();
        if (tid == 0) {
          const int64_t rows_b = (int64_t)(p.img_off ? p.img_off[b + 1] - p.img_off[b]
                                                     : p.rows_total);
          const unsigned tiles_b = (unsigned)((rows_b + BM - 1) / BM);
          __threadfence();
          s_last = atomicAdd(p.done + (size_t)it * p.batch + b, 1u) + 1 == tiles_b;
        }
        __syncthreads();
        if (s_last) {
          __threadfence();
          finalize_prototypes(p, sums_b, kb, p.protos + (size_t)it * per_iter + b * per_img);
          __threadfence();
          __syncthreads();
          if (tid == 0) atomicExch(p.ready + (size_t)it * p.batch + b, 1u);
        }
        KM_TRACE(4);
      }
    }
  }
}

// A5 alone: one tile per CTA against ready prototypes.
__global__ void __launch_bounds__(kGemmThreads) nearest_prototype_kernel(KmeansArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* At = smem;
  float* Bt = At + (size_t)p.dpad * LDA;
  __shared__ int s_lab[BM];
  Tile tile;
  if (!tile_of(p, blockIdx.x, tile)) return;
  load_rows_transposed<BM, LDA>(At, p.x, p.dim, nullptr, tile.row0, tile.rows, p.dim, p.dpad);
  assign_tile(p, p.protos_in, p.num_clusters, At, Bt, s_lab);
  if (threadIdx.x < tile.rows) p.labels_out64[tile.row0 + threadIdx.x] = s_lab[threadIdx.x];
}

__global__ void copy_labels_kernel(const int32_t* in, int64_t n, int32_t* out, int64_t* out64) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (out) out[i] = in[i];
  if (out64) out64[i] = in[i];
}

// The M-step saw an input outside the fixed-point range (|x| > 8 or NaN): the sums are
// meaningless, so the ids are replaced by -1 (kmeans_small.cu does this in its last pass).
__global__ void kmeans_poison_kernel(const int* __restrict__ poison, int64_t n,
                                     int32_t* __restrict__ out, int64_t* __restrict__ out64) {
  if (!*poison) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (out) out[i] = -1;
  if (out64) out64[i] = -1;
}

static size_t kmeans_smem_bytes(int dpad) {
  return (size_t)dpad * (LDA + LDB) * sizeof(float);
}

}  // namespace spml

extern "C" {

// workspace: [poison | done, ready counters | fixed-point sums] (all zeroed) | unit prototypes
static size_t kmeans_zeroed_bytes(int batch, int num_clusters, int dim, int iterations,
                                  int replicas) {
  return 16 + spml::align_up((size_t)2 * iterations * batch * sizeof(unsigned), 16) +
         (size_t)iterations * replicas * batch * num_clusters * dim * sizeof(long long);
}

// offset of the split (bf16) prototypes of the tensor-core path: behind the zeroed block (sized
// for its replicated sums) and the fp32 prototypes
static size_t kmeans_split_offset(int batch, int num_clusters, int dim, int iterations) {
  return spml::align_up(
      kmeans_zeroed_bytes(batch, num_clusters, dim, iterations, spml::kKmReplicas) +
          (size_t)iterations * batch * num_clusters * dim * sizeof(float),
      256);
}

// Which kernel runs (all three return the same labels, bit for bit).
//  * K <= 128 (every shipped configuration): kmeans_small.cu, the tcgen05 E-step with the
//    prototypes rebuilt from the sums inside every CTA (no finalising CTA / flag / TMA chain).
//  * otherwise, measured on B200 (profiles/r1c_*): the tcgen05 kernel of kmeans_tc.cu wins when
//    the assignment GEMM is big (K >= 256: 1.2-2.8x) and whenever its second fp32 tile buffer
//    (cp.async prefetch) fits, i.e. up to ~96 channels; the corner "many tiles per SM, D > 96"
//    stays on the fp32 kernel, whose 2-4 co-resident CTAs per SM hide the per-tile latencies.
// SPML_B200_KMEANS=fp32|tc|small overrides (read per call so that tests can compare them).
//  * kmeans_cluster.cu: one thread-block cluster of up to 16 CTAs per image, everything in shared
//    memory, no global traffic between the passes.  Its time hardly depends on the batch (images
//    are independent clusters), while the small-K kernel serialises the tiles of a CTA.  Measured
//    (VOC shape, K = 36, us per call, profiles/r2b_kmeans_paths.txt): batch 1: 197 vs 94 (16 SMs
//    per image are issue-bound in the argmax epilogue and the DSMEM exchange pays per 16-byte
//    request), batch 2: 232 vs 161, batch 3: 231 vs 212, batch 4: 231 vs 266, batch 8: 404 vs 456,
//    batch 16: 636 vs 959; with K = 64 the two are level up to batch 8.  Default from batch 4 on
//    when K <= 48; SPML_B200_KMEANS=cluster forces it wherever it fits.
enum KmeansPath { kPathFp32, kPathTc, kPathSmall, kPathCluster };

static KmeansPath kmeans_path(int dim, int num_clusters, int batch, int max_rows, int64_t tiles,
                              int sms) {
  const bool small_ok = spml::kmeans_small_supported(dim, num_clusters, batch, tiles * spml::BM);
  const bool tc_ok = spml::kmeans_tc_supported(dim);
  const char* e = getenv("SPML_B200_KMEANS");
  const bool forced = e && *e;
  if (e && !strcmp(e, "fp32")) return kPathFp32;
  if (e && !strcmp(e, "tc") && tc_ok) return kPathTc;
  if (e && !strcmp(e, "small") && small_ok) return kPathSmall;
  if (((forced && !strcmp(e, "cluster")) || (!forced && batch >= 4 && num_clusters <= 48)) &&
      spml::kmeans_cluster_supported(dim, num_clusters, batch, max_rows))
    return kPathCluster;
  if (small_ok) return kPathSmall;
  if (!tc_ok) return kPathFp32;
  return (tiles <= sms || num_clusters >= 256 || dim <= 96) ? kPathTc : kPathFp32;
}

int spml_debug_kmeans_path(int batch, int max_rows_per_image, int dim, int num_clusters) {
  int device = 0, sms = 0;
  if (cudaGetDevice(&device) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return -1;
  return (int)kmeans_path(dim, num_clusters, batch, max_rows_per_image,
                          (int64_t)batch * spml::ceil_div(max_rows_per_image, spml::BM), sms);
}

size_t spml_kmeans_workspace_bytes(int batch, int num_clusters, int dim, int iterations) {
  if (batch <= 0 || num_clusters <= 0 || dim <= 0 || iterations <= 0) return 16;
  return kmeans_split_offset(batch, num_clusters, dim, iterations) +
         spml::kmeans_tc_split_bytes(batch, num_clusters, dim, iterations);
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
  int device = 0, sms = 0, per_sm = 0;
  SPML_CUDA(cudaGetDevice(&device));
  SPML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  KmeansPath path = kmeans_path(dim, num_clusters, batch, max_rows_per_image,
                                (int64_t)batch * ceil_div(max_rows_per_image, BM), sms);
  // (the cluster kernel keeps the fp32 prototypes of a pass in the prototype scratch, padded to
  // 16-byte blocks: the scratch of a single iteration has no room for the padding)
  if (path == kPathCluster && iterations < 2) path = kPathSmall;
  const bool use_tc = path == kPathTc;
  const int replicas = path == kPathTc ? kKmReplicas : 1;   // the small-K kernel pre-reduces per CTA
  // (the cluster kernel keeps its sums in shared memory: only the poison flag is cleared)
  const size_t zeroed = kmeans_zeroed_bytes(batch, num_clusters, dim, iterations, replicas);
  SPML_CUDA(cudaMemsetAsync(workspace, 0, path == kPathCluster ? 16 : zeroed, st));

  KmeansArgs p{};
  p.x = x;
  p.img_off = img_off;
  p.batch = batch;
  p.tiles_per_img = (int)ceil_div(max_rows_per_image, BM);
  p.dim = dim;
  p.dpad = pad4(dim);
  p.num_clusters = num_clusters;
  p.k_per_image = k_per_image;
  char* base = reinterpret_cast<char*>(workspace);
  const size_t counters = align_up((size_t)2 * iterations * batch * sizeof(unsigned), 16);
  p.poison = reinterpret_cast<int*>(base);
  p.done = reinterpret_cast<unsigned*>(base + 16);
  p.ready = p.done + (size_t)iterations * batch;
  p.sums = reinterpret_cast<long long*>(base + 16 + counters);
  p.replicas = replicas;
  p.protos = reinterpret_cast<float*>(base + zeroed);
  p.iterations = iterations;
  p.labels_in = init_labels;
  p.labels_out = labels_out;
  p.labels_out64 = labels_out_i64;
  p.eps = 1e-12f;

  if (path == kPathCluster) return kmeans_cluster_launch(p, max_rows_per_image, st);
  if (path == kPathSmall) return kmeans_small_launch(p, sms, st);
  const int64_t cap_rows = (int64_t)batch * max_rows_per_image;
  if (use_tc) {
    int rc = kmeans_tc_launch(p, base + kmeans_split_offset(batch, num_clusters, dim, iterations),
                              sms, st);
    if (rc != SPML_OK) return rc;
    kmeans_poison_kernel<<<(unsigned)ceil_div(cap_rows, 256), 256, 0, st>>>(
        p.poison, cap_rows, labels_out, labels_out_i64);
    SPML_LAUNCH_CHECK("kmeans_poison_kernel");
    return SPML_OK;
  }
  const size_t smem = kmeans_smem_bytes(p.dpad);
  SPML_CUDA(cudaFuncSetAttribute(kmeans_persistent_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SPML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmeans_persistent_kernel,
                                                          kGemmThreads, smem));
  SPML_CHECK_SUPPORTED(per_sm >= 1, "kmeans: kernel does not fit on an SM (dim %d)", dim);
  const int64_t tiles = (int64_t)batch * p.tiles_per_img;
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sms * per_sm);
  void* args[] = {&p};
  SPML_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kmeans_persistent_kernel),
                                        dim3(grid), dim3(kGemmThreads), args, smem, st));
  SPML_LAUNCH_CHECK("kmeans_persistent_kernel");
  kmeans_poison_kernel<<<(unsigned)ceil_div(cap_rows, 256), 256, 0, st>>>(
      p.poison, cap_rows, labels_out, labels_out_i64);
  SPML_LAUNCH_CHECK("kmeans_poison_kernel");
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
  KmeansArgs p{};
  p.x = x;
  p.rows_total = rows;
  p.batch = 1;
  p.tiles_per_img = (int)ceil_div(rows, BM);
  p.dim = dim;
  p.dpad = pad4(dim);
  p.num_clusters = num_protos;
  p.protos_in = protos;
  p.labels_out64 = out;
  p.eps = 1e-12f;
  const size_t smem = kmeans_smem_bytes(p.dpad);
  SPML_CUDA(cudaFuncSetAttribute(nearest_prototype_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nearest_prototype_kernel<<<(unsigned)p.tiles_per_img, kGemmThreads, smem, as_stream(stream)>>>(p);
  SPML_LAUNCH_CHECK("nearest_prototype_kernel");
  return SPML_OK;
}

}  // extern "C"

#ifdef SPML_KM_TRACE
extern "C" int spml_debug_km_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, spml::g_km_trace, sizeof(long long) * 16 * 16) == cudaSuccess
             ? 0 : -2;
}
#endif
