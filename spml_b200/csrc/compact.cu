// A7 / A8 tail / B1: torch.unique(key, return_inverse=True) on the device, with the
// reference's numbering (rank of the key in ascending order; reference
// spml/utils/segsort/common.py:213-214,398-401, spml/models/utils.py:95-108).
//
// The number of distinct keys on this path is the number of segments (hundreds to a
// few thousand) while the number of keys is the number of pixels, so the keys are
// deduplicated through an open-addressing hash table (one 64-bit CAS per pixel) and
// only the distinct keys are ranked (all-pairs count, shared-memory tiled).
#include "common.cuh"

namespace spml {

constexpr unsigned long long kEmptyKey = 0x8080808080808080ull;  // memset(0x80) pattern

struct UniqueWs {
  long long* max_lo;           // [1] running max of lo (bound == 0)
  unsigned long long* table;   // [cap]
  int32_t* rank_of_slot;       // [cap]
  int32_t* slot_of;            // [n]
  int32_t* distinct;           // [n] slots of the distinct keys, arbitrary order
  size_t bytes;
  int64_t cap;
};

static int64_t table_capacity(int64_t n) {
  int64_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

static UniqueWs carve(void* base, int64_t n) {
  UniqueWs w{};
  w.cap = table_capacity(n);
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  w.max_lo = reinterpret_cast<long long*>(p + off);
  off += 16;
  w.table = reinterpret_cast<unsigned long long*>(p + off);
  off += (size_t)w.cap * 8;
  w.rank_of_slot = reinterpret_cast<int32_t*>(p + off);
  off += (size_t)w.cap * 4;
  w.slot_of = reinterpret_cast<int32_t*>(p + off);
  off += align_up((size_t)n * 4, 16);
  w.distinct = reinterpret_cast<int32_t*>(p + off);
  off += align_up((size_t)n * 4, 16);
  w.bytes = off;
  return w;
}

__global__ void max_lo_kernel(const int64_t* __restrict__ lo, int64_t n_cap,
                              const int32_t* n_dev, long long* out) {
  long long m = (long long)kEmptyKey;  // very negative
  const int64_t n = n_dev ? min(n_cap, (int64_t)*n_dev) : n_cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (long long)lo[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ long long resolve_bound(int64_t bound, const long long* max_lo) {
  return bound > 0 ? bound : *max_lo + 1;
}

__device__ __forceinline__ int64_t live_keys(int64_t n, const int32_t* n_dev) {
  return n_dev ? min(n, (int64_t)*n_dev) : n;
}

__global__ void unique_insert_kernel(const int64_t* __restrict__ hi, const int64_t* __restrict__ lo,
                                     int64_t n, const int32_t* n_dev, int64_t bound,
                                     const long long* max_lo,
                                     unsigned long long* table, int64_t cap_mask,
                                     int32_t* __restrict__ slot_of, int32_t* distinct,
                                     int32_t* count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= live_keys(n, n_dev)) return;
  long long key = lo[i];
  if (hi) key += hi[i] * resolve_bound(bound, max_lo);
  const unsigned long long ukey = (unsigned long long)key;
  int64_t h = (int64_t)(mix64(ukey) & (unsigned long long)cap_mask);
  while (true) {
    const unsigned long long prev = atomicCAS(&table[h], kEmptyKey, ukey);
    if (prev == kEmptyKey) {
      distinct[atomicAdd(count, 1)] = (int32_t)h;
      break;
    }
    if (prev == ukey) break;
    h = (h + 1) & cap_mask;
  }
  slot_of[i] = (int32_t)h;
}

// unique_insert_kernel for the tail of the clustering stage (pipeline.cu), with what used to be
// two more launches on the way to the step's host read-back folded in:
//  * the key is built on the fly: high part (image, k-means cluster), low part the pixel's label
//    (common.py:398-405), and the label is decoded into its semantic / instance parts
//    (resnet_deeplab.py:134-135: floor division / modulo);
//  * the LAST block to finish publishes {rows kept, segments, status, 0} - the number of distinct
//    keys is final then - into device memory and, with a sequence number behind a system-scope
//    fence, into pinned host memory that the host polls.  The ticket counter is the spare half of
//    the workspace header (filled with 0x80 by the same memset as the table).
__global__ void cluster_insert_kernel(const int32_t* __restrict__ km, const int64_t* __restrict__ batch,
                                      const int64_t* __restrict__ lo, int64_t n,
                                      const int32_t* __restrict__ n_dev, int64_t num_clusters,
                                      const long long* max_lo, unsigned long long* table,
                                      int64_t cap_mask, int32_t* __restrict__ slot_of,
                                      int32_t* distinct, int32_t* count, int64_t divisor,
                                      int64_t* __restrict__ sem_out, int64_t* __restrict__ inst_out,
                                      unsigned* ticket, int32_t* status, int32_t* counts_out,
                                      int32_t* host_out, int32_t seq) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < live_keys(n, n_dev)) {
    const long long lab = lo[i];
    const long long key = lab + ((long long)km[i] + batch[i] * num_clusters) * (*max_lo + 1);
    const unsigned long long ukey = (unsigned long long)key;
    int64_t h = (int64_t)(mix64(ukey) & (unsigned long long)cap_mask);
    while (true) {
      const unsigned long long prev = atomicCAS(&table[h], kEmptyKey, ukey);
      if (prev == kEmptyKey) {
        distinct[atomicAdd(count, 1)] = (int32_t)h;
        break;
      }
      if (prev == ukey) break;
      h = (h + 1) & cap_mask;
    }
    slot_of[i] = (int32_t)h;
    if (sem_out) {
      int64_t q = lab / divisor, m = lab % divisor;
      if (m != 0 && ((m < 0) != (divisor < 0))) --q, m += divisor;
      sem_out[i] = q;
      inst_out[i] = m;
    }
  }
  if (!counts_out) return;
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(ticket, 1u) != 0x80808080u + gridDim.x - 1u) return;
  __threadfence();
  const int32_t r = *reinterpret_cast<const volatile int32_t*>(n_dev);
  const int32_t m = *reinterpret_cast<volatile int32_t*>(count);
  const int32_t st = status ? atomicExch(status, 0) : 0;
  counts_out[0] = r, counts_out[1] = m, counts_out[2] = st, counts_out[3] = 0;
  if (host_out) {
    volatile int32_t* hp = host_out;
    hp[0] = r, hp[1] = m, hp[2] = st, hp[3] = 0;
    __threadfence_system();
    hp[4] = seq;
  }
}

// rank of every distinct key = number of distinct keys that are smaller (signed order)
__global__ void unique_rank_kernel(const unsigned long long* __restrict__ table,
                                   const int32_t* __restrict__ distinct,
                                   const int32_t* __restrict__ count, int has_hi, int64_t bound,
                                   const long long* max_lo, int32_t* __restrict__ rank_of_slot,
                                   int64_t* __restrict__ uniq_hi, int64_t* __restrict__ uniq_lo,
                                   int64_t* bound_out) {
  __shared__ long long s_keys[256];
  const int u_total = *count;
  if ((int64_t)blockIdx.x * blockDim.x >= u_total) return;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = u < u_total;
  const int slot = active ? distinct[u] : 0;
  const long long key = active ? (long long)table[slot] : 0;
  int rank = 0;
  for (int t0 = 0; t0 < u_total; t0 += 256) {
    const int v = t0 + threadIdx.x;
    __syncthreads();
    s_keys[threadIdx.x] = v < u_total ? (long long)table[distinct[v]] : 0;
    __syncthreads();
    const int cnt = min(256, u_total - t0);
    if (active)
      for (int j = 0; j < cnt; ++j) rank += s_keys[j] < key;
  }
  if (!active) return;
  rank_of_slot[slot] = rank;
  const long long bnd = resolve_bound(bound, max_lo);
  if (has_hi) {
    if (uniq_hi) uniq_hi[rank] = key / bnd;
    if (uniq_lo) uniq_lo[rank] = key % bnd;
  } else {
    if (uniq_hi) uniq_hi[rank] = 0;
    if (uniq_lo) uniq_lo[rank] = key;
  }
  if (u == 0 && bound_out) *bound_out = bnd;
}

__global__ void unique_inverse_kernel(const int32_t* __restrict__ slot_of,
                                      const int32_t* __restrict__ rank_of_slot, int64_t n,
                                      const int32_t* n_dev, int64_t* __restrict__ inverse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < live_keys(n, n_dev)) inverse[i] = rank_of_slot[slot_of[i]];
}

// ---- the three phases of spml_unique_inverse, separately, for a caller that overlaps them with
// other work (pipeline.cu: the preparation runs beside the k-means, the count of distinct keys
// is final after the insertion and goes to the host while the ranking still runs)

// count = 0, empty table, running maximum of lo (when the bound is to be derived from it)
int unique_prepare(bool has_hi, const int64_t* lo, int64_t n, const int32_t* n_dev, int64_t bound,
                   int32_t* count, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  SPML_CHECK_ARG(n >= 0 && count && bound >= 0, "unique_inverse: bad arguments");
  SPML_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
  if (n == 0) return SPML_OK;
  SPML_CHECK_ARG(lo && workspace, "unique_inverse: null pointer");
  SPML_CHECK_SUPPORTED(n < (1ll << 30), "unique_inverse: more than 2^30 keys");
  UniqueWs w = carve(workspace, n);
  if (workspace_bytes < w.bytes) {
    set_error("unique_inverse: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    return SPML_E_WORKSPACE;
  }
  // max_lo and the table share the 0x80 fill: "very negative" / "empty"
  SPML_CUDA(cudaMemsetAsync(workspace, 0x80, 16 + (size_t)w.cap * 8, st));
  if (bound == 0 && has_hi) {
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    max_lo_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, st>>>(lo, n, n_dev, w.max_lo);
    SPML_LAUNCH_CHECK("max_lo_kernel");
  }
  return SPML_OK;
}

// distinct keys into the table; *count is final when this kernel is
int unique_insert(const int64_t* hi, const int64_t* lo, int64_t n, const int32_t* n_dev,
                  int64_t bound, int32_t* count, void* workspace, cudaStream_t st) {
  UniqueWs w = carve(workspace, n);
  unique_insert_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
      hi, lo, n, n_dev, bound, w.max_lo, w.table, w.cap - 1, w.slot_of, w.distinct, count);
  SPML_LAUNCH_CHECK("unique_insert_kernel");
  return SPML_OK;
}

int cluster_insert(const int32_t* km, const int64_t* batch, const int64_t* labels, int64_t n,
                   const int32_t* n_dev, int64_t num_clusters, int32_t* count, int64_t divisor,
                   int64_t* sem_out, int64_t* inst_out, int32_t* status, int32_t* counts_out,
                   int32_t* host_out, int32_t seq, void* workspace, cudaStream_t st) {
  UniqueWs w = carve(workspace, n);
  cluster_insert_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
      km, batch, labels, n, n_dev, num_clusters, w.max_lo, w.table, w.cap - 1, w.slot_of, w.distinct,
      count, divisor, sem_out, inst_out, reinterpret_cast<unsigned*>(w.max_lo + 1), status,
      counts_out, host_out, seq);
  SPML_LAUNCH_CHECK("cluster_insert_kernel");
  return SPML_OK;
}

// rank of the distinct keys, inverse map
int unique_finish(bool has_hi, int64_t n, const int32_t* n_dev, int64_t bound, int64_t* inverse,
                  int64_t* uniq_hi, int64_t* uniq_lo, const int32_t* count, int64_t* bound_out,
                  void* workspace, cudaStream_t st) {
  UniqueWs w = carve(workspace, n);
  const unsigned blocks = (unsigned)ceil_div(n, 256);
  unique_rank_kernel<<<blocks, 256, 0, st>>>(w.table, w.distinct, count, has_hi, bound, w.max_lo,
                                             w.rank_of_slot, uniq_hi, uniq_lo, bound_out);
  SPML_LAUNCH_CHECK("unique_rank_kernel");
  unique_inverse_kernel<<<blocks, 256, 0, st>>>(w.slot_of, w.rank_of_slot, n, n_dev, inverse);
  SPML_LAUNCH_CHECK("unique_inverse_kernel");
  return SPML_OK;
}

}  // namespace spml

extern "C" {

size_t spml_unique_workspace_bytes(int64_t n) {
  if (n <= 0) return 16;
  return spml::carve(nullptr, n).bytes;
}

int spml_unique_inverse(const int64_t* hi, const int64_t* lo, int64_t n, const int32_t* n_dev,
                        int64_t bound, int64_t* inverse, int64_t* uniq_hi, int64_t* uniq_lo, int32_t* count,
                        int64_t* bound_out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  using namespace spml;
  cudaStream_t st = as_stream(stream);
  int rc = unique_prepare(hi != nullptr, lo, n, n_dev, bound, count, workspace, workspace_bytes, st);
  if (rc != SPML_OK || n == 0) return rc;
  SPML_CHECK_ARG(inverse, "unique_inverse: null pointer");
  if ((rc = unique_insert(hi, lo, n, n_dev, bound, count, workspace, st)) != SPML_OK) return rc;
  return unique_finish(hi != nullptr, n, n_dev, bound, inverse, uniq_hi, uniq_lo, count, bound_out,
                       workspace, st);
}

}  // extern "C"
