// Stage-group entry points: one C-ABI call per stage of the reference's training step
// (pyscripts/train/train.py:167-219), each enqueueing all of its kernels, so that the host
// side of the drop-in API issues a handful of calls per step instead of ~100.
//
//   spml_segment_by_kmeans        A8  spml/utils/segsort/common.py:270-408
//   spml_gather_prototypes_fwd/bwd B1  spml/models/utils.py:95-116 (ids fresh from A8)
//   spml_head_fwd/bwd             C4  spml/models/predictions/segsort*.py losses()
//
// Independent kernels of a stage (the three losses, the two prototype sets, the retrieval
// accuracy) run on a small per-thread pool of side streams that fork from and join back into
// the caller's stream, so the call is still "everything enqueued on `stream`" for the caller.
#include <algorithm>
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"

namespace spml {

// ------------------------------------------------------------------------- side streams

constexpr int kSideStreams = 3;
constexpr int kMaxDevices = 16;

struct StreamPool {
  bool ready;
  cudaStream_t side[kSideStreams];
  cudaEvent_t fork;
  cudaEvent_t join[kSideStreams];
  // count read-back of the clustering stage: the kernel stores {rows, segments, status, 0, seq}
  // straight into pinned host memory and the host polls `seq` (a 16-byte cudaMemcpyAsync into
  // pageable memory + synchronise costs ~10 us of copy-engine and driver latency on the one
  // host-device round trip of the step)
  cudaEvent_t counts_ev;      // (blocking variant: the count copy has landed)
  int32_t* host_counts;       // pinned, mapped: 8 words
  int32_t* host_counts_dev;   // the same words as the device sees them
  int32_t seq;
};

// One pool per (host thread, device): the reference drives every GPU from its own Python
// thread (lib/nn/parallel/data_parallel.py:105), so pools are never shared between threads
// and an event is only ever re-recorded by the thread that waits on it.
static thread_local StreamPool g_pools[kMaxDevices];
// (host thread, device) pairs that drive this library in the process: more than one means the
// reference's thread-per-GPU DataParallel, where the count read-back must not spin (below)
static std::atomic<int> g_pool_count{0};

static int get_pool(StreamPool** out) {
  int device = 0;
  SPML_CUDA(cudaGetDevice(&device));
  SPML_CHECK_SUPPORTED(device >= 0 && device < kMaxDevices, "device index %d not supported", device);
  StreamPool& p = g_pools[device];
  if (!p.ready) {
    int lo = 0, hi = 0;
    SPML_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < kSideStreams; ++i) {
      SPML_CUDA(cudaStreamCreateWithPriority(&p.side[i], cudaStreamNonBlocking, hi));
      SPML_CUDA(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
    }
    SPML_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    SPML_CUDA(cudaEventCreateWithFlags(&p.counts_ev, cudaEventDisableTiming));
    SPML_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p.host_counts), 8 * sizeof(int32_t),
                            cudaHostAllocMapped));
    for (int i = 0; i < 8; ++i) p.host_counts[i] = 0;
    SPML_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p.host_counts_dev), p.host_counts, 0));
    p.seq = 0;
    p.ready = true;
    g_pool_count.fetch_add(1, std::memory_order_relaxed);
  }
  *out = &p;
  return SPML_OK;
}

// side streams [0, count) start after everything enqueued on `main` so far
static int fork_streams(StreamPool& p, cudaStream_t main, int count) {
  SPML_CUDA(cudaEventRecord(p.fork, main));
  for (int i = 0; i < count; ++i) SPML_CUDA(cudaStreamWaitEvent(p.side[i], p.fork, 0));
  return SPML_OK;
}

// `waiter` continues after everything enqueued on side stream i so far
static int join_stream(StreamPool& p, int i, cudaStream_t waiter) {
  SPML_CUDA(cudaEventRecord(p.join[i], p.side[i]));
  SPML_CUDA(cudaStreamWaitEvent(waiter, p.join[i], 0));
  return SPML_OK;
}

#define SPML_TRY(call)              \
  do {                              \
    int rc__ = (call);              \
    if (rc__ != SPML_OK) return rc__; \
  } while (0)

// Carves 256-byte aligned regions out of a workspace.  With a null base (size queries) the
// regions still get distinct NON-NULL addresses that are never dereferenced: the plans test
// pointers for null-ness (proto_valid, row_index, ...) and must size themselves identically
// in both modes.
struct Carver {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>((base ? base : reinterpret_cast<char*>(size_t(1) << 20)) + off);
    off += align_up(std::max<size_t>(count, 1) * sizeof(T), 256);
    return p;
  }
  void* take_bytes(size_t bytes) { return take<char>(bytes); }
};

// ------------------------------------------------------------------------- A8 kernels

constexpr int64_t kDroppedLabel = 0x7fffffffffffffffll;

// resnet_deeplab.py:112-117 without the max-reduction: the label of a dropped pixel never
// leaves the clustering, so any value no kept pixel can have does.
__global__ void pack_labels_kernel(const int64_t* __restrict__ sem, const int64_t* __restrict__ inst,
                                   int64_t count, int64_t divisor, int64_t semantic_ignore,
                                   int64_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t s = sem[i];
  labels[i] = s == semantic_ignore ? kDroppedLabel : s * divisor + inst[i];
}

// ------------------------------------------------------------------------- B1 kernels

__global__ void scatter_segment_labels_kernel(const int64_t* __restrict__ seg,
                                              const int64_t* __restrict__ batch,
                                              const int64_t* __restrict__ sem,
                                              const int64_t* __restrict__ inst, int64_t rows,
                                              int64_t m, int64_t* __restrict__ p_sem,
                                              int64_t* __restrict__ p_inst,
                                              int64_t* __restrict__ p_batch, int32_t* status) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t s = seg[r];
  if (s < 0 || s >= m) {
    atomicOr(status, 2);
    return;
  }
  // every pixel of a segment carries the same labels (checked below), so the race is benign
  if (p_sem) p_sem[s] = sem[r];
  if (p_inst) p_inst[s] = inst[r];
  if (p_batch) p_batch[s] = batch[r];
}

__global__ void check_segment_labels_kernel(const int64_t* __restrict__ seg,
                                            const int64_t* __restrict__ batch,
                                            const int64_t* __restrict__ sem,
                                            const int64_t* __restrict__ inst, int64_t rows,
                                            int64_t m, const int64_t* __restrict__ p_sem,
                                            const int64_t* __restrict__ p_inst,
                                            const int64_t* __restrict__ p_batch, int32_t* status) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t s = seg[r];
  if (s < 0 || s >= m) return;
  if ((p_sem && p_sem[s] != sem[r]) || (p_inst && p_inst[s] != inst[r]) ||
      (p_batch && p_batch[s] != batch[r]))
    atomicOr(status, 1);
}

// ------------------------------------------------------------------------- C4 kernels

// dst[off + i] = src[i] for a list of segments (current prototypes + memory-bank entries)
struct ConcatList {
  const void* src[SPML_MAX_BANK + 1];
  int64_t count[SPML_MAX_BANK + 1];   // elements
  int64_t first[SPML_MAX_BANK + 2];   // prefix sums
  int n;
};

template <typename T>
__global__ void concat_kernel(ConcatList l, T* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= l.first[l.n]) return;
  int s = 0;
  while (s + 1 < l.n && i >= l.first[s + 1]) ++s;
  dst[i] = reinterpret_cast<const T*>(l.src[s])[i - l.first[s]];
}

// bit masks of tag rows: bit (c - col0) set iff tags[r, c] != 0 for c in [col0, col1)
struct TagList {
  const int64_t* src[SPML_MAX_BANK + 1];
  int64_t ld[SPML_MAX_BANK + 1];
  int64_t first[SPML_MAX_BANK + 2];   // rows
  int n;
};

__global__ void gather_mask_kernel(const int64_t* __restrict__ table, int64_t table_rows,
                                   const int64_t* __restrict__ index, int64_t rows,
                                   int64_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t i = index[r];
  out[r] = (i >= 0 && i < table_rows) ? table[i] : 0;
}

// labelled-prototype filter of sem_ann (segsort.py:184-194) and densepose's "no tag ->
// every tag" rule (segsort_softmax_densepose.py:186-189)
__global__ void proto_flags_kernel(const int64_t* __restrict__ psem, int64_t m, int64_t num_classes,
                                   uint8_t* __restrict__ valid, int64_t* __restrict__ masks,
                                   int fill_empty_masks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  valid[i] = psem[i] < num_classes ? 1 : 0;
  if (fill_empty_masks && masks[i] == 0)
    masks[i] = num_classes >= 64 ? -1ll : (int64_t)((1ull << num_classes) - 1ull);
}

// One pass over everything the losses need besides the embeddings: the prototype bank of the
// step (current prototypes + memory-bank entries, segsort.py:153-182) with its class labels,
// labelled-prototype flags (segsort.py:184-194) and image-tag masks, and the per-pixel tag
// masks (segsort.py:146-150).  Thread i serves element i of the bank matrix, prototype i and
// pixel i.
struct PrepArgs {
  ConcatList protos;      // floats
  ConcatList psem;        // int64
  TagList ptags;          // rows of tag matrices (VOC heads)
  const int64_t* img_tags;
  int64_t img_tags_ld, tag_rows;
  const int64_t* bid;
  int64_t n, m_all, num_classes;
  int dim, col0, col1, tag_mode;   // tag_mode 0: no masks, 1: image tags
};

__global__ void head_prep_kernel(PrepArgs a, float* __restrict__ p_all,
                                 int64_t* __restrict__ psem_all, uint8_t* __restrict__ pvalid,
                                 int64_t* __restrict__ pmask_all, int64_t* __restrict__ pix_mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.m_all * a.dim) {
    int s = 0;
    while (s + 1 < a.protos.n && i >= a.protos.first[s + 1]) ++s;
    p_all[i] = reinterpret_cast<const float*>(a.protos.src[s])[i - a.protos.first[s]];
  }
  if (i < a.m_all) {
    int s = 0;
    while (s + 1 < a.psem.n && i >= a.psem.first[s + 1]) ++s;
    const int64_t sem = reinterpret_cast<const int64_t*>(a.psem.src[s])[i - a.psem.first[s]];
    psem_all[i] = sem;
    pvalid[i] = sem < a.num_classes ? 1 : 0;
    unsigned long long m = 0;
    if (a.tag_mode == 1) {
      const int64_t* row = a.ptags.src[s] + (i - a.ptags.first[s]) * a.ptags.ld[s];
      for (int c = a.col0; c < a.col1; ++c)
        if (row[c] != 0) m |= 1ull << (c - a.col0);
    }
    pmask_all[i] = (int64_t)m;
  }
  if (a.tag_mode == 1 && i < a.n) {
    const int64_t img = a.bid[i];
    unsigned long long m = 0;
    if (img >= 0 && img < a.tag_rows) {
      const int64_t* row = a.img_tags + img * a.img_tags_ld;
      for (int c = a.col0; c < a.col1; ++c)
        if (row[c] != 0) m |= 1ull << (c - a.col0);
    }
    pix_mask[i] = (int64_t)m;
  }
}

// group g = image bid0 + g: first pixel row / first prototype column of each image (both
// lists are sorted by image); one thread per boundary, binary search.
__global__ void group_offsets_kernel(const int64_t* __restrict__ bid, int64_t rows,
                                     const int64_t* __restrict__ pbid, int64_t m, int groups,
                                     int32_t* __restrict__ group_off, int32_t* __restrict__ col_off,
                                     int32_t* status) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > groups) return;
  const int64_t base = rows > 0 ? bid[0] : 0;
  // more images than the caller's bound: the last group would mix images
  if (g == 0 && status && rows > 0 && bid[rows - 1] - base >= groups) atomicOr(status, 4);
  const int64_t want = base + g;   // first index whose image id is >= want
  for (int pass = 0; pass < 2; ++pass) {
    const int64_t* v = pass == 0 ? bid : pbid;
    const int64_t n = pass == 0 ? rows : m;
    int64_t lo = 0, hi = n;
    // rows: the last group takes everything that is left.  Columns: the bank may continue with
    // the prototypes of other ranks' images (cross-GPU exchange), which belong to no group.
    if (g == groups && pass == 0) lo = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (v[mid] < want) lo = mid + 1; else hi = mid;
    }
    (pass == 0 ? group_off : col_off)[g] = (int32_t)lo;
  }
}

// the three losses from their per-tile partial sums (one warp each), weights, accuracy
struct FinishArgs {
  spml_segsort_desc desc[3];
  const float* partial[3];
  int tiles_x[3];
};

__global__ void head_finish_kernel(FinishArgs f, const int32_t* __restrict__ hits, float w_ann,
                                   float w_occ, float w_sim, int k, unsigned enable,
                                   float* __restrict__ out) {
  __shared__ float s_loss[3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 3) {
    const float w = warp == 0 ? w_ann : (warp == 1 ? w_occ : w_sim);
    float v = 0.f;
    if (enable & (1u << warp)) v = segsort_finalize_loss(f.desc[warp], f.partial[warp], f.tiles_x[warp]) * w;
    if (lane == 0) out[warp] = v, s_loss[warp] = v;
  } else if (lane == 0) {
    out[3] = (enable & 8u) ? (float)hits[0] / ((float)hits[1] * (float)k) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // Python's sum([...]) over the enabled losses: ((0 + a) + b) + c, 0 + a being exact
    float total = 0.f;
    bool first = true;
    for (int i = 0; i < 3; ++i)
      if (enable & (1u << i)) {
        total = first ? s_loss[i] : total + s_loss[i];
        first = false;
      }
    out[4] = total;
  }
}

__global__ void scale_grads_kernel(const float* g_ann, const float* g_occ, const float* g_sim,
                                   const float* g_total, float w_ann, float w_occ, float w_sim,
                                   float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float t = g_total ? *g_total : 0.f;
  out[0] = ((g_ann ? *g_ann : 0.f) + t) * w_ann;
  out[1] = ((g_occ ? *g_occ : 0.f) + t) * w_occ;
  out[2] = ((g_sim ? *g_sim : 0.f) + t) * w_sim;
}

__global__ void add_rows_kernel(const float* __restrict__ a, int64_t count, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] += a[i];
}

// ------------------------------------------------------------------------- C4 plan

constexpr unsigned kAnn = 1u, kOcc = 2u, kSim = 4u, kAcc = 8u;

struct HeadPlan {
  int64_t m_all;
  float* p_all;           // [m_all, dim] current + bank prototypes
  float* pl_all;          // [m_all, dim_loc] (densepose tag propagation only)
  int64_t* psem_all;      // [m_all]
  int64_t* pbid_all;      // [m_all] (densepose)
  int64_t* pmask_all;     // [m_all] tag bit masks
  uint8_t* pvalid_ann;    // [m_all] psem < num_classes
  int64_t* img_mask;      // [tag_rows]
  int64_t* pix_mask;      // [n]
  int64_t* pinst;         // [m] when the caller has no prototype_instance_label
  int32_t* ann_dst;       // [n] scan scratch
  int32_t* ann_rows;      // [n] labelled pixel rows
  int32_t* ann_off;       // [2]
  int32_t* group_off;     // [groups + 1]
  int32_t* col_off;       // [groups + 1]
  float* sim_protos;      // [m, dim_sim] per-image prototypes of the img_sim embeddings
  float* sim_norms;       // [m]
  float* d_sim_protos;    // [m, dim_sim] (backward)
  float* stats[3];        // [n, 3] each
  float* raw;             // [4] unweighted losses
  float* gw;              // [4] weighted incoming gradients (backward)
  int32_t* hits;          // [2]
  int64_t* topk_labels;   // [m_all, 5]
  void* scan_ws;
  size_t scan_ws_bytes;
  void* proto_ws;
  size_t proto_ws_bytes;
  void* nn_ws;
  size_t nn_ws_bytes;
  void* seg_ws[3];
  size_t seg_ws_bytes[3];
  spml_segsort_desc desc[3];
  size_t bytes;
};

static int head_dim_sim(const spml_head_args& a) { return a.img_sim_on_plain ? a.dim : a.dim_loc; }

static HeadPlan head_plan(const spml_head_args& a, void* base) {
  HeadPlan p{};
  Carver c{reinterpret_cast<char*>(base), 0};
  int64_t m_all = a.m;
  for (int i = 0; i < a.num_bank; ++i) m_all += a.bank_m[i];
  p.m_all = m_all;
  const int64_t n = a.n;
  const int dsim = head_dim_sim(a);
  p.p_all = c.take<float>((size_t)m_all * a.dim);
  p.pl_all = c.take<float>(a.nn_tags ? (size_t)m_all * a.dim_loc : 1);
  p.psem_all = c.take<int64_t>(m_all);
  p.pbid_all = c.take<int64_t>(m_all);
  p.pmask_all = c.take<int64_t>(m_all);
  p.pvalid_ann = c.take<uint8_t>(m_all);
  p.img_mask = c.take<int64_t>(a.tag_rows);
  p.pix_mask = c.take<int64_t>(n);
  p.pinst = c.take<int64_t>(a.m);
  p.ann_dst = c.take<int32_t>(n);
  p.ann_rows = c.take<int32_t>(n);
  p.ann_off = c.take<int32_t>(4);
  p.group_off = c.take<int32_t>(a.max_groups + 1);
  p.col_off = c.take<int32_t>(a.max_groups + 1);
  p.sim_protos = c.take<float>((size_t)a.m * dsim);
  p.sim_norms = c.take<float>(a.m);
  p.d_sim_protos = c.take<float>((size_t)a.m * dsim);
  for (int i = 0; i < 3; ++i) p.stats[i] = c.take<float>((size_t)n * 3);
  p.raw = c.take<float>(4);
  p.gw = c.take<float>(4);
  p.hits = c.take<int32_t>(4);
  p.topk_labels = c.take<int64_t>((size_t)m_all * 8);
  p.scan_ws_bytes = spml_valid_scan_workspace_bytes(1, (int)std::max<int64_t>(n, 1));
  p.scan_ws = c.take_bytes(p.scan_ws_bytes);
  p.proto_ws_bytes = spml_segment_prototypes_workspace_bytes(a.m, dsim);
  p.proto_ws = c.take_bytes(p.proto_ws_bytes);
  p.nn_ws_bytes = a.nn_tags ? spml_nn_multiset_labels_workspace_bytes(m_all, 1) : 16;
  p.nn_ws = c.take_bytes(p.nn_ws_bytes);

  // sem_ann: labelled pixels x labelled prototypes (segsort.py:184-201)
  spml_segsort_desc& ann = p.desc[0];
  ann.emb = a.e; ann.ld_emb = a.dim; ann.dim = a.dim; ann.num_groups = 1;
  ann.row_index = p.ann_rows; ann.group_off = p.ann_off; ann.col_off = nullptr;
  ann.n_rows = n; ann.max_rows_per_group = n;
  ann.pix_code = a.sem; ann.seg = a.seg;
  ann.protos = p.p_all; ann.ld_protos = a.dim; ann.m = m_all;
  ann.proto_code = p.psem_all; ann.proto_valid = p.pvalid_ann;
  ann.kappa = a.kappa_ann; ann.mode = SPML_MODE_CLASS; ann.reduction = SPML_REDUCE_MEAN;
  ann.reserved = 0;
  // sem_occ: all pixels x all prototypes, tag-set masks (segsort.py:204-210)
  spml_segsort_desc& occ = p.desc[1];
  occ = ann;
  occ.row_index = nullptr; occ.group_off = nullptr;
  occ.pix_code = p.pix_mask; occ.proto_code = p.pmask_all; occ.proto_valid = nullptr;
  occ.kappa = a.kappa_occ; occ.mode = SPML_MODE_TAGS;
  occ.reserved = a.wide_tags ? 1 : 0;      // > 32 tag bits: the tcgen05 path compares 32-bit codes
  // img_sim: per image, pixels x the image's own prototypes (segsort.py:220-240)
  spml_segsort_desc& sim = p.desc[2];
  sim = ann;
  sim.emb = a.img_sim_on_plain ? a.e : a.el; sim.ld_emb = dsim; sim.dim = dsim;
  sim.num_groups = a.max_groups; sim.row_index = nullptr;
  sim.group_off = p.group_off; sim.col_off = p.col_off;
  sim.max_rows_per_group = std::min<int64_t>(n, a.max_rows_per_group > 0 ? a.max_rows_per_group : n);
  sim.pix_code = a.inst; sim.protos = p.sim_protos; sim.ld_protos = dsim; sim.m = a.m;
  sim.proto_code = a.pinst ? a.pinst : p.pinst; sim.proto_valid = nullptr;
  sim.kappa = a.kappa_sim; sim.reduction = SPML_REDUCE_GROUP_MEAN;
  for (int i = 0; i < 3; ++i) {
    p.seg_ws_bytes[i] = spml_segsort_workspace_bytes(&p.desc[i]);
    p.seg_ws[i] = c.take_bytes(p.seg_ws_bytes[i]);
  }
  p.bytes = c.off;
  return p;
}

static int check_head(const spml_head_args* a, const char* who) {
  SPML_CHECK_ARG(a, "%s: null arguments", who);
  SPML_CHECK_ARG(a->n > 0 && a->m > 0 && a->dim > 0 && a->num_bank >= 0 &&
                     a->num_bank <= SPML_MAX_BANK && a->max_groups >= 1,
                 "%s: bad sizes", who);
  SPML_CHECK_ARG(a->e && a->seg && a->protos && a->psem, "%s: null pointer", who);
  SPML_CHECK_ARG(!(a->enable & kAnn) || a->sem, "%s: sem_ann needs the pixel labels", who);
  SPML_CHECK_ARG(!(a->enable & kSim) || (a->inst && a->bid && a->pbid &&
                                        (a->img_sim_on_plain || (a->el && a->dim_loc > 0))),
                 "%s: img_sim needs instance labels, batch indices and its embeddings", who);
  if (a->enable & kOcc) {
    if (a->nn_tags)
      SPML_CHECK_ARG(a->protos_loc && a->pbid && a->dim_loc > 0,
                     "%s: tag propagation needs prototype_with_loc and batch indices", who);
    else
      SPML_CHECK_ARG(a->img_tags && a->ptags && a->bid && a->tag_rows > 0 &&
                         a->tag_col1 > a->tag_col0 && a->tag_col1 - a->tag_col0 <= 64,
                     "%s: sem_occ needs image tags (<= 64 columns)", who);
  }
  for (int i = 0; i < a->num_bank; ++i) {
    SPML_CHECK_ARG(a->bank_m[i] > 0 && a->bank_protos[i] && a->bank_psem[i],
                   "%s: bad memory-bank entry %d", who, i);
    if ((a->enable & kOcc) && !a->nn_tags)
      SPML_CHECK_ARG(a->bank_tags[i], "%s: memory-bank entry %d has no tags", who, i);
    if ((a->enable & kOcc) && a->nn_tags)
      SPML_CHECK_ARG(a->bank_protos_loc[i] && a->bank_pbid[i],
                     "%s: memory-bank entry %d has no prototype_with_loc / batch index", who, i);
  }
  return SPML_OK;
}

static unsigned blocks_of(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div(n, 256)); }

}  // namespace spml

extern "C" {

// sizes of the argument structs, for bindings to check their own layout against
size_t spml_sizeof_struct(int which) {
  switch (which) {
    case 0: return sizeof(spml_segsort_desc);
    case 1: return sizeof(spml_cluster_args);
    case 2: return sizeof(spml_head_args);
    default: return 0;
  }
}

// =========================================================================== A8

size_t spml_segment_by_kmeans_workspace_bytes(int batch, int n, int dim_total, int num_clusters,
                                              int iterations) {
  using namespace spml;
  if (batch <= 0 || n <= 0) return 256;
  Carver c{nullptr, 0};
  c.take_bytes(spml_valid_scan_workspace_bytes(batch, n));
  c.take_bytes(spml_kmeans_workspace_bytes(batch, num_clusters, dim_total, std::max(iterations, 1)));
  c.take_bytes(spml_unique_workspace_bytes((int64_t)batch * n));
  c.take<int64_t>((size_t)batch * n);
  c.take<int64_t>((size_t)batch * n);
  return c.off;
}

int spml_segment_by_kmeans(const spml_cluster_args* a, void* workspace, size_t workspace_bytes,
                           void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(a && workspace, "segment_by_kmeans: null pointer");
  SPML_CHECK_ARG(a->batch > 0 && a->n > 0 && a->dim > 0 && a->num_clusters > 0 &&
                     a->iterations >= 0 && a->loc_ch >= 0,
                 "segment_by_kmeans: bad sizes");
  SPML_CHECK_ARG(a->labels || (a->sem && a->inst && a->label_divisor > 0),
                 "segment_by_kmeans: labels, or semantic + instance labels and a divisor");
  SPML_CHECK_ARG((!a->sem_out && !a->inst_out) || (a->sem_out && a->inst_out && a->label_divisor > 0),
                 "segment_by_kmeans: sem_out / inst_out come together and need the divisor");
  SPML_CHECK_ARG(a->emb && a->seeds && a->e && a->el && a->nx && a->nc &&
                     a->labels_out && a->batch_out && a->segment_ids && a->dst && a->img_off &&
                     a->kmeans_labels && a->seed_out && a->num_segments,
                 "segment_by_kmeans: null pointer");
  const int dl = a->dim + a->loc_ch;
  const size_t need = spml_segment_by_kmeans_workspace_bytes(a->batch, a->n, dl, a->num_clusters,
                                                             a->iterations);
  if (workspace_bytes < need) {
    set_error("segment_by_kmeans: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  Carver c{reinterpret_cast<char*>(workspace), 0};
  const size_t scan_bytes = spml_valid_scan_workspace_bytes(a->batch, a->n);
  void* scan_ws = c.take_bytes(scan_bytes);
  const size_t km_bytes =
      spml_kmeans_workspace_bytes(a->batch, a->num_clusters, dl, std::max(a->iterations, 1));
  void* km_ws = c.take_bytes(km_bytes);
  const int64_t cap = (int64_t)a->batch * a->n;
  const size_t uq_bytes = spml_unique_workspace_bytes(cap);
  void* uq_ws = c.take_bytes(uq_bytes);
  (void)c.take<int64_t>((size_t)cap);   // (room of the former key array: the keys are built on the fly)
  int64_t* packed = c.take<int64_t>((size_t)cap);
  cudaStream_t st = as_stream(stream);

  // label packing (sem * divisor + inst, ignored pixels dropped) + valid-pixel scan + normalise /
  // pack: one kernel (normalize.cu::scan_normalize_pack_kernel)
  static const bool fused_prepass = []() {
    const char* e = getenv("SPML_B200_FUSED_PREPASS");   // =0: the three round-1 kernels (A/B timing)
    return !(e && e[0] == '0');
  }();
  if (fused_prepass) {
  SPML_TRY(scan_normalize_pack(a->emb, a->loc, a->loc_batch_stride, a->loc_ch, a->labels,
                               a->has_ignore, a->ignore_index, a->ignore_index_dev, a->sem, a->inst,
                               a->label_divisor, a->semantic_ignore, kDroppedLabel, a->seeds,
                               a->seed_batch_stride, a->batch, a->dim, a->n, a->batch_index_offset,
                               a->eps, a->dst, a->img_off, a->e, a->el, a->nx, a->nc, a->labels_out,
                               a->batch_out, a->seed_out, scan_ws, scan_bytes, st));
  } else {
    const int64_t* labels = a->labels;
    int has_ignore = a->has_ignore;
    int64_t ignore_index = a->ignore_index;
    const int64_t* ignore_dev = a->ignore_index_dev;
    if (!labels) {
      pack_labels_kernel<<<blocks_of(cap), 256, 0, st>>>(a->sem, a->inst, cap, a->label_divisor,
                                                        a->semantic_ignore, packed);
      SPML_LAUNCH_CHECK("pack_labels_kernel");
      labels = packed;
      has_ignore = 1;
      ignore_index = kDroppedLabel;
      ignore_dev = nullptr;
    }
    SPML_TRY(spml_valid_scan(labels, has_ignore, ignore_index, ignore_dev, a->batch, a->n, a->dst,
                             nullptr, a->img_off, scan_ws, scan_bytes, stream));
    SPML_TRY(spml_normalize_pack_fwd(a->emb, a->loc, a->loc_batch_stride, a->loc_ch, labels,
                                     a->seeds, a->seed_batch_stride, a->dst, a->batch, a->dim, a->n,
                                     a->batch_index_offset, a->eps, a->e, a->el, a->nx, a->nc,
                                     a->labels_out, a->batch_out, a->seed_out, stream));
  }
  // The table of the id numbering is cleared, and the largest label found, beside the k-means
  // (side stream); the number of segments is final once the distinct (image, cluster, label)
  // keys are in the table, so it goes to the host right then and the ranking of the keys runs
  // while the host already prepares the next stage.
  const int32_t* rows_dev = a->img_off + a->batch;
  StreamPool* pool = nullptr;
  SPML_TRY(get_pool(&pool));
  SPML_TRY(fork_streams(*pool, st, 1));
  SPML_TRY(unique_prepare(true, a->labels_out, cap, rows_dev, 0, a->num_segments, uq_ws, uq_bytes,
                          pool->side[0]));
  SPML_TRY(spml_kmeans(a->el, a->img_off, a->batch, a->n, dl, a->num_clusters, a->k_per_image,
                       a->iterations, a->seed_out, a->kmeans_labels, nullptr, km_ws, km_bytes,
                       stream));
  SPML_TRY(join_stream(*pool, 0, st));
  if (a->counts_host) {
    SPML_CHECK_ARG(a->counts_dev, "segment_by_kmeans: counts_host needs counts_dev");
    pool->seq = pool->seq == 0x7fffffff ? 1 : pool->seq + 1;
  }
  // keys (image, cluster, label) into the table, labels decoded, counts published by the last block
  SPML_TRY(cluster_insert(a->kmeans_labels, a->batch_out, a->labels_out, cap, rows_dev,
                          a->num_clusters, a->num_segments, a->label_divisor, a->sem_out, a->inst_out,
                          a->status, a->counts_host ? a->counts_dev : nullptr,
                          a->counts_host ? pool->host_counts_dev : nullptr, pool->seq, uq_ws, st));
  // How the host waits for the counts.  Alone on the host it polls the pinned words (below).
  // With several ranks per host (torchrun sets LOCAL_WORLD_SIZE / WORLD_SIZE) the polling thread
  // was measured to cost a neighbouring rank up to ~100 us per step every other run (shared
  // cores), so there the driver's own wait is used: a 16-byte copy into the caller's memory,
  // which returns once the data is there; likewise when several host threads of this process
  // drive GPUs (thread-per-GPU DataParallel).  SPML_B200_COUNT_WAIT=poll|block overrides the
  // rank heuristic.
  static const bool poll_counts = []() {
    const char* e = getenv("SPML_B200_COUNT_WAIT");
    if (e && !strcmp(e, "poll")) return true;
    if (e && !strcmp(e, "block")) return false;
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const char* w = getenv("WORLD_SIZE");
    const int ranks = lw ? atoi(lw) : (w ? atoi(w) : 1);
    return ranks <= 1;
  }();
  const bool poll = poll_counts && g_pool_count.load(std::memory_order_relaxed) <= 1;
  if (a->counts_host && !poll) {
    SPML_CUDA(cudaMemcpyAsync(a->counts_host, a->counts_dev, 4 * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, st));
    SPML_CUDA(cudaEventRecord(pool->counts_ev, st));
  }
  SPML_TRY(unique_finish(true, cap, rows_dev, 0, a->segment_ids, nullptr, nullptr, a->num_segments,
                         nullptr, uq_ws, st));
  if (a->counts_host && !poll) {
    SPML_CUDA(cudaEventSynchronize(pool->counts_ev));
    return SPML_OK;
  }
  if (a->counts_host) {
    // the step's one wait for the device: poll the sequence number
    volatile int32_t* h = pool->host_counts;
    for (unsigned spins = 1; h[4] != pool->seq; ++spins) {
      // PAUSE: a bare load loop starves the sibling hyper-thread; the stream is looked at every
      // 65 536 polls (the query takes a driver lock) so that a failed kernel surfaces as an
      // error instead of a hang
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((spins & 0xffffu) == 0) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) {
          if (h[4] == pool->seq) break;
          set_error("segment_by_kmeans: the stream drained without publishing the counts");
          return SPML_E_CUDA;
        }
        if (q != cudaErrorNotReady) return cuda_fail(q, "segment_by_kmeans: waiting for the counts");
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    for (int i = 0; i < 4; ++i) a->counts_host[i] = h[i];
  }
  return SPML_OK;
}

// =========================================================================== B1

size_t spml_gather_prototypes_workspace_bytes(int64_t m, int dim, int dim_loc) {
  spml::Carver c{nullptr, 0};
  c.take_bytes(spml_segment_prototypes_workspace_bytes(m, dim));
  c.take_bytes(spml_segment_prototypes_workspace_bytes(m, std::max(dim_loc, 1)));
  return c.off;
}

int spml_gather_prototypes_fwd(const float* e, const float* el, int64_t rows, int dim, int dim_loc,
                               const int64_t* seg, const int64_t* batch, const int64_t* sem,
                               const int64_t* inst, int64_t m, float eps, float* protos,
                               float* protos_loc, float* norms, float* norms_loc, int64_t* p_sem,
                               int64_t* p_inst, int64_t* p_batch, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(rows > 0 && m > 0 && dim > 0 && e && seg && protos && norms && status && workspace,
                 "gather_prototypes_fwd: bad arguments");
  SPML_CHECK_ARG(!el || (dim_loc > 0 && protos_loc && norms_loc),
                 "gather_prototypes_fwd: embeddings_with_loc need their outputs");
  SPML_CHECK_ARG((!p_sem || sem) && (!p_inst || inst) && (!p_batch || batch),
                 "gather_prototypes_fwd: label outputs need label inputs");
  const size_t need = spml_gather_prototypes_workspace_bytes(m, dim, dim_loc);
  if (workspace_bytes < need) {
    set_error("gather_prototypes_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  Carver c{reinterpret_cast<char*>(workspace), 0};
  const size_t b0 = spml_segment_prototypes_workspace_bytes(m, dim);
  void* ws0 = c.take_bytes(b0);
  const size_t b1 = spml_segment_prototypes_workspace_bytes(m, std::max(dim_loc, 1));
  void* ws1 = c.take_bytes(b1);
  cudaStream_t st = as_stream(stream);
  StreamPool* pool = nullptr;
  SPML_TRY(get_pool(&pool));
  if (el) {
    SPML_TRY(fork_streams(*pool, st, 1));
    SPML_TRY(spml_segment_prototypes_fwd(el, rows, nullptr, dim_loc, seg, m, eps, protos_loc,
                                         norms_loc, ws1, b1, pool->side[0]));
  }
  SPML_TRY(spml_segment_prototypes_fwd(e, rows, nullptr, dim, seg, m, eps, protos, norms, ws0, b0,
                                       stream));
  if (p_sem || p_inst || p_batch) {
    if (p_sem && p_inst == p_sem + m && p_batch == p_inst + m) {   // one [3, m] buffer: one fill
      SPML_CUDA(cudaMemsetAsync(p_sem, 0xff, (size_t)m * 24, st));
    } else {
      if (p_sem) SPML_CUDA(cudaMemsetAsync(p_sem, 0xff, (size_t)m * 8, st));
      if (p_inst) SPML_CUDA(cudaMemsetAsync(p_inst, 0xff, (size_t)m * 8, st));
      if (p_batch) SPML_CUDA(cudaMemsetAsync(p_batch, 0xff, (size_t)m * 8, st));
    }
    scatter_segment_labels_kernel<<<blocks_of(rows), 256, 0, st>>>(seg, batch, sem, inst, rows, m,
                                                                   p_sem, p_inst, p_batch, status);
    SPML_LAUNCH_CHECK("scatter_segment_labels_kernel");
    check_segment_labels_kernel<<<blocks_of(rows), 256, 0, st>>>(seg, batch, sem, inst, rows, m,
                                                                 p_sem, p_inst, p_batch, status);
    SPML_LAUNCH_CHECK("check_segment_labels_kernel");
  }
  if (el) SPML_TRY(join_stream(*pool, 0, st));
  return SPML_OK;
}

int spml_gather_prototypes_bwd(const float* dprotos, const float* dprotos_loc, const float* protos,
                               const float* protos_loc, const float* norms, const float* norms_loc,
                               const int64_t* seg, int64_t rows, int dim, int dim_loc, int64_t m,
                               float eps, float* de, float* del, void* stream) {
  using namespace spml;
  SPML_CHECK_ARG(rows > 0 && m > 0 && seg, "gather_prototypes_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  StreamPool* pool = nullptr;
  SPML_TRY(get_pool(&pool));
  const bool both = dprotos && dprotos_loc;
  if (both) SPML_TRY(fork_streams(*pool, st, 1));
  if (dprotos_loc)
    SPML_TRY(spml_segment_prototypes_bwd(dprotos_loc, protos_loc, norms_loc, seg, rows, nullptr,
                                         dim_loc, m, eps, 0.f, del,
                                         both ? (void*)pool->side[0] : stream));
  if (dprotos)
    SPML_TRY(spml_segment_prototypes_bwd(dprotos, protos, norms, seg, rows, nullptr, dim, m, eps,
                                         0.f, de, stream));
  if (both) SPML_TRY(join_stream(*pool, 0, st));
  return SPML_OK;
}

// =========================================================================== C4

size_t spml_head_workspace_bytes(const spml_head_args* a) {
  if (!a || a->n <= 0 || a->m <= 0) return 256;
  return spml::head_plan(*a, nullptr).bytes;
}

int spml_head_fwd(const spml_head_args* a, void* state, size_t state_bytes, float* out,
                  void* stream) {
  using namespace spml;
  SPML_TRY(check_head(a, "head_fwd"));
  SPML_CHECK_ARG(state && out, "head_fwd: null pointer");
  HeadPlan p = head_plan(*a, state);
  if (state_bytes < p.bytes) {
    set_error("head_fwd: state buffer %zu < %zu bytes", state_bytes, p.bytes);
    return SPML_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  StreamPool* pool = nullptr;
  SPML_TRY(get_pool(&pool));
  const unsigned en = a->enable;
  const int64_t n = a->n, m_all = p.m_all;
  const bool contrast = (en & (kAnn | kOcc | kAcc)) != 0;
  FinishArgs fin{};
  for (int i = 0; i < 3; ++i) fin.desc[i] = p.desc[i];

  // ---- img_sim branch (side stream 0): per-image prototypes of its embeddings, then the loss
  SPML_TRY(fork_streams(*pool, st, 1));
  if (en & kSim) {
    cudaStream_t s0 = pool->side[0];
    const int dsim = head_dim_sim(*a);
    SPML_TRY(spml_segment_prototypes_fwd(a->img_sim_on_plain ? a->e : a->el, n, nullptr, dsim,
                                         a->seg, a->m, a->eps, p.sim_protos, p.sim_norms,
                                         p.proto_ws, p.proto_ws_bytes, s0));
    // the prototypes' instance labels (segsort.py:229-231 re-derives them per image): taken
    // from the caller when it has them, else scattered from the pixels; either way every pixel
    // must carry its segment's label (a violation is reported through the status word)
    int32_t* status = a->status ? a->status : p.hits + 2;
    if (!a->pinst) {
      SPML_CUDA(cudaMemsetAsync(p.pinst, 0xff, (size_t)a->m * 8, s0));
      scatter_segment_labels_kernel<<<blocks_of(n), 256, 0, s0>>>(
          a->seg, nullptr, nullptr, a->inst, n, a->m, nullptr, p.pinst, nullptr, status);
      SPML_LAUNCH_CHECK("scatter_segment_labels_kernel");
    }
    check_segment_labels_kernel<<<blocks_of(n), 256, 0, s0>>>(
        a->seg, nullptr, nullptr, a->inst, n, a->m, nullptr, a->pinst ? a->pinst : p.pinst,
        nullptr, status);
    SPML_LAUNCH_CHECK("check_segment_labels_kernel");
    group_offsets_kernel<<<(unsigned)ceil_div(a->max_groups + 1, 64), 64, 0, s0>>>(
        a->bid, n, a->pbid, a->m, a->max_groups, p.group_off, p.col_off, a->status);
    SPML_LAUNCH_CHECK("group_offsets_kernel");
    SPML_TRY(segsort_fwd_partial(&p.desc[2], p.stats[2], nullptr, p.seg_ws[2], p.seg_ws_bytes[2],
                                 s0, &fin.partial[2], &fin.tiles_x[2]));
  }

  // ---- the prototype bank of this step: current prototypes + memory bank (segsort.py:153-182)
  if (contrast) {
    ConcatList lp{}, ls{}, lb{}, ll{};
    TagList lt{};
    int64_t rows = 0;
    for (int i = 0; i <= a->num_bank; ++i) {
      const int64_t mi = i == 0 ? a->m : a->bank_m[i - 1];
      lp.src[i] = i == 0 ? a->protos : a->bank_protos[i - 1];
      ls.src[i] = i == 0 ? a->psem : a->bank_psem[i - 1];
      lb.src[i] = i == 0 ? a->pbid : a->bank_pbid[i - 1];
      ll.src[i] = i == 0 ? a->protos_loc : a->bank_protos_loc[i - 1];
      lt.src[i] = i == 0 ? a->ptags : a->bank_tags[i - 1];
      lt.ld[i] = i == 0 ? a->ptags_ld : a->bank_tags_ld[i - 1];
      lp.first[i] = rows * a->dim;
      ll.first[i] = rows * a->dim_loc;
      ls.first[i] = lb.first[i] = lt.first[i] = rows;
      rows += mi;
    }
    const int cnt = a->num_bank + 1;
    lp.n = ls.n = lb.n = ll.n = lt.n = cnt;
    lp.first[cnt] = rows * a->dim;
    ll.first[cnt] = rows * a->dim_loc;
    ls.first[cnt] = lb.first[cnt] = lt.first[cnt] = rows;
    const bool voc_tags = (en & kOcc) && !a->nn_tags;
    PrepArgs pa{};
    pa.protos = lp;
    pa.psem = ls;
    pa.ptags = lt;
    pa.img_tags = a->img_tags;
    pa.img_tags_ld = a->img_tags_ld;
    pa.tag_rows = a->tag_rows;
    pa.bid = a->bid;
    pa.n = n;
    pa.m_all = m_all;
    pa.num_classes = a->num_classes;
    pa.dim = a->dim;
    pa.col0 = a->tag_col0;
    pa.col1 = a->tag_col1;
    pa.tag_mode = voc_tags ? 1 : 0;
    head_prep_kernel<<<blocks_of(std::max<int64_t>(rows * a->dim, n)), 256, 0, st>>>(
        pa, p.p_all, p.psem_all, p.pvalid_ann, p.pmask_all, p.pix_mask);
    SPML_LAUNCH_CHECK("head_prep_kernel");
    if ((en & kOcc) && a->nn_tags) {
      // segsort_softmax_densepose.py:174-191: tags of a prototype = class of its most similar
      // labelled prototype of the same image (>= threshold), no tag -> every tag
      concat_kernel<float><<<blocks_of(rows * a->dim_loc), 256, 0, st>>>(ll, p.pl_all);
      SPML_LAUNCH_CHECK("concat_kernel");
      concat_kernel<int64_t><<<blocks_of(rows), 256, 0, st>>>(lb, p.pbid_all);
      SPML_LAUNCH_CHECK("concat_kernel");
      SPML_TRY(spml_nn_multiset_labels(p.pl_all, m_all, p.pl_all, m_all, a->dim_loc, p.psem_all,
                                       p.pbid_all, p.pbid_all, (int)a->num_classes, 1,
                                       a->nn_threshold, nullptr, p.pmask_all, p.nn_ws,
                                       p.nn_ws_bytes, stream));
      proto_flags_kernel<<<blocks_of(m_all), 256, 0, st>>>(p.psem_all, m_all, a->num_classes,
                                                           p.pvalid_ann, p.pmask_all, 1);
      SPML_LAUNCH_CHECK("proto_flags_kernel");
      gather_mask_kernel<<<blocks_of(n), 256, 0, st>>>(p.pmask_all, m_all, a->seg, n, p.pix_mask);
      SPML_LAUNCH_CHECK("gather_mask_kernel");
    }

    // ---- fork: sem_occ on the caller's stream, sem_ann on side 1, accuracy on side 2
    SPML_CUDA(cudaEventRecord(pool->fork, st));
    if (en & kAnn) {
      cudaStream_t s1 = pool->side[1];
      SPML_CUDA(cudaStreamWaitEvent(s1, pool->fork, 0));
      SPML_TRY(spml_valid_scan(a->sem, 2, a->num_classes, nullptr, 1, (int)n, p.ann_dst, p.ann_rows,
                               p.ann_off, p.scan_ws, p.scan_ws_bytes, s1));
      SPML_TRY(segsort_fwd_partial(&p.desc[0], p.stats[0], nullptr, p.seg_ws[0],
                                   p.seg_ws_bytes[0], s1, &fin.partial[0], &fin.tiles_x[0]));
    }
    if (en & kAcc) {
      cudaStream_t s2 = pool->side[2];
      SPML_CUDA(cudaStreamWaitEvent(s2, pool->fork, 0));
      TopkExtra none{};
      SPML_TRY(topk_launch(p.p_all, m_all, p.p_all, m_all, a->dim, p.psem_all, p.psem_all, nullptr,
                           nullptr, (int)std::min<int64_t>(5, m_all), p.topk_labels, nullptr,
                           p.hits, none, s2));
    }
    if (en & kOcc)
      SPML_TRY(segsort_fwd_partial(&p.desc[1], p.stats[1], nullptr, p.seg_ws[1],
                                   p.seg_ws_bytes[1], st, &fin.partial[1], &fin.tiles_x[1]));
    if (en & kAnn) SPML_TRY(join_stream(*pool, 1, st));
    if (en & kAcc) SPML_TRY(join_stream(*pool, 2, st));
  }
  SPML_TRY(join_stream(*pool, 0, st));
  head_finish_kernel<<<1, 128, 0, st>>>(fin, p.hits, a->weight_ann, a->weight_occ, a->weight_sim,
                                        (int)std::min<int64_t>(5, m_all), en, out);
  SPML_LAUNCH_CHECK("head_finish_kernel");
  return SPML_OK;
}

int spml_head_bwd(const spml_head_args* a, void* state, size_t state_bytes, const float* g_ann,
                  const float* g_occ, const float* g_sim, const float* g_total, float* de,
                  float* del, float* dprotos, void* stream) {
  using namespace spml;
  SPML_TRY(check_head(a, "head_bwd"));
  SPML_CHECK_ARG(state, "head_bwd: null pointer");
  HeadPlan p = head_plan(*a, state);
  if (state_bytes < p.bytes) {
    set_error("head_bwd: state buffer %zu < %zu bytes", state_bytes, p.bytes);
    return SPML_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  StreamPool* pool = nullptr;
  SPML_TRY(get_pool(&pool));
  const bool ann = (a->enable & kAnn) && (g_ann || g_total);
  const bool occ = (a->enable & kOcc) && (g_occ || g_total);
  const bool sim = (a->enable & kSim) && (g_sim || g_total);
  const int64_t n = a->n;
  const int dsim = head_dim_sim(*a);
  float* d_sim_emb = a->img_sim_on_plain ? de : del;
  SPML_CHECK_ARG((!(ann || occ) || de) && (!sim || d_sim_emb),
                 "head_bwd: missing gradient output");
  // which buffers end up untouched by a kernel with beta = 0
  const bool de_written = occ || (sim && a->img_sim_on_plain && !occ && !ann);
  if (de && !de_written) SPML_CUDA(cudaMemsetAsync(de, 0, (size_t)n * a->dim * 4, st));
  if (del && !(sim && !a->img_sim_on_plain))
    SPML_CUDA(cudaMemsetAsync(del, 0, (size_t)n * a->dim_loc * 4, st));
  if (dprotos && !(ann || occ))
    SPML_CUDA(cudaMemsetAsync(dprotos, 0, (size_t)a->m * a->dim * 4, st));
  scale_grads_kernel<<<1, 32, 0, st>>>(ann ? g_ann : nullptr, occ ? g_occ : nullptr,
                                       sim ? g_sim : nullptr, g_total, a->weight_ann,
                                       a->weight_occ, a->weight_sim, p.gw);
  SPML_LAUNCH_CHECK("scale_grads_kernel");
  for (int i = 0; i < 3; ++i) p.desc[i].reserved |= 4;   // operands prepared by the forward

  SPML_TRY(fork_streams(*pool, st, kSideStreams));
  // d(prototypes) of sem_occ / sem_ann on side streams 1 / 2: only the current step's
  // prototypes carry a gradient (the memory bank behind them is detached); the per-chunk
  // partial sums of both problems stay in their workspaces and are added once after the join
  const float *part_occ = nullptr, *part_ann = nullptr;
  int chunks_occ = 0, chunks_ann = 0;
  if (occ && dprotos)
    SPML_TRY(segsort_bwd_impl(&p.desc[1], p.stats[1], p.gw + 1, 0.f, nullptr, a->dim, nullptr,
                              a->m, &part_occ, &chunks_occ, p.seg_ws[1], p.seg_ws_bytes[1],
                              pool->side[1]));
  if (ann && dprotos)
    SPML_TRY(segsort_bwd_impl(&p.desc[0], p.stats[0], p.gw + 0, 0.f, nullptr, a->dim, nullptr,
                              a->m, &part_ann, &chunks_ann, p.seg_ws[0], p.seg_ws_bytes[0],
                              pool->side[2]));
  // img_sim: d(embedding) and d(per-image prototypes) -> segment-prototype backward
  if (sim && !a->img_sim_on_plain) {
    cudaStream_t s0 = pool->side[0];
    SPML_TRY(spml_segsort_bwd(&p.desc[2], p.stats[2], p.gw + 2, 0.f, del, dsim, p.d_sim_protos,
                              p.seg_ws[2], p.seg_ws_bytes[2], s0));
    SPML_TRY(spml_segment_prototypes_bwd(p.d_sim_protos, p.sim_protos, p.sim_norms, a->seg, n,
                                         nullptr, dsim, a->m, a->eps, 1.f, del, s0));
  }
  // d(embedding) of sem_occ (all rows, overwrites) then sem_ann (labelled rows, accumulates)
  if (occ)
    SPML_TRY(spml_segsort_bwd(&p.desc[1], p.stats[1], p.gw + 1, 0.f, de, a->dim, nullptr,
                              p.seg_ws[1], p.seg_ws_bytes[1], stream));
  if (ann)
    SPML_TRY(spml_segsort_bwd(&p.desc[0], p.stats[0], p.gw + 0, 1.f, de, a->dim, nullptr,
                              p.seg_ws[0], p.seg_ws_bytes[0], stream));
  if (sim && a->img_sim_on_plain) {
    // densepose: img_sim works on the plain embeddings too, so it queues behind the others
    SPML_TRY(spml_segsort_bwd(&p.desc[2], p.stats[2], p.gw + 2, (occ || ann) ? 1.f : 0.f, de, dsim,
                              p.d_sim_protos, p.seg_ws[2], p.seg_ws_bytes[2], stream));
    SPML_TRY(spml_segment_prototypes_bwd(p.d_sim_protos, p.sim_protos, p.sim_norms, a->seg, n,
                                         nullptr, dsim, a->m, a->eps, 1.f, de, stream));
  }
  for (int i = 0; i < kSideStreams; ++i) SPML_TRY(join_stream(*pool, i, st));
  if (dprotos && (ann || occ))
    SPML_TRY(segsort_reduce_two(part_occ, chunks_occ, part_ann, chunks_ann, a->m * a->dim, dprotos,
                                st));
  return SPML_OK;
}

}  // extern "C"
