/*
 * spml_b200.h - C ABI of libspml_b200.so: the pixel-to-segment contrastive hot
 * path of SPML (twke18/SPML) as hand-written sm_100a CUDA.
 *
 * The reference has no native code and therefore no FFI of its own; its
 * "plugin interface" for this path is a set of Python functions.  Each entry
 * point below names the reference function(s) it replaces (paths relative to the
 * reference checkout).  spml_b200/_lib.py binds exactly these symbols with
 * ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - embeddings are fp32, row-major, rows contiguous (ld == dim) unless an
 *    ld_* argument says otherwise; labels / indices at the API edge are int64
 *    (the reference's dtype); image offsets and internal ids are int32;
 *  - the caller allocates every input, output and workspace buffer and owns it;
 *    the library never allocates or frees device memory, keeps no pointer after
 *    a call returns and has no global mutable state (re-entrant per stream);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *    no call synchronises with the host;
 *  - return value: 0 on success, a negative SPML_E_* code otherwise, with a
 *    thread-local message available from spml_last_error().
 */
#ifndef SPML_B200_H_
#define SPML_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPML_OK 0
#define SPML_E_INVALID (-1)     /* bad argument (null pointer, size, dim)    */
#define SPML_E_CUDA (-2)        /* a CUDA runtime call or launch failed       */
#define SPML_E_UNSUPPORTED (-3) /* shape outside what the kernels are built for */
#define SPML_E_WORKSPACE (-4)   /* workspace too small                       */

#define SPML_MAX_DIM 136   /* largest embedding dim (incl. location channels) */
#define SPML_MAX_TOPK 32
#define SPML_MAX_BANK 8    /* memory-bank entries one spml_head_* call can take */

/* thread-local text of the last error on this thread ("" if none). */
const char* spml_last_error(void);
/* ABI version, bumped on any signature change. */
int spml_abi_version(void);
/* diagnostics: number of kernels this library has launched (all threads) since it
 * was loaded (memsets are not counted). */
uint64_t spml_debug_launch_count(void);
/* diagnostics: which k-means kernel spml_kmeans runs for this shape on the current device
 * (0 fp32 CUDA cores, 1 tcgen05 tiles, 2 tcgen05 small-K, 3 one thread-block cluster per image). */
int spml_debug_kmeans_path(int batch, int max_rows_per_image, int dim, int num_clusters);
/* sizeof the argument structs below (0: spml_segsort_desc, 1: spml_cluster_args, 2:
 * spml_head_args), so that a binding can check its own struct layout at load time. */
size_t spml_sizeof_struct(int which);

/* ---------------------------------------------------------------------------
 * A1. spml/utils/general/common.py:101-120 normalize_embedding
 *     y = x / max(||x||, eps) row-wise; norms_out[i] = ||x_i|| if >= eps else
 *     -eps (the sign marks the clamped branch for the backward).
 *     bwd: dx = (dy - y (y.dy)) / ||x||   (dy / eps on the clamped branch).
 */
int spml_normalize_rows_fwd(const float* x, int64_t rows, int dim, float eps,
                            float* y, float* norms_out, void* stream);
int spml_normalize_rows_bwd(const float* dy, const float* y, const float* norms,
                            int64_t rows, int dim, float* dx, void* stream);

/* ---------------------------------------------------------------------------
 * A8 (front half). spml/utils/segsort/common.py:306-310,339-365,376-381:
 *     NCHW->NHWC, L2-normalise, concatenate the local (location / colour)
 *     features and re-normalise, drop pixels whose label == ignore_index,
 *     keep raster order.
 *
 * spml_valid_scan: dst[b*n+p] = output row of pixel p of image b, or -1 when
 *     it is ignored; img_off[b] = first output row of image b, img_off[batch] =
 *     number of kept pixels.  has_ignore == 0 keeps everything, 1 drops the pixels whose
 *     label == ignore_index, 2 keeps only the pixels whose label < ignore_index (the
 *     labelled-pixel filter of segsort.py:184-185).  src (nullable,
 *     capacity batch*n) is the inverse map: src[row] = b*n+p.  ignore_index_dev
 *     (nullable, device) overrides ignore_index when the value only exists on
 *     the device (resnet_deeplab.py:113 computes it as labels.max() + 1).
 *     workspace: spml_valid_scan_workspace_bytes(batch, n).
 */
size_t spml_valid_scan_workspace_bytes(int batch, int n);
int spml_valid_scan(const int64_t* labels, int has_ignore, int64_t ignore_index,
                    const int64_t* ignore_index_dev, int batch, int n, int32_t* dst,
                    int32_t* src, int32_t* img_off, void* workspace,
                    size_t workspace_bytes, void* stream);

/* spml_normalize_pack_fwd: for every kept pixel (row r = dst[pixel]):
 *     e[r]  = x / max(||x||, eps)                       [rows, dim]
 *     el[r] = cat(e[r], loc) / max(||cat||, eps)        [rows, dim + loc_ch]
 *     nx[r], nc[r] = the two norms (negative eps when clamped)
 *     labels_out[r] = labels[pixel]; batch_out[r] = b + batch_index_offset;
 *     seed_out[r] = seeds[b * seed_batch_stride + p]  (int32, for k-means)
 *   emb is [batch, dim, n] (NCHW with n = H*W), loc is [*, n, loc_ch] with a
 *   batch stride in elements (0 = the same map for every image).  labels,
 *   seeds, labels_out, batch_out, seed_out may be NULL.
 * spml_normalize_pack_bwd: d(emb) [batch, dim, n] from d(e), d(el) (either may
 *   be NULL); ignored pixels get 0.  Follows autograd of common.py:306-365.
 */
int spml_normalize_pack_fwd(const float* emb, const float* loc, int64_t loc_batch_stride,
                            int loc_ch, const int64_t* labels, const int64_t* seeds,
                            int64_t seed_batch_stride, const int32_t* dst, int batch,
                            int dim, int n, int64_t batch_index_offset, float eps,
                            float* e, float* el, float* nx, float* nc,
                            int64_t* labels_out, int64_t* batch_out, int32_t* seed_out,
                            void* stream);
int spml_normalize_pack_bwd(const float* de, const float* del, const float* e,
                            const float* el, const float* nx, const float* nc,
                            const int32_t* dst, int batch, int dim, int loc_ch, int n,
                            float eps, float* demb, void* stream);

/* ---------------------------------------------------------------------------
 * A4-A6. spml/utils/segsort/common.py:11-97: spherical k-means with initial
 *     labels, batched over images.  x is [rows, dim] unit vectors grouped by
 *     image (img_off, as written by spml_valid_scan); image b clusters into
 *     k_per_image[b] (NULL: num_clusters) prototypes.  Each iteration is the
 *     reference's M-step (segment sum, L2-normalise; an empty cluster is the
 *     zero vector) followed by the E-step (argmax of x.P^T, first index on
 *     ties).  Segment sums are accumulated in 64-bit fixed point (2^-32), so
 *     the result does not depend on the order of accumulation.
 *     The E-step runs as a tcgen05 GEMM (bf16 hi/lo split or, in the kernel that
 *     keeps an image inside one thread-block cluster, a single fp16 product; near-ties
 *     re-scored in fp32) or as an fp32 CUDA-core GEMM; all kernels return the same ids
 *     bit for bit (the choice is made per call from the shape).
 *     labels_out (int32) and labels_out_i64 (nullable) receive the final ids.
 *     workspace: spml_kmeans_workspace_bytes(batch, num_clusters, dim, iterations),
 *     16-byte aligned; the call is one launch: cooperative (all its CTAs must be
 *     able to be resident: it waits for the SMs of concurrently running kernels) or
 *     one cluster of up to 16 CTAs per image.
 */
size_t spml_kmeans_workspace_bytes(int batch, int num_clusters, int dim, int iterations);
int spml_kmeans(const float* x, const int32_t* img_off, int batch,
                int max_rows_per_image, int dim, int num_clusters,
                const int32_t* k_per_image, int iterations, const int32_t* init_labels,
                int32_t* labels_out, int64_t* labels_out_i64, void* workspace,
                size_t workspace_bytes, void* stream);

/* A5 alone. common.py:44-64 find_nearest_prototypes: out[i] = argmax_k x_i.p_k. */
int spml_nearest_prototype(const float* x, int64_t rows, int dim, const float* protos,
                           int num_protos, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------
 * A7 / A8 tail / B1. torch.unique(key, return_inverse=True) as used in
 *     common.py:213-214,341,398-401 and models/utils.py:95-108:
 *     key[i] = hi[i] * bound + lo[i]  (hi NULL: key = lo).  bound > 0 is used as
 *     given; bound == 0 means max(lo) + 1 computed on the device (the
 *     reference's `labels.max() + 1`).  Outputs: inverse[i] = rank of key[i]
 *     among the distinct keys in ascending order; uniq_hi/uniq_lo[r] = the
 *     decoded r-th distinct key (capacity n, nullable); count[0] = number of
 *     distinct keys; bound_out[0] (nullable) = the bound used.  n is the
 *     capacity of the key arrays; n_dev (nullable, device) holds the number of
 *     live keys when that is only known on the device (min(n, *n_dev) are used).
 *     workspace: spml_unique_workspace_bytes(n).
 */
size_t spml_unique_workspace_bytes(int64_t n);
int spml_unique_inverse(const int64_t* hi, const int64_t* lo, int64_t n,
                        const int32_t* n_dev, int64_t bound, int64_t* inverse, int64_t* uniq_hi, int64_t* uniq_lo,
                        int32_t* count, int64_t* bound_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * A4 / B1. common.py:11-41 calculate_prototypes_from_labels (also
 *     models/utils.py:113-116): protos[m] = normalize(sum_{seg[i]==m} x[i]).
 *     norms[m] = ||sum|| (negative eps when clamped).  Inputs must satisfy
 *     |x| <= 1024 (they are unit vectors on this path); a violation (or a NaN)
 *     turns every prototype of the call into NaN.
 *     bwd: dx[i] = beta*dx[i] + (dp - p (p.dp)) / ||sum||  gathered at seg[i].
 *     rows_dev (nullable, device) holds the live row count when it only exists on the
 *     device; rows is then the capacity.
 *     workspace: spml_segment_prototypes_workspace_bytes(m, dim).
 */
size_t spml_segment_prototypes_workspace_bytes(int64_t m, int dim);
int spml_segment_prototypes_fwd(const float* x, int64_t rows, const int32_t* rows_dev, int dim,
                                const int64_t* seg, int64_t m, float eps, float* protos,
                                float* norms, void* workspace, size_t workspace_bytes,
                                void* stream);
int spml_segment_prototypes_bwd(const float* dprotos, const float* protos,
                                const float* norms, const int64_t* seg, int64_t rows,
                                const int32_t* rows_dev, int dim, int64_t m, float eps,
                                float beta, float* dx, void* stream);

/* B1 (shortcut). spml/models/utils.py:100-111 when the segment ids come straight from
 *     segment_by_kmeans (dense ranks of (image, cluster, label)): decodes the packed
 *     labels per pixel (sem = label / divisor, inst = label % divisor, keep = live &&
 *     sem < num_classes) and scatters them to the segments (p_sem, p_inst, p_batch,
 *     p_live; capacity m_cap).  Rows >= *rows_dev are padding.  Segments nobody maps to
 *     get p_sem = dead_label, p_batch = -1, p_live = 0.  A segment id >= m_cap sets
 *     overflow[0] = 1 (never reset here) and is skipped. */
int spml_segment_labels(const int64_t* labels, const int64_t* batch, const int64_t* seg,
                        int64_t cap, const int32_t* rows_dev, int64_t divisor,
                        int64_t num_classes, int64_t m_cap, int64_t dead_label, int64_t* sem,
                        int64_t* inst, int64_t* keep, int64_t* p_sem, int64_t* p_inst,
                        int64_t* p_batch, uint8_t* p_live, int32_t* overflow, void* stream);

/* ---------------------------------------------------------------------------
 * C1 / C2. spml/utils/segsort/loss.py:15-130 (+ SegSortLoss / SetSegSortLoss
 *     forward, :133-251), group_mode 'segsort+'.
 *
 *   S_ij = exp(kappa * e_i . p_j);  same_ij = code match, diff_ij = !same_ij
 *   mode SPML_MODE_CLASS: match = (pix_code[i] == proto_code[j])      (C1)
 *   mode SPML_MODE_TAGS : match = (pix_code[i] & proto_code[j]) != 0  (C2; codes
 *        are bit masks packed by spml_pack_tags)
 *   o_i = sum_j S_ij same_ij - S_{i,seg_i};  num_i = o_i > 0 ? o_i : S_{i,seg_i}
 *   den_i = sum_j S_ij diff_ij + num_i;      nll_i = -log(num_i / den_i)
 *
 *   Rows: row r of the problem is embedding row row_index[r] (NULL: r).  Rows
 *   are partitioned into num_groups consecutive groups by group_off (device,
 *   num_groups+1 entries; NULL: one group [0, n_rows)); group g only sees the
 *   prototype columns [col_off[g], col_off[g+1]) (NULL: all m columns), which
 *   is how the per-image img_sim loss (segsort.py:220-240) runs as one launch.
 *   proto_valid (nullable) masks prototype columns out entirely (sem_ann's
 *   labelled-prototype filter, segsort.py:184-201).  pix_code and seg are
 *   indexed by embedding row; seg holds absolute prototype columns.
 *
 *   reduction SPML_REDUCE_MEAN       : loss = mean over all rows
 *             SPML_REDUCE_GROUP_MEAN : mean over non-empty groups of group means
 *             SPML_REDUCE_SUM        : sum
 *   An empty problem gives NaN for the means (torch.mean of an empty tensor).
 *
 *   fwd writes stats[r] = {num, den, o_i > 0 ? 1 : 0}, nll[r] (nullable) and
 *   loss[0].  bwd consumes stats and grad_loss[0] (device scalar) and writes
 *   demb (embedding rows, scattered through row_index; beta = 0 overwrites the
 *   rows of this problem, beta = 1 accumulates) and dprotos [m, dim] (nullable;
 *   always overwritten).  workspace: spml_segsort_workspace_bytes(&desc).
 */
#define SPML_MODE_CLASS 0
#define SPML_MODE_TAGS 1
#define SPML_REDUCE_MEAN 0
#define SPML_REDUCE_GROUP_MEAN 1
#define SPML_REDUCE_SUM 2

typedef struct spml_segsort_desc {
  const float* emb;          /* [*, ld_emb] */
  int64_t ld_emb;
  int32_t dim;
  int32_t num_groups;        /* >= 1 */
  const int32_t* row_index;  /* [n_rows] or NULL */
  const int32_t* group_off;  /* [num_groups+1] device, or NULL */
  const int32_t* col_off;    /* [num_groups+1] device, or NULL */
  int64_t n_rows;            /* upper bound on the number of rows (exact if group_off == NULL) */
  int64_t max_rows_per_group;/* upper bound used to size the grid */
  const int64_t* pix_code;
  const int64_t* seg;
  const float* protos;       /* [m, ld_protos] */
  int64_t ld_protos;
  int64_t m;
  const int64_t* proto_code; /* [m] */
  const uint8_t* proto_valid;/* [m] or NULL */
  float kappa;
  int32_t mode;
  int32_t reduction;
  int32_t reserved;          /* bit 0: force the fp32 kernels, bit 1: force the tcgen05 kernels,
                                bit 2 (bwd): the workspace is the one the forward call of
                                this same problem used, its prepared operands are re-used */
} spml_segsort_desc;

size_t spml_segsort_workspace_bytes(const spml_segsort_desc* desc);
int spml_segsort_fwd(const spml_segsort_desc* desc, float* stats, float* nll, float* loss,
                     void* workspace, size_t workspace_bytes, void* stream);
int spml_segsort_bwd(const spml_segsort_desc* desc, const float* stats,
                     const float* grad_loss, float beta, float* demb, int64_t ld_demb,
                     float* dprotos, void* workspace, size_t workspace_bytes,
                     void* stream);
/* The same with the prototype gradient limited to the first dprotos_rows prototypes
 * (dprotos is [dprotos_rows, dim]): the rows behind them are a detached memory bank
 * (train.py:276-293), whose gradient the reference computes and throws away. */
int spml_segsort_bwd_rows(const spml_segsort_desc* desc, const float* stats,
                          const float* grad_loss, float beta, float* demb, int64_t ld_demb,
                          float* dprotos, int64_t dprotos_rows, void* workspace,
                          size_t workspace_bytes, void* stream);

/* packs rows of a [rows, cols] int64 0/1 matrix (cols <= 64, row stride ld) into
 * bit masks: bit c set iff tags[r, c] != 0  (loss.py:107-109 uses (a.b) > 0). */
int spml_pack_tags(const int64_t* tags, int64_t rows, int cols, int64_t ld,
                   int64_t* masks, void* stream);

/* ---------------------------------------------------------------------------
 * C3. spml/utils/segsort/eval.py:9-52 top_k_ranking: for each query row the k
 *     prototypes of largest q.p (descending, lowest index first on ties);
 *     topk_labels[q, r] = plab[index]; hit_count[0] = number of (q, r) with
 *     qlab[q] == plab[index], hit_count[1] = number of queries that took part
 *     (accuracy = hit_count[0] / (hit_count[1] * k)).  qvalid / pvalid (nullable byte
 *     masks) drop query rows / prototype columns, for fixed-capacity buffers.
 *     A large prototype bank without masks (retrieval inference,
 *     predictions/segsort.py:68-125) runs on the tensor cores: tcgen05 scores select
 *     k + 8 candidates per query, which are re-scored exactly; same results.  That
 *     path keeps a scratch buffer per host thread and device (allocated on first use,
 *     never while the stream is being captured: then the FMA kernel runs).
 */
int spml_topk_ranking(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                      const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                      const uint8_t* pvalid, int k, int64_t* topk_labels,
                      int64_t* topk_index, int32_t* hit_count, void* stream);

/* ---------------------------------------------------------------------------
 * f3. spml/models/utils.py:157-223 gather_multiset_labels_per_batch_by_nearest_neighbor:
 *     for every query row the top_k most similar prototypes (q.p) among those with
 *     pgroup == qgroup[row] (nullable: no group restriction) and plab < num_classes;
 *     tags[row, c] = 1 iff one of them has plab == c and similarity >= threshold
 *     ([nq, num_classes] int64, nullable); masks[row] = the same as a bit mask (nullable,
 *     num_classes <= 64).  workspace: spml_nn_multiset_labels_workspace_bytes(nq, top_k).
 */
size_t spml_nn_multiset_labels_workspace_bytes(int64_t nq, int top_k);
int spml_nn_multiset_labels(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                            const int64_t* plab, const int64_t* qgroup, const int64_t* pgroup,
                            int num_classes, int top_k, float threshold, int64_t* tags,
                            int64_t* masks, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ===========================================================================
 * Stage-group entry points: one call per stage of the reference's training step
 * (pyscripts/train/train.py:167-219), each enqueueing every kernel of its stage.  Kernels
 * of a stage that do not depend on each other run on side streams owned by the library
 * (a lazily created pool per host thread and device; they fork from and join back into
 * `stream`, so for the caller all work is ordered on `stream`).
 *
 * A8 whole. spml/utils/segsort/common.py:270-408 segment_by_kmeans (+ the label packing
 *     of resnet_deeplab.py:112-117 on the caller's side): valid-pixel scan, NCHW -> packed
 *     normalised rows (+ location features), spherical k-means per image, final segment
 *     ids = rank of (image, cluster, label).  Buffers have the capacity batch * n rows;
 *     img_off[batch] (rows kept) and num_segments[0] stay on the device.  `seeds` must be
 *     dense ids in [0, num_clusters) (k_per_image: per-image cluster counts, nullable).
 *     workspace: spml_segment_by_kmeans_workspace_bytes(batch, n, dim + loc_ch,
 *     num_clusters, iterations), 256-byte aligned.
 */
typedef struct spml_cluster_args {
  const float* emb;             /* [batch, dim, n] */
  const float* loc;             /* [*, n, loc_ch], nullable when loc_ch == 0 */
  int64_t loc_batch_stride;     /* elements; 0 = one map for every image */
  const int64_t* labels;        /* [batch, n]; or NULL with sem / inst below */
  /* resnet_deeplab.py:112-117 inside the call: label = sem * label_divisor + inst, pixels
     with sem == semantic_ignore are dropped (labels == NULL; has_ignore etc. unused) */
  const int64_t* sem;           /* [batch, n] or NULL */
  const int64_t* inst;          /* [batch, n] or NULL */
  int64_t label_divisor;
  int64_t semantic_ignore;
  const int64_t* ignore_index_dev; /* nullable: overrides ignore_index */
  int64_t ignore_index;
  const int64_t* seeds;         /* [*, n] */
  int64_t seed_batch_stride;    /* elements; 0 = one map for every image */
  const int32_t* k_per_image;   /* [batch] or NULL */
  int64_t batch_index_offset;
  int32_t batch, dim, n, loc_ch;
  int32_t num_clusters, iterations, has_ignore, reserved;
  float eps;
  float reserved_f;
  /* outputs (capacity batch * n rows) */
  float* e;                     /* [cap, dim] */
  float* el;                    /* [cap, dim + loc_ch] */
  float* nx;                    /* [cap] */
  float* nc;                    /* [cap] */
  int64_t* labels_out;          /* [cap] */
  int64_t* batch_out;           /* [cap] */
  int64_t* segment_ids;         /* [cap] */
  int64_t* sem_out;             /* [cap] labels_out / label_divisor (nullable) */
  int64_t* inst_out;            /* [cap] labels_out % label_divisor (nullable) */
  int32_t* dst;                 /* [batch * n] pixel -> row or -1 */
  int32_t* img_off;             /* [batch + 1] */
  int32_t* kmeans_labels;       /* [cap] */
  int32_t* seed_out;            /* [cap] */
  int32_t* num_segments;        /* [1] */
  /* The one host read-back of a step, inside the call (nullable: nothing is synchronised):
     counts_host[0..2] = {rows kept, segments, *status}; *status (device word, nullable) is
     reset to 0 once read.  The call then returns after `stream` has finished. */
  int32_t* counts_host;         /* HOST pointer, [4] */
  int32_t* status;              /* device word kernels OR error bits into */
  int32_t* counts_dev;          /* device scratch [4] for the read-back (with counts_host) */
} spml_cluster_args;

size_t spml_segment_by_kmeans_workspace_bytes(int batch, int n, int dim_total, int num_clusters,
                                              int iterations);
int spml_segment_by_kmeans(const spml_cluster_args* args, void* workspace,
                           size_t workspace_bytes, void* stream);

/* B1 for ids fresh from A8. spml/models/utils.py:95-116: the ids are already the dense
 *     ranks of (image, cluster, label), so both re-numberings are the identity; what is
 *     left are the per-segment labels (p_sem / p_inst / p_batch [m], nullable; every pixel
 *     of a segment must carry the same values: a violation sets bit 0 of status[0], an id
 *     outside [0, m) bit 1; the word is never cleared here) and the two prototype sets
 *     (el / protos_loc nullable).  bwd: de / del from dprotos / dprotos_loc (either NULL).
 *     workspace: spml_gather_prototypes_workspace_bytes(m, dim, dim_loc).
 */
size_t spml_gather_prototypes_workspace_bytes(int64_t m, int dim, int dim_loc);
int spml_gather_prototypes_fwd(const float* e, const float* el, int64_t rows, int dim, int dim_loc,
                               const int64_t* seg, const int64_t* batch, const int64_t* sem,
                               const int64_t* inst, int64_t m, float eps, float* protos,
                               float* protos_loc, float* norms, float* norms_loc, int64_t* p_sem,
                               int64_t* p_inst, int64_t* p_batch, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream);
int spml_gather_prototypes_bwd(const float* dprotos, const float* dprotos_loc, const float* protos,
                               const float* protos_loc, const float* norms, const float* norms_loc,
                               const int64_t* seg, int64_t rows, int dim, int dim_loc, int64_t m,
                               float eps, float* de, float* del, void* stream);

/* C4. losses() of spml/models/predictions/segsort.py:127-243, segsort_softmax.py:133-242
 *     and segsort_softmax_densepose.py:134-252 (everything but the conv classifier):
 *     memory-bank concatenation, tag masks, the labelled-pixel / labelled-prototype filter
 *     of sem_ann, per-image prototypes and row groups of img_sim, the three SegSort
 *     losses and the top-5 retrieval accuracy.
 *       enable: bit 0 sem_ann, 1 sem_occ, 2 img_sim, 3 accuracy.
 *       tags:   nn_tags == 0: image-tag columns [tag_col0, tag_col1) of img_tags
 *               [tag_rows, img_tags_ld] (row = the pixel's batch index) and of ptags /
 *               bank_tags (one row per prototype);  nn_tags == 1 (DensePose): 1-NN
 *               propagation over protos_loc / bank_protos_loc within an image (pbid /
 *               bank_pbid), threshold nn_threshold, no tag -> every tag.
 *       img_sim_on_plain: img_sim runs on e (DensePose) instead of el.
 *       max_groups: upper bound on the images the rows span (bid[last] - bid[0] + 1); a
 *               violation sets bit 2 of status[0] (nullable).
 *       wide_tags: more than 32 tag columns (keeps sem_occ off the tcgen05 kernels, whose
 *               epilogue compares 32-bit codes).
 *     fwd writes out[0..4] = {w_ann sem_ann, w_occ sem_occ, w_sim img_sim, accuracy, the
 *     sum of the enabled losses in that order (train.py:213-219)} and
 *     leaves what the backward needs in `state` (spml_head_workspace_bytes(args),
 *     256-byte aligned, untouched between the two calls).  bwd takes the incoming gradients
 *     of the three losses and of their sum as device scalars (NULL = zero; g_total counts
 *     for every enabled loss) and writes de [n, dim], del [n, dim_loc]
 *     (nullable when unused) and dprotos [m, dim] (nullable; memory-bank prototypes are
 *     detached, train.py:280).
 */
typedef struct spml_head_args {
  const float* e;               /* [n, dim]      cluster_embedding */
  const float* el;              /* [n, dim_loc]  cluster_embedding_with_loc (nullable) */
  const int64_t* seg;           /* [n] cluster_index (ids in [0, m)) */
  const int64_t* bid;           /* [n] cluster_batch_index */
  const int64_t* sem;           /* [n] cluster_semantic_label */
  const int64_t* inst;          /* [n] cluster_instance_label */
  const float* protos;          /* [m, dim]      prototype */
  const float* protos_loc;      /* [m, dim_loc]  prototype_with_loc (nn_tags only) */
  const int64_t* psem;          /* [m] */
  const int64_t* pinst;         /* [m] nullable */
  const int64_t* pbid;          /* [m] */
  const int64_t* img_tags;      /* [tag_rows, img_tags_ld] */
  const int64_t* ptags;         /* [m, ptags_ld] */
  int64_t img_tags_ld, ptags_ld, tag_rows;
  int64_t n, m, num_classes;
  int64_t max_rows_per_group;   /* upper bound on the pixels of one image (0: n) */
  int32_t dim, dim_loc;
  int32_t tag_col0, tag_col1;
  int32_t num_bank, max_groups;
  uint32_t enable;
  int32_t nn_tags, img_sim_on_plain, wide_tags;
  float kappa_ann, kappa_occ, kappa_sim;
  float weight_ann, weight_occ, weight_sim;
  float nn_threshold, eps;
  int32_t* status;              /* device word, nullable */
  const float* bank_protos[SPML_MAX_BANK];
  const float* bank_protos_loc[SPML_MAX_BANK];
  const int64_t* bank_psem[SPML_MAX_BANK];
  const int64_t* bank_pbid[SPML_MAX_BANK];
  const int64_t* bank_tags[SPML_MAX_BANK];
  int64_t bank_tags_ld[SPML_MAX_BANK];
  int64_t bank_m[SPML_MAX_BANK];
} spml_head_args;

size_t spml_head_workspace_bytes(const spml_head_args* args);
int spml_head_fwd(const spml_head_args* args, void* state, size_t state_bytes, float* out,
                  void* stream);
int spml_head_bwd(const spml_head_args* args, void* state, size_t state_bytes,
                  const float* g_ann, const float* g_occ, const float* g_sim,
                  const float* g_total, float* de, float* del, float* dprotos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPML_B200_H_ */
