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

/* thread-local text of the last error on this thread ("" if none). */
const char* spml_last_error(void);
/* ABI version, bumped on any signature change. */
int spml_abi_version(void);
/* diagnostics: number of kernels this library has launched (all threads) since it
 * was loaded (memsets are not counted). */
uint64_t spml_debug_launch_count(void);

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
 *     number of kept pixels.  has_ignore == 0 keeps everything.  src (nullable,
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
 *     The E-step runs as a TMA-fed tcgen05 GEMM (bf16 hi/lo split, near-ties
 *     re-scored in fp32) or as an fp32 CUDA-core GEMM; both return the same ids
 *     bit for bit (the choice is made per call from the shape).
 *     labels_out (int32) and labels_out_i64 (nullable) receive the final ids.
 *     workspace: spml_kmeans_workspace_bytes(batch, num_clusters, dim, iterations),
 *     16-byte aligned; the call is one cooperative launch (all its CTAs must be
 *     able to be resident: it waits for the SMs of concurrently running kernels).
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
 */
int spml_topk_ranking(const float* q, int64_t nq, const float* p, int64_t m, int dim,
                      const int64_t* qlab, const int64_t* plab, const uint8_t* qvalid,
                      const uint8_t* pvalid, int k, int64_t* topk_labels,
                      int64_t* topk_index, int32_t* hit_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPML_B200_H_ */
