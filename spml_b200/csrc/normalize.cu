// A8 front half: valid-pixel scan, NCHW -> packed NHWC rows with L2 normalisation and
// location-feature concatenation (reference spml/utils/segsort/common.py:306-310,
// 339-365), and its backward.  HBM-bound: every embedding element is read once
// (coalesced along the pixel axis), transposed through shared memory and written
// once (coalesced along the channel axis).
#include "common.cuh"

namespace spml {

#ifndef SPML_PACK_TILE
#define SPML_PACK_TILE 64
#endif
constexpr int kTile = SPML_PACK_TILE;   // pixels per CTA (also the scan granularity)
constexpr int kTileLd = kTile + 1;  // odd stride: conflict-free in both directions
constexpr int kPackThreads = 256;

__device__ __forceinline__ bool pixel_kept(const int64_t* labels, int has_ignore,
                                           int64_t ignore_index, int64_t pix) {
  // has_ignore: 0 keeps everything, 1 drops label == ignore_index, 2 keeps label < ignore_index
  if (!has_ignore) return true;
  const int64_t v = labels[pix];
  return has_ignore == 2 ? v < ignore_index : v != ignore_index;
}

// the ignore index may live on the device (generate_clusters passes labels.max() + 1)
__device__ __forceinline__ int64_t resolve_ignore(int64_t host_value, const int64_t* dev_value) {
  return dev_value ? *dev_value : host_value;
}

// Exclusive prefix of a tile's count over all tiles with a smaller ticket: every thread of
// the CTA polls the aggregates of some of the preceding tiles (state[t] = {flag:32 | count:32},
// published by publish_tile_count) and the CTA adds them up.  A FLAT look-back: the classic
// decoupled look-back walks the chain tile by tile from one thread, ~0.1 us per tile, which
// is 12 us for the 128 tiles of one 512 x 512 image; here a tile waits for the slowest of its
// predecessors plus one L2 round trip.  Tiles take their number from a ticket, so every
// predecessor is already running (or done) and publishes without waiting for anybody.
// Cost: tile t reads t words, T^2 / 2 in total - 0.5 M words at batch 4 of 128 x 128 maps (1 024
// tiles), nothing next to the embedding traffic; beyond ~10^4 tiles per call a windowed variant
// that stops at the first published inclusive prefix would be the thing to write.
__device__ __forceinline__ void publish_tile_count(unsigned long long* state, int tile, int total) {
  // one 64-bit word: flag and count arrive together
  reinterpret_cast<volatile unsigned long long*>(state)[tile] = (1ull << 32) | (unsigned)total;
}
__device__ __forceinline__ int tile_prefix(unsigned long long* state, int tile, int* s_red) {
  volatile unsigned long long* vs = state;
  int sum = 0;
  for (int j = threadIdx.x; j < tile; j += blockDim.x) {
    unsigned long long s;
    do {
      s = vs[j];
    } while ((unsigned)(s >> 32) == 0);
    sum += (int)(unsigned)(s & 0xffffffffu);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  int prefix = 0;
  for (int w = 0; w < nwarps; ++w) prefix += s_red[w];
  return prefix;
}

__global__ void valid_scan_kernel(const int64_t* __restrict__ labels, int has_ignore,
                                  int64_t ignore_host, const int64_t* __restrict__ ignore_dev,
                                  int batch, int n, int tiles_per_img,
                                  int32_t* __restrict__ dst, int32_t* __restrict__ src,
                                  int32_t* __restrict__ img_off, unsigned long long* state,
                                  int* ticket) {
  __shared__ int s_tile;
  __shared__ int s_warp[kTile / 32];
  __shared__ int s_red[kTile / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int64_t ignore_index = has_ignore ? resolve_ignore(ignore_host, ignore_dev) : 0;
  const int b = tile / tiles_per_img;
  const int p = (tile % tiles_per_img) * kTile + threadIdx.x;
  const int64_t pix = (int64_t)b * n + p;
  const bool keep = p < n && pixel_kept(labels, has_ignore, ignore_index, pix);
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kTile / 32; ++w) {
    if (w < warp) before += s_warp[w];
    total += s_warp[w];
  }
  if (threadIdx.x == 0) publish_tile_count(state, tile, total);
  const int prefix = tile_prefix(state, tile, s_red);
  if (threadIdx.x == 0) {
    if (tile % tiles_per_img == 0) img_off[b] = prefix;
    if (tile == batch * tiles_per_img - 1) img_off[batch] = prefix + total;
  }
  if (p < n) {
    const int row = keep ? prefix + before + __popc(ballot & ((1u << lane) - 1)) : -1;
    dst[pix] = row;
    if (src && keep) src[row] = (int32_t)pix;
  }
}

struct PackArgs {
  const float* emb;          // [batch, dim, n]
  const float* loc;          // [.., n, loc_ch]
  int64_t loc_batch_stride;
  int loc_ch;
  const int64_t* seeds;
  int64_t seed_batch_stride;
  int dim, n, tiles_per_img;
  int64_t batch_index_offset;
  float eps;
  float* e;                  // [rows, dim]
  float* el;                 // [rows, dim + loc_ch]
  float* nx;
  float* nc;
  int64_t* labels_out;
  int64_t* batch_out;
  int32_t* seed_out;
};

// the tile's embedding columns -> shared memory [dim][kTileLd] (coalesced along the pixels)
__device__ __forceinline__ void pack_load_tile(const PackArgs& a, int b, int p0, int np, float* tile) {
  const int tid = threadIdx.x;
  const int px = tid % kTile;
  for (int d = tid / kTile; d < a.dim; d += kPackThreads / kTile)
    tile[d * kTileLd + px] = px < np ? a.emb[((int64_t)b * a.dim + d) * a.n + p0 + px] : 0.f;
}

// normalise the tile's pixels, concatenate the location channels, normalise again and write the
// kept pixels to their rows (s_row[px], -1 = dropped); s_lab[px] = the pixel's label.  Needs a
// __syncthreads() between pack_load_tile and this.
__device__ __forceinline__ void pack_normalize_store(const PackArgs& a, int b, int p0, int np,
                                                     float* tile, const int* s_row,
                                                     const int64_t* s_lab, float* s_nx, float* s_nc) {
  const int tid = threadIdx.x;
  const int dim = a.dim, loc_ch = a.loc_ch, dp = dim + loc_ch;
  const float eps = a.eps;
  if (tid < kTile) {
    float ss = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float v = tile[d * kTileLd + tid];
      ss += v * v;
    }
    const float n1 = sqrtf(ss);
    const bool ok1 = n1 >= eps;
    const float div1 = ok1 ? n1 : eps;
    float ss2 = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float v = tile[d * kTileLd + tid] / div1;
      tile[d * kTileLd + tid] = v;
      ss2 += v * v;
    }
    if (tid < np) {
      const float* lp = a.loc + (int64_t)b * a.loc_batch_stride + (int64_t)(p0 + tid) * loc_ch;
      for (int c = 0; c < loc_ch; ++c) ss2 += lp[c] * lp[c];
    }
    const float n2 = sqrtf(ss2);
    s_nx[tid] = ok1 ? n1 : -eps;
    s_nc[tid] = n2 >= eps ? n2 : -eps;
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5;
  for (int px = warp; px < np; px += kPackThreads / 32) {
    const int r = s_row[px];
    if (r < 0) continue;
    const float div2 = fabsf(s_nc[px]);
    for (int d = lane; d < dim; d += 32) {
      const float v = tile[d * kTileLd + px];
      a.e[(int64_t)r * dim + d] = v;
      a.el[(int64_t)r * dp + d] = v / div2;
    }
    if (lane < loc_ch)
      a.el[(int64_t)r * dp + dim + lane] =
          a.loc[(int64_t)b * a.loc_batch_stride + (int64_t)(p0 + px) * loc_ch + lane] / div2;
    if (lane == 0) {
      a.nx[r] = s_nx[px];
      a.nc[r] = s_nc[px];
      if (a.labels_out) a.labels_out[r] = s_lab[px];
      if (a.batch_out) a.batch_out[r] = b + a.batch_index_offset;
      if (a.seed_out) a.seed_out[r] = (int32_t)a.seeds[(int64_t)b * a.seed_batch_stride + p0 + px];
    }
  }
}

// emb [batch, dim, n] -> e [rows, dim], el [rows, dim+loc_ch]
__global__ void __launch_bounds__(kPackThreads)
normalize_pack_fwd_kernel(const PackArgs a, const int64_t* __restrict__ labels,
                          const int32_t* __restrict__ dst) {
  extern __shared__ float tile[];  // [dim][kTileLd]
  __shared__ float s_nx[kTile], s_nc[kTile];
  __shared__ int s_row[kTile];
  __shared__ int64_t s_lab[kTile];
  const int b = blockIdx.x / a.tiles_per_img;
  const int p0 = (blockIdx.x % a.tiles_per_img) * kTile;
  const int np = min(kTile, a.n - p0);
  const int tid = threadIdx.x;
  if (tid < kTile) {
    s_row[tid] = tid < np ? dst[(int64_t)b * a.n + p0 + tid] : -1;
    s_lab[tid] = (tid < np && a.labels_out) ? labels[(int64_t)b * a.n + p0 + tid] : 0;
  }
  pack_load_tile(a, b, p0, np, tile);
  __syncthreads();
  pack_normalize_store(a, b, p0, np, tile, s_row, s_lab, s_nx, s_nc);
}

// A8 front half in ONE kernel: label packing (sem * divisor + inst, ignored pixels dropped;
// resnet_deeplab.py:118-137), the valid-pixel scan and the normalise / pack above.  The
// embedding tile is requested before the CTA waits for the counts of the tiles in front of it,
// so the scan costs no time of its own.  Replaces pack_labels_kernel + valid_scan_kernel +
// normalize_pack_fwd_kernel (three dependent launches on the way to the k-means) in
// spml_segment_by_kmeans.
struct FusedLabelArgs {
  const int64_t* labels;      // ready labels, or nullptr: sem * divisor + inst
  int has_ignore;
  int64_t ignore_host;
  const int64_t* ignore_dev;
  const int64_t* sem;
  const int64_t* inst;
  int64_t divisor, semantic_ignore, dropped;
  int batch;
  int32_t* dst;
  int32_t* img_off;
  unsigned long long* state;
  int* ticket;
};

__global__ void __launch_bounds__(kPackThreads)
scan_normalize_pack_kernel(const PackArgs a, const FusedLabelArgs f) {
  extern __shared__ float tile[];  // [dim][kTileLd]
  __shared__ float s_nx[kTile], s_nc[kTile];
  __shared__ int s_row[kTile];
  __shared__ int64_t s_lab[kTile];
  __shared__ int s_tile;
  __shared__ int s_warp[kTile / 32];
  __shared__ int s_red[kPackThreads / 32];
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(f.ticket, 1);
  __syncthreads();
  const int t = s_tile;
  const int b = t / a.tiles_per_img;
  const int p0 = (t % a.tiles_per_img) * kTile;
  const int np = min(kTile, a.n - p0);
  bool keep = false;
  unsigned ballot = 0;
  if (tid < kTile) {   // whole warps
    const int64_t pix = (int64_t)b * a.n + p0 + tid;
    int64_t lab = 0;
    if (tid < np) {
      if (f.labels) {
        const int64_t ignore_index = f.has_ignore ? resolve_ignore(f.ignore_host, f.ignore_dev) : 0;
        lab = f.labels[pix];
        keep = !f.has_ignore || (f.has_ignore == 2 ? lab < ignore_index : lab != ignore_index);
      } else {
        const int64_t sv = f.sem[pix];
        lab = sv == f.semantic_ignore ? f.dropped : sv * f.divisor + f.inst[pix];
        keep = lab != f.dropped;
      }
    }
    s_lab[tid] = lab;
    ballot = __ballot_sync(0xffffffffu, keep);
    if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(ballot);
  }
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int w = 0; w < kTile / 32; ++w) total += s_warp[w];
  if (tid == 0) publish_tile_count(f.state, t, total);
  pack_load_tile(a, b, p0, np, tile);              // in flight while the predecessors publish
  const int prefix = tile_prefix(f.state, t, s_red);
  if (tid == 0) {
    if (t % a.tiles_per_img == 0) f.img_off[b] = prefix;
    if (t == f.batch * a.tiles_per_img - 1) f.img_off[f.batch] = prefix + total;
  }
  if (tid < kTile) {
    int before = 0;
#pragma unroll
    for (int w = 0; w < kTile / 32; ++w)
      if (w < (tid >> 5)) before += s_warp[w];
    const int row = keep ? prefix + before + __popc(ballot & ((1u << (tid & 31)) - 1)) : -1;
    s_row[tid] = row;
    if (tid < np) f.dst[(int64_t)b * a.n + p0 + tid] = row;
  }
  __syncthreads();
  pack_normalize_store(a, b, p0, np, tile, s_row, s_lab, s_nx, s_nc);
}

constexpr int kMaxPerLane = (SPML_MAX_DIM + 31) / 32;  // channel slots per lane

// d(emb) from d(e), d(el).  One warp per pixel for the dot products, then a
// transposed, pixel-coalesced store.
constexpr int kBwdTile = 32;            // pixels per CTA in the backward: 4 per warp
constexpr int kBwdTileLd = kBwdTile + 1;

__global__ void __launch_bounds__(kPackThreads)
normalize_pack_bwd_kernel(const float* __restrict__ de, const float* __restrict__ del,
                          const float* __restrict__ e, const float* __restrict__ el,
                          const float* __restrict__ nx, const float* __restrict__ nc,
                          const int32_t* __restrict__ dst, int dim, int loc_ch, int n,
                          int tiles_per_img, float eps, float* __restrict__ demb) {
  extern __shared__ float tile[];  // [dim][kBwdTileLd]
  const int b = blockIdx.x / tiles_per_img;
  const int p0 = (blockIdx.x % tiles_per_img) * kBwdTile;
  const int np = min(kBwdTile, n - p0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dp = dim + loc_ch;

  for (int px = warp; px < np; px += kPackThreads / 32) {
    const int r = dst[(int64_t)b * n + p0 + px];
    float g[kMaxPerLane], ev[kMaxPerLane];
#pragma unroll
    for (int s = 0; s < kMaxPerLane; ++s) g[s] = 0.f, ev[s] = 0.f;
    if (r >= 0) {
      // through el = cat(e, loc) / max(||cat||, eps)
      if (del) {
        float t1 = 0.f;
        for (int d = lane; d < dp; d += 32) t1 += el[(int64_t)r * dp + d] * del[(int64_t)r * dp + d];
        t1 = warp_sum(t1);
        const float n2 = nc[r];
#pragma unroll
        for (int s = 0; s < kMaxPerLane; ++s) {
          const int d = lane + 32 * s;
          if (d < dim) {
            const float gl = del[(int64_t)r * dp + d];
            g[s] = n2 > 0.f ? (gl - el[(int64_t)r * dp + d] * t1) / n2 : gl / eps;
          }
        }
      }
      float t2 = 0.f;
#pragma unroll
      for (int s = 0; s < kMaxPerLane; ++s) {
        const int d = lane + 32 * s;
        if (d < dim) {
          if (de) g[s] += de[(int64_t)r * dim + d];
          ev[s] = e[(int64_t)r * dim + d];
          t2 += ev[s] * g[s];
        }
      }
      t2 = warp_sum(t2);
      const float n1 = nx[r];
#pragma unroll
      for (int s = 0; s < kMaxPerLane; ++s)
        g[s] = n1 > 0.f ? (g[s] - ev[s] * t2) / n1 : g[s] / eps;
    }
#pragma unroll
    for (int s = 0; s < kMaxPerLane; ++s) {
      const int d = lane + 32 * s;
      if (d < dim) tile[d * kBwdTileLd + px] = g[s];
    }
  }
  __syncthreads();
  const int px = tid % kBwdTile;
  if (px < np)
    for (int d = tid / kBwdTile; d < dim; d += kPackThreads / kBwdTile)
      demb[((int64_t)b * dim + d) * n + p0 + px] = tile[d * kBwdTileLd + px];
}

static int tiles_per_image(int n) { return (int)ceil_div(n, kTile); }

}  // namespace spml

namespace spml {

// pipeline.cu: labels (or sem / inst) -> dst, img_off, packed rows in one launch
int scan_normalize_pack(const float* emb, const float* loc, int64_t loc_batch_stride, int loc_ch,
                        const int64_t* labels, int has_ignore, int64_t ignore_index,
                        const int64_t* ignore_dev, const int64_t* sem, const int64_t* inst,
                        int64_t divisor, int64_t semantic_ignore, int64_t dropped,
                        const int64_t* seeds, int64_t seed_batch_stride, int batch, int dim, int n,
                        int64_t batch_index_offset, float eps, int32_t* dst, int32_t* img_off,
                        float* e, float* el, float* nx, float* nc, int64_t* labels_out,
                        int64_t* batch_out, int32_t* seed_out, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
  SPML_CHECK_SUPPORTED(loc_ch >= 0 && loc_ch <= 32 && dim + loc_ch <= SPML_MAX_DIM,
                       "normalize_pack_fwd: dim %d + loc_ch %d exceeds %d", dim, loc_ch,
                       SPML_MAX_DIM);
  SPML_CHECK_SUPPORTED((int64_t)batch * n < (1ll << 31), "valid_scan: more than 2^31 pixels");
  const size_t need = spml_valid_scan_workspace_bytes(batch, n);
  if (workspace_bytes < need) {
    set_error("valid_scan: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  SPML_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  const int tpi = tiles_per_image(n);
  PackArgs pa{emb, loc, loc_batch_stride, loc_ch, seeds, seed_batch_stride, dim, n, tpi,
              batch_index_offset, eps, e, el, nx, nc, labels_out, batch_out, seed_out};
  FusedLabelArgs fa{labels, has_ignore, ignore_index, ignore_dev, sem, inst, divisor,
                    semantic_ignore, dropped, batch, dst, img_off,
                    reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 16),
                    reinterpret_cast<int*>(workspace)};
  const size_t smem = (size_t)dim * kTileLd * sizeof(float);
  SPML_CUDA(cudaFuncSetAttribute(scan_normalize_pack_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_normalize_pack_kernel<<<batch * tpi, kPackThreads, smem, st>>>(pa, fa);
  SPML_LAUNCH_CHECK("scan_normalize_pack_kernel");
  return SPML_OK;
}

}  // namespace spml

extern "C" {

size_t spml_valid_scan_workspace_bytes(int batch, int n) {
  if (batch <= 0 || n <= 0) return 16;
  const size_t tiles = (size_t)batch * spml::tiles_per_image(n);
  return 16 + tiles * sizeof(unsigned long long);
}

int spml_valid_scan(const int64_t* labels, int has_ignore, int64_t ignore_index,
                    const int64_t* ignore_index_dev, int batch, int n, int32_t* dst, int32_t* src, int32_t* img_off, void* workspace,
                    size_t workspace_bytes, void* stream) {
  SPML_CHECK_ARG(batch > 0 && n > 0 && dst && img_off && workspace,
                 "valid_scan: bad arguments");
  SPML_CHECK_ARG(!has_ignore || labels, "valid_scan: labels required with an ignore index");
  SPML_CHECK_SUPPORTED((int64_t)batch * n < (1ll << 31), "valid_scan: more than 2^31 pixels");
  const size_t need = spml_valid_scan_workspace_bytes(batch, n);
  if (workspace_bytes < need) {
    spml::set_error("valid_scan: workspace %zu < %zu bytes", workspace_bytes, need);
    return SPML_E_WORKSPACE;
  }
  cudaStream_t st = spml::as_stream(stream);
  const int tpi = spml::tiles_per_image(n);
  SPML_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  int* ticket = reinterpret_cast<int*>(workspace);
  unsigned long long* state =
      reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 16);
  spml::valid_scan_kernel<<<batch * tpi, spml::kTile, 0, st>>>(
      labels, has_ignore, ignore_index, ignore_index_dev, batch, n, tpi, dst, src, img_off, state,
      ticket);
  SPML_LAUNCH_CHECK("valid_scan_kernel");
  return SPML_OK;
}

int spml_normalize_pack_fwd(const float* emb, const float* loc, int64_t loc_batch_stride,
                            int loc_ch, const int64_t* labels, const int64_t* seeds,
                            int64_t seed_batch_stride, const int32_t* dst, int batch, int dim,
                            int n, int64_t batch_index_offset, float eps, float* e, float* el,
                            float* nx, float* nc, int64_t* labels_out, int64_t* batch_out,
                            int32_t* seed_out, void* stream) {
  SPML_CHECK_ARG(emb && dst && e && el && nx && nc && batch > 0 && n > 0 && dim > 0,
                 "normalize_pack_fwd: bad arguments");
  SPML_CHECK_ARG(loc_ch == 0 || loc, "normalize_pack_fwd: loc required when loc_ch > 0");
  SPML_CHECK_ARG(!labels_out || labels, "normalize_pack_fwd: labels_out needs labels");
  SPML_CHECK_ARG(!seed_out || seeds, "normalize_pack_fwd: seed_out needs seeds");
  SPML_CHECK_SUPPORTED(loc_ch >= 0 && loc_ch <= 32 && dim + loc_ch <= SPML_MAX_DIM,
                       "normalize_pack_fwd: dim %d + loc_ch %d exceeds %d", dim, loc_ch,
                       SPML_MAX_DIM);
  const size_t smem = (size_t)dim * spml::kTileLd * sizeof(float);
  SPML_CUDA(cudaFuncSetAttribute(spml::normalize_pack_fwd_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tpi = spml::tiles_per_image(n);
  spml::PackArgs pa{emb, loc, loc_batch_stride, loc_ch, seeds, seed_batch_stride, dim, n, tpi,
                    batch_index_offset, eps, e, el, nx, nc, labels_out, batch_out, seed_out};
  spml::normalize_pack_fwd_kernel<<<batch * tpi, spml::kPackThreads, smem,
                                    spml::as_stream(stream)>>>(pa, labels, dst);
  SPML_LAUNCH_CHECK("normalize_pack_fwd_kernel");
  return SPML_OK;
}

int spml_normalize_pack_bwd(const float* de, const float* del, const float* e, const float* el,
                            const float* nx, const float* nc, const int32_t* dst, int batch,
                            int dim, int loc_ch, int n, float eps, float* demb, void* stream) {
  SPML_CHECK_ARG(e && el && nx && nc && dst && demb && batch > 0 && n > 0 && dim > 0,
                 "normalize_pack_bwd: bad arguments");
  SPML_CHECK_SUPPORTED(loc_ch >= 0 && dim + loc_ch <= SPML_MAX_DIM,
                       "normalize_pack_bwd: dim %d + loc_ch %d exceeds %d", dim, loc_ch,
                       SPML_MAX_DIM);
  const size_t smem = (size_t)dim * spml::kBwdTileLd * sizeof(float);
  SPML_CUDA(cudaFuncSetAttribute(spml::normalize_pack_bwd_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tpi = (int)spml::ceil_div(n, spml::kBwdTile);
  spml::normalize_pack_bwd_kernel<<<batch * tpi, spml::kPackThreads, smem,
                                    spml::as_stream(stream)>>>(
      de, del, e, el, nx, nc, dst, dim, loc_ch, n, tpi, eps, demb);
  SPML_LAUNCH_CHECK("normalize_pack_bwd_kernel");
  return SPML_OK;
}

}  // extern "C"
