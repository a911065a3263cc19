// Declarations shared by the two k-means kernels (kmeans.cu: fp32 CUDA-core E-step and the
// host entry points; kmeans_tc.cu: tcgen05 E-step).
#pragma once

#include "tile_gemm.cuh"

namespace spml {

struct KmeansArgs {
  const float* x;            // [rows, dim]
  const int32_t* img_off;    // [batch + 1] or nullptr (one image of `rows_total` rows)
  int64_t rows_total;
  int batch, tiles_per_img;
  int dim, dpad;
  int num_clusters;          // stride of the per-image prototype arrays
  const int32_t* k_per_image;
  long long* sums;           // [iterations][replicas][batch][K][dim] fixed point, zeroed
  int replicas;              // copies of the sums (CTA c adds into copy c % replicas)
  float* protos;             // [iterations][batch][K][dim] unit prototypes (published per image)
  unsigned* done;            // [iterations][batch] tiles that have accumulated, zeroed
  unsigned* ready;           // [iterations][batch] prototypes published, zeroed
  const float* protos_in;    // A5: ready prototypes [K, dim] instead of sums
  int* poison;
  int iterations;
  const int32_t* labels_in;  // initial labels
  int32_t* labels_out;       // nullable
  int64_t* labels_out64;     // nullable
  float eps;
};

struct Tile {
  int b;
  int64_t row0;
  int rows;
};

__device__ __forceinline__ bool tile_of(const KmeansArgs& p, int t, Tile& tile) {
  tile.b = t / p.tiles_per_img;
  const int64_t first = p.img_off ? p.img_off[tile.b] : 0;
  const int64_t last = p.img_off ? p.img_off[tile.b + 1] : p.rows_total;
  tile.row0 = first + (int64_t)(t % p.tiles_per_img) * BM;
  tile.rows = (int)min((int64_t)BM, last - tile.row0);
  return tile.row0 < last;
}

// value of a 2^-32 fixed-point sum with 32-bit conversions only
__device__ __forceinline__ float fixed_to_float(long long s) {
  const int hi = (int)(s >> 32);
  const unsigned lo = (unsigned)(s & 0xffffffffll);
  return fmaf((float)lo, 2.3283064365386963e-10f, (float)hi);
}


// round(v * 2^32) = hi * 2^16 + lo for |v| <= 8 on the FMA pipe only (no F2I / FRND, which
// run at 1/8 rate): adding 1.5 * 2^23 rounds to the nearest-even integer and leaves it in the
// low mantissa bits.  Same integers as rintf / (int).
__device__ __forceinline__ void split_fixed(float v, int& hi, int& lo) {
  constexpr float kMagic = 12582912.f;
  const float t = fmaf(v, 65536.f, kMagic);
  hi = __float_as_int(t) - 0x4B400000;
  const float u = fmaf(v, 65536.f, kMagic - t) * 65536.f;   // exact remainder, |u| <= 2^15
  lo = __float_as_int(u + kMagic) - 0x4B400000;
}

constexpr int kAccSlots = (SPML_MAX_DIM + 31) / 32;
constexpr int kKmReplicas = 2;   // copies of the segment sums in the tensor-core path

// kmeans_small.cu (K <= 128: prototypes rebuilt from the sums inside every CTA)
bool kmeans_small_supported(int dim, int num_clusters, int batch, int64_t rows);
int kmeans_small_launch(const KmeansArgs& p, int sms, cudaStream_t st);

// kmeans_cluster.cu (an image fits in the shared memory of one thread-block cluster)
bool kmeans_cluster_supported(int dim, int num_clusters, int batch, int max_rows);
int kmeans_cluster_launch(const KmeansArgs& p, int max_rows, cudaStream_t st);

// kmeans_tc.cu
bool kmeans_tc_supported(int dim);
size_t kmeans_tc_split_bytes(int batch, int num_clusters, int dim, int iterations);
int kmeans_tc_launch(const KmeansArgs& p, void* split_protos, int sms, cudaStream_t st);

}  // namespace spml
