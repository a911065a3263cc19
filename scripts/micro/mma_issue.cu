// Microbenchmark: cost of back-to-back tcgen05.mma (kind::f16, M = 128, K = 16, SS) issued by
// one thread, as a function of N and of whether consecutive MMAs share the accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I spml_b200/csrc -o /tmp/mma_issue scripts/micro/mma_issue.cu
#include <cstdio>
#include "tc_common.cuh"
namespace spml { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } void count_launch() {} }
using namespace spml;

__global__ void __launch_bounds__(128, 1) bench(int n, int accs, int count, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::umma_idesc_bf16(128, n, 0, 0);
    const uint32_t hi = tc::umma_desc_hi_sw128(1024);
    const uint32_t a_lo = tc::umma_desc_lo(tc::smem_u32(smem), 16);
    const uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(smem + 16384), 16);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i)
      tc::umma_bf16_words(tmem + (i % accs) * 256, a_lo + (i & 3) * 2, hi, b_lo + (i & 3) * 2, hi, idesc, i >= accs);
    const long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int accs : {1, 2}) for (int n : {64, 128, 256}) {
    if (accs == 2 && n == 256) { /* 2 x 256 columns = 512: fits */ }
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<1, 128, 100 * 1024>>>(n, accs, 64, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (rep) printf("N=%3d accumulators=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", n, accs, h[0] / 64.0, h[1] / 64.0, cudaGetErrorString(e));
    }
  }
  return 0;
}
