// Microbenchmark: cost of back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) as a function of
// N, of the A source (smem descriptor "SS" or TMEM "TS") and of HOW the single issuing thread is
// selected: `if (threadIdx.x == 0)` (the compiler wraps every UTCHMMA in an ELECT / R2UR /
// BRA.U.ANY waterfall loop because it cannot prove the operands warp-uniform) versus a
// warp-uniform loop with the MMAs under `elect.sync`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I spml_b200/csrc -o scripts/micro/mma_issue scripts/micro/mma_issue.cu
#include <cstdio>
#include "tc_common.cuh"
namespace spml { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } void count_launch() {} }
using namespace spml;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <bool kTs, bool kElect, int kN>
__global__ void __launch_bounds__(128, 1) bench(int count, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  constexpr uint32_t idesc = tc::umma_idesc_bf16(128, kN, 0, 0);
  const uint32_t hi = tc::umma_desc_hi_sw128(1024);
  const uint32_t a_lo = tc::umma_desc_lo(tc::smem_u32(smem), 16);
  const uint32_t b_lo = tc::umma_desc_lo(tc::smem_u32(smem + 16384), 16);
  const uint32_t a_tmem = tmem + 448;
  auto body = [&](int i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (kTs) tc::umma_bf16_ts_words(tmem, a_tmem + k * 8, b_lo + k * 2, hi, idesc, (i | k) != 0);
      else tc::umma_bf16_words(tmem, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, (i | k) != 0);
    }
  };
  if (kElect) {
    if (warp == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < count; i += 4) {
        if (elect_one()) body(i);
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (threadIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t0;
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < count; i += 4) body(i);
    const long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0, out[1] = t2 - t0;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <bool kTs, bool kElect, int kN>
void run(long long* d) {
  const int count = 256;
  cudaFuncSetAttribute(bench<kTs, kElect, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h[2];
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 2; ++rep) {
    bench<kTs, kElect, kN><<<1, 128, 100 * 1024>>>(count, d);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  }
  printf("%s %-22s N=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA, floor %3d (%s)\n", kTs ? "TS" : "SS",
         kElect ? "elect.sync uniform loop" : "if (threadIdx.x == 0)", kN, h[0] / (double)count,
         h[1] / (double)count, kN / 2, cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<false, false, 64>(d); run<false, false, 128>(d); run<false, false, 256>(d);
  run<false, true, 32>(d); run<false, true, 64>(d); run<false, true, 128>(d); run<false, true, 256>(d);
  run<true, false, 64>(d); run<true, false, 256>(d);
  run<true, true, 32>(d); run<true, true, 64>(d); run<true, true, 128>(d); run<true, true, 256>(d);
  return 0;
}
