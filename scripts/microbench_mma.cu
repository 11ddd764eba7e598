// tcgen05.mma issue/throughput micro-benchmark on B200: cycles per M=128 x N x K=16 bf16 MMA (A, B from smem) as a
// function of N, with all MMAs accumulating into ONE TMEM accumulator vs round-robin over several.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench_mma scripts/microbench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../avici_b200/csrc/tc_ptx.cuh"
using namespace avb;

__global__ void mma_bench(int N, int naccum, int iters, int a_mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u == 1) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
    const uint64_t da = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024, ptx::kSwz128);
    const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(smem) + 32768, 16, 1024, ptx::kSwz128);
    long long t0 = clock64();
    if (ptx::elect_one()) {  // back-to-back issue from one thread, as the kernels do: 8 MMAs per loop trip, fully unrolled
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = tmem + (uint32_t)(j & (naccum - 1)) * 64;
          if (a_mode == 0) ptx::umma_bf16(d, da + (uint64_t)((j & 3) * 2), db + (uint64_t)((j & 3) * 2), idesc, 1);
          else ptx::umma_bf16_ts(d, tmem + 448 + (uint32_t)((j & 3) * 8), db + (uint64_t)((j & 3) * 2), idesc, 1);  // A from TMEM
        }
      }
    }
    __syncwarp();
    if (ptx::elect_one()) ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 4096;
  for (int a_mode : {0, 1})
  for (int naccum : {1, 2, 4})
    for (int N : {16, 32, 48, 64, 96, 128, 256}) {
      if (naccum * 64 < N && naccum > 1 && N > 64) continue;
      if (a_mode == 1 && N > 64 && naccum > 1) continue;
      mma_bench<<<148, 64, 65536>>>(N, naccum, iters, a_mode, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      printf("%s M=128 N=%3d K=16, %d accumulator(s): %.1f cycles/MMA (ideal N/2 = %d) (%s)\n", a_mode ? "A=tmem" : "A=smem", N, naccum,
             (double)h[0] / iters, N / 2, cudaGetErrorString(e));
    }
  return 0;
}
