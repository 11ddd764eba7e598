// Micro-benchmarks that settle design questions for the attention kernel on B200:
//   * tcgen05.ld throughput (TMEM -> registers) with 4 and 8 warps
//   * MUFU ex2 throughput: ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../avici_b200/csrc/tc_ptx.cuh"
using namespace avb;

__global__ void tmem_ld_bench(int iters, long long* cycles, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ptx::tmem_ld32(base + ((i * 4 + c) & 15) * 32, r);
      ptx::tmem_wait_ld();
      acc += r[0] ^ r[31];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

template <int MODE>
__global__ void ex2_bench(int iters, long long* cycles, float* sink) {
  float x[8];
  for (int e = 0; e < 8; ++e) x[e] = -0.001f * (threadIdx.x + e);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[e]));
      } else if (MODE == 2) {
        uint32_t u = __float_as_uint(x[e]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
        x[e] = __uint_as_float(u);
      } else if (MODE == 3) {  // packed fp32 FMA: 2 FMAs per instruction
        unsigned long long v = ((unsigned long long)__float_as_uint(x[e]) << 32) | __float_as_uint(x[e]);
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        x[e] = __uint_as_float((uint32_t)v);
      } else if (MODE == 4) {
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[e]));
      } else {
        uint32_t u = __float_as_uint(x[e]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u));
        x[e] = __uint_as_float(u);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0; for (int e = 0; e < 8; ++e) s += x[e];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* cyc; unsigned* sink; float* fs;
  cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&sink, 1 << 22); cudaMalloc(&fs, 1 << 22);
  long long h[148];
  const int iters = 2000;
  for (int threads : {128, 256}) {
    tmem_ld_bench<<<148, threads>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double bytes = (double)threads * 32 * 4 * 4 * iters;  // per CTA
    printf("tcgen05.ld 32x32b.x32: %d threads/CTA: %lld cycles for %.0f bytes -> %.1f B/clk/SM (%s)\n", threads, h[0], bytes,
           bytes / h[0], cudaGetErrorString(e));
  }
  for (int threads : {128, 256, 512}) {
    ex2_bench<0><<<148, threads>>>(iters, cyc, fs);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("ex2.f32:    %d threads/CTA: %.2f ex2/clk/SM\n", threads, (double)threads * 8 * iters / h[0]);
    ex2_bench<1><<<148, threads>>>(iters, cyc, fs);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("ex2.bf16x2: %d threads/CTA: %.2f instr/clk/SM = %.2f exps/clk/SM (%s)\n", threads, (double)threads * 8 * iters / h[0],
           2.0 * threads * 8 * iters / h[0], cudaGetErrorString(e));
    ex2_bench<2><<<148, threads>>>(iters, cyc, fs);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("ex2.f16x2:  %d threads/CTA: %.2f instr/clk/SM = %.2f exps/clk/SM (%s)\n", threads, (double)threads * 8 * iters / h[0],
           2.0 * threads * 8 * iters / h[0], cudaGetErrorString(e));
    ex2_bench<3><<<148, threads>>>(iters, cyc, fs);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("fma.f32x2:  %d threads/CTA: %.2f instr/clk/SM (%s)\n", threads, (double)threads * 8 * iters / h[0], cudaGetErrorString(e));
    ex2_bench<4><<<148, threads>>>(iters, cyc, fs);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("fma.f32:    %d threads/CTA: %.2f instr/clk/SM (%s)\n", threads, (double)threads * 8 * iters / h[0], cudaGetErrorString(e));
  }
  return 0;
}
