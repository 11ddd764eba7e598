// TMA row-rate micro-benchmark: how many cycles does one SM's TMA unit need per box row, as a function
// of the row width (64 B / 128 B) and of the gather pattern used by the axial attention kernel?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench_tma scripts/microbench_tma.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../avici_b200/csrc/tc_ptx.cuh"
using namespace avb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one thread issues `iters` loads of one box each into a ring of `stages` buffers
__global__ void tma_bench(const __grid_constant__ CUtensorMap tm, int iters, int box_bytes, int axis, int n, int d,
                          int cstep, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  constexpr int ST = 4;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int st = i % ST;
      if (i >= ST) ptx::mbar_wait(&bars[st], ((i / ST) - 1) & 1);
      ptx::mbar_expect_tx(&bars[st], box_bytes);
      const int seq = (blockIdx.x * 131 + i * 7);
      const int c0 = ((i * 5) % 12) * cstep;
      if (axis == 0) ptx::tma_load_4d(smem + st * box_bytes, &tm, &bars[st], c0, 0, seq % n, 0);
      else           ptx::tma_load_4d(smem + st * box_bytes, &tm, &bars[st], c0, seq % d, ((i * 3) % 7) * 128, 0);
    }
    for (int i = iters; i < iters + ST; ++i) ptx::mbar_wait(&bars[i % ST], ((i / ST) - 1) & 1);
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  void* fnp; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int B = 4, n = 1000, d = 100, C = 768;
  void* buf; cudaMalloc(&buf, (size_t)B * n * d * C * 2); cudaMemset(buf, 0, (size_t)B * n * d * C * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  for (int axis = 0; axis < 2; ++axis)
    for (int w : {32, 64, 128, 256}) {  // box inner elements (bf16): 64 B, 128 B, 256 B, 512 B rows
      CUtensorMap tm;
      cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)d, (cuuint64_t)n, (cuuint64_t)B};
      cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)d * C * 2, (cuuint64_t)n * d * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)w, axis == 0 ? 128u : 1u, axis == 0 ? 1u : 128u, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUtensorMapSwizzle sw = w == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (w == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("axis %d w %d: encode failed %d\n", axis, w, (int)r); continue; }
      const int iters = 4000, box_bytes = w * 2 * 128;
      tma_bench<<<148, 32, 4 * box_bytes>>>(tm, iters, box_bytes, axis, n, d, w, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("axis=%d row=%4d B: %.0f cycles/box (128 rows) = %.2f cycles/row, %.1f B/clk/SM, agg %.2f TB/s @1.9GHz (%s)\n", axis, w * 2,
             c / iters, c / iters / 128, box_bytes / (c / iters), box_bytes / (c / iters) * 148 * 1.9e9 / 1e12, cudaGetErrorString(e));
    }
  return 0;
}
