// Graph metrics on the device: the step AFTER the forward, over a whole batch of
// predicted edge-probability matrices [B, d, d] that already sit in HBM.
//   reference: avici/metrics.py:5-24 (shd), :27-32 (n_edges), :35-43 (is_acyclic), :54-87 (classification_metrics),
//   called per dataset from the eval loop avici/train.py:125-151 on numpy copies.
// Everything here is integer work except the acyclicity test, which the reference DEFINES through an fp64 matrix
// power (trace((I + G/d)^d) - d isclose 0): a cycle longer than ~12 edges at d = 100 falls below its 1e-8 tolerance
// and is reported acyclic.  To return the reference's answer, not the graph-theoretic one, the same fp64 expression is
// evaluated here, with numpy's matrix_power multiplication order (the host loop in avici_b200.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace avb {

constexpr int kGmFields = 8;  // shd, n_edges_pred, n_edges_true, tp, fp, fn, cyclic, reserved (AVB_GM_* in the header)

// One CTA per dataset.  pred = probs > threshold (train.py:63); counts over all d*d entries (diagonal included, as the
// reference's flattened arrays do); shd over the upper triangle INCLUDING the diagonal (onp.triu, metrics.py:24).
__global__ void __launch_bounds__(256) graph_counts_kernel(const float* __restrict__ probs, const int32_t* __restrict__ g_true,
                                                           float threshold, int d, long long* __restrict__ out) {
  const int b = blockIdx.x;
  const size_t base = (size_t)b * d * d;
  const float* p = probs + base;
  const int32_t* g = g_true ? g_true + base : nullptr;
  long long c[6] = {0, 0, 0, 0, 0, 0};  // shd, n_pred, n_true, tp, fp, fn
  const int total = d * d;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = e / d, j = e - i * d;
    const int pij = p[e] > threshold ? 1 : 0;
    const int gij = g ? g[e] : 0;
    c[1] += pij;
    c[2] += gij;
    c[3] += (pij != 0 && gij != 0);
    c[4] += (pij != 0 && gij == 0);
    c[5] += (pij == 0 && gij != 0);
    if (g && j >= i) {  // mistakes = |g - h| + |g - h|^T, clipped at 1, upper triangle (metrics.py:17-24)
      const int e2 = j * d + i;
      const int pji = p[e2] > threshold ? 1 : 0;
      const int m = abs(gij - pij) + abs(g[e2] - pji);
      c[0] += m > 1 ? 1 : m;
    }
  }
  __shared__ long long red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    long long v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    long long v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    out[(size_t)b * kGmFields + threadIdx.x] = v;
  }
  if (threadIdx.x == 6) out[(size_t)b * kGmFields + 7] = 0;
}

// mat = I + G/d in fp64 (metrics.py:41), G = probs > threshold
__global__ void acyc_init_kernel(const float* __restrict__ probs, float threshold, int d, size_t total, double* __restrict__ mat) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int dd = d * d, r = (int)(e % dd), i = r / d, j = r - i * d;
  const double g = probs[e] > threshold ? 1.0 : 0.0;
  mat[e] = (i == j ? 1.0 : 0.0) + g / (double)d;
}

// C[b] = A[b] . B[b], fp64, [d, d] row-major, 16x16 output tile per CTA (grid: tiles x tiles x batch).
// The matrices are 80 KB each at d = 100 and live in L2; this is launch- and latency-bound, not a roofline kernel.
__global__ void __launch_bounds__(256) bmm_f64_kernel(const double* __restrict__ A, const double* __restrict__ Bm,
                                                      double* __restrict__ Cm, int d) {
  __shared__ double sa[16][17], sb[16][17];
  const size_t base = (size_t)blockIdx.z * d * d;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < d; k0 += 16) {
    sa[ty][tx] = (row < d && k0 + tx < d) ? A[base + (size_t)row * d + k0 + tx] : 0.0;
    sb[ty][tx] = (k0 + ty < d && col < d) ? Bm[base + (size_t)(k0 + ty) * d + col] : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fma(sa[ty][k], sb[k][tx], acc);
    __syncthreads();
  }
  if (row < d && col < d) Cm[base + (size_t)row * d + col] = acc;
}

// cyclic = !isclose(trace(mat_pow) - d, 0)  with numpy's defaults rtol 1e-5, atol 1e-8 against b = 0: |x| <= 1e-8;
// a non-finite trace is "not close", i.e. cyclic (metrics.py:42-43, :46-51).  One warp per dataset.
__global__ void acyc_trace_kernel(const double* __restrict__ mat_pow, int d, int B, long long* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const double* m = mat_pow + (size_t)b * d * d;
  double tr = 0.0;  // numpy's trace sums the diagonal in index order; the lanes' partial sums differ from that only
                    // in rounding, far below the 1e-8 tolerance unless the trace itself is astronomically large
  for (int i = lane; i < d; i += 32) tr += m[(size_t)i * d + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  if (lane == 0) {
    const double x = tr - (double)d;
    const bool acyclic = isfinite(x) && fabs(x) <= 1e-8;
    out[(size_t)b * kGmFields + 6] = acyclic ? 0 : 1;
  }
}

}  // namespace avb
