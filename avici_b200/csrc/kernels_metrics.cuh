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

// Threshold-free metrics of one prediction against the truth: auroc, auprc (trapezoid over the precision-recall curve)
// and average precision, as avici/metrics.py:90-127 computes them with sklearn (roc_curve + auc,
// precision_recall_curve + auc, average_precision_score), including its three corner cases.
// sklearn sorts the scores; here every quantity is written as a sum over the POSITIVE entries of counts over all
// entries, which needs no sort and no workspace (ties included; checked against sklearn 1.9.0 to 3e-15):
//   with, for a positive with score s:  TP>= / TP> = positives with score >= s / > s,  N>= / N> the same for negatives,
//   auroc = 1/(P N) sum_pos [ (N - N>=) + (N>= - N>)/2 ]                      (Mann-Whitney = trapezoid under the ROC curve)
//   ap    = 1/P     sum_pos  TP>=/(TP>= + N>=)                                (precision at the positive's own threshold)
//   auprc = 1/P     sum_pos [ TP>=/(TP>= + N>=) + prec> ]/2,  prec> = TP>/(TP> + N>) or 1 when nothing scores higher
//                                                                           (the curve's previous point; (0, 1) at the start)
// One CTA per dataset; thread t owns entries t, t + blockDim, ...; fp64 sums, fixed reduction order (deterministic).
// out[b] = {auroc, auprc, ap, 0}.
__global__ void __launch_bounds__(1024) threshold_metrics_kernel(const float* __restrict__ probs, const int32_t* __restrict__ g_true,
                                                                 int d, double* __restrict__ out) {
  const int b = blockIdx.x, total = d * d;
  const float* p = probs + (size_t)b * total;
  const int32_t* g = g_true + (size_t)b * total;
  __shared__ double red[3][32];
  __shared__ double tot[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  auto block_sum3 = [&](double a0, double a1, double a2) {
    double v[3] = {a0, a1, a2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if (lane == 0) red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double acc = 0.0;
      for (int w = 0; w < nwarp; ++w) acc += red[threadIdx.x][w];
      tot[threadIdx.x] = acc;
    }
    __syncthreads();
  };
  double npos = 0.0, psum = 0.0;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    npos += g[e] != 0 ? 1.0 : 0.0;
    psum += (double)p[e];
  }
  block_sum3(npos, psum, 0.0);
  const double P = tot[0], S = tot[1], N = (double)total - P;
  __syncthreads();
  if (!(S > 0.0 && P > 0.0)) {  // metrics.py:112-127
    if (threadIdx.x == 0) {
      const bool none = (S == 0.0 && P == 0.0);
      out[(size_t)b * 4 + 0] = none ? 1.0 : 0.5;
      out[(size_t)b * 4 + 1] = none ? 1.0 : 0.0;
      out[(size_t)b * 4 + 2] = none ? 1.0 : 0.0;
      out[(size_t)b * 4 + 3] = 0.0;
    }
    return;
  }
  double au = 0.0, pr = 0.0, ap = 0.0;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    if (g[e] == 0) continue;
    const float s = p[e];
    int tpge = 0, tpgt = 0, nge = 0, ngt = 0;
    for (int f = 0; f < total; ++f) {
      const float sf = __ldg(p + f);
      const int lab = __ldg(g + f) != 0;
      const int ge = sf >= s, gt = sf > s;
      tpge += ge & lab; tpgt += gt & lab;
      nge += ge & (lab ^ 1); ngt += gt & (lab ^ 1);
    }
    au += (N - (double)nge) + 0.5 * (double)(nge - ngt);
    const double pge = (double)tpge / (double)(tpge + nge);
    const double pgt = (tpgt + ngt) > 0 ? (double)tpgt / (double)(tpgt + ngt) : 1.0;
    ap += pge;
    pr += 0.5 * (pge + pgt);
  }
  block_sum3(au, pr, ap);
  if (threadIdx.x == 0) {
    out[(size_t)b * 4 + 0] = tot[0] / (P * N);  // N = 0: nan, as sklearn's 0/0 false-positive rate
    out[(size_t)b * 4 + 1] = tot[1] / P;
    out[(size_t)b * 4 + 2] = tot[2] / P;
    out[(size_t)b * 4 + 3] = 0.0;
  }
}

}  // namespace avb
