// fp32 SIMT kernels of the AVICI forward (sm_100a).
//
//  * prologue / epilogue kernels shared by both compute modes: standardizers
//    (avici/utils/data_jax.py:48-59,72-79), embedding Linear (avici/model.py:95), final LayerNorm +
//    max over n (model.py:86-89), u/v heads + cosine logits + sigmoid (model.py:121-141,218-239);
//  * the true-fp32 block kernels of the parity mode (AVB_DTYPE_F32): LayerNorm-fused tiled SGEMM
//    (model.py:56-58,68-73) and a flash-style fp32 axial attention (hk.MultiHeadAttention at
//    model.py:59-64) that never materialises the T x T logits.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace avb {

constexpr float kLnEps = 1e-5f;  // hk.LayerNorm default eps

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// Standardizers
// ------------------------------------------------------------------------------------------

// z-score per (dataset, variable) over the n axis, population std, std==0 -> 1
// (data_jax.py:72-79).  grid (ceil(d/32), B), block (32, 8): threadIdx.x walks variables
// (coalesced), threadIdx.y strides observations; statistics in fp64.
__global__ void zscore_stats_kernel(const float* __restrict__ x, int n, int d,
                                    float* __restrict__ mean_out, float* __restrict__ std_out) {
  __shared__ double s_a[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * n * d;
  double acc = 0.0;
  if (j < d)
    for (int i = threadIdx.y; i < n; i += 8) acc += (double)xb[(size_t)i * d + j];
  s_a[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  double mean = 0.0;
  for (int r = 0; r < 8; ++r) mean += s_a[r][threadIdx.x];
  mean /= (double)n;
  __syncthreads();
  acc = 0.0;
  if (j < d)
    for (int i = threadIdx.y; i < n; i += 8) {
      double t = (double)xb[(size_t)i * d + j] - mean;
      acc += t * t;
    }
  s_a[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < d) {
    double var = 0.0;
    for (int r = 0; r < 8; ++r) var += s_a[r][threadIdx.x];
    float stdv = (float)sqrt(var / (double)n);
    mean_out[(size_t)b * d + j] = (float)mean;
    std_out[(size_t)b * d + j] = (stdv == 0.0f) ? 1.0f : stdv;
  }
}

__global__ void zscore_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ stdv, int n, int d, size_t total,
                                    float* __restrict__ xs) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int j = (int)(idx % d);
  size_t b = idx / ((size_t)n * d);
  xs[idx] = (x[idx] - mean[b * d + j]) / stdv[b * d + j];  // std == 0 was replaced by 1
}

// log2-CPM per cell (data_jax.py:10-17): one warp per observation row; zero entries -> NaN.
__global__ void cpm_kernel(const float* __restrict__ x, size_t rows, int d, float* __restrict__ out) {
  size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * d;
  float lib = 0.f;
  for (int j = lane; j < d; j += 32) lib += xr[j];
  lib = warp_sum(lib);
  for (int j = lane; j < d; j += 32) {
    float v = xr[j];
    // jnp.isclose(x, 0.0): |x| <= atol (1e-8)
    out[row * d + j] = (fabsf(v) <= 1e-8f) ? __int_as_float(0x7fc00000) : log2f(v / (lib * 1e-6f));
  }
}

// Per-dataset nan-min and nan-std over all cells (data_jax.py:53-55).  One block per dataset.
// stats[b*2+0] = shift (0 when NaN), stats[b*2+1] = scale (1 when NaN).
__global__ void cpm_stats_kernel(const float* __restrict__ cpm, size_t per_ds,
                                 float* __restrict__ stats) {
  __shared__ double s_sum[32], s_cnt[32];
  __shared__ float s_min[32];
  __shared__ double s_mean;
  const float* p = cpm + (size_t)blockIdx.x * per_ds;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double sum = 0.0, cnt = 0.0;
  float mn = INFINITY;
  for (size_t i = threadIdx.x; i < per_ds; i += blockDim.x) {
    float v = p[i];
    if (v == v) { sum += v; cnt += 1.0; mn = fminf(mn, v); }
  }
  sum = warp_sum_d(sum); cnt = warp_sum_d(cnt); mn = -warp_max(-mn);
  if (lane == 0) { s_sum[wid] = sum; s_cnt[wid] = cnt; s_min[wid] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0, C = 0; float M = INFINITY;
    for (int w = 0; w < nw; ++w) { S += s_sum[w]; C += s_cnt[w]; M = fminf(M, s_min[w]); }
    s_mean = (C > 0) ? S / C : 0.0;
    s_sum[0] = C; s_min[0] = M;
  }
  __syncthreads();
  const double mean = s_mean, C = s_sum[0];
  const float M = s_min[0];
  __syncthreads();
  double sq = 0.0;
  for (size_t i = threadIdx.x; i < per_ds; i += blockDim.x) {
    float v = p[i];
    if (v == v) { double t = (double)v - mean; sq += t * t; }
  }
  sq = warp_sum_d(sq);
  if (lane == 0) s_sum[wid] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double Q = 0; for (int w = 0; w < nw; ++w) Q += s_sum[w];
    // all-NaN dataset: nanmin/nanstd are NaN -> shift 0, scale 1 (data_jax.py:22)
    stats[blockIdx.x * 2 + 0] = (C > 0) ? M : 0.0f;
    stats[blockIdx.x * 2 + 1] = (C > 0) ? (float)sqrt(Q / C) : 1.0f;
  }
}

__global__ void cpm_apply_kernel(float* __restrict__ cpm, const float* __restrict__ stats,
                                 size_t per_ds, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  size_t b = idx / per_ds;
  float v = (cpm[idx] - stats[b * 2]) / stats[b * 2 + 1];
  cpm[idx] = (v == v) ? v : 0.0f;  // NaN -> 0 (data_jax.py:26-27)
}

// ------------------------------------------------------------------------------------------
// Embedding: z[t,c] = xs[t]*W[0,c] + iv[t]*W[1,c] + b[c]     (model.py:95), dim = 128.
// One warp per token, float4 per lane; write-bound (512 B per token).
// ------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ xs, const float* __restrict__ iv,
                             const float* __restrict__ w, const float* __restrict__ bias,
                             size_t ntok, float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const float4 w0 = reinterpret_cast<const float4*>(w)[lane];
  const float4 w1 = reinterpret_cast<const float4*>(w + 128)[lane];
  const float4 bb = reinterpret_cast<const float4*>(bias)[lane];
  size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t t = warp; t < ntok; t += nwarps) {
    const float a = xs[t];
    const float m = iv ? iv[t] : 0.0f;
    float4 o;
    o.x = fmaf(m, w1.x, fmaf(a, w0.x, bb.x));
    o.y = fmaf(m, w1.y, fmaf(a, w0.y, bb.y));
    o.z = fmaf(m, w1.z, fmaf(a, w0.z, bb.z));
    o.w = fmaf(m, w1.w, fmaf(a, w0.w, bb.w));
    reinterpret_cast<float4*>(z + t * 128)[lane] = o;
  }
}

// ------------------------------------------------------------------------------------------
// Final LayerNorm + max over n  (model.py:86-89).  grid (d, B), 8 warps; each warp normalises
// one token at a time (float4 per lane) and keeps a running max for its 4 channels.
// ------------------------------------------------------------------------------------------
__global__ void ln_max_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int n, int d,
                              float* __restrict__ pooled, int accumulate) {
  __shared__ float4 s_m[8][32];
  const int j = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float4 g = reinterpret_cast<const float4*>(gamma)[lane];
  const float4 be = reinterpret_cast<const float4*>(beta)[lane];
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int i = wid; i < n; i += 8) {
    const float4 v = reinterpret_cast<const float4*>(z + (((size_t)b * n + i) * d + j) * 128)[lane];
    float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
    float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
    float var = warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.0f / 128.0f);
    float r = 1.0f / sqrtf(var + kLnEps);
    m.x = fmaxf(m.x, fmaf(c.x * r, g.x, be.x));
    m.y = fmaxf(m.y, fmaf(c.y * r, g.y, be.y));
    m.z = fmaxf(m.z, fmaf(c.z * r, g.z, be.z));
    m.w = fmaxf(m.w, fmaf(c.w * r, g.w, be.w));
  }
  s_m[wid][lane] = m;
  __syncthreads();
  if (wid == 0) {
    for (int w = 1; w < 8; ++w) {
      float4 o = s_m[w][lane];
      m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
    }
    float4* dst = reinterpret_cast<float4*>(pooled + ((size_t)b * d + j) * 128) + lane;
    if (accumulate) {
      float4 o = *dst;
      m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
    }
    *dst = m;
  }
}

// ------------------------------------------------------------------------------------------
// Head (model.py:121-141): per pooled row, LN_u/LN_v -> Linear -> L2-normalise (no eps).
// One block of 128 threads per (dataset, variable) row; thread c owns output channel c.
// dim = 128, out_dim <= 128 handled by guarding c < out_dim.
// ------------------------------------------------------------------------------------------
__global__ void head_uv_kernel(const float* __restrict__ pooled,
                               const float* __restrict__ gu, const float* __restrict__ bu,
                               const float* __restrict__ wu, const float* __restrict__ cu,
                               const float* __restrict__ gv, const float* __restrict__ bv,
                               const float* __restrict__ wv, const float* __restrict__ cv,
                               int out_dim, float* __restrict__ u_out, float* __restrict__ v_out) {
  __shared__ float s_xu[128], s_xv[128], s_red[8];
  const size_t row = blockIdx.x;
  const int c = threadIdx.x, lane = c & 31, wid = c >> 5;
  const float x = pooled[row * 128 + c];
  float s = warp_sum(x);
  if (lane == 0) s_red[wid] = s;
  __syncthreads();
  const float mean = (s_red[0] + s_red[1] + s_red[2] + s_red[3]) * (1.0f / 128.0f);
  const float xc = x - mean;
  float q = warp_sum(xc * xc);
  if (lane == 0) s_red[4 + wid] = q;
  __syncthreads();
  const float var = (s_red[4] + s_red[5] + s_red[6] + s_red[7]) * (1.0f / 128.0f);
  const float r = 1.0f / sqrtf(var + kLnEps);
  s_xu[c] = fmaf(xc * r, gu[c], bu[c]);
  s_xv[c] = fmaf(xc * r, gv[c], bv[c]);
  __syncthreads();
  float au = 0.f, av = 0.f;
  if (c < out_dim) {
    au = cu[c]; av = cv[c];
    for (int k = 0; k < 128; ++k) {
      au = fmaf(s_xu[k], wu[(size_t)k * out_dim + c], au);
      av = fmaf(s_xv[k], wv[(size_t)k * out_dim + c], av);
    }
  }
  float su = warp_sum(au * au), sv = warp_sum(av * av);
  __syncthreads();
  if (lane == 0) { s_red[wid] = su; s_red[4 + wid] = sv; }
  __syncthreads();
  const float nu = sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  const float nv = sqrtf(s_red[4] + s_red[5] + s_red[6] + s_red[7]);
  if (c < out_dim) {
    u_out[row * out_dim + c] = au / nu;
    v_out[row * out_dim + c] = av / nv;
  }
}

// p[b,i,j] = exp(log_sigmoid(<u_i,v_j> * exp(temp) + bias)), diag = 0 (model.py:135-141,218-239); with log_probs the
// log_sigmoid itself, diag = -inf (infer_edge_logprobs, model.py:218-220).
// grid (ceil(d/16), ceil(d/16), B), block (16,16).
__global__ void head_logits_kernel(const float* __restrict__ u, const float* __restrict__ v, int d,
                                   int out_dim, const float* __restrict__ temp,
                                   const float* __restrict__ bias, int mask_diag, int log_probs,
                                   float* __restrict__ probs) {
  __shared__ float s_u[16][129], s_v[16][129];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const int tid = threadIdx.y * 16 + threadIdx.x;
  for (int e = tid; e < 16 * out_dim; e += 256) {
    int r = e / out_dim, c = e % out_dim;
    s_u[r][c] = (i0 + r < d) ? u[((size_t)b * d + i0 + r) * out_dim + c] : 0.f;
    s_v[r][c] = (j0 + r < d) ? v[((size_t)b * d + j0 + r) * out_dim + c] : 0.f;
  }
  __syncthreads();
  const int i = i0 + threadIdx.y, j = j0 + threadIdx.x;
  if (i >= d || j >= d) return;
  float acc = 0.f;
  for (int c = 0; c < out_dim; ++c) acc = fmaf(s_u[threadIdx.y][c], s_v[threadIdx.x][c], acc);
  const float l = acc * expf(temp[0]) + bias[0];
  // exp(log_sigmoid(l)) with log_sigmoid(l) = -(max(-l,0) + log1p(exp(-|l|)))
  const float logp = -(fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l))));
  float p = log_probs ? logp : expf(logp);
  if (mask_diag && i == j) p = log_probs ? -INFINITY : 0.f;
  probs[((size_t)b * d + i) * d + j] = p;
}

// ------------------------------------------------------------------------------------------
// fp32 tiled SGEMM with optional exact LayerNorm prologue and bias/ReLU/residual epilogue.
//   C[m, n] = epi( sum_k LN?(A)[m,k] * W[k,n] + bias[n] )          W: [K,N] row-major (Haiku layout)
//   LN (K must be 128): a' = (a - mean_m) * rsqrt(var_m + eps) * gamma[k] + beta[k]
//   EPI 0: store, 1: relu + store, 2: C = R + val (R may alias C).
// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------
template <int EPI, bool LN>
__global__ void __launch_bounds__(256) sgemm_kernel(
    const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
    const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* R, int ldr, float* C, int ldc, size_t M, int N, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64];
  __shared__ float s_mean[64], s_rstd[64];
  const int tid = threadIdx.x;
  const size_t m0 = (size_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;

  if (LN) {
    // 4 threads per row, each 32 strided elements of the K=128 row
    const int r = tid >> 2, q = tid & 3;
    const size_t m = m0 + r;
    float vals[32];
    float s = 0.f;
    if (m < M) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float4 t = *reinterpret_cast<const float4*>(A + m * lda + (e * 4 + q) * 4);
        vals[e * 4 + 0] = t.x; vals[e * 4 + 1] = t.y; vals[e * 4 + 2] = t.z; vals[e * 4 + 3] = t.w;
        s += t.x + t.y + t.z + t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) vals[e] = 0.f;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float mean = s * (1.0f / 128.0f);
    float qv = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) { float t = vals[e] - mean; qv = fmaf(t, t, qv); }
    qv += __shfl_xor_sync(0xffffffffu, qv, 1);
    qv += __shfl_xor_sync(0xffffffffu, qv, 2);
    if (q == 0) { s_mean[r] = mean; s_rstd[r] = 1.0f / sqrtf(qv * (1.0f / 128.0f) + kLnEps); }
    __syncthreads();
  }

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ar = tid >> 2, ak = (tid & 3) * 4;   // A tile: row ar, k offset ak..ak+3
  const int bk = tid >> 4, bn = (tid & 15) * 4;  // B tile: k row bk, n offset bn..bn+3
  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + ar < M) a = *reinterpret_cast<const float4*>(A + (m0 + ar) * lda + k0 + ak);
    if (LN) {
      const float mu = s_mean[ar], rs = s_rstd[ar];
      const float4 g = *reinterpret_cast<const float4*>(gamma + k0 + ak);
      const float4 be = *reinterpret_cast<const float4*>(beta + k0 + ak);
      a.x = fmaf((a.x - mu) * rs, g.x, be.x);
      a.y = fmaf((a.y - mu) * rs, g.y, be.y);
      a.z = fmaf((a.z - mu) * rs, g.z, be.z);
      a.w = fmaf((a.w - mu) * rs, g.w, be.w);
    }
    As[ak + 0][ar] = a.x; As[ak + 1][ar] = a.y; As[ak + 2][ar] = a.z; As[ak + 3][ar] = a.w;
    *reinterpret_cast<float4*>(&Bs[bk][bn]) =
        *reinterpret_cast<const float4*>(W + (size_t)(k0 + bk) * ldw + n0 + bn);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
    if (EPI == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (EPI == 2) {
      const float4 r = *reinterpret_cast<const float4*>(R + m * ldr + n0 + tx * 4);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4*>(C + m * ldc + n0 + tx * 4) = o;
  }
}

// ------------------------------------------------------------------------------------------
// fp32 flash-style axial attention (hk.MultiHeadAttention core, model.py:59-64), key_size 32.
//   qkv: [ntok, 3*H*32] fp32 (q | k | v, head-major columns c = h*32 + e); out: [ntok, H*32].
//   sequence s, position t  ->  token  (s / inner) * outer_stride + (s % inner) * inner_stride
//                                       + t * tok_stride
//   (axis d: inner = 1 -> token = s*d + t;   axis n: inner = d -> token = b*n*d + j + t*d)
//   dg > 0 (column-blocked n-shard of avb_forward_sharded, axis d): token = (t / dg) * blk_stride + s * dg + t % dg
// grid (ceil(T/128), H, S); 128 threads, one query per thread; keys streamed through smem in
// tiles of 32 with an online softmax (tile-wise rescale).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_f32_kernel(
    const float* __restrict__ qkv, float* __restrict__ out, int T, int H, long long s_off, long long inner,
    long long outer_stride, long long inner_stride, long long tok_stride, float scale, int dg, long long blk_stride) {
  __shared__ float4 Ks[32][8];
  __shared__ float4 Vs[32][8];
  const int h = blockIdx.y;
  const long long s = s_off + blockIdx.z;
  const long long base = dg > 0 ? s * dg : (s / inner) * outer_stride + (s % inner) * inner_stride;
  auto tok_of = [&](int t) -> long long {
    return dg > 0 ? base + (long long)(t / dg) * blk_stride + (t % dg) : base + (long long)t * tok_stride;
  };
  const int ld = 3 * H * 32;
  const int tq = blockIdx.x * 128 + threadIdx.x;
  const bool valid = tq < T;

  float q[32], o[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) { q[e] = 0.f; o[e] = 0.f; }
  if (valid) {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (size_t)tok_of(tq) * ld + h * 32);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float4 t = qp[e];
      q[e * 4 + 0] = t.x * scale; q[e * 4 + 1] = t.y * scale; q[e * 4 + 2] = t.z * scale; q[e * 4 + 3] = t.w * scale;
    }
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < T; k0 += 32) {
    // cooperative K/V tile load: 32 keys x 8 float4, 256 float4 per matrix, 2 per thread
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int idx = threadIdx.x + rep * 128;
      const int kr = idx >> 3, kc = idx & 7;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + kr < T) {
        const float* rowp = qkv + (size_t)tok_of(k0 + kr) * ld + h * 32;
        kk = reinterpret_cast<const float4*>(rowp + H * 32)[kc];
        vv = reinterpret_cast<const float4*>(rowp + 2 * H * 32)[kc];
      }
      Ks[kr][kc] = kk;
      Vs[kr][kc] = vv;
    }
    __syncthreads();
    const int nk = min(32, T - k0);
    float sc[32];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 kk = Ks[j][e];
        a = fmaf(q[e * 4 + 0], kk.x, a); a = fmaf(q[e * 4 + 1], kk.y, a);
        a = fmaf(q[e * 4 + 2], kk.z, a); a = fmaf(q[e * 4 + 3], kk.w, a);
      }
      sc[j] = (j < nk) ? a : -INFINITY;
      tmax = fmaxf(tmax, sc[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float alpha = expf(m - m_new);  // m = -inf on the first tile -> 0
    l *= alpha;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] *= alpha;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float p = expf(sc[j] - m_new);  // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 vv = Vs[j][e];
        o[e * 4 + 0] = fmaf(p, vv.x, o[e * 4 + 0]); o[e * 4 + 1] = fmaf(p, vv.y, o[e * 4 + 1]);
        o[e * 4 + 2] = fmaf(p, vv.z, o[e * 4 + 2]); o[e * 4 + 3] = fmaf(p, vv.w, o[e * 4 + 3]);
      }
    }
    m = m_new;
    __syncthreads();
  }
  if (valid) {
    const float inv = 1.0f / l;
    float4* op = reinterpret_cast<float4*>(out + (size_t)tok_of(tq) * (H * 32) + h * 32);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      op[e] = make_float4(o[e * 4 + 0] * inv, o[e * 4 + 1] * inv, o[e * 4 + 2] * inv, o[e * 4 + 3] * inv);
  }
}

}  // namespace avb
