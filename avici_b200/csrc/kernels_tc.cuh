// bf16 tensor-core kernels (sm_100a): tcgen05.mma with TMEM accumulators, operands staged in
// shared memory by TMA, warp-specialised producer / MMA-issuer / epilogue roles.
//
//   gemm_tc_kernel       C = A @ Wt^T + bias  (QKV / out / FFN projections, model.py:59-64,69-73)
//   attention_tc_kernel  flash-style axial attention for one (sequence, head, 128-query tile)
//                        work item at a time, both attended axes via one 4-D tensor map
//   ln_bf16_kernel       LayerNorm statistics + normalise to bf16 (affine folded into the GEMM)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_f32.cuh"
#include "tc_ptx.cuh"

namespace avb {

// ------------------------------------------------------------------------------------------
// LayerNorm (no affine) fp32 -> bf16, one warp per token of 128 channels.
// ------------------------------------------------------------------------------------------
__global__ void ln_bf16_kernel(const float* __restrict__ z, size_t ntok, __nv_bfloat16* __restrict__ xhat) {
  const int lane = threadIdx.x & 31;
  size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t t = warp; t < ntok; t += nwarps) {
    const float4 v = reinterpret_cast<const float4*>(z + t * 128)[lane];
    const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
    const float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
    const float var = warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.0f / 128.0f);
    const float r = rsqrtf(var + kLnEps);
    uint2 o;
    o.x = ptx::pack_bf16(c.x * r, c.y * r);
    o.y = ptx::pack_bf16(c.z * r, c.w * r);
    reinterpret_cast<uint2*>(xhat + t * 128)[lane] = o;
  }
}

// ------------------------------------------------------------------------------------------
// Persistent warp-specialised GEMM:  C[M,N] = A[M,K] @ Wt[N,K]^T + bias[N]
//   A, Wt: bf16, K-major, fetched by TMA in 128B-swizzled [rows x 64] boxes.
//   tile 128 x BN, BK = 64; accumulators double-buffered in TMEM (2 x BN columns).
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-7: epilogue
//   EPI 0: bf16 store   1: ReLU + bf16 store   2: fp32 residual add in place (out += acc + bias)
// ------------------------------------------------------------------------------------------
constexpr int kGemmThreads = 256;
template <int BN> struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const float* __restrict__ bias, void* __restrict__ out, int ldo, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * Cfg::kStageBytes);
  uint64_t* full = bars;            // [ST]
  uint64_t* empty = bars + ST;      // [ST]
  uint64_t* tfull = bars + 2 * ST;  // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128, num_n = N / BN, num_k = K / 64;
  const int num_tiles = num_m * num_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < ST; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 2 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          ptx::mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          ptx::tma_load_2d(sa, &tm_a, &full[stage], kb * 64, m_blk * 128);
          ptx::tma_load_2d(sa + Cfg::kABytes, &tm_b, &full[stage], kb * 64, n_blk * BN);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        ptx::mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&full[stage], phase, 3);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc(sa + k * 32, 16, 1024, ptx::kSwz128);
            const uint64_t db = ptx::make_smem_desc(sb + k * 32, 16, 1024, ptx::kSwz128);
            ptx::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty[stage]);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int as = it & 1;
      ptx::mbar_wait(&tfull[as], (it >> 1) & 1, 4);
      ptx::tc_fence_after();
      const int row = m_blk * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c * 32, r);
        ptx::tmem_wait_ld();
        const int col0 = n_blk * BN + c * 32;
        if (row < M) {
          if (EPI == 2) {
            float* op = reinterpret_cast<float*>(out) + (size_t)row * ldo + col0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float4 o = *reinterpret_cast<float4*>(op + e * 4);
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0) + e);
              o.x += __uint_as_float(r[e * 4 + 0]) + bb.x;
              o.y += __uint_as_float(r[e * 4 + 1]) + bb.y;
              o.z += __uint_as_float(r[e * 4 + 2]) + bb.z;
              o.w += __uint_as_float(r[e * 4 + 3]) + bb.w;
              *reinterpret_cast<float4*>(op + e * 4) = o;
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * ldo + col0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0) + e * 2 + u);
                v[u * 4 + 0] = __uint_as_float(r[e * 8 + u * 4 + 0]) + bb.x;
                v[u * 4 + 1] = __uint_as_float(r[e * 8 + u * 4 + 1]) + bb.y;
                v[u * 4 + 2] = __uint_as_float(r[e * 8 + u * 4 + 2]) + bb.z;
                v[u * 4 + 3] = __uint_as_float(r[e * 8 + u * 4 + 3]) + bb.w;
              }
              if (EPI == 1) {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
              }
              uint4 o;
              o.x = ptx::pack_bf16(v[0], v[1]); o.y = ptx::pack_bf16(v[2], v[3]);
              o.z = ptx::pack_bf16(v[4], v[5]); o.w = ptx::pack_bf16(v[6], v[7]);
              *reinterpret_cast<uint4*>(op + e * 8) = o;
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[as]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------------------
// Flash-style axial attention on tensor cores.  head dim 32, 128-query x 128-key tiles.
//   qkv viewed through ONE 4-D tensor map {3*H*32, d, n, B}; a work item (s, h, qt) loads its
//   [128 x 32] Q tile and streams [128 x 32] K / V tiles with a box of {32,128,1,1} (attend over d)
//   or {32,1,128,1} (attend over n): the attended axis is only a different TMA stride, z is never
//   transposed.  Rows beyond the sequence end are zero-filled by TMA and masked in the softmax.
//   S = Q K^T : M=128, N<=128, K=32   (A,B K-major, 64B swizzle)        -> TMEM S[2] (fp32)
//   P = exp2(S - m) (q is pre-scaled by log2e/sqrt(32)) -> bf16 smem (128B swizzle)
//   O += P V : M=128, N=32, K<=128    (A K-major smem, B = V MN-major)  -> TMEM O[2] (fp32)
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc
//   warps 4-7: softmax, O rescale, epilogue (one query row per thread = one TMEM lane)
// ------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 256;
constexpr int kAttnKvStages = 4;
constexpr int kAttnTileBytes = 128 * 32 * 2;  // 8 KB: Q, K or V tile
constexpr int kAttnPBytes = 128 * 128 * 2;    // 32 KB
constexpr int kAttnSmemBytes = 2 * kAttnTileBytes + kAttnKvStages * 2 * kAttnTileBytes + 2 * kAttnPBytes + 1024 + 256;

struct AttnStep {  // one (item, kv tile) step of a CTA's work list
  long long item;  // global item index, -1 = end
  int it;          // per-CTA item iteration
  int j;           // kv tile within the item
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int B, int n,
                    int d, int H, int axis) {
  constexpr int NS = kAttnKvStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                   // [2][8K]
  uint8_t* sKV = sQ + 2 * kAttnTileBytes;               // [NS][K 8K | V 8K]
  uint8_t* sP = sKV + NS * 2 * kAttnTileBytes;          // [2][32K]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttnPBytes);
  uint64_t* q_full = bars;             // [2]
  uint64_t* q_empty = bars + 2;        // [2]
  uint64_t* kv_full = bars + 4;        // [NS]
  uint64_t* kv_empty = bars + 4 + NS;  // [NS]
  uint64_t* s_full = bars + 4 + 2 * NS;  // [2]
  uint64_t* p_full = s_full + 2;       // [2]
  uint64_t* pv_done = p_full + 2;      // [2]
  uint64_t* o_empty = pv_done + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = axis == 0 ? d : n;
  const long long S = axis == 0 ? (long long)B * n : (long long)B * d;
  const int nqt = (T + 127) / 128;  // query tiles == kv tiles per sequence
  const long long num_items = S * H * nqt;
  const int HD = H * 32;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 128);
      ptx::mbar_init(&pv_done[i], 1); ptx::mbar_init(&o_empty[i], 128);
    }
    for (int i = 0; i < NS; ++i) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 32 columns

  // item -> (sequence, head, query tile); TMA coordinates of a [128 x 32] tile of that sequence
  auto decode = [&](long long item, long long& s, int& h, int& qt) {
    qt = (int)(item % nqt);
    const long long sh = item / nqt;
    h = (int)(sh % H);
    s = sh / H;
  };

  if (warp == 0) {
    if (lane == 0) {
      long long g = 0;
      int it = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        long long s; int h, qt;
        decode(item, s, h, qt);
        int c1, c2, c3;  // coordinates along d, n, B; the attended one is filled per tile
        if (axis == 0) { c3 = (int)(s / n); c2 = (int)(s % n); c1 = 0; }
        else           { c3 = (int)(s / d); c1 = (int)(s % d); c2 = 0; }
        const int qb = it & 1;
        ptx::mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1, 10);
        ptx::mbar_expect_tx(&q_full[qb], kAttnTileBytes);
        if (axis == 0) ptx::tma_load_4d(sQ + qb * kAttnTileBytes, &tm_qkv, &q_full[qb], h * 32, qt * 128, c2, c3);
        else           ptx::tma_load_4d(sQ + qb * kAttnTileBytes, &tm_qkv, &q_full[qb], h * 32, c1, qt * 128, c3);
        for (int j = 0; j < nqt; ++j, ++g) {
          const int st = (int)(g % NS);
          ptx::mbar_wait(&kv_empty[st], (uint32_t)((g / NS) & 1) ^ 1, 11);
          uint8_t* sk = sKV + st * 2 * kAttnTileBytes;
          ptx::mbar_expect_tx(&kv_full[st], 2 * kAttnTileBytes);
          if (axis == 0) {
            ptx::tma_load_4d(sk, &tm_qkv, &kv_full[st], HD + h * 32, j * 128, c2, c3);
            ptx::tma_load_4d(sk + kAttnTileBytes, &tm_qkv, &kv_full[st], 2 * HD + h * 32, j * 128, c2, c3);
          } else {
            ptx::tma_load_4d(sk, &tm_qkv, &kv_full[st], HD + h * 32, c1, j * 128, c3);
            ptx::tma_load_4d(sk + kAttnTileBytes, &tm_qkv, &kv_full[st], 2 * HD + h * 32, c1, j * 128, c3);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // number of key columns the MMAs of kv tile j touch (multiple of 16)
      auto kcols = [&](int j) { int v = T - j * 128; v = v > 128 ? 128 : v; return (v + 15) & ~15; };
      auto issue_S = [&](long long g, int it, int j) {
        if (j == 0) { ptx::mbar_wait(&q_full[it & 1], (it >> 1) & 1, 20); }
        const int st = (int)(g % NS);
        ptx::mbar_wait(&kv_full[st], (uint32_t)((g / NS) & 1), 21);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::make_idesc_bf16(128, kcols(j), 0, 0);
        const uint32_t sq = ptx::smem_u32(sQ + (it & 1) * kAttnTileBytes);
        const uint32_t sk = ptx::smem_u32(sKV + st * 2 * kAttnTileBytes);
        const uint32_t dS = tmem_S + (uint32_t)(g & 1) * 128;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t da = ptx::make_smem_desc(sq + k * 32, 16, 512, ptx::kSwz64);
          const uint64_t db = ptx::make_smem_desc(sk + k * 32, 16, 512, ptx::kSwz64);
          ptx::umma_bf16(dS, da, db, idesc, k != 0);
        }
        ptx::umma_commit(&s_full[g & 1]);
      };
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
      long long g = 0;
      int it = 0;
      long long item = blockIdx.x;
      if (item < num_items) issue_S(0, 0, 0);
      while (item < num_items) {
        for (int j = 0; j < nqt; ++j, ++g) {
          // look ahead: S of the next step overlaps this step's softmax
          if (j + 1 < nqt) issue_S(g + 1, it, j + 1);
          else if (item + gridDim.x < num_items) issue_S(g + 1, it + 1, 0);
          if (j == 0 && it >= 2) { ptx::mbar_wait(&o_empty[it & 1], (uint32_t)(((it >> 1) - 1) & 1), 22); }
          ptx::mbar_wait(&p_full[g & 1], (uint32_t)((g >> 1) & 1), 23);
          ptx::tc_fence_after();
          const int st = (int)(g % NS);
          const uint32_t sp = ptx::smem_u32(sP + (g & 1) * kAttnPBytes);
          const uint32_t sv = ptx::smem_u32(sKV + st * 2 * kAttnTileBytes + kAttnTileBytes);
          const uint32_t dO = tmem_O + (uint32_t)(it & 1) * 32;
          const int ksteps = kcols(j) >> 4;
          for (int k = 0; k < ksteps; ++k) {
            // P: [128 x 128] as two 128B-swizzled [128 x 64] K blocks; V: [keys x 32], 16 keys per step
            const uint64_t da = ptx::make_smem_desc(sp + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024, ptx::kSwz128);
            const uint64_t db = ptx::make_smem_desc(sv + k * (16 * 64), 16, 512, ptx::kSwz64);
            ptx::umma_bf16(dO, da, db, idesc_pv, (j | k) != 0);
          }
          ptx::umma_commit(&kv_empty[st]);
          ptx::umma_commit(&pv_done[g & 1]);
          if (j == nqt - 1) ptx::umma_commit(&q_empty[it & 1]);
        }
        item += gridDim.x;
        ++it;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    long long g = 0;
    int it = 0;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      long long s; int h, qt;
      decode(item, s, h, qt);
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < nqt; ++j, ++g) {
        const int sb = (int)(g & 1);
        ptx::mbar_wait(&s_full[sb], (uint32_t)((g >> 1) & 1), 30);
        ptx::tc_fence_after();
        uint32_t sv[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) ptx::tmem_ld32(tmem_S + sb * 128 + c * 32 + lane_off, sv[c]);
        ptx::tmem_wait_ld();
        const int nvalid = T - j * 128;  // >= 1
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float v = __uint_as_float(sv[c][e]);
            v = (c * 32 + e < nvalid) ? v : -INFINITY;
            sv[c][e] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
        const float m_new = fmaxf(m, mx);
        const float alpha = fast_exp2(m - m_new);
        if (j > 0) {
          // rescale the running output once every PV MMA up to the previous step has landed
          ptx::mbar_wait(&pv_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1), 31);
          ptx::tc_fence_after();
          uint32_t o[32];
          ptx::tmem_ld32(tmem_O + (it & 1) * 32 + lane_off, o);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          ptx::tmem_st32(tmem_O + (it & 1) * 32 + lane_off, o);
          ptx::tmem_wait_st();
        }
        if (g >= 2) { ptx::mbar_wait(&pv_done[sb], (uint32_t)(((g >> 1) - 1) & 1), 32); }  // P[sb] free
        float sum = 0.f;
        uint8_t* prow = sP + sb * kAttnPBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c16 = 0; c16 < 16; ++c16) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = fast_exp2(__uint_as_float(sv[c16 >> 2][(c16 & 3) * 8 + e]) - m_new);
          }
          uint4 pk;
          pk.x = ptx::pack_bf16(p[0], p[1]); pk.y = ptx::pack_bf16(p[2], p[3]);
          pk.z = ptx::pack_bf16(p[4], p[5]); pk.w = ptx::pack_bf16(p[6], p[7]);
          sum += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
          *reinterpret_cast<uint4*>(prow + (c16 >> 3) * (128 * 128) + (((c16 & 7) ^ (r & 7)) << 4)) = pk;
        }
        l = l * alpha + sum;
        m = m_new;
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_full[sb]);
      }
      // epilogue: O / l -> bf16 -> out[token, h*32 ..]
      const long long g_last = g - 1;
      ptx::mbar_wait(&pv_done[g_last & 1], (uint32_t)((g_last >> 1) & 1), 33);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld32(tmem_O + (it & 1) * 32 + lane_off, o);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&o_empty[it & 1]);
      const int t = qt * 128 + r;
      if (t < T) {
        const long long tok = axis == 0 ? s * d + t : (s / d) * (long long)n * d + (long long)t * d + (s % d);
        const float inv = 1.0f / l;
        uint4* op = reinterpret_cast<uint4*>(out + (size_t)tok * HD + h * 32);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 pk;
          pk.x = ptx::pack_bf16(__uint_as_float(o[e * 8 + 0]) * inv, __uint_as_float(o[e * 8 + 1]) * inv);
          pk.y = ptx::pack_bf16(__uint_as_float(o[e * 8 + 2]) * inv, __uint_as_float(o[e * 8 + 3]) * inv);
          pk.z = ptx::pack_bf16(__uint_as_float(o[e * 8 + 4]) * inv, __uint_as_float(o[e * 8 + 5]) * inv);
          pk.w = ptx::pack_bf16(__uint_as_float(o[e * 8 + 6]) * inv, __uint_as_float(o[e * 8 + 7]) * inv);
          op[e] = pk;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace avb
