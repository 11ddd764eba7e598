// fp32-accurate GEMMs on the bf16 tensor cores for the parity (fp32) mode: the projections and the FFN of a block
// (model.py:59-64, 69-73) as split-precision products.
//
//     C[M, N] = epi( A'[M, K] . W[K, N] + bias ),   A' = LayerNorm(A) without affine (kLn) or A itself
//
// Every fp32 operand x is split into two bf16 numbers, x_hi = bf16(x) and x_lo = bf16(x - x_hi) (x = x_hi + x_lo up to
// 2^-17 |x|), and a product a.b becomes three tensor-core products accumulated in fp32 in TMEM:
//     a_hi b_hi + a_hi b_lo + a_lo b_hi          (the dropped a_lo b_lo is 2^-18 relative)
// i.e. ~1e-5 relative per product against 2e-3 for plain bf16 operands - the measured bf16-mode error of the forward
// (5e-4) scales to a few 1e-6, two orders inside the 1e-4 criterion of the fp32 mode (TF32's 2^-11 would not be: ~1.3e-4).
// LayerNorm statistics, the softmax and the residual stream stay in fp32 (the attention itself stays on the SIMT
// kernel).  The LayerNorm affine and the softmax scale are folded into W / bias in fp64 at load time, like in the bf16
// mode, so one normalised A tile serves the q, k and v projections.
//
// Per CTA and 128-row tile: four producer warps normalise / split the fp32 rows once into K/64 pairs of 128B-swizzled
// [128 x 64] bf16 tiles (hi, lo), which stay resident for all N/128 column chunks; the weight tiles (hi, lo, [128 x 64]
// each, pre-split at load time) stream through a TMA ring; 12 MMAs (M 128, N 128, K 16) per K block; accumulators
// double-buffered in TMEM; four epilogue warps add the bias, apply ReLU or the residual and write full 128-byte fp32
// segments, one row per thread.  K is 128 or 256 per launch (the 512-wide FFN down-projection runs as two launches, the
// second accumulating onto the first through the residual epilogue).
//   warp 0: weight TMA   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-7: A producers   warps 8-11: epilogue
#pragma once
#include "kernels_tc.cuh"

namespace avb {

constexpr int kSplitThreads = 384;
constexpr int kSplitBStages = 2;
constexpr int kSplitTile = 128 * 64 * 2;  // one [128 x 64] bf16 tile, 16 KB
template <int KB>
struct SplitCfg {
  static constexpr int kABytes = KB * 2 * kSplitTile;             // resident A: (hi, lo) per K block
  static constexpr int kBBytes = kSplitBStages * 2 * kSplitTile;  // weight ring: (hi, lo) per stage
  static constexpr int kSmemBytes = kABytes + kBBytes + 1024 + 256;
};

enum { kSplitEpiBias = 0, kSplitEpiRelu = 1, kSplitEpiResidual = 2 };

// 32 fp32 rows of width 64*KB (row pitch lda floats) -> hi / lo bf16 K-major tiles; eight lanes share a row (128-byte
// coalesced segments), optional LayerNorm without affine (exact two-pass statistics in fp32, K = 128 only).
template <int KB, bool kLn>
__device__ __forceinline__ void split_rows_to_tiles(const float* __restrict__ a, int lda, int M, int grow0, uint32_t a_tiles,
                                                    int row0, int lane) {
  constexpr int NK = 2 * KB;  // float4 per lane and row: columns 4 (j + 8 k)
  const int rs = lane >> 3, j = lane & 7;
#pragma unroll 1
  for (int r0 = 0; r0 < 32; r0 += 4) {
    const int grow = grow0 + r0 + rs;
    float4 v[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k)
      v[k] = grow < M ? __ldcg(reinterpret_cast<const float4*>(a + (size_t)grow * lda) + j + 8 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (kLn) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float mean = sum * (1.0f / (64 * KB));
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
        sq += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      const float rstd = rsqrtf(sq * (1.0f / (64 * KB)) + kLnEps);
#pragma unroll
      for (int k = 0; k < NK; ++k) { v[k].x *= rstd; v[k].y *= rstd; v[k].z *= rstd; v[k].w *= rstd; }
    }
    const int r = row0 + r0 + rs;
    const uint32_t rbase = a_tiles + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (j & 1) * 8);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      // columns 4 (j + 8k) .. +3: K block k >> 1, 16-byte chunk (j + 8 (k & 1)) >> 1 of its 128-byte row
      const int chunk = (j + 8 * (k & 1)) >> 1;
      const uint32_t addr = rbase + (uint32_t)((k >> 1) * (2 * kSplitTile) + ((chunk ^ (r & 7)) << 4));
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[k].x, v[k].y), h1 = __floats2bfloat162_rn(v[k].z, v[k].w);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const uint32_t l0 = ptx::pack_bf16(v[k].x - f0.x, v[k].y - f0.y), l1 = ptx::pack_bf16(v[k].z - f1.x, v[k].w - f1.y);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h0)),
                   "r"(*reinterpret_cast<const uint32_t*>(&h1)) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + (uint32_t)kSplitTile), "r"(l0), "r"(l1) : "memory");
    }
  }
}

// tm_hi / tm_lo: tensor maps of the pre-split weight W^T[N, Ktot] (K-major bf16) with [128 rows x 64] boxes; k_off: first K
// column of this launch inside Ktot.  C / R may alias (kSplitEpiResidual: C = R + A' W + bias, in place on the residual
// stream).  N must be a multiple of 128.
template <int KB, bool kLn, int EPI>
__global__ void __launch_bounds__(kSplitThreads, 1)
gemm_split_tc_kernel(const float* __restrict__ A, int lda, const __grid_constant__ CUtensorMap tm_hi,
                     const __grid_constant__ CUtensorMap tm_lo, int k_off, const float* __restrict__ bias,
                     const float* R, int ldr, float* C, int ldc, int M, int N) {
  using Cfg = SplitCfg<KB>;
  constexpr int ST = kSplitBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                    // [KB][hi 16K | lo 16K]
  uint8_t* sB = sA + Cfg::kABytes;       // [ST][hi 16K | lo 16K]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::kBBytes);
  uint64_t* a_full = bars;               // 128 producer threads
  uint64_t* a_empty = bars + 1;          // commit
  uint64_t* b_full = bars + 2;           // [ST] tx
  uint64_t* b_empty = b_full + ST;       // [ST] commit
  uint64_t* t_full = b_empty + ST;       // [2] commit
  uint64_t* t_empty = t_full + 2;        // [2] 128 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128, num_c = N / 128;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_hi); ptx::prefetch_tmap(&tm_lo); }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(a_full, 128); ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < ST; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight TMA producer: per tile, per column chunk, per K block: hi and lo [128 x 64]
    int stage = 0; uint32_t phase = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x)
      for (int c = 0; c < num_c; ++c)
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1, 60, 40u);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&b_full[stage], 2 * kSplitTile);
            ptx::tma_load_2d(sB + stage * 2 * kSplitTile, &tm_hi, &b_full[stage], k_off + kb * 64, c * 128);
            ptx::tma_load_2d(sB + stage * 2 * kSplitTile + kSplitTile, &tm_lo, &b_full[stage], k_off + kb * 64, c * 128);
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
  } else if (warp == 1) {
    // ===================== MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
    const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024, ptx::kSwz128);
    const uint64_t dB0 = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 1024, ptx::kSwz128);
    constexpr uint64_t kLo = kSplitTile >> 4;
    int stage = 0; uint32_t phase = 0;
    int it = 0, acc_it = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x, ++it) {
      ptx::mbar_wait(a_full, (uint32_t)(it & 1), 61, 40u);
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_empty[as], (uint32_t)((acc_it >> 1) & 1) ^ 1, 62, 40u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 128u;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&b_full[stage], phase, 63, 40u);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t da = dA0 + (uint64_t)kb * (2 * kLo), db = dB0 + (uint64_t)stage * (2 * kLo);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);              // hi hi
              ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + kLo + (uint64_t)(k * 2), idesc, 1);                     // hi lo
              ptx::umma_bf16(d_tmem, da + kLo + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1);                     // lo hi
            }
            ptx::umma_commit(&b_empty[stage]);
            if (kb == KB - 1) {
              ptx::umma_commit(&t_full[as]);
              if (c == num_c - 1) ptx::umma_commit(a_empty);
            }
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A producers: warp w splits rows 32w .. 32w+31 of each tile
    const int w = warp - 4;
    int it = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x, ++it) {
      ptx::mbar_wait(a_empty, (uint32_t)(it & 1) ^ 1, 64);
      split_rows_to_tiles<KB, kLn>(A, lda, M, mt * 128 + w * 32, ptx::smem_u32(sA), w * 32, lane);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(a_full);
    }
  } else if (warp >= 8) {
    // ===================== epilogue: one row per thread, 32 columns (one full 128-byte line) at a time
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int acc_it = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x) {
      const int row = mt * 128 + q * 32 + lane;
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_full[as], (uint32_t)((acc_it >> 1) & 1), 65);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col = c * 128 + cc * 32;
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (uint32_t)as * 128u + cc * 32 + lane_off, v);
          float4 rr[8];
          if (EPI == kSplitEpiResidual && row < M) {
#pragma unroll
            for (int e = 0; e < 8; ++e) rr[e] = __ldcg(reinterpret_cast<const float4*>(R + (size_t)row * ldr + col) + e);
          }
          ptx::tmem_wait_ld();
          if (cc == 3) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&t_empty[as]);
          }
          if (row < M) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col) + e);
              float4 o;
              o.x = __uint_as_float(v[e * 4 + 0]) + b4.x; o.y = __uint_as_float(v[e * 4 + 1]) + b4.y;
              o.z = __uint_as_float(v[e * 4 + 2]) + b4.z; o.w = __uint_as_float(v[e * 4 + 3]) + b4.w;
              if (EPI == kSplitEpiRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (EPI == kSplitEpiResidual) { o.x += rr[e].x; o.y += rr[e].y; o.z += rr[e].z; o.w += rr[e].w; }
              reinterpret_cast<float4*>(C + (size_t)row * ldc + col)[e] = o;
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace avb
