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
// Per CTA and 128-row tile: four producer warps normalise / split the fp32 rows into A units (128 columns = two K blocks
// of 128B-swizzled [128 x 64] bf16 tiles, hi and lo), double-buffered so the next tile (K = 128) or the next half tile
// (K = 256) is split while the tensor core works, and kept for all N/128 column chunks; the weight tiles (hi, lo, [128 x
// 64] each, pre-split at load time) stream through a TMA ring; 12 MMAs (M 128, N 128, K 16) per K block; accumulators
// double-buffered in TMEM; four epilogue warps add the bias (from shared memory), apply ReLU or the residual and write
// full 128-byte fp32 segments, one row per thread.  K is 128 or 256 per launch (the 512-wide FFN down-projection runs as two launches, the
// second accumulating onto the first through the residual epilogue).
//   warp 0: weight TMA   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-7: A producers   warps 8-11: epilogue
#pragma once
#include "kernels_tc.cuh"

namespace avb {

constexpr int kSplitThreads = 384;
constexpr int kSplitBStages = 2;
constexpr int kSplitTile = 128 * 64 * 2;  // one [128 x 64] bf16 tile, 16 KB
constexpr int kSplitUnit = 2 * 2 * kSplitTile;  // an A unit: two K blocks x (hi, lo) = 64 KB = 128 columns of 128 rows
template <int KB>
struct SplitCfg {
  static constexpr int kABytes = 2 * kSplitUnit;                  // two A units: double-buffered by tile (KB 2) / half tile (KB 4)
  static constexpr int kBBytes = kSplitBStages * 2 * kSplitTile;  // weight ring: (hi, lo) per stage
  static constexpr int kBiasBytes = 768 * 4;
  static constexpr int kStageBytes = 4 * 4096;                    // per epilogue warp: [32 rows x 128 B] transposition stage
  static constexpr int kSmemBytes = kABytes + kBBytes + kBiasBytes + kStageBytes + 1024 + 256;
};

enum { kSplitEpiBias = 0, kSplitEpiRelu = 1, kSplitEpiResidual = 2, kSplitEpiQkvOperands = 3 };

// kSplitEpiQkvOperands: the QKV projection writes hi / lo bf16 halves straight into the operand layout of
// attention_split_tc_kernel (below) instead of fp32 rows; the row order of the attended axis comes from RowMap.
struct SplitSeq { int n, d, axis, dg, H; };

// 32 fp32 rows x 128 columns (row pitch lda floats) -> one A unit (two K blocks of hi / lo bf16 K-major tiles); eight
// lanes share a row (128-byte coalesced segments), optional LayerNorm without affine over those 128 columns (exact
// two-pass statistics in fp32).
template <bool kLn>
__device__ __forceinline__ void split_rows_to_unit(const float* __restrict__ a, int lda, int M, int grow0, uint32_t a_unit,
                                                   int row0, int lane, const RowMap rm = RowMap{0, 0, 0}) {
  constexpr int NK = 4;  // float4 per lane and row: columns 4 (j + 8 k)
  const int rs = lane >> 3, j = lane & 7;
  float4 nxt[NK];
  auto load = [&](float4 (&dst)[NK], int r0) {
    const int grow = grow0 + r0 + rs;
    const int tok = grow < M ? rm.token(grow) : 0;  // the QKV projection walks its rows in sequence-major order
#pragma unroll
    for (int k = 0; k < NK; ++k)
      dst[k] = grow < M ? __ldcg(reinterpret_cast<const float4*>(a + (size_t)tok * lda) + j + 8 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  load(nxt, 0);
#pragma unroll 1
  for (int r0 = 0; r0 < 32; r0 += 4) {
    float4 v[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) v[k] = nxt[k];
    if (r0 + 4 < 32) load(nxt, r0 + 4);  // the next four rows are in flight while these are split
    if (kLn) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float mean = sum * (1.0f / 128.0f);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
        sq += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      const float rstd = rsqrtf(sq * (1.0f / 128.0f) + kLnEps);
#pragma unroll
      for (int k = 0; k < NK; ++k) { v[k].x *= rstd; v[k].y *= rstd; v[k].z *= rstd; v[k].w *= rstd; }
    }
    const int r = row0 + r0 + rs;
    const uint32_t rbase = a_unit + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (j & 1) * 8);
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
                     const float* R, int ldr, float* C, int ldc, int M, int N, SplitSeq sq, float* __restrict__ norms) {
  // norms (kSplitEpiQkvOperands, may be null, zeroed by the caller): {max |q_h|^2, max |k_h|^2} over the (row, head) vectors
  // written - the score bound with which attention_split_tc_kernel skips its row-maximum pass (as in the bf16 mode)
  static_assert(!kLn || KB == 2, "the LayerNorm prologue normalises over one A unit (128 columns)");
  using Cfg = SplitCfg<KB>;
  constexpr int ST = kSplitBStages, UNITS = KB / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                    // [2 units][2 K blocks][hi 16K | lo 16K]
  uint8_t* sB = sA + Cfg::kABytes;       // [ST][hi 16K | lo 16K]
  float* s_bias = reinterpret_cast<float*>(sB + Cfg::kBBytes);  // [N <= 768]: a __ldg would queue behind the producers' row loads
  uint8_t* sStage = sB + Cfg::kBBytes + Cfg::kBiasBytes;        // [4 epilogue warps][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::kStageBytes);
  uint64_t* a_full = bars;               // [2] 128 producer threads
  uint64_t* a_empty = bars + 2;          // [2] commit
  uint64_t* b_full = bars + 4;           // [ST] tx
  uint64_t* b_empty = b_full + ST;       // [ST] commit
  uint64_t* t_full = b_empty + ST;       // [2] commit
  uint64_t* t_empty = t_full + 2;        // [2] 128 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128, num_c = N / 128;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_hi); ptx::prefetch_tmap(&tm_lo); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&a_full[i], 128); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < ST; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < N; i += kSplitThreads) s_bias[i] = bias[i];
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
    // ===================== MMA issuer.  An A unit (128 columns) is waited for in the first column chunk that uses it and
    // released after the last one: with KB = 2 the tile's single unit alternates between the two buffers (the producers
    // split the next tile meanwhile), with KB = 4 and one chunk the two halves of a tile do.
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
    const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024, ptx::kSwz128);
    const uint64_t dB0 = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 1024, ptx::kSwz128);
    constexpr uint64_t kLo = kSplitTile >> 4, kUnit = kSplitUnit >> 4;
    int stage = 0; uint32_t phase = 0;
    int acc_it = 0;
    uint32_t uc0 = 0;  // running unit count at the start of the tile
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x, uc0 += UNITS) {
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_empty[as], (uint32_t)((acc_it >> 1) & 1) ^ 1, 62, 40u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 128u;
        for (int u = 0; u < UNITS; ++u) {
          const uint32_t uc = uc0 + (uint32_t)u, ab = uc & 1u;
          if (c == 0) ptx::mbar_wait(&a_full[ab], (uc >> 1) & 1u, 61, 40u);
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            const int kb = u * 2 + kb2;
            ptx::mbar_wait(&b_full[stage], phase, 63, 40u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t da = dA0 + (uint64_t)ab * kUnit + (uint64_t)kb2 * (2 * kLo), db = dB0 + (uint64_t)stage * (2 * kLo);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);              // hi hi
                ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + kLo + (uint64_t)(k * 2), idesc, 1);                     // hi lo
                ptx::umma_bf16(d_tmem, da + kLo + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1);                     // lo hi
              }
              ptx::umma_commit(&b_empty[stage]);
              if (kb2 == 1 && c == num_c - 1) ptx::umma_commit(&a_empty[ab]);
              if (kb == KB - 1) ptx::umma_commit(&t_full[as]);
            }
            __syncwarp();
            if (++stage == ST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A producers: warp w splits rows 32w .. 32w+31 of each unit
    const int w = warp - 4;
    uint32_t uc = 0;
    const RowMap prm = EPI == kSplitEpiQkvOperands ? RowMap{sq.n, sq.d, sq.axis, sq.dg} : RowMap{0, 0, 0};
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x) {
      // L2 prefetch of this warp's 32 rows of the NEXT tile: a warp keeps only 2 KB of row loads in flight, far below
      // HBM latency x bandwidth (measured without it: 42K clocks per tile, the whole kernel latency-bound on these loads)
      {
        const int nrow0 = (mt + (int)gridDim.x) * 128 + w * 32;
#pragma unroll
        for (int i = 0; i < 2 * KB; ++i) {
          const int line = i * 32 + lane, row = nrow0 + line / (2 * KB), seg = line % (2 * KB);
          if (row < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)prm.token(row) * lda + seg * 32));
        }
      }
      for (int u = 0; u < UNITS; ++u, ++uc) {
        const uint32_t ab = uc & 1u;
        ptx::mbar_wait(&a_empty[ab], ((uc >> 1) & 1u) ^ 1u, 64);
        split_rows_to_unit<kLn>(A + u * 128, lda, M, mt * 128 + w * 32, ptx::smem_u32(sA) + ab * kSplitUnit, w * 32, lane, prm);
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&a_full[ab]);
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: one row per thread, 32 columns (one full 128-byte line) at a time
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int acc_it = 0;
    [[maybe_unused]] float nrm_q = 0.f, nrm_k = 0.f;
    const uint32_t a_bias = ptx::smem_u32(s_bias);
    const RowMap rm{sq.n, sq.d, sq.axis, sq.dg};
    const int seq_T = sq.axis == 0 ? sq.d : sq.n;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x) {
      const int row = mt * 128 + q * 32 + lane;
      [[maybe_unused]] unsigned seq = 0, pos = 0;
      if (EPI == kSplitEpiQkvOperands && row < M) {  // GEMM rows are in sequence-major order (the producers gather them)
        seq = (unsigned)row / (unsigned)seq_T; pos = (unsigned)row - seq * (unsigned)seq_T;
      }
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_full[as], (uint32_t)((acc_it >> 1) & 1), 65);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col = c * 128 + cc * 32;
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (uint32_t)as * 128u + cc * 32 + lane_off, v);
          ptx::tmem_wait_ld();
          if (cc == 3) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&t_empty[as]);
          }
          if (EPI == kSplitEpiQkvOperands) {
            if (row < M) {  // 32 columns = one (operand w, head): 64 bytes of hi and of lo at [seq][w][part][head][pos]
              const int w = col >> 8, hd = (col & 255) >> 5;
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(C) +
                                   ((((size_t)seq * 3 + w) * 2 * sq.H + hd) * seq_T + pos) * 32;
              const size_t part = (size_t)sq.H * seq_T * 32;
              const unsigned swz = (pos >> 1) & 3;
              float ss = 0.f;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 b0 = ptx::ld_shared_v4f(a_bias + (uint32_t)((col + 8 * e) * 4));
                const float4 b1 = ptx::ld_shared_v4f(a_bias + (uint32_t)((col + 8 * e + 4) * 4));
                float f[8];
                f[0] = __uint_as_float(v[e * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[e * 8 + 1]) + b0.y;
                f[2] = __uint_as_float(v[e * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[e * 8 + 3]) + b0.w;
                f[4] = __uint_as_float(v[e * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[e * 8 + 5]) + b1.y;
                f[6] = __uint_as_float(v[e * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[e * 8 + 7]) + b1.w;
                if (w < 2) {
#pragma unroll
                  for (int t = 0; t < 8; ++t) ss = fmaf(f[t], f[t], ss);
                }
                uint4 hi, lo;
                uint32_t* hp = &hi.x; uint32_t* lp = &lo.x;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const uint32_t h2 = ptx::pack_bf16(f[2 * t], f[2 * t + 1]);
                  hp[t] = h2;
                  lp[t] = ptx::pack_bf16(f[2 * t] - __uint_as_float(h2 << 16), f[2 * t + 1] - __uint_as_float(h2 & 0xffff0000u));
                }
                *reinterpret_cast<uint4*>(dst + (((unsigned)e ^ swz) << 3)) = hi;
                *reinterpret_cast<uint4*>(dst + part + (((unsigned)e ^ swz) << 3)) = lo;
              }
              if (w == 0) nrm_q = fmaxf(nrm_q, ss); else if (w == 1) nrm_k = fmaxf(nrm_k, ss);
            }
          } else {
            // A row-per-thread global access touches 32 different lines per warp instruction (bound by L1 wavefronts, ~11K
            // clocks per 128 x 128 chunk measured): the [32 rows x 32 columns] block goes through a 4 KB XOR-swizzled stage
            // owned by the warp, after which every instruction covers 4 rows x 128 contiguous bytes.
            const uint32_t stage = ptx::smem_u32(sStage) + (uint32_t)((warp - 8) * 4096);
#pragma unroll
            for (int e = 0; e < 8; ++e) {  // thread = row `lane`: 16-byte chunk e goes to slot e ^ (lane & 7)
              const float4 b4 = ptx::ld_shared_v4f(a_bias + (uint32_t)((col + 4 * e) * 4));
              float4 o;
              o.x = __uint_as_float(v[e * 4 + 0]) + b4.x; o.y = __uint_as_float(v[e * 4 + 1]) + b4.y;
              o.z = __uint_as_float(v[e * 4 + 2]) + b4.z; o.w = __uint_as_float(v[e * 4 + 3]) + b4.w;
              if (EPI == kSplitEpiRelu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              ptx::st_shared_v4(stage + (uint32_t)(lane * 128 + ((e ^ (lane & 7)) << 4)),
                                make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
            }
            __syncwarp();
            const int sub = lane >> 3, l8 = lane & 7, brow0 = mt * 128 + q * 32;
            float4 rr2[8];
            if (EPI == kSplitEpiResidual) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rw = brow0 + i * 4 + sub;
                rr2[i] = rw < M ? __ldcg(reinterpret_cast<const float4*>(R + (size_t)rw * ldr + col) + l8) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub, rw = brow0 + rl;
              float4 o = ptx::ld_shared_v4f(stage + (uint32_t)(rl * 128 + ((l8 ^ (rl & 7)) << 4)));
              if (EPI == kSplitEpiResidual) { o.x += rr2[i].x; o.y += rr2[i].y; o.z += rr2[i].z; o.w += rr2[i].w; }
              if (rw < M) reinterpret_cast<float4*>(C + (size_t)rw * ldc + col)[l8] = o;
            }
            __syncwarp();
          }
        }
      }
    }
    if (EPI == kSplitEpiQkvOperands && norms != nullptr) {  // one atomic per warp and operand (non-negative floats order like ints)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        nrm_q = fmaxf(nrm_q, __shfl_xor_sync(0xffffffffu, nrm_q, o));
        nrm_k = fmaxf(nrm_k, __shfl_xor_sync(0xffffffffu, nrm_k, o));
      }
      if (lane == 0) {
        atomicMax(reinterpret_cast<int*>(norms), __float_as_int(nrm_q));
        atomicMax(reinterpret_cast<int*>(norms) + 1, __float_as_int(nrm_k));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace avb

namespace avb {

// ------------------------------------------------------------------------------------------
// Split-precision axial attention for the fp32 mode (hk.MultiHeadAttention core, model.py:59-64), head dim 32.
//   Operands come from the QKV projection (gemm_split_tc_kernel, kSplitEpiQkvOperands) as hi / lo bf16 halves in the
//   tensor-core attention's operand layout, qkvs[seq][w = q,k,v][part = hi,lo][head][pos][32 ch] with the 16-byte chunks
//   of a 64-byte row at chunk ^ ((pos >> 1) & 3); q carries log2e / sqrt(32).
//     S = Q K^T  as  Qhi Khi + Qhi Klo + Qlo Khi                 (6 MMAs M 128, N <= 128, K 16, fp32 in TMEM)
//     exact online softmax in fp32, one query row per thread: running maximum, p = exp2(s - m), row sum in fp32
//     O += P V   as  Phi Vhi + Phi Vlo + Plo Vhi                 (A = P from TMEM, 3 MMAs N 32 per 16 keys)
//   P (hi | lo, packed bf16 pairs) overwrites the S tile it came from, half by half - a thread owns its whole row - so a
//   CTA needs 160 TMEM columns and 80 KB of shared memory and two CTAs share an SM: within a CTA the steps S -> softmax ->
//   PV run strictly in sequence (the tensor pipe's in-order execution replaces most barriers), the second CTA fills the gaps.
//   out: fp32 attn[token][H * 32] (token order, the A operand of the split-precision out-projection).
//   warp 0: bulk-copy producer   warp 1: MMA issuer + TMEM alloc   warps 2-5: softmax / epilogue (TMEM lane quarter = warp & 3)
// ------------------------------------------------------------------------------------------
constexpr int kSplitAttnThreads = 192;
constexpr int kSplitAttnStages = 2;
constexpr int kSplitAttnTile = 128 * 32 * 2;  // 8 KB: one [128 x 32] bf16 tile
constexpr int kSplitAttnSmemBytes = 2 * kSplitAttnTile + kSplitAttnStages * 4 * kSplitAttnTile + 1024 + 256;  // Q hi|lo, stages of K hi|lo|V hi|lo

__global__ void __launch_bounds__(kSplitAttnThreads, 2)
attention_split_tc_kernel(const __nv_bfloat16* __restrict__ qkvs, float* __restrict__ out, int B, int n, int d, int H, int axis,
                          int dg, const float* __restrict__ norms) {
  // norms (may be null): {max |q_h|^2, max |k_h|^2} of the projection that wrote qkvs.  With |score| <= |q||k| <= 100 (log2
  // units) exp2(s) needs no reference: fp32 (and the bf16 halves of P) carry an 8-bit exponent and O / l cancels the missing
  // shift exactly - the row-maximum pass over S and the rescaling of O drop out.  Without the bound: exact online softmax.
  constexpr int NS = kSplitAttnStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                           // hi 8K | lo 8K
  uint8_t* sKV = sQ + 2 * kSplitAttnTile;       // [NS][K hi | K lo | V hi | V lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NS * 4 * kSplitAttnTile);
  uint64_t* q_full = bars;             // tx
  uint64_t* q_empty = bars + 1;        // commit
  uint64_t* kv_full = bars + 2;        // [NS] tx
  uint64_t* kv_empty = kv_full + NS;   // [NS] commit
  uint64_t* s_full = kv_empty + NS;    // commit
  uint64_t* p_full = s_full + 1;       // 128 softmax threads
  uint64_t* pv_done = p_full + 1;      // commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int T = axis == 0 ? d : n;
  const int S = axis == 0 ? B * n : B * d;
  const int nqt = (T + 127) / 128;
  const int items = S * H * nqt;  // the host guarantees < 2^31
  const size_t part_elems = (size_t)H * T * 32;  // one (sequence, operand, hi|lo) block

  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1); ptx::mbar_init(q_empty, 1);
    for (int i = 0; i < NS; ++i) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 1); }
    ptx::mbar_init(s_full, 1); ptx::mbar_init(p_full, 128); ptx::mbar_init(pv_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  // ragged tiles copy only their valid rows: whatever else the tensor core reads in a slot must be finite
  for (int i = threadIdx.x; i < (2 + NS * 4) * kSplitAttnTile / 16; i += kSplitAttnThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    // ======================= producer
    uint32_t g = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int qt = item % nqt, pair = item / nqt, h = pair % H, s = pair / H;
      const __nv_bfloat16* base = qkvs + (size_t)s * 6 * part_elems + (size_t)h * T * 32;  // (s, q, hi, h)
      ptx::mbar_wait(q_empty, (uint32_t)(it & 1) ^ 1, 50, 100u);
      if (ptx::elect_one()) {
        const uint32_t bytes = (uint32_t)min(128, T - qt * 128) * 64u;
        ptx::mbar_expect_tx(q_full, 2 * bytes);
        ptx::bulk_load(sQ, base + (size_t)qt * 128 * 32, bytes, q_full);
        ptx::bulk_load(sQ + kSplitAttnTile, base + part_elems + (size_t)qt * 128 * 32, bytes, q_full);
      }
      __syncwarp();
      for (int j = 0; j < nqt; ++j, ++g) {
        const int st = (int)(g % NS);
        ptx::mbar_wait(&kv_empty[st], (uint32_t)((g / NS) & 1) ^ 1, 51, 100u);
        if (ptx::elect_one()) {
          uint8_t* sk = sKV + st * 4 * kSplitAttnTile;
          const uint32_t bytes = (uint32_t)min(128, T - j * 128) * 64u;
          ptx::mbar_expect_tx(&kv_full[st], 4 * bytes);
#pragma unroll
          for (int t = 0; t < 4; ++t)  // K hi, K lo, V hi, V lo: operand blocks 2..5 of the sequence
            ptx::bulk_load_hint(sk + t * kSplitAttnTile, base + (size_t)(2 + t) * part_elems + (size_t)j * 128 * 32, bytes, &kv_full[st],
                                ptx::kL2EvictLast);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer.  One thread issues every MMA, the tensor pipe executes them in order: S of a step
    // is behind the PV of the previous one, so the P tile it overwrites has been consumed, and O is complete when S is.
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
    const uint64_t dQ = ptx::make_smem_desc(ptx::smem_u32(sQ), 16, 512, ptx::kSwz64);
    const uint64_t dKV = ptx::make_smem_desc(ptx::smem_u32(sKV), 16, 512, ptx::kSwz64);
    constexpr uint64_t kT = kSplitAttnTile >> 4;
    uint32_t g = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      ptx::mbar_wait(q_full, (uint32_t)(it & 1), 52, 40u);
      for (int j = 0; j < nqt; ++j, ++g) {
        const int st = (int)(g % NS);
        int kc = T - j * 128; kc = kc > 128 ? 128 : kc; kc = (kc + 15) & ~15;  // key columns the MMAs touch
        const uint32_t idesc_s = ptx::make_idesc_bf16(128, kc, 0, 0);
        ptx::mbar_wait(&kv_full[st], (uint32_t)((g / NS) & 1), 53, 40u);
        ptx::tc_fence_after();
        const uint64_t dk = dKV + (uint64_t)st * (4 * kT);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            ptx::umma_bf16(tS, dQ + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);            // Qhi Khi
            ptx::umma_bf16(tS, dQ + (uint64_t)(k * 2), dk + kT + (uint64_t)(k * 2), idesc_s, 1);             // Qhi Klo
            ptx::umma_bf16(tS, dQ + kT + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, 1);             // Qlo Khi
          }
          ptx::umma_commit(s_full);
          if (j == nqt - 1) ptx::umma_commit(q_empty);
        }
        __syncwarp();
        ptx::mbar_wait(p_full, g & 1u, 54, 40u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dvh = dk + 2 * kT, dvl = dk + 3 * kT;
          const int ksteps = kc >> 4;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ph = tS + (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8), pl = ph + 32;
            ptx::umma_bf16_ts(tO, ph, dvh + (uint64_t)(ks * 64), idesc_pv, (j | ks) != 0);                   // Phi Vhi
            ptx::umma_bf16_ts(tO, ph, dvl + (uint64_t)(ks * 64), idesc_pv, 1);                                // Phi Vlo
            ptx::umma_bf16_ts(tO, pl, dvh + (uint64_t)(ks * 64), idesc_pv, 1);                                // Plo Vhi
          }
          ptx::umma_commit(&kv_empty[st]);
          ptx::umma_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= softmax / epilogue: thread = query row r of the tile = TMEM lane
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const RowMap rm{n, d, axis, dg};
    const bool bounded = norms != nullptr && norms[0] * norms[1] <= 1.0e4f;  // false also for NaN
    uint32_t g = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int qt = item % nqt, pair = item / nqt, h = pair % H, s = pair / H;
      float m = bounded ? 0.f : -INFINITY, l = 0.f;
      for (int j = 0; j < nqt; ++j, ++g) {
        ptx::mbar_wait(s_full, g & 1u, 55);
        ptx::tc_fence_after();
        const int nv = T - j * 128;  // valid key columns of this tile (>= 1)
        // ---- pass 1: row maximum (skipped under the score bound: m stays 0)
        float mx = m;
#pragma unroll 1
        for (int c = 0; c < (bounded ? 0 : 4); ++c) {
          if (c * 32 < nv) {
            uint32_t sv[32];
            ptx::tmem_ld32(tS + c * 32 + lane_off, sv);
            ptx::tmem_wait_ld();
            const int rem = nv - c * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, e < rem ? __uint_as_float(sv[e]) : -INFINITY);
          }
        }
        // ---- rescale O and the row sum (O is complete: S of this step was issued behind the previous PV)
        if (!bounded && j > 0 && __any_sync(0xffffffffu, mx > m)) {  // warp-uniform: tcgen05.ld / st are warp-collective
          const float alpha = fast_exp2(m - mx);          // 1 for the rows whose maximum did not move
          uint32_t o[32];
          ptx::tmem_ld32(tO + lane_off, o);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          ptx::tmem_st32(tO + lane_off, o);
          l *= alpha;
        }
        m = mx;
        // ---- pass 2: p = exp2(s - m), split into hi | lo bf16 pairs over the S columns they came from
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t ph[32], pl[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t sv[32];
            ptx::tmem_ld32(tS + hf * 64 + c * 32 + lane_off, sv);
            ptx::tmem_wait_ld();
            const int rem = nv - hf * 64 - c * 32;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float p0 = 2 * e < rem ? fast_exp2(__uint_as_float(sv[2 * e]) - m) : 0.f;
              const float p1 = 2 * e + 1 < rem ? fast_exp2(__uint_as_float(sv[2 * e + 1]) - m) : 0.f;
              l += p0 + p1;
              const uint32_t hi = ptx::pack_bf16(p0, p1);
              ph[c * 16 + e] = hi;
              pl[c * 16 + e] = ptx::pack_bf16(p0 - __uint_as_float(hi << 16), p1 - __uint_as_float(hi & 0xffff0000u));
            }
          }
          ptx::tmem_st32(tS + hf * 64 + lane_off, ph);
          ptx::tmem_st32(tS + hf * 64 + 32 + lane_off, pl);
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_full);
      }
      // ---- epilogue: O / l -> fp32 attn[token][h * 32 ..]
      ptx::mbar_wait(pv_done, (g - 1) & 1u, 56);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld32(tO + lane_off, o);
      ptx::tmem_wait_ld();
      const int pos = qt * 128 + r;
      if (pos < T) {
        const float inv = 1.0f / l;
        float4* op = reinterpret_cast<float4*>(out + (size_t)rm.token(s * T + pos) * (H * 32) + h * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          op[e] = make_float4(__uint_as_float(o[e * 4 + 0]) * inv, __uint_as_float(o[e * 4 + 1]) * inv,
                              __uint_as_float(o[e * 4 + 2]) * inv, __uint_as_float(o[e * 4 + 3]) * inv);
      }
      ptx::tc_fence_before();  // O has been read before this thread's next p_full arrival lets the first PV of the next item start
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace avb

namespace avb {

// ------------------------------------------------------------------------------------------
// Edge-logit contraction of the head on the tensor cores (model.py:133-141, 218-239):
//     p[b, i, j] = sigmoid( <u_i, v_j> * exp(temp) + bias ),   diag = 0      (u, v: L2-normalised rows from head_uv_kernel)
//   One CTA per (dataset, 128 x 128 tile of the [d, d] matrix).  Both operands are [rows x 128] fp32 matrices that the
//   producer warps split into hi / lo bf16 K-major tiles (the same A-unit image for both: u rows are the M side, v rows
//   the N side), 24 MMAs (M 128, N 128, K 16: hi hi + hi lo + lo hi) into 128 TMEM columns, epilogue one row per thread.
//   The cosines carry ~1e-5 relative error (split precision), far inside the bf16 mode's budget; the fp32 mode keeps the
//   SIMT head.  d^2 * 256 FLOP per dataset - 2.6 MFLOP at d = 100, 256 MFLOP at d = 1000: this is about where the
//   contraction runs, not about time (< 0.1 % of a forward either way).
//   warps 0-3: u tile, then epilogue (TMEM lane quarter = warp)   warps 4-7: v tile (warp 4 also allocates TMEM, issues)
// ------------------------------------------------------------------------------------------
constexpr int kHeadTcThreads = 256;
constexpr int kHeadTcSmemBytes = 2 * kSplitUnit + 1024 + 256;

__global__ void __launch_bounds__(kHeadTcThreads, 1)
head_logits_tc_kernel(const float* __restrict__ u, const float* __restrict__ v, int d, const float* __restrict__ temp,
                      const float* __restrict__ bias, int mask_diag, int log_probs, float* __restrict__ probs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // u tile: [2 K blocks][hi 16K | lo 16K]
  uint8_t* sB = smem + kSplitUnit;    // v tile, same image
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 2 * kSplitUnit);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int b = blockIdx.z, i0 = blockIdx.y * 128, j0 = blockIdx.x * 128;
  const float* ub = u + (size_t)b * d * 128;
  const float* vb = v + (size_t)b * d * 128;

  if (warp == 4 && lane == 0) { ptx::mbar_init(done, 1); ptx::fence_barrier_init(); }
  if (warp == 4) { ptx::tmem_alloc(tmem_slot, 128); ptx::tmem_relinquish(); }
  if (warp < 4) split_rows_to_unit<false>(ub, 128, d, i0 + warp * 32, ptx::smem_u32(sA), warp * 32, lane);
  else split_rows_to_unit<false>(vb, 128, d, j0 + (warp - 4) * 32, ptx::smem_u32(sB), (warp - 4) * 32, lane);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 4) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024, ptx::kSwz128);
      const uint64_t dB0 = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 1024, ptx::kSwz128);
      constexpr uint64_t kLo = kSplitTile >> 4;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = dA0 + (uint64_t)kb * (2 * kLo) + (uint64_t)(k * 2), db = dB0 + (uint64_t)kb * (2 * kLo) + (uint64_t)(k * 2);
          ptx::umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
          ptx::umma_bf16(tmem_base, da, db + kLo, idesc, 1);
          ptx::umma_bf16(tmem_base, da + kLo, db, idesc, 1);
        }
      ptx::umma_commit(done);
    }
    __syncwarp();
  }
  if (warp < 4) {
    ptx::mbar_wait(done, 0, 66);
    ptx::tc_fence_after();
    const int i = i0 + warp * 32 + lane;
    const float scale = expf(temp[0]), bs = bias[0];
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t acc[32];
      ptx::tmem_ld32(tmem_base + cc * 32 + ((uint32_t)(warp * 32) << 16), acc);
      ptx::tmem_wait_ld();
      if (i < d) {
        float* prow = probs + ((size_t)b * d + i) * d;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + cc * 32 + e;
          if (j < d) {
            const float l = __uint_as_float(acc[e]) * scale + bs;
            const float logp = -(fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l))));  // log_sigmoid(l)
            float p = log_probs ? logp : expf(logp);
            if (mask_diag && i == j) p = log_probs ? -INFINITY : 0.f;
            prow[j] = p;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, 128);
}

}  // namespace avb
