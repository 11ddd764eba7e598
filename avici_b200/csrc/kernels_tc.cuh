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


// Debug timeline: when set (avb_debug_set_trace), block 0 records clock64() stamps of the first steps of each
// role of stream 0 into trace[role * 4096 + step * 8 + event].
__device__ long long* g_attn_trace = nullptr;
#ifdef AVB_TRACE_ENABLE
#define AVB_TRACE(role, step, ev)                                                      \
  do {                                                                                 \
    if (trace && (step) < 500) trace[(role) * 4096 + (step) * 8 + (ev)] = clock64();  \
  } while (0)
#else  // production builds carry no tracing code (it costs registers and issue slots in the softmax loop)
#define AVB_TRACE(role, step, ev) do { (void)trace; } while (0)
#endif

// Sequence-major row order of the attention sub-layer (QKV projection ln_gemm_tc_kernel<2>, attention, out-projection): GEMM row m = seq * T + pos, where a
// sequence is what the following attention attends over.  axis 0 (attend over d): seq = (b, i), pos = j, i.e. m is the
// token index itself.  axis 1 (attend over n): seq = (b, j), pos = i, token = (b*n + i)*d + j.
// Row order of the out-projection's bf16 delta buffer: 1 = token order (the out-projection scatters its 256-byte rows,
// the FFN reads them next to z; measured 2.4 ms per step faster in the FFN for 0.3 ms in the out-projection),
// 0 = sequence-major (the FFN gathers through RowMap::seq_rows); kept as an A/B build (-DAVB_DELTA_TOKENS=0).
#ifndef AVB_DELTA_TOKENS
#define AVB_DELTA_TOKENS 1
#endif
// Column-blocked storage (dg > 0, axis 0 only, one dataset): the n-shard of an axially sharded forward
// (avb_forward_sharded) keeps its rows in G = d / dg column blocks, token = (q * n + i) * dg + jj for row i and column
// j = q * dg + jj, so that block q is exactly the contiguous chunk the all-to-all sends to rank q - the exchange between
// an n-sharded and a d-sharded block then needs no pack or unpack pass.  `n` is the number of rows of the shard.
struct RowMap {
  int n, d, axis;
  int dg = 0;
  __device__ __forceinline__ int token(int m) const {
    if (axis == 0) {
      if (dg == 0) return m;
      const unsigned i = (unsigned)m / (unsigned)d, j = (unsigned)m - i * (unsigned)d;
      const unsigned q = j / (unsigned)dg, jj = j - q * (unsigned)dg;
      return (int)((q * (unsigned)n + i) * (unsigned)dg + jj);
    }
    const unsigned nd = (unsigned)n * (unsigned)d, b = (unsigned)m / nd, rem = (unsigned)m - b * nd;
    const unsigned j = rem / (unsigned)n, i = rem - j * (unsigned)n;
    return (int)((b * (unsigned)n + i) * (unsigned)d + j);
  }
  // inverse: token index -> sequence-major row
  __device__ __forceinline__ int seq_row(int t) const {
    if (axis == 0) {
      if (dg == 0) return t;
      const unsigned blk = (unsigned)n * (unsigned)dg, q = (unsigned)t / blk, rem = (unsigned)t - q * blk;
      const unsigned i = rem / (unsigned)dg, jj = rem - i * (unsigned)dg;
      return (int)(i * (unsigned)d + q * (unsigned)dg + jj);
    }
    const unsigned nd = (unsigned)n * (unsigned)d, b = (unsigned)t / nd, rem = (unsigned)t - b * nd;
    const unsigned i = rem / (unsigned)d, jj = rem - i * (unsigned)d;
    return (int)((b * (unsigned)d + jj) * (unsigned)n + i);
  }
  // sequence-major rows of tokens t, t + step, ... (count of them) without a division per token; -1 past the end
  template <int COUNT>
  __device__ __forceinline__ void seq_rows(int t, int step, int M, int (&out)[COUNT]) const {
    if (axis == 0) {
#pragma unroll
      for (int k = 0; k < COUNT; ++k) out[k] = t + k * step < M ? (dg == 0 ? t + k * step : seq_row(t + k * step)) : -1;
      return;
    }
    const unsigned nd = (unsigned)n * (unsigned)d;
    unsigned b = (unsigned)t / nd;
    const unsigned rem = (unsigned)t - b * nd;
    unsigned i = rem / (unsigned)d, j = rem - i * (unsigned)d;
#pragma unroll
    for (int k = 0; k < COUNT; ++k) {
      out[k] = t + k * step < M ? (int)((b * (unsigned)d + j) * (unsigned)n + i) : -1;
      j += (unsigned)step;
      while (j >= (unsigned)d) {
        j -= (unsigned)d;
        if (++i >= (unsigned)n) { i = 0; ++b; }
      }
    }
  }
  // tokens of rows m, m + step, m + 2*step, ... (count of them) without a division per row: the position inside the
  // sequence advances by `step` and carries into (j, b)
  template <int COUNT>
  __device__ __forceinline__ void tokens(int m, int step, int M, int (&tok)[COUNT]) const {
    if (axis == 0) {
      if (dg == 0) {
#pragma unroll
        for (int k = 0; k < COUNT; ++k) tok[k] = m + k * step < M ? m + k * step : -1;
        return;
      }
      unsigned i = (unsigned)m / (unsigned)d;
      const unsigned j = (unsigned)m - i * (unsigned)d;
      unsigned q = j / (unsigned)dg, jj = j - q * (unsigned)dg;
      const unsigned G = (unsigned)d / (unsigned)dg;
#pragma unroll
      for (int k = 0; k < COUNT; ++k) {
        tok[k] = m + k * step < M ? (int)((q * (unsigned)n + i) * (unsigned)dg + jj) : -1;
        jj += (unsigned)step;
        while (jj >= (unsigned)dg) {
          jj -= (unsigned)dg;
          if (++q >= G) { q = 0; ++i; }
        }
      }
      return;
    }
    const unsigned nd = (unsigned)n * (unsigned)d;
    unsigned b = (unsigned)m / nd;
    const unsigned rem = (unsigned)m - b * nd;
    unsigned j = rem / (unsigned)n, i = rem - j * (unsigned)n;
#pragma unroll
    for (int k = 0; k < COUNT; ++k) {
      tok[k] = m + k * step < M ? (int)((b * (unsigned)n + i) * (unsigned)d + j) : -1;
      i += (unsigned)step;
      while (i >= (unsigned)n) {
        i -= (unsigned)n;
        if (++j >= (unsigned)d) { j = 0; ++b; }
      }
    }
  }
};

// Packed fp32 pairs (sm_100 add/fma.f32x2): two results per issued instruction.  The softmax warps are bound by
// issue slots and the MUFU pipe together, so everything that is not an ex2 is done two elements at a time.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// ------------------------------------------------------------------------------------------
// Residual update of a [32 rows x 32 cols] fp32 block held row-per-thread in registers (the tcgen05.ld layout):
//     zblk[r][c] += acc[r][c] + bias[c]
// A row-per-thread global access touches 32 different 128-byte lines per warp instruction and is bound by L1
// wavefronts (measured: ~11k cycles per 128x128 tile).  The block is therefore transposed through a 4 KB
// XOR-swizzled smem stage owned by the warp, after which every warp instruction covers 4 rows x 128 contiguous bytes.
//   stage: smem address of this warp's 4 KB stage; zblk: &z[row0][col0]; ldz: row pitch in floats;
//   rows_valid: number of rows of the block that exist (<= 32)
// ------------------------------------------------------------------------------------------
template <class RowPtr>  // rowptr(row) -> float* of the block's row `row` (first of its 32 columns), nullptr if absent
__device__ __forceinline__ void residual_add_block_32x32_rows(const uint32_t (&acc)[32], const float* __restrict__ bias,
                                                              uint32_t stage, RowPtr rowptr, int lane) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // thread = row `lane`: 16-byte chunk c goes to slot c ^ (lane & 7)
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + c);
    uint4 pk;
    pk.x = __float_as_uint(__uint_as_float(acc[c * 4 + 0]) + b4.x);
    pk.y = __float_as_uint(__uint_as_float(acc[c * 4 + 1]) + b4.y);
    pk.z = __float_as_uint(__uint_as_float(acc[c * 4 + 2]) + b4.z);
    pk.w = __float_as_uint(__uint_as_float(acc[c * 4 + 3]) + b4.w);
    ptx::st_shared_v4(stage + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4)), pk);
  }
  __syncwarp();
  const int sub = lane >> 3, l8 = lane & 7;
  float4 zr[8];
  float4* zp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* rp = rowptr(i * 4 + sub);
    zp[i] = rp ? reinterpret_cast<float4*>(rp) + l8 : nullptr;
    zr[i] = zp[i] ? __ldcg(zp[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sub;
    uint4 a;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                 : "r"(stage + (uint32_t)(row * 128 + ((l8 ^ (row & 7)) << 4))));
    float4 o = zr[i];
    o.x += __uint_as_float(a.x); o.y += __uint_as_float(a.y); o.z += __uint_as_float(a.z); o.w += __uint_as_float(a.w);
    if (zp[i]) *zp[i] = o;
  }
  __syncwarp();
}

__device__ __forceinline__ void residual_add_block_32x32(const uint32_t (&acc)[32], const float* __restrict__ bias,
                                                         uint32_t stage, float* zblk, int ldz, int rows_valid, int lane) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // thread = row `lane`: 16-byte chunk c goes to slot c ^ (lane & 7)
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + c);
    uint4 pk;
    pk.x = __float_as_uint(__uint_as_float(acc[c * 4 + 0]) + b4.x);
    pk.y = __float_as_uint(__uint_as_float(acc[c * 4 + 1]) + b4.y);
    pk.z = __float_as_uint(__uint_as_float(acc[c * 4 + 2]) + b4.z);
    pk.w = __float_as_uint(__uint_as_float(acc[c * 4 + 3]) + b4.w);
    ptx::st_shared_v4(stage + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4)), pk);
  }
  __syncwarp();
  const int sub = lane >> 3, l8 = lane & 7;
  float4 zr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sub;
    zr[i] = row < rows_valid ? __ldcg(reinterpret_cast<const float4*>(zblk + (size_t)row * ldz) + l8) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sub;
    uint4 a;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                 : "r"(stage + (uint32_t)(row * 128 + ((l8 ^ (row & 7)) << 4))));
    float4 o = zr[i];
    o.x += __uint_as_float(a.x); o.y += __uint_as_float(a.y); o.z += __uint_as_float(a.z); o.w += __uint_as_float(a.w);
    if (row < rows_valid) reinterpret_cast<float4*>(zblk + (size_t)row * ldz)[l8] = o;
  }
  __syncwarp();
}


// ------------------------------------------------------------------------------------------
// Attention output projection (model.py:59-65) as a bf16 update of the residual stream:
//     delta[token][128] = attn[M,K] @ Wo_t[128,K]^T + bo            (K = heads * 32 = 256)
//   The FFN kernel that follows adds delta to z while it reads z for its LayerNorm anyway (ffn2_tc_kernel), so this
//   kernel moves 768 B/token and the residual stream makes one HBM round trip per block.  The update is rounded to
//   bf16 (2^-9 relative, the same size as the bf16 rounding of the attention output that feeds it); z itself stays fp32.
//   The [128 x K] weight stays resident in shared memory for the whole (persistent) CTA; only the attention
//   output streams through a ring.  The attention kernel writes it in this kernel's A-operand layout,
//       attn[tile = m >> 7][head][m & 127][32 ch], 16-byte chunks at chunk ^ ((m >> 1) & 3)   (64B swizzle image)
//   with m the SEQUENCE-MAJOR row index (RowMap), so a stage (two heads = one K block of 64) is one contiguous
//   16 KB bulk copy.  Only the epilogue knows about tokens: row m writes the 256-byte row delta[token(m)]
//   (AVB_DELTA_TOKENS; scattered when the block attends over n).  Accumulators are double-buffered in TMEM; eight
//   epilogue warps (TMEM lane quarter x 64-column half) pack them through a per-warp 4 KB shared-memory stage.
//   run_if: when not null the kernel returns at once unless *run_if != 0 (the guarded fallback behind dattn_tc_kernel).
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-11: epilogue
// ------------------------------------------------------------------------------------------
constexpr int kOutProjThreads = 384;
constexpr int kOutProjStages = 6;
constexpr int kOutProjKB = 4;  // K blocks of 64 (K = 256)
constexpr int kOutProjSmemBytes = kOutProjKB * 128 * 128 + kOutProjStages * 128 * 128 + 8 * 4096 + 1024 + 256;

__global__ void __launch_bounds__(kOutProjThreads, 1)
outproj_tc_kernel(const __nv_bfloat16* __restrict__ attn, const __grid_constant__ CUtensorMap tm_w,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ delta, int M, int seq_n,
                  int seq_d, int seq_axis, int seq_dg, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;
  const RowMap rm{seq_n, seq_d, seq_axis, seq_dg};
  constexpr int ST = kOutProjStages, KB = kOutProjKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                          // [KB][128 x 64] bf16, resident
  uint8_t* sA = sW + KB * 128 * 128;           // [ST][128 x 64]
  uint8_t* sStage = sA + ST * 128 * 128;       // [8 warps][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 8 * 4096);
  uint64_t* full = bars;            // [ST]
  uint64_t* empty = bars + ST;      // [ST]
  uint64_t* tfull = bars + 2 * ST;  // [2]
  uint64_t* tempty = tfull + 2;     // [2] 256 epilogue threads
  uint64_t* w_full = tempty + 2;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tm_w);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < ST; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 256); }
    ptx::mbar_init(w_full, 1);
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
    if (ptx::elect_one()) {  // the weight: once
      ptx::mbar_expect_tx(w_full, KB * 128 * 128);
      for (int kb = 0; kb < KB; ++kb) ptx::tma_load_2d(sW + kb * 128 * 128, &tm_w, w_full, kb * 64, 0);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&empty[stage], phase ^ 1, 70, 100u);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full[stage], 128 * 128);  // heads 2kb, 2kb+1 of tile mt: 2 x [128 x 32] = 16 KB contiguous
          ptx::bulk_load(sA + stage * 128 * 128, attn + ((size_t)mt * (2 * KB) + 2 * kb) * (128 * 32), 128 * 128, &full[stage]);
        }
        __syncwarp();
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
    const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 512, ptx::kSwz64);  // [128 x 32] head blocks
    const uint64_t dW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024, ptx::kSwz128);
    ptx::mbar_wait(w_full, 0, 71, 40u);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x, ++it) {
      const int as = it & 1;
      ptx::mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1, 72, 40u);
      ptx::tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&full[stage], phase, 73, 40u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t da = dA0 + (uint64_t)stage * ((128 * 128) >> 4);
          const uint64_t dw = dW0 + (uint64_t)kb * ((128 * 128) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // K slice k of the block: head block k >> 1 (8 KB apart), 32-byte half k & 1 of its rows
            ptx::umma_bf16(tmem_base + as * 128, da + (uint64_t)((k >> 1) * (8192 >> 4) + (k & 1) * 2), dw + (uint64_t)(k * 2), idesc,
                           (kb | k) != 0);
          ptx::umma_commit(&empty[stage]);
          if (kb == KB - 1) ptx::umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t stage_addr = ptx::smem_u32(sStage) + (uint32_t)((warp - 4) * 4096);
    int it = 0;
    for (int mt = blockIdx.x; mt < num_m; mt += gridDim.x, ++it) {
      const int as = it & 1;
      ptx::mbar_wait(&tfull[as], (uint32_t)((it >> 1) & 1), 74);
      ptx::tc_fence_after();
      const int row0 = mt * 128 + q * 32;
      {
        // ---- delta rows: this warp's 32 rows x 64 columns -> bf16 -> [32 rows x 128 B] stage -> 4 rows per store
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + as * 128 + half * 64 + cc * 32 + lane_off, v);
          ptx::tmem_wait_ld();
          if (cc == 1) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[as]);
          }
          const float* bp = bias + half * 64 + cc * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + c);
            pk[cc * 16 + c * 2 + 0] = ptx::pack_bf16(__uint_as_float(v[c * 4 + 0]) + b4.x, __uint_as_float(v[c * 4 + 1]) + b4.y);
            pk[cc * 16 + c * 2 + 1] = ptx::pack_bf16(__uint_as_float(v[c * 4 + 2]) + b4.z, __uint_as_float(v[c * 4 + 3]) + b4.w);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)  // thread = row `lane`: 16-byte chunk c (8 bf16) goes to slot c ^ (lane & 7)
          ptx::st_shared_v4(stage_addr + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4)),
                            make_uint4(pk[c * 4 + 0], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]));
        __syncwarp();
        const int l8 = lane & 7;
#if AVB_DELTA_TOKENS
        int dtok[8];  // delta rows in TOKEN order: the scatter happens here, the FFN reads them next to z
        rm.tokens<8>(row0 + (lane >> 3), 4, M, dtok);
#endif
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          uint4 a;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                       : "r"(stage_addr + (uint32_t)(row * 128 + ((l8 ^ (row & 7)) << 4))));
#if AVB_DELTA_TOKENS
          if (dtok[i] >= 0) *reinterpret_cast<uint4*>(delta + (size_t)dtok[i] * 128 + half * 64 + l8 * 8) = a;
#else
          if (row0 + row < M) *reinterpret_cast<uint4*>(delta + (size_t)(row0 + row) * 128 + half * 64 + l8 * 8) = a;
#endif
        }
        __syncwarp();
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

// L2 prefetch of the 32 fp32 rows (16 KB = 128 lines) a LayerNorm warp will normalise in its NEXT tile: the producer
// warps can only keep ~8 rows of loads in flight in registers, far less than HBM latency x bandwidth needs, so the
// next tile is pulled into L2 while the current one is processed.
__device__ __forceinline__ void prefetch_rows_l2(const float* __restrict__ z, int M, int grow0, int lane,
                                                 const RowMap rm = RowMap{0, 0, 0}) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int line = i * 32 + lane;           // 128-byte line index inside the 32 x 512 B block
    const int row = grow0 + (line >> 2);
    if (row < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(z + (size_t)rm.token(row) * 128 + (line & 3) * 32));
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm (no affine) of 32 consecutive 128-channel fp32 rows by one warp, written as bf16 into a
// 128B-swizzled K-major UMMA A tile ([128 rows x 128 ch] = two K blocks of [128 x 64], 16 KB apart).
// Eight lanes share a row (16 channels each, 3 shuffle steps per statistic) and a warp instruction covers
// four rows, so the global loads are 128-byte segments and the reductions cost 6 shuffles per 4 rows.
//   a_tile: smem address of the A tile; row0: first row of this warp inside the tile (multiple of 32)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_warp_rows_to_umma_tile(const float* __restrict__ z, int M, int grow0, uint32_t a_tile,
                                                          int row0, int lane, const RowMap rm = RowMap{0, 0, 0}) {
  const int rs = lane >> 3, j = lane & 7;
  const float4* zp = reinterpret_cast<const float4*>(z) + j;
  float4 cur[2][4], nxt[2][4];
  auto load = [&](float4 (&dst)[2][4], int r0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int grow = grow0 + r0 + u * 4 + rs;
      const int tok = grow < M ? rm.token(grow) : 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dst[u][k] = grow < M ? __ldcg(zp + (size_t)tok * 32 + k * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load(cur, 0);
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    if (r0 + 8 < 32) load(nxt, r0 + 8);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) sum += (cur[u][k].x + cur[u][k].y) + (cur[u][k].z + cur[u][k].w);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float mean = sum * (1.0f / 128.0f);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cur[u][k].x -= mean; cur[u][k].y -= mean; cur[u][k].z -= mean; cur[u][k].w -= mean;
        sq += (cur[u][k].x * cur[u][k].x + cur[u][k].y * cur[u][k].y) + (cur[u][k].z * cur[u][k].z + cur[u][k].w * cur[u][k].w);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      const float rstd = rsqrtf(sq * (1.0f / 128.0f) + kLnEps);
      const int r = row0 + r0 + u * 4 + rs;
      const uint32_t rbase = a_tile + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (j & 1) * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // channels 4(j+8k)..+3: K block (k >> 1), 16-byte chunk (j + 8(k & 1)) >> 1 inside its 128-byte row
        const int chunk = (j + 8 * (k & 1)) >> 1;
        const uint32_t addr = rbase + (uint32_t)((k >> 1) * (128 * 128) + ((chunk ^ (r & 7)) << 4));
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ptx::pack_bf16(cur[u][k].x * rstd, cur[u][k].y * rstd)),
                     "r"(ptx::pack_bf16(cur[u][k].z * rstd, cur[u][k].w * rstd))
                     : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[u][k] = nxt[u][k];
  }
}

// Same, and the RAW fp32 rows are also parked in a per-warp 16 KB shared-memory stage ([32 rows][32 x 16-byte chunks],
// chunk index XOR-ed with row & 7) from which the warp reads them back one row per thread - the tcgen05.st layout.
// Used by the CTA-pair FFN kernel to preload its residual accumulator with z (see ffn2_tc_kernel).
__device__ __forceinline__ void ln_warp_rows_to_umma_tile_stash(const float* __restrict__ z, int M, int grow0, uint32_t a_tile,
                                                                int row0, int lane, uint32_t stash,
                                                                const __nv_bfloat16* __restrict__ delta = nullptr,
                                                                const RowMap rm = RowMap{0, 0, 0}) {
  const int rs = lane >> 3, j = lane & 7;
  const float4* zp = reinterpret_cast<const float4*>(z) + j;
  float4 cur[2][4], nxt[2][4];
  uint2 dcur[2][4], dnxt[2][4];  // the out-projection's update of the same rows (bf16), added when the row is processed
  int mrow[8];                    // its sequence-major row for each of this lane's 8 rows (grow0 + rs + 4q)
#if AVB_DELTA_TOKENS
#pragma unroll
  for (int q = 0; q < 8; ++q) mrow[q] = grow0 + rs + 4 * q;
#else
  if (delta != nullptr) rm.seq_rows<8>(grow0 + rs, 4, M, mrow);
#endif
  auto load = [&](float4 (&dst)[2][4], uint2 (&ddst)[2][4], int r0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int grow = grow0 + r0 + u * 4 + rs;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        dst[u][k] = grow < M ? __ldcg(zp + (size_t)grow * 32 + k * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) ddst[u][k] = make_uint2(0u, 0u);
      if (delta != nullptr && grow < M) {
        // z1 = z + delta is formed here so that the residual stream is read once for both sub-layers; the loads are
        // only consumed one row group later, like the z loads
        const uint2* dp = reinterpret_cast<const uint2*>(delta + (size_t)mrow[(r0 >> 2) + u] * 128) + j;
#pragma unroll
        for (int k = 0; k < 4; ++k) ddst[u][k] = __ldcg(dp + k * 8);
      }
    }
  };
  load(cur, dcur, 0);
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    if (r0 + 8 < 32) load(nxt, dnxt, r0 + 8);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rl = r0 + u * 4 + rs;  // row inside the warp's 32
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cur[u][k].x += __uint_as_float(dcur[u][k].x << 16); cur[u][k].y += __uint_as_float(dcur[u][k].x & 0xffff0000u);
        cur[u][k].z += __uint_as_float(dcur[u][k].y << 16); cur[u][k].w += __uint_as_float(dcur[u][k].y & 0xffff0000u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)  // raw values: chunk (j + 8k) of row rl
        ptx::st_shared_v4(stash + (uint32_t)(rl * 512 + (((j + 8 * k) ^ (rl & 7)) << 4)),
                          make_uint4(__float_as_uint(cur[u][k].x), __float_as_uint(cur[u][k].y), __float_as_uint(cur[u][k].z),
                                     __float_as_uint(cur[u][k].w)));
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) sum += (cur[u][k].x + cur[u][k].y) + (cur[u][k].z + cur[u][k].w);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float mean = sum * (1.0f / 128.0f);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cur[u][k].x -= mean; cur[u][k].y -= mean; cur[u][k].z -= mean; cur[u][k].w -= mean;
        sq += (cur[u][k].x * cur[u][k].x + cur[u][k].y * cur[u][k].y) + (cur[u][k].z * cur[u][k].z + cur[u][k].w * cur[u][k].w);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      const float rstd = rsqrtf(sq * (1.0f / 128.0f) + kLnEps);
      const int r = row0 + rl;
      const uint32_t rbase = a_tile + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (j & 1) * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int chunk = (j + 8 * (k & 1)) >> 1;
        const uint32_t addr = rbase + (uint32_t)((k >> 1) * (128 * 128) + ((chunk ^ (r & 7)) << 4));
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ptx::pack_bf16(cur[u][k].x * rstd, cur[u][k].y * rstd)),
                     "r"(ptx::pack_bf16(cur[u][k].z * rstd, cur[u][k].w * rstd))
                     : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) { cur[u][k] = nxt[u][k]; dcur[u][k] = dnxt[u][k]; }
  }
}

// ------------------------------------------------------------------------------------------
// Fused LayerNorm -> GEMM:  out[M, N] = act( LN(z)[M,128] @ Wt[N,128]^T + bias[N] ),  bf16 out.
//   (QKV projection: N = 768, act = id;  FFN up-projection: N = 512, act = ReLU.  LayerNorm gamma/beta
//    are folded into Wt / bias at load time, so the prologue only normalises.)
//   The A operand never exists in HBM: 4 "LN" warps read the fp32 residual rows (one coalesced 512 B row
//   per warp-load), normalise with warp shuffles and write bf16 straight into the 128B-swizzled K-major
//   smem layout the UMMA descriptors expect; the [128 x 128] A tile is reused by all N/256 column chunks.
//   Weights stream through a TMA ring; accumulators are double-buffered in TMEM (2 x 256 columns).
//   The epilogue stages [32 x 64] pieces in smem and writes them out as 128-byte row segments, four rows per
//   warp store (the 16-byte-per-lane pieces of a row-per-thread store waste HBM sectors).
//   warps 0-3: LN producers   warp 4: weight TMA   warp 5: MMA issuer   warp 6: TMEM alloc
//   warps 8-15: epilogue (warp = TMEM lane quarter x 128-column half of the 256-column accumulator)
// ------------------------------------------------------------------------------------------
constexpr int kLnGemmThreads = 512;
constexpr int kLnGemmBStages = 3;
constexpr int kLnGemmABytes = 128 * 128 * 2;               // 32 KB: two 128B-swizzled K blocks of 64
constexpr int kLnGemmBBytes = 256 * 64 * 2;                // 32 KB per (256-column chunk, K block)
constexpr int kLnGemmStagePitch = 64 * 2 + 16;             // bytes per staged row (64 bf16 + pad)
constexpr int kLnGemmStageBytes = 8 * 2 * 4096;  // 8 epilogue warps x 2 buffers x 4 KB (pair kernel); the single-CTA kernel
                                                 // uses 8 x 32 rows x pitch = 36 KB of it
constexpr int kLnGemmSmemBytes = 2 * kLnGemmABytes + kLnGemmBStages * kLnGemmBBytes + kLnGemmStageBytes + 1024 + 256;

//   EPI 2 (the QKV projection in front of the tensor-core attention): rows are processed in SEQUENCE-MAJOR order
//   (RowMap: the LN producers gather the fp32 rows, a 512-byte segment each, in whatever order) and the output is
//   written in the attention kernel's operand layout
//       out[seq][w = q,k,v][head][pos][32 ch]   with the four 16-byte chunks of a 64-byte (pos, head) row stored at
//       chunk ^ ((pos >> 1) & 3)                 (the UMMA 64B-swizzle pattern, applied here once)
//   so that a [128 positions x 32 ch] Q / K / V tile is one contiguous 8 KB block whose byte image IS the swizzled
//   shared-memory tile: the attention producer fetches it with a single 1-D bulk copy (no gather, no per-row TMA
//   cost, every DRAM sector used), for both attention axes.  N must be 3 * H * 32 with 32-channel heads.

// ------------------------------------------------------------------------------------------
// CTA-pair version of ln_gemm_tc_kernel (cluster of 2, tcgen05 cta_group::2): the pair works on 256 rows, each CTA
// LayerNorm-s its own 128-row A tile and runs its own epilogue exactly as above, but every weight stage is fetched
// half by each CTA ([128 of the 256 N-rows] x 64 K, 16 KB instead of 32 KB) and ONE M = 256 MMA issued by the leader
// feeds both SMs.  The single-CTA kernel re-streams all 192 KB of QKV weights out of L2 for every 128 tokens and is
// bound by the L2 -> SM fabric (measured ~29 B/clk/SM, reads + writes); the pair halves the weight share.
// tm_b: tensor map of Wt with [128 rows x 64] boxes.
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLnGemmThreads, 1)
ln_gemm2_tc_kernel(const float* __restrict__ z, const __grid_constant__ CUtensorMap tm_b,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int ldo, int M, int N,
                  int seq_n, int seq_d, int seq_axis, int seq_dg, const int* __restrict__ run_if, float* __restrict__ norms) {
  // run_if: when not null the kernel returns at once unless *run_if != 0 (the guarded fallback behind dattn_tc_kernel)
  // norms (may be null, zeroed by the caller): {max |q_h|^2, max |k_h|^2} over the (row, head) vectors this launch writes,
  //   the score bound that lets the attention kernel drop the range clamps of its polynomial exp2 lanes
  if (run_if != nullptr && *run_if == 0) return;
  const RowMap rm{seq_n, seq_d, EPI == 2 ? seq_axis : 0, EPI == 2 ? seq_dg : 0};
  constexpr int ST = kLnGemmBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // [2][32K]
  uint8_t* sB = sA + 2 * kLnGemmABytes;                 // [ST][32K]
  uint8_t* sStage = sB + ST * kLnGemmBBytes;            // [128][pitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kLnGemmStageBytes);
  uint64_t* a_full = bars;             // [2] leader: 2 x 128 LN threads
  uint64_t* a_empty = bars + 2;        // [2] per CTA (multicast commit)
  uint64_t* b_full = bars + 4;         // [ST]
  uint64_t* b_empty = bars + 4 + ST;   // [ST]
  uint64_t* t_full = bars + 4 + 2 * ST;  // [2]
  uint64_t* t_empty = t_full + 2;      // [2] leader: 2 x 256 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128, num_c = N / 256;
  const uint32_t rank = ptx::cluster_ctarank();
  const int num_pairs = (num_m + 1) / 2, pair0 = (int)blockIdx.x >> 1, pair_stride = (int)gridDim.x >> 1;

  if (warp == 4 && lane == 0) ptx::prefetch_tmap(&tm_b);
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 256); ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 512);
    }
    for (int i = 0; i < ST; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 6) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives on the peer's
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== LN producers: warp w normalises rows 32w .. 32w+31 of each tile
    int it = 0;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride, ++it) {
      const int buf = it & 1;
      prefetch_rows_l2(z, M, (mt + 2 * pair_stride) * 128 + warp * 32, lane, rm);
      ptx::mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1, 40);
      ln_warp_rows_to_umma_tile(z, M, mt * 128 + warp * 32, ptx::smem_u32(sA + buf * kLnGemmABytes), warp * 32, lane, rm);
      ptx::fence_proxy_async();
      ptx::mbar_arrive_cluster(&a_full[buf], 0);
    }
  } else if (warp == 4) {
    // ===================== weight TMA producer: for every m tile, N/256 chunks x 2 K blocks of [256 x 64]
    int stage = 0; uint32_t phase = 0;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride) {
      for (int c = 0; c < num_c; ++c) {
        for (int kb = 0; kb < 2; ++kb) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1, 41);
          if (ptx::elect_one()) {  // this CTA's half of the stage: N-rows [c*256 + rank*128, +128)
            if (rank == 0) ptx::mbar_expect_tx(&b_full[stage], kLnGemmBBytes);  // both halves complete on the leader's barrier
            ptx::tma_load_2d_2sm(sB + stage * kLnGemmBBytes, &tm_b, &b_full[stage], kb * 64, c * 256 + (int)rank * 128);
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5 && rank == 0) {
    // ===================== MMA issuer (leader CTA only): M = 256 over the pair
    constexpr uint32_t idesc = ptx::make_idesc_bf16(256, 256, 0, 0);
    const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024, ptx::kSwz128);
    const uint64_t dB0 = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 1024, ptx::kSwz128);
    int stage = 0; uint32_t phase = 0;
    int it = 0, acc_it = 0;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride, ++it) {
      const int buf = it & 1;
      ptx::mbar_wait(&a_full[buf], (it >> 1) & 1, 42);
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_empty[as], ((acc_it >> 1) & 1) ^ 1, 43);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < 2; ++kb) {
          ptx::mbar_wait(&b_full[stage], phase, 44);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t da = dA0 + (uint64_t)(buf * (kLnGemmABytes >> 4) + kb * ((128 * 128) >> 4));
            const uint64_t db = dB0 + (uint64_t)stage * (kLnGemmBBytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            ptx::umma_commit_2sm(&b_empty[stage], 3u);
            if (kb == 1) {
              ptx::umma_commit_2sm(&t_full[as], 3u);
              if (c == num_c - 1) ptx::umma_commit_2sm(&a_empty[buf], 3u);
            }
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: TMEM -> +bias -> bf16 -> smem stage -> bulk stores (attention operand layout)
    // A piece of 64 accumulator columns = heads (hd0, hd0+1) of q, k or v for this warp's 32 rows.  It is staged as
    // [2 heads][32 rows][64 B] with the 16-byte chunks already at chunk ^ ((pos >> 1) & 3): that is both the byte image
    // of the output (each head's rows are one contiguous run per sequence) and a conflict-free shared-memory pattern
    // for the row-per-thread writes.  One lane then issues a bulk copy per head and run: the LSU / L1 path, which the
    // ld.shared + st.global copy-out kept ~80 % busy (ncu l1tex throughput), only sees the staging writes.
    static_assert(EPI == 2, "the CTA-pair kernel is built for the QKV projection");
    const int q = warp & 3, hf = (warp - 8) >> 2;
    const uint32_t stage_warp = ptx::smem_u32(sStage) + (uint32_t)((warp - 8) * 2 * 4096);
    const int seq_T = seq_axis == 0 ? seq_d : seq_n, seq_H = N / 96;  // sequence length, heads (N = 3 * H * 32)
    int acc_it = 0;
    uint32_t piece = 0;  // stage buffer parity
    float nrm_q = 0.f, nrm_k = 0.f;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride) {
      const int grow0 = mt * 128 + q * 32;                    // first row (sequence-major index) of this warp
      const int rows_valid = M - grow0 < 32 ? M - grow0 : 32;  // may be <= 0
      const unsigned my_m = (unsigned)(grow0 + lane);
      const unsigned my_pos = my_m % (unsigned)seq_T;
      const uint32_t my_swz = (my_pos >> 1) & 3;
      for (int c = 0; c < num_c; ++c, ++acc_it) {
        const int as = acc_it & 1;
        ptx::mbar_wait(&t_full[as], (acc_it >> 1) & 1, 45);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + as * 256 + hf * 128 + ((uint32_t)(q * 32) << 16);
        const int col0 = c * 256 + hf * 128;
#pragma unroll 1
        for (int pc = 0; pc < 2; ++pc, piece ^= 1u) {  // two pieces of 64 columns
          const uint32_t stage = stage_warp + piece * 4096;
          // the bulk stores that read this buffer two pieces ago must be done with it
          if (lane == 0) ptx::bulk_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float4 bb[8];  // bias of these 32 columns, requested before the TMEM load so the latencies overlap
#pragma unroll
            for (int e = 0; e < 8; ++e) bb[e] = __ldg(reinterpret_cast<const float4*>(bias + col0 + pc * 64 + cc * 32) + e);
            uint32_t v[32];
            ptx::tmem_ld32(taddr + pc * 64 + cc * 32, v);
            ptx::tmem_wait_ld();
            if (pc == 1 && cc == 1) {  // accumulator fully read by this warp: hand it back to the MMA warp
              ptx::tc_fence_before();
              ptx::mbar_arrive_cluster(&t_empty[as], 0);
            }
            uint64_t ss2 = pack2(0.f, 0.f);  // |row . head|^2 of these 32 channels (q and k chunks only), two partial sums
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f[8];
              f[0] = __uint_as_float(v[e * 8 + 0]) + bb[e * 2].x;     f[1] = __uint_as_float(v[e * 8 + 1]) + bb[e * 2].y;
              f[2] = __uint_as_float(v[e * 8 + 2]) + bb[e * 2].z;     f[3] = __uint_as_float(v[e * 8 + 3]) + bb[e * 2].w;
              f[4] = __uint_as_float(v[e * 8 + 4]) + bb[e * 2 + 1].x; f[5] = __uint_as_float(v[e * 8 + 5]) + bb[e * 2 + 1].y;
              f[6] = __uint_as_float(v[e * 8 + 6]) + bb[e * 2 + 1].z; f[7] = __uint_as_float(v[e * 8 + 7]) + bb[e * 2 + 1].w;
              if (c < 2) {  // columns [0, 512): q and k (N = 768, 256-column chunks); packed f32x2: half the issue slots
#pragma unroll
                for (int t = 0; t < 8; t += 2) {
                  const uint64_t f2 = pack2(f[t], f[t + 1]);
                  ss2 = fma2(f2, f2, ss2);
                }
              }
              uint4 pk;
              pk.x = ptx::pack_bf16(f[0], f[1]); pk.y = ptx::pack_bf16(f[2], f[3]);
              pk.z = ptx::pack_bf16(f[4], f[5]); pk.w = ptx::pack_bf16(f[6], f[7]);
              ptx::st_shared_v4(stage + (uint32_t)(cc * 2048 + lane * 64 + (((uint32_t)e ^ my_swz) << 4)), pk);
            }
            if (c < 2 && lane < rows_valid) {
              float ss0, ss1;
              unpack2(ss2, ss0, ss1);
              if (c == 0) nrm_q = fmaxf(nrm_q, ss0 + ss1); else nrm_k = fmaxf(nrm_k, ss0 + ss1);
            }
          }
          ptx::fence_proxy_async();  // staging writes (generic proxy) -> visible to the bulk copy engine
          __syncwarp();
          if (lane == 0 && rows_valid > 0) {
            const int colp = col0 + pc * 64, w = colp >> 8, hd0 = (colp & 255) >> 5;
            int r = 0;
            while (r < rows_valid) {  // one run per sequence the 32 rows touch (usually one)
              const unsigned m = (unsigned)(grow0 + r);
              const unsigned sq = m / (unsigned)seq_T, pos = m - sq * (unsigned)seq_T;
              const int len = min(rows_valid - r, seq_T - (int)pos);
              __nv_bfloat16* dst = out + ((((size_t)sq * 3 + w) * seq_H + hd0) * seq_T + pos) * 32;
              ptx::bulk_store(dst, stage + (uint32_t)(r * 64), (uint32_t)len * 64u);
              ptx::bulk_store(dst + (size_t)seq_T * 32, stage + 2048u + (uint32_t)(r * 64), (uint32_t)len * 64u);
              r += len;
            }
          }
          if (lane == 0) ptx::bulk_commit();
        }
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
    __syncwarp();
    if (norms != nullptr) {  // one atomic per warp and operand (non-negative floats order like their bit patterns)
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
  ptx::cluster_sync_all();  // the peer's shared memory and barriers stay valid until both CTAs are done
  if (warp == 6) ptx::tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// Fused FFN sub-layer, in place on the fp32 residual stream (model.py:67-75):
//     z += relu( LN(z) @ W1' + b1' ) @ W2 + b2          (LayerNorm affine folded into W1' / b1')
//   Neither the LayerNorm output nor the 512-wide hidden activation ever touch HBM: per 128-token tile the
//   hidden layer is produced in four 128-unit chunks,
//     G1_c : D1[128x128 fp32, TMEM] = xhat[smem, bf16] . W1c^T[smem]            (8 MMAs 128x128x16)
//     E1_c : D1 -> +b1, ReLU -> bf16 pairs -> H[TMEM]                            (8 epilogue warps)
//     G2_c : D2[128x128 fp32, TMEM] += H[TMEM, A operand] . W2c[smem]            (8 MMAs, A from TMEM)
//   and after the fourth chunk  E2 : z += D2 + b2.  H is double-buffered (G1_{c+1} is issued as soon as E1_c has read
//   D1, so the tensor core works while E1_c packs H_c) and D2 is double-buffered by tile, so E2 of a tile is
//   deferred behind the first E1 of the next tile.  TMEM: 128 (D1) + 2x64 (H) + 2x128 (D2) = 512 columns.
//   Weights stream from L2 through a 2-stage TMA ring (64 KB per chunk: W1c and W2c).
//   warps 0-3: LN producers   warp 4: weight TMA   warp 5: MMA issuer   warp 6: TMEM alloc   warps 8-15: epilogues
// ------------------------------------------------------------------------------------------
constexpr int kFfnThreads = 512;
constexpr int kFfnWBytes = 128 * 128 * 2;  // 32 KB: one [128 x 128] bf16 weight chunk = two 128B-swizzled K blocks
constexpr int kFfnSmemBytes = 2 * kLnGemmABytes + 2 * 2 * kFfnWBytes + 8 * 4096 + 1024 + 256;  // + 8 x 4 KB epilogue stages


// ------------------------------------------------------------------------------------------
// CTA-pair version of ffn_tc_kernel (cluster of 2, tcgen05 cta_group::2).  Each CTA keeps its own 128-token tile,
// LayerNorm, E1 / E2 epilogues and TMEM exactly as above; the pair shares every weight chunk: each CTA fetches 64 of
// the 128 N-rows of W1c and of W2c (half the L2 -> SM weight traffic, which is 2/3 of this kernel's reads and what
// bounds it) and the leader issues M = 256 MMAs for both SMs.  Barriers the MMA issuer waits on live in the leader
// and count the arrivals of both CTAs; its commits are multicast to both.
// Differences from the single-CTA schedule above (each measured, profiles/r1f_summary.md, r1g_summary.md):
//  * the shared memory the halved weight ring frees carries the raw z rows from the LayerNorm warps into TMEM: D2 starts
//    as the residual, G2 always accumulates, and the final epilogue is write-only (no second read of z) and leaves
//    through 128B-swizzled tensor-map stores;
//  * D1 is double-buffered by chunk and H is packed IN PLACE over the D1 columns its warp has just read, so G1 of the
//    next chunk overlaps E1 entirely.  TMEM: 2x128 (D1, H inside) + 2x128 (D2) = 512 columns.
// tm_w1 / tm_w2: tensor maps with [64 rows x 64] boxes; tm_z: the residual stream, [32 rows x 32 cols] fp32 boxes.
// ------------------------------------------------------------------------------------------
// kResident (the default, AVB_FFN_RESIDENT=0 selects the ring version): ALL of this CTA's weight halves (4 chunks x
// [W1c half 16K | W2c half 16K] = 128 KB) are fetched ONCE and stay in shared memory, so the only per-tile traffic
// through the L2 -> SM fabric (what bounds the ring version) is the tile's own z and delta rows.  The 64 KB this needs
// is the stash: the LayerNorm warps instead read their rows one ROW PER THREAD (256-bit loads, a full sector per
// lane), write z1 = z + delta straight into the D2 accumulator (tcgen05.st: the thread's row is its TMEM lane) while
// accumulating sum and sum of squares, then read the row back from TMEM to normalise it into the A tile: TMEM is the
// stash.  No shuffles, no shared-memory round trip.
// AVB_FFN_DECOUPLE=1 (A/B build, not measured yet): breaks the serial chain LN(t) -> G1(t,0) -> E1(t,0) -> E2(t-1) ->
// LN(t+1) found by the trace (profiles/r1j_summary.md) at both ends: the MMA issuer never blocks on the next tile's
// LayerNorm while down-projections of the current tile are still unissued, and the epilogue warps run E2 of the
// previous tile first whenever the current tile's first D1 is not ready.  (E2-first alone was measured slower: with the
// blocking issuer, d2_full of the previous tile itself waited for the LayerNorm.)
#ifndef AVB_FFN_DECOUPLE
#define AVB_FFN_DECOUPLE 0
#endif
#ifndef AVB_FFN_E1_X64
#define AVB_FFN_E1_X64 0  // 1: E1 reads its 64 D1 columns with one tcgen05.ld.x64 (A/B build, not measured yet)
#endif
#ifndef AVB_FFN_E2_EARLY
#define AVB_FFN_E2_EARLY 1  // 1: E2 reads both halves of D2 and releases the buffer before staging / storing them
#endif
#ifndef AVB_FFN_E2_PLAIN
#define AVB_FFN_E2_PLAIN 0  // 1: final epilogue leaves through plain coalesced stores instead of tensor-map stores (A/B build)
#endif
// The FFN biases of the resident-weight variant (F = 512) travel as a kernel parameter and are read from the constant
// bank: a __ldg from the epilogue warps queues in the L1 pipeline behind the LayerNorm warps' streaming row loads (found
// on the fused d-axis kernel, r2b: ~500 clocks of long-scoreboard stall per 32-column chunk), and there are 1.7 KB of
// shared memory left in this kernel, less than b1.  Measured in the bench step: FFN 33.7 -> 30.6 ms (standalone, without
// the delta rows and the other kernels' L2 traffic, no difference).  The same change in the QKV projection (HBM-bound)
// made no difference and was not kept.
struct FfnBias { float b1[512]; float b2[128]; };
template <bool kResident>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFfnThreads, 1)
ffn2_tc_kernel(float* __restrict__ z, const __nv_bfloat16* __restrict__ delta, int seq_n, int seq_d, int seq_axis,
               const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2,
              const float* __restrict__ b1, const float* __restrict__ b2, int M, int F,
               const __grid_constant__ FfnBias cb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                        // [2][32K] xhat
  uint8_t* sW = sA + 2 * kLnGemmABytes;      // [2 stages][W1c half 16K | W2c half 16K]: 64 N-rows per CTA
  uint8_t* sZ = sW + 2 * kFfnWBytes;         // [4 LN warps][16 KB]: raw z rows on their way into TMEM (the half of the
                                             // weight ring the pair does not need)
  uint8_t* sStage = sW + 4 * kFfnWBytes;     // [8 epilogue warps][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 8 * 4096);
  uint64_t* a_full = bars;          // [2] 128 LN threads
  uint64_t* a_empty = bars + 2;     // [2]
  constexpr int WS = kResident ? 4 : 2;  // weight stages: ring of 2, or the four chunks resident
  uint64_t* w1_full = bars + 4;     // [WS] leader
  uint64_t* w1_empty = bars + 8;    // [WS]
  uint64_t* w2_full = bars + 12;    // [WS] leader
  uint64_t* w2_empty = bars + 16;   // [WS]
  uint64_t* d1_full = bars + 20;    // [2] by chunk parity
  uint64_t* h_full = bars + 22;     // [2] leader: 2 x 256 epilogue threads
  uint64_t* h_empty = bars + 24;    // [2]
  uint64_t* d2_full = bars + 26;    // [2]
  uint64_t* d2_empty = bars + 28;   // [2] local: 256 epilogue threads have read D2[i] (the LN warps may preload it again)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_m = (M + 127) / 128, num_c = F / 128;
  const uint32_t rank = ptx::cluster_ctarank();
  const int num_pairs = (num_m + 1) / 2, pair0 = (int)blockIdx.x >> 1, pair_stride = (int)gridDim.x >> 1;
  long long* trace = (blockIdx.x == 0 && (warp == 5 || warp == 8 || warp == 0) && lane == 0) ? g_attn_trace : nullptr;

  if (warp == 4 && lane == 0) { ptx::prefetch_tmap(&tm_w1); ptx::prefetch_tmap(&tm_w2); }
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < WS; ++i) {
      ptx::mbar_init(&w1_full[i], 1); ptx::mbar_init(&w1_empty[i], 1);
      ptx::mbar_init(&w2_full[i], 1); ptx::mbar_init(&w2_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 256); ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&h_full[i], 512); ptx::mbar_init(&h_empty[i], 1);
      ptx::mbar_init(&d2_full[i], 1); ptx::mbar_init(&d2_empty[i], 256);
    }
    ptx::mbar_init(&d1_full[0], 1); ptx::mbar_init(&d1_full[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 6) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D1 is double-buffered by chunk parity and the packed hidden activation H is written IN PLACE over the columns its
  // warp has just read (half 0: D1 columns 0-31, half 1: columns 64-95), so G1 of the next chunk runs while E1 is
  // still reading / packing this one, without a separate H buffer.  MMAs execute in issue order, which is all the
  // reuse of a D1 buffer (G1 of chunk g+2 after G2 of chunk g) needs.
  const uint32_t tmem_D1 = tmem_base;        // 2 x 128 (by chunk parity); H inside
  const uint32_t tmem_D2 = tmem_base + 256;  // 2 x 128 (by tile parity)

  if (warp < 4) {
    // ===================== LN producers (same as ln_gemm_tc_kernel)
    int it = 0;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride, ++it) {
      const int buf = it & 1;
      AVB_TRACE(0, it, 0);
      prefetch_rows_l2(z, M, (mt + 2 * pair_stride) * 128 + warp * 32, lane);
      if constexpr (kResident) {
        // ---- row per thread: lane = row = TMEM lane
        const int r = warp * 32 + lane, grow = mt * 128 + r;
        const bool valid = grow < M;
        const float* zr = z + (size_t)(valid ? grow : 0) * 128;
        const __nv_bfloat16* dr = delta ? delta + (size_t)(valid ? grow : 0) * 128 : nullptr;
        if (delta != nullptr) {  // next tile's delta rows towards L2 (z: prefetch_rows_l2 above)
          const int nrow = (mt + 2 * pair_stride) * 128 + warp * 32 + (lane >> 1);  // 16 rows x 2 lines per pass
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (nrow + i * 16 < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(delta + (size_t)(nrow + i * 16) * 128 + (lane & 1) * 64));
        }
        // 8 blocks of 16 columns, loads two blocks ahead in a rotating 3-slot register buffer (72 registers):
        // ~6 KB per warp in flight, enough for L2-latency x the 10 B/clk/SM this stream needs
        float zb[3][2][8];
        uint32_t db[3][8];
        auto load = [&](int slot, int blk) {
          if (valid) {
            ptx::ldg_v8_f32(zr + blk * 16, zb[slot][0]);
            ptx::ldg_v8_f32(zr + blk * 16 + 8, zb[slot][1]);
            if (dr != nullptr) ptx::ldg_v8_b32(dr + blk * 16, db[slot]);
          }
        };
#pragma unroll
        for (int sl = 0; sl < 3; ++sl)
#pragma unroll
          for (int e = 0; e < 8; ++e) { zb[sl][0][e] = 0.f; zb[sl][1][e] = 0.f; db[sl][e] = 0u; }
        load(0, 0);
        load(1, 1);
        if (it >= 2) ptx::mbar_wait(&d2_empty[buf], (uint32_t)(((it >> 1) - 1) & 1), 62);  // E2 of tile it-2 has read D2[buf]
        AVB_TRACE(0, it, 1);
        ptx::tc_fence_after();
        const uint32_t trow = tmem_D2 + buf * 128 + ((uint32_t)(warp * 32) << 16);
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int blk = 0; blk < 8; ++blk) {
          const int sl = blk % 3;
          if (blk + 2 < 8) load((blk + 2) % 3, blk + 2);
          uint32_t v[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {  // columns 2e, 2e+1 of this 16-column block; delta word e holds both
            const uint32_t dw = db[sl][e];
            const float x0 = zb[sl][e >> 2][(e & 3) * 2] + __uint_as_float(dw << 16);
            const float x1 = zb[sl][e >> 2][(e & 3) * 2 + 1] + __uint_as_float(dw & 0xffff0000u);
            sum += x0 + x1;
            sq = fmaf(x0, x0, sq);
            sq = fmaf(x1, x1, sq);
            v[2 * e] = __float_as_uint(x0);
            v[2 * e + 1] = __float_as_uint(x1);
          }
          ptx::tmem_st16(trow + blk * 16, v);
        }
        ptx::tmem_wait_st();
        const float mean = sum * (1.0f / 128.0f);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f) + kLnEps);
        const float shift = -mean * rstd;
        AVB_TRACE(0, it, 3);
        ptx::mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1, 50);  // G1 of tile it-2 has read this A tile
        AVB_TRACE(0, it, 4);
        const uint32_t rbase = ptx::smem_u32(sA + buf * kLnGemmABytes) + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          uint32_t v[32];
          ptx::tmem_ld32(trow + blk * 32, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {  // 16-byte chunk (blk & 1) * 4 + c4 of K block blk >> 1
            uint4 pk;
            pk.x = ptx::pack_bf16(fmaf(__uint_as_float(v[c4 * 8 + 0]), rstd, shift), fmaf(__uint_as_float(v[c4 * 8 + 1]), rstd, shift));
            pk.y = ptx::pack_bf16(fmaf(__uint_as_float(v[c4 * 8 + 2]), rstd, shift), fmaf(__uint_as_float(v[c4 * 8 + 3]), rstd, shift));
            pk.z = ptx::pack_bf16(fmaf(__uint_as_float(v[c4 * 8 + 4]), rstd, shift), fmaf(__uint_as_float(v[c4 * 8 + 5]), rstd, shift));
            pk.w = ptx::pack_bf16(fmaf(__uint_as_float(v[c4 * 8 + 6]), rstd, shift), fmaf(__uint_as_float(v[c4 * 8 + 7]), rstd, shift));
            const int chunk = (blk & 1) * 4 + c4;
            ptx::st_shared_v4(rbase + (uint32_t)((blk >> 1) * (128 * 128) + ((chunk ^ (r & 7)) << 4)), pk);
          }
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();
        ptx::mbar_arrive_cluster(&a_full[buf], 0);
        AVB_TRACE(0, it, 2);
        continue;
      }
      ptx::mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1, 50);
      AVB_TRACE(0, it, 1);
      const uint32_t stash = ptx::smem_u32(sZ) + (uint32_t)(warp * 16384);
      ln_warp_rows_to_umma_tile_stash(z, M, mt * 128 + warp * 32, ptx::smem_u32(sA + buf * kLnGemmABytes), warp * 32, lane, stash,
                                      delta, RowMap{seq_n, seq_d, seq_axis});
      // D2[buf] := z rows of this tile (fp32, exact): the down-projection then accumulates onto the residual and the
      // final epilogue only adds b2 and stores - the second read of z (64 KB per tile through L2) is gone.
      if (it >= 2) ptx::mbar_wait(&d2_empty[buf], (uint32_t)(((it >> 1) - 1) & 1), 62);  // E2 of tile it-2 has read D2[buf]
      ptx::tc_fence_after();
      __syncwarp();
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 q4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(q4.x), "=r"(q4.y), "=r"(q4.z), "=r"(q4.w)
                       : "r"(stash + (uint32_t)(lane * 512 + (((blk * 8 + c) ^ (lane & 7)) << 4))));
          v[c * 4 + 0] = q4.x; v[c * 4 + 1] = q4.y; v[c * 4 + 2] = q4.z; v[c * 4 + 3] = q4.w;
        }
        ptx::tmem_st32(tmem_D2 + buf * 128 + blk * 32 + ((uint32_t)(warp * 32) << 16), v);
      }
      ptx::tmem_wait_st();
      __syncwarp();  // the stash is rewritten by the next tile
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      ptx::mbar_arrive_cluster(&a_full[buf], 0);
      AVB_TRACE(0, it, 2);
    }
  } else if (warp == 4) {
    // ===================== weight TMA producer: per chunk W1c = W1t[128c.., 0..127], W2c = W2t[0..127, 128c..]
    int g = 0;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride) {
      if (kResident && g >= num_c) break;  // resident: the four chunks are fetched once, stage = chunk
      for (int c = 0; c < num_c; ++c, ++g) {
        const int st = kResident ? c : (g & 1);
        const uint32_t ph = (uint32_t)(((kResident ? 0 : g >> 1)) & 1) ^ 1;
        uint8_t* w1 = sW + st * kFfnWBytes;       // 32 KB per stage: W1c half | W2c half
        uint8_t* w2 = w1 + kFfnWBytes / 2;
        ptx::mbar_wait(&w1_empty[st], ph, 51);
        if (ptx::elect_one()) {  // this CTA's 64 of the chunk's 128 N-rows (hidden units), both K blocks
          if (rank == 0) ptx::mbar_expect_tx(&w1_full[st], kFfnWBytes);  // both halves complete on the leader's barrier
          ptx::tma_load_2d_2sm(w1, &tm_w1, &w1_full[st], 0, c * 128 + (int)rank * 64);
          ptx::tma_load_2d_2sm(w1 + 64 * 128, &tm_w1, &w1_full[st], 64, c * 128 + (int)rank * 64);
        }
        __syncwarp();
        ptx::mbar_wait(&w2_empty[st], ph, 52);
        if (ptx::elect_one()) {  // this CTA's 64 of the 128 N-rows (output channels)
          if (rank == 0) ptx::mbar_expect_tx(&w2_full[st], kFfnWBytes);
          ptx::tma_load_2d_2sm(w2, &tm_w2, &w2_full[st], c * 128, (int)rank * 64);
          ptx::tma_load_2d_2sm(w2 + 64 * 128, &tm_w2, &w2_full[st], c * 128 + 64, (int)rank * 64);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5 && rank == 0) {
    // ===================== MMA issuer (leader CTA): G1(g+1) is issued before G2(g) so the tensor core overlaps E1
    constexpr uint32_t idesc = ptx::make_idesc_bf16(256, 128, 0, 0);
    const uint64_t dA0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 1024, ptx::kSwz128);
    const uint64_t dW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024, ptx::kSwz128);
    const int my_tiles = pair0 < num_pairs ? (num_pairs - 1 - pair0) / pair_stride + 1 : 0;
    const int total = my_tiles * num_c;
    auto issue_G1 = [&](int g, bool block = true) -> bool {
      const int tile_it = g / num_c, c = g % num_c, buf = tile_it & 1, st = kResident ? c : (g & 1);
      if (c == 0) {
        if (block) ptx::mbar_wait(&a_full[buf], (uint32_t)((tile_it >> 1) & 1), 53);
        else if (!ptx::mbar_test_warp(&a_full[buf], (uint32_t)((tile_it >> 1) & 1))) return false;
      }
      ptx::mbar_wait(&w1_full[st], kResident ? 0u : (uint32_t)((g >> 1) & 1), 54);  // resident: completes once, stays complete
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t da = dA0 + (uint64_t)(buf * (kLnGemmABytes >> 4));
        const uint64_t dw = dW0 + (uint64_t)(st * (kFfnWBytes >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // K block (kk >> 2): 16 KB further in A (128 rows), 8 KB in the 64-row W half
          const uint64_t offa = (uint64_t)((kk >> 2) * ((128 * 128) >> 4) + (kk & 3) * 2);
          const uint64_t offw = (uint64_t)((kk >> 2) * ((64 * 128) >> 4) + (kk & 3) * 2);
          ptx::umma_bf16_2sm(tmem_D1 + (g & 1) * 128, da + offa, dw + offw, idesc, kk != 0);
        }
        if (!kResident) ptx::umma_commit_2sm(&w1_empty[st], 3u);
        ptx::umma_commit_2sm(&d1_full[g & 1], 3u);
        if (c == num_c - 1) ptx::umma_commit_2sm(&a_empty[buf], 3u);
      }
      __syncwarp();
      return true;
    };
    if (total > 0) issue_G1(0);
    if (total > 1) issue_G1(1);
    [[maybe_unused]] int next_g1 = 2;  // AVB_FFN_DECOUPLE: first up-projection not issued yet
    for (int g = 0; g < total; ++g) {
#if AVB_FFN_DECOUPLE
      while (next_g1 <= g) { issue_G1(next_g1, true); ++next_g1; }  // this chunk's own up-projection is due now
#endif
      AVB_TRACE(1, g, 0);
      AVB_TRACE(1, g, 1);
      const int tile_it = g / num_c, c = g % num_c, st = kResident ? c : (g & 1), hs = g & 1;
      ptx::mbar_wait(&w2_full[st], kResident ? 0u : (uint32_t)((g >> 1) & 1), 56);
      AVB_TRACE(1, g, 2);
      ptx::mbar_wait(&h_full[hs], (uint32_t)((g >> 1) & 1), 57);
      AVB_TRACE(1, g, 3);
      // D2[tile_it & 1] already holds the tile's residual rows (LN warps, ordered by a_full): always accumulate
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dw = dW0 + (uint64_t)(st * (kFfnWBytes >> 4) + ((kFfnWBytes / 2) >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t off = (uint64_t)((kk >> 2) * ((64 * 128) >> 4) + (kk & 3) * 2);
          // H of this chunk: 16 hidden units (8 packed columns) per K slice; slices 0-3 at D1 columns 0-31, 4-7 at 64-95
          const uint32_t ah = tmem_D1 + hs * 128 + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8);
          ptx::umma_bf16_ts_2sm(tmem_D2 + (tile_it & 1) * 128, ah, dw + off, idesc, 1u);
        }
        if (!kResident) ptx::umma_commit_2sm(&w2_empty[st], 3u);
        if (c == num_c - 1) ptx::umma_commit_2sm(&d2_full[tile_it & 1], 3u);
      }
      __syncwarp();
      AVB_TRACE(1, g, 4);
#if AVB_FFN_DECOUPLE
      // up to two chunks ahead, but never WAITING for the next tile's LayerNorm here: the remaining down-projections of
      // this tile (and with them d2_full -> E2 -> the D2 buffer the LayerNorm warps need) must not queue behind it
      while (next_g1 <= g + 2 && next_g1 < total) {
        if (!issue_G1(next_g1, false)) break;
        ++next_g1;
      }
#else
      if (g + 2 < total) issue_G1(g + 2);  // into the D1 buffer G2(g) has just been issued on (MMAs run in issue order)
#endif
    }
  } else if (warp >= 8) {
    // ===================== epilogue warps: (TMEM lane quarter q) x (64-column half)
    const int q = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t stage = ptx::smem_u32(sStage) + (uint32_t)((warp - 8) * 4096);
    // E2 of a tile (z += D2 + b2) is deferred until after the first E1 of the NEXT tile: D2 is double-buffered by tile
    // parity, so the wait for the tile's last G2 MMAs and the z round trip overlap the next tile's MMAs.
    [[maybe_unused]] int e2_g = 0;       // chunk index the trace stamps of E2 are filed under
    auto E2 = [&](int t_it, int t_mt) {  // z_new = D2 + b2 (D2 = z + FFN already): write-only
      // Each [32 rows x 32 cols] fp32 block is staged with the 128B-swizzle pattern (chunk ^ (row & 7): conflict-free
      // for the row-per-thread writes) and leaves through ONE tensor-map store; rows past M are clipped by the TMA.
      ptx::mbar_wait(&d2_full[t_it & 1], (uint32_t)((t_it >> 1) & 1), 61);
      AVB_TRACE(2, e2_g, 5);
      ptx::tc_fence_after();
      const int row0 = t_mt * 128 + q * 32;
#if AVB_FFN_E2_EARLY
      // both 32-column halves leave TMEM first and the D2 buffer is handed back at once: the LayerNorm warps of the tile after
      // next wait for exactly this (r2f trace: 3.6K of their 8.7K clocks per tile), not for the staging and stores below
      uint32_t vv[2][32];
      ptx::tmem_ld32(tmem_D2 + (t_it & 1) * 128 + half * 64 + lane_off, vv[0]);
      ptx::tmem_ld32(tmem_D2 + (t_it & 1) * 128 + half * 64 + 32 + lane_off, vv[1]);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&d2_empty[t_it & 1]);
      AVB_TRACE(2, e2_g, 7);
#endif
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
#if AVB_FFN_E2_EARLY
        uint32_t (&v)[32] = vv[cc];
#else
        uint32_t v[32];
        ptx::tmem_ld32(tmem_D2 + (t_it & 1) * 128 + half * 64 + cc * 32 + lane_off, v);
        ptx::tmem_wait_ld();
        if (cc == 0) AVB_TRACE(2, e2_g, 7);
        if (cc == 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&d2_empty[t_it & 1]);
        }
#endif
#if !AVB_FFN_E2_PLAIN
        if (lane == 0) ptx::bulk_wait_read<0>();  // the previous block's store has read the stage
#endif
        __syncwarp();
        const float* bias = b2 + half * 64 + cc * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // thread = row `lane`: 16-byte chunk c goes to slot c ^ (lane & 7)
          const float4 b4 = kResident ? *reinterpret_cast<const float4*>(&cb.b2[half * 64 + cc * 32 + c * 4])
                                      : __ldg(reinterpret_cast<const float4*>(bias) + c);
          uint4 pk;
          pk.x = __float_as_uint(__uint_as_float(v[c * 4 + 0]) + b4.x);
          pk.y = __float_as_uint(__uint_as_float(v[c * 4 + 1]) + b4.y);
          pk.z = __float_as_uint(__uint_as_float(v[c * 4 + 2]) + b4.z);
          pk.w = __float_as_uint(__uint_as_float(v[c * 4 + 3]) + b4.w);
          ptx::st_shared_v4(stage + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4)), pk);
        }
#if AVB_FFN_E2_PLAIN
        // no async proxy: fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC (it drains every outstanding memory
        // operation of the thread) and sits in front of each tensor-map store.  Read the staged block back 8 lanes per
        // row and leave with plain coalesced 16-byte stores.
        __syncwarp();
        {
          const int l8 = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3);
            uint4 a;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                         : "r"(stage + (uint32_t)(row * 128 + ((l8 ^ (row & 7)) << 4))));
            if (row0 + row < M) *reinterpret_cast<uint4*>(z + (size_t)(row0 + row) * 128 + half * 64 + cc * 32 + l8 * 4) = a;
          }
        }
#else
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tm_z, stage, half * 64 + cc * 32, row0);
          ptx::bulk_commit();
        }
#endif
      }
    };
    int g = 0, tile_it = 0, prev_mt = -1;
    for (int pr = pair0, mt = 2 * pair0 + (int)rank; pr < num_pairs; pr += pair_stride, mt += 2 * pair_stride, ++tile_it) {
      for (int c = 0; c < num_c; ++c, ++g) {
        // ---- E1: D1 -> +b1 -> ReLU -> bf16 pairs -> H
        const int hs = g & 1;
#if AVB_FFN_DECOUPLE
        // E2 of the previous tile goes FIRST when this tile's first D1 is not there yet (the issuer is waiting for the
        // LayerNorm warps, which in turn wait for the D2 buffer E2 releases)
        bool e2_early = false;
        if (c == 0 && prev_mt >= 0 && !ptx::mbar_test_warp(&d1_full[g & 1], (uint32_t)((g >> 1) & 1))) {
          e2_g = g;
          E2(tile_it - 1, prev_mt);
          e2_early = true;
        }
#endif
#if AVB_FFN_E1_X64
        // A/B build: the warp's 64 D1 columns in ONE tcgen05.ld (one TMEM round trip instead of two serial ones); the
        // bias then comes from L1 (b1 is 2 KB and hot) instead of 64 preloaded registers
        const float* bp = b1 + c * 128 + half * 64;
        AVB_TRACE(2, g, 0);
        ptx::mbar_wait(&d1_full[g & 1], (uint32_t)((g >> 1) & 1), 59);
        AVB_TRACE(2, g, 1);
        ptx::tc_fence_after();
        uint32_t pk[32];
        {
          uint32_t v[64];
          ptx::tmem_ld64(tmem_D1 + hs * 128 + half * 64 + lane_off, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + e);
            const float f0 = fmaxf(__uint_as_float(v[e * 4 + 0]) + b4.x, 0.f), f1 = fmaxf(__uint_as_float(v[e * 4 + 1]) + b4.y, 0.f);
            const float f2 = fmaxf(__uint_as_float(v[e * 4 + 2]) + b4.z, 0.f), f3 = fmaxf(__uint_as_float(v[e * 4 + 3]) + b4.w, 0.f);
            pk[e * 2 + 0] = ptx::pack_bf16(f0, f1);
            pk[e * 2 + 1] = ptx::pack_bf16(f2, f3);
          }
        }
#else
        const float* bp = b1 + c * 128 + half * 64;
        float4 bb[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          bb[e] = kResident ? *reinterpret_cast<const float4*>(&cb.b1[c * 128 + half * 64 + e * 4]) : __ldg(reinterpret_cast<const float4*>(bp) + e);
        AVB_TRACE(2, g, 0);
        ptx::mbar_wait(&d1_full[g & 1], (uint32_t)((g >> 1) & 1), 59);
        AVB_TRACE(2, g, 1);
        ptx::tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_D1 + hs * 128 + half * 64 + cc * 32 + lane_off, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 b4 = bb[cc * 8 + e];
            const float f0 = fmaxf(__uint_as_float(v[e * 4 + 0]) + b4.x, 0.f), f1 = fmaxf(__uint_as_float(v[e * 4 + 1]) + b4.y, 0.f);
            const float f2 = fmaxf(__uint_as_float(v[e * 4 + 2]) + b4.z, 0.f), f3 = fmaxf(__uint_as_float(v[e * 4 + 3]) + b4.w, 0.f);
            pk[cc * 16 + e * 2 + 0] = ptx::pack_bf16(f0, f1);
            pk[cc * 16 + e * 2 + 1] = ptx::pack_bf16(f2, f3);
          }
        }
#endif
        AVB_TRACE(2, g, 2);
        AVB_TRACE(2, g, 3);
        ptx::tmem_st32(tmem_D1 + hs * 128 + half * 64 + lane_off, pk);  // in place: the first 32 of this warp's own 64 columns
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive_cluster(&h_full[hs], 0);
        AVB_TRACE(2, g, 4);
#if AVB_FFN_DECOUPLE
        if (c == 0 && prev_mt >= 0 && !e2_early) {
#else
        if (c == 0 && prev_mt >= 0) {
#endif
          e2_g = g;
          E2(tile_it - 1, prev_mt);
          AVB_TRACE(2, g, 6);
        }
      }
      prev_mt = mt;
    }
    if (prev_mt >= 0) E2(tile_it - 1, prev_mt);
    if (lane == 0) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 6) ptx::tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// Flash-style axial attention on tensor cores.  head dim 32, 128-query x 128-key tiles.
//   qkv comes from ln_gemm_tc_kernel<2> in the operand layout [seq][q,k,v][head][pos][32 ch] (pre-swizzled):
//   a work item (s, h, qt) fetches its [128 x 32] Q tile and streams the [128 x 32] K / V tiles of its sequence,
//   each ONE contiguous 1-D bulk copy (<= 8 KB) whose bytes are the 64B-swizzled shared-memory image.  The
//   attended axis only changes the row order of the projection that wrote qkv; z is never transposed.  A ragged
//   last tile copies only its valid rows: the rest of the slot holds older (finite) rows, which the softmax masks.
//   S = Q K^T : M=128, N<=128, K=32   (A,B K-major, 64B swizzle)        -> TMEM S (fp32)
//   P = exp2(S - m) (q is pre-scaled by log2e/sqrt(32)) -> bf16 pairs in TMEM (A operand from TMEM:
//              no shared-memory round trip; measured: smem bandwidth was the limiter with P in smem)
//   O += P V : M=128, N=32, K<=128    (A = P in TMEM, B = V MN-major smem)  -> TMEM O (fp32)
//   L += P 1 : M=128, N=16             (B = a constant tile of ones)     -> TMEM L: the softmax
//              denominator comes from the tensor core, summing exactly the bf16 P the PV MMA sees
//
//   Head dim 32 makes the softmax (16 ex2/clk/SM on the MUFU pipe), not the MMA, the floor, so the
//   CTA runs TWO independent streams of work items (ping-pong): each stream owns a TMA-producer
//   warp, a single-thread MMA issuer, 4 softmax warps (one query row per thread = one TMEM lane),
//   its own smem ring and TMEM columns.  While one stream's softmax warps wait for their S tile,
//   the other stream's warps keep the MUFU/FMA pipes busy and the tensor core alternates streams.
//   The softmax reads S twice from TMEM (row max, then exp) to stay inside 168 registers, rescales
//   O lazily (only when the running max grew by > 2^8, warp-uniform vote) and defers the epilogue
//   of an item until the first softmax step of the next one has been handed to the tensor core.
//   warp 0/1: TMA producers   warp 2/3: MMA issuers (warp 2 allocates TMEM)
//   warps 4-7 / 8-11: softmax + O rescale + epilogue of stream 0 / 1
// ------------------------------------------------------------------------------------------
constexpr int kAttnStreams = 2;
constexpr int kAttnSoftmaxWarps = 8;  // per stream: 4 TMEM lane quarters x 2 column halves
constexpr int kAttnSoftmaxThreads = kAttnSoftmaxWarps * 32;
constexpr int kAttnThreads = 128 + kAttnStreams * kAttnSoftmaxThreads;  // 640
constexpr int kAttnKvStages = 5;
constexpr int kAttnTileBytes = 128 * 32 * 2;  // 8 KB: Q, K or V tile
constexpr int kAttnStreamBytes = 2 * kAttnTileBytes + kAttnKvStages * 2 * kAttnTileBytes;  // Q[2] | KV ring = 96 KB
constexpr int kAttnXchgBytes = 2 /*parity*/ * kAttnStreams * 2 /*halves*/ * 128 * 4;  // row-max exchange, 4 KB
constexpr int kAttnSmemBytes = kAttnStreams * kAttnStreamBytes + 1024 + 512 + 512 + kAttnXchgBytes;  // + align slack, barriers, ones tile
constexpr int kAttnTmemStreamCols = 256;  // S 128 (fp32) | P 64 (bf16 pairs) | O 32 | L 16 | spare
constexpr float kAttnRescaleThreshold = 8.0f;  // log2 units
constexpr int kAttnPolyEvery = 4;  // 1 of every kAttnPolyEvery exponent pairs is evaluated on the FMA pipe
#ifndef AVB_ATTN_POLY_EVERY
#define AVB_ATTN_POLY_EVERY 3
#endif
#ifndef AVB_ATTN_POLY_MASK  // which of the 16 pairs of a 32-score chunk take the polynomial (default: every 3rd, 5 of 16)
#define AVB_ATTN_POLY_MASK 0
#endif
#ifndef AVB_ATTN_PIPE_LD  // 1: the softmax warps prefetch the next S tile behind the P store (A/B build, measured r2c:
#define AVB_ATTN_PIPE_LD 0  // 15.4 vs 11.05 ms - S, P and the next S do not fit the 104 registers and the loop spills)
#endif
#ifndef AVB_ATTN_KV_POLICY  // L2 eviction priority of the K / V tile loads
#define AVB_ATTN_KV_POLICY (nqt > 1 ? ptx::kL2EvictLast : ptx::kL2EvictFirst)
#endif
constexpr float kAttnOverflowGuard = 100.0f;  // log2 units: scores beyond this on a polynomial lane flag the launch

// exp2 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = i + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.5e-5, far below the bf16 rounding of P) and i added into the exponent.
// The attention kernel is bound by the 16 ex2/clk/SM MUFU pipe, so a fixed fraction of the exponentials takes
// this path (the trick FlashAttention-4 uses on Blackwell).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.055171650f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// 32 scores (one TMEM chunk of a query row) -> 16 packed bf16 pairs of exp2(s).  No reference is subtracted (see the
// fast softmax section).  One pair in kAttnFastPolyEvery goes through the FMA-pipe polynomial (same polynomial as
// poly_exp2, evaluated on packed pairs), the others through MUFU.  pmax accumulates the maximum of the raw scores on
// the polynomial lanes: MUFU saturates to +inf when a score is out of range (the epilogue sees it in the row sum) but
// the polynomial's exponent arithmetic wraps, so those lanes are range-checked by the caller.
// kClamp = false: the caller guarantees |s| <= 100 for every score of the launch (the QKV projection records the largest
// |q_h|^2 and |k_h|^2 it writes: |s| <= |q||k|), so the polynomial lanes need neither the lower clamp nor the running
// maximum - 9 instead of 12 issue slots per pair, which is what lets a larger share of the exponentials leave the MUFU.
constexpr int kAttnFastPolyEvery = AVB_ATTN_POLY_EVERY;
#ifndef AVB_ATTN_POLY_DEG  // degree of the exp2 polynomial of the FMA-pipe lanes (3: rel. error 7.5e-5; 2: 1.7e-3, one packed FMA less).
#define AVB_ATTN_POLY_DEG 2  // P is rounded to bf16 right behind it (half an ulp = 2.0e-3) and the row sum L is taken over the same
#endif                       // rounded P, so degree 2 costs nothing that the rounding does not: r2g, same box, n-axis launch 10.26 ->
                             // 10.06 ms (8/16 share 10.08, 9/16 10.25); bench 333.8 -> 338.1 datasets/s, max|dp| 5.4e-4 -> 6.2e-4 (tol 1e-2)
#ifndef AVB_ATTN_POLY_MASK_NC  // polynomial pairs of the unclamped variant: 7 of 16.  Measured (r2d, n-axis launch, B = 64, ms):
#define AVB_ATTN_POLY_MASK_NC 0x52A5  // clamped 5/16 11.05 | unclamped 5/16 10.61, 6/16 10.15-10.18 (three patterns),
#endif                                // 7/16 {0,2,5,7,9,12,14} 9.98, 7/16 {1,3,5,8,10,12,15} 10.33-10.41, 8/16 ~10.5
template <bool kClamp = true, int kEvery = kAttnFastPolyEvery>
__device__ __forceinline__ void exp_chunk32(const uint32_t (&s)[32], uint32_t* pk, float& pmax) {
  const uint64_t magic2 = pack2(12582912.0f, 12582912.0f), nmagic2 = pack2(-12582912.0f, -12582912.0f);
  const uint64_t c3 = pack2(0.055171650f, 0.055171650f), c2 = pack2(0.24261113f, 0.24261113f);
  const uint64_t c1 = pack2(0.69326097f, 0.69326097f), c0 = pack2(0.99992806f, 0.99992806f);
  const uint64_t neg1 = pack2(-1.0f, -1.0f);
#if AVB_ATTN_POLY_DEG == 2  // degree-2 minimax of 2^f on [-0.5, 0.5]: relative error 1.7e-3 (half a bf16 ulp is 2.0e-3)
  const uint64_t q2 = pack2(0.23842894f, 0.23842894f), q1 = pack2(0.70344801f, 0.70344801f), q0 = pack2(1.00044314f, 1.00044314f);
#endif
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float a = __uint_as_float(s[2 * e]), b = __uint_as_float(s[2 * e + 1]);
    constexpr unsigned kMask = kClamp ? (unsigned)(AVB_ATTN_POLY_MASK) : (unsigned)(AVB_ATTN_POLY_MASK_NC);
    if (kMask ? ((kMask >> e) & 1) : ((e % kEvery) == kEvery - 1)) {
      if (kClamp) pmax = max3(pmax, a, b);
      const uint64_t x2 = kClamp ? pack2(fmaxf(a, -126.0f), fmaxf(b, -126.0f)) : pack2(a, b);
      const uint64_t t2 = add2(x2, magic2);           // integer part lands in the low mantissa bits
      const uint64_t f2 = fma2(add2(t2, nmagic2), neg1, x2);  // f = x - round(x), in [-0.5, 0.5]
#if AVB_ATTN_POLY_DEG == 2
      uint64_t p2 = fma2(f2, q2, q1);
      p2 = fma2(p2, f2, q0);
#else
      uint64_t p2 = fma2(f2, c3, c2);
      p2 = fma2(p2, f2, c1);
      p2 = fma2(p2, f2, c0);
#endif
      float p0, p1, t0, t1;
      unpack2(p2, p0, p1);
      unpack2(t2, t0, t1);
      p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
      p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
      pk[e] = ptx::pack_bf16(p0, p1);
    } else {
#ifdef AVB_ATTN_NOEXP  // timing experiment only: everything but the exponentials
      pk[e] = ptx::pack_bf16(a, b);
#else
      pk[e] = ptx::pack_bf16(fast_exp2(a), fast_exp2(b));
#endif
    }
  }
}

template <bool kFast>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int n, int d, int H,
                    int axis, int* __restrict__ guard, const float* __restrict__ norms, int qsplit) {
  // norms (fast variant, may be null): {max |q_h|^2, max |k_h|^2} over every (row, head) of this launch, recorded by the
  // QKV projection that wrote qkv.  |score| <= |q||k| <= 100 licenses the unclamped polynomial lanes (exp_chunk32).
  // guard: one int per launch pair.  The fast variant (stale row maximum, see the softmax section) raises it when a
  // score tile exceeds the running maximum by more than kAttnOverflowGuard; the exact variant, launched right after
  // it on the same stream, returns immediately unless the guard is up and otherwise recomputes the whole launch.
  if (!kFast && guard != nullptr && *guard == 0) return;
  constexpr int NS = kAttnKvStages;
  constexpr int kBarsPerStream = 2 + 2 + NS + NS + 1 + 1 + 1 + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + kAttnStreams * kAttnStreamBytes);
  uint64_t* x_full_all = bars_all + kAttnStreams * kBarsPerStream;  // [2][4]: row-max exchange of (stream, row quarter)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full_all + 8);
  uint32_t* s_ones = reinterpret_cast<uint32_t*>(smem + kAttnStreams * kAttnStreamBytes + 512);  // 512 B of bf16 1.0
  float* s_xchg = reinterpret_cast<float*>(smem + kAttnStreams * kAttnStreamBytes + 1024);      // [2][streams][2][128]

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the
  // MMA / TMA operands in uniform registers (no per-instruction R2UR chains on the single issuing thread)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  long long* trace = (blockIdx.x == 0 && (warp == 0 || warp == 2 || warp == 4 || warp == 12) && lane == 0) ? g_attn_trace : nullptr;
  [[maybe_unused]] const int trole = warp == 12 ? 3 : 2;  // trace builds only
  const int T = axis == 0 ? d : n;
  const int S = axis == 0 ? B * n : B * d;  // sequences; the host guarantees S * H * nqt < 2^31
  const int nqt = (T + 127) / 128;  // query tiles == kv tiles per sequence

  if (warp == 1 && lane == 0) {
    for (int sid = 0; sid < kAttnStreams; ++sid) {
      uint64_t* b = bars_all + sid * kBarsPerStream;
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&b[i], 1); ptx::mbar_init(&b[2 + i], 1); }  // q_full (expect_tx), q_empty
      for (int i = 0; i < NS; ++i) { ptx::mbar_init(&b[4 + i], 1); ptx::mbar_init(&b[4 + NS + i], 1); }  // kv_full, kv_empty
      ptx::mbar_init(&b[4 + 2 * NS + 0], 1);                    // s_full
      ptx::mbar_init(&b[4 + 2 * NS + 1], kAttnSoftmaxThreads);  // s_free
      ptx::mbar_init(&b[4 + 2 * NS + 2], kAttnSoftmaxThreads);  // p_full
      ptx::mbar_init(&b[4 + 2 * NS + 3], 1);                    // pv_done
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&x_full_all[i], 64);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x >= 128 && threadIdx.x < 256) s_ones[threadIdx.x - 128] = 0x3F803F80u;
  // the Q / K / V slots start as zeros: a ragged tile leaves the tail of its slot untouched, and whatever the tensor
  // core reads there must be finite (0 x NaN would poison O through the masked, zero-probability keys)
  for (int i = threadIdx.x; i < kAttnStreams * kAttnStreamBytes / 16; i += kAttnThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- per-stream view
  const int sid = warp < 2 ? warp : (warp < 4 ? warp - 2 : (warp - 4) / kAttnSoftmaxWarps);
  uint8_t* sQ = smem + sid * kAttnStreamBytes;          // [2][8K]
  uint8_t* sKV = sQ + 2 * kAttnTileBytes;               // [NS][K 8K | V 8K]
  uint64_t* bars = bars_all + sid * kBarsPerStream;
  uint64_t* q_full = bars;        // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 4 + NS;
  uint64_t* s_full = bars + 4 + 2 * NS;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_done = s_full + 3;
  const uint32_t tmem_S = tmem_base + sid * kAttnTmemStreamCols;
  const uint32_t tmem_P = tmem_S + 128;  // 64 columns: P as packed bf16 pairs, A operand of the PV / L MMAs
  const uint32_t tmem_O = tmem_S + 192;  // O 32 | L 16 columns
  // Work items of a stream.  A stream owns whole (sequence, head) pairs - pair = pair0 + i * pair_stride - and walks
  // the query tiles of a pair back to back, so the pair's K / V (re-read by every query tile) stays in L2 between its
  // own passes instead of depending on how well eight different streams keep time (measured: 2.2x the algorithmic
  // DRAM reads with the query tiles spread over streams).  Coordinates are carried incrementally: a division by a
  // run-time value costs a long dependent MUFU.RCP / I2F chain on the softmax warps' critical path per item.
  // qsplit (a divisor of nqt, chosen by the host) cuts every pair's query tiles into qsplit contiguous parts that different
  // streams own: with few pairs per stream (one large dataset sharded over 8 GPUs: 1000 pairs for 296 streams) whole pairs
  // leave the last round 38 % full; the parts of a pair run at the same time and share its K / V through L2.
  const int num_pairs = S * H * qsplit;                 // (sequence, head, part) triples
  const int qn = nqt / qsplit;                          // query tiles per part
  const int pair0 = (int)blockIdx.x * kAttnStreams + sid;
  const int pair_stride = (int)gridDim.x * kAttnStreams;
  const int my_items = pair0 < num_pairs ? ((num_pairs - pair0 + pair_stride - 1) / pair_stride) * qn : 0;
  struct ItemPos { int s, h, qt, v, left; };
  // coordinates of triple v: two divisions per (sequence, head, part) - amortised over qn * nqt softmax steps (per query
  // tile they were ~1000 dependent cycles on the softmax warps' critical path)
  auto locate = [&](ItemPos& p) {
    const int pair = p.v / qsplit, part = p.v - pair * qsplit;
    p.s = pair / H; p.h = pair - p.s * H; p.qt = part * qn; p.left = qn;
  };
  auto first_item = [&]() { ItemPos p; p.v = pair0; locate(p); return p; };
  auto advance = [&](ItemPos& p) {  // next query tile of the part, or the first tile of the stream's next triple
    p.qt += 1;
    if (--p.left == 0) { p.v += pair_stride; locate(p); }
  };

  // Register budget per role (warpgroup-wide, setmaxnreg): 640 threads leave 96 registers each, and the softmax loop
  // (64 scores + 32 packed results + its loop state) spilled a handful of them - ncu attributed ~20 % of the d-axis
  // launch's stall samples to the reloads.  The four role warps give registers up, the softmax warpgroups take them.
#ifndef AVB_ATTN_NO_SETMAXNREG
  // (the pool only holds what this CTA released: 128 x (96 - 64) == 512 x (104 - 96))
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
#endif
  if (warp < 2) {
    // ======================= producer warp of stream `sid`: one bulk copy per [128 x 32] bf16 tile
    // (Measured on B200 before the operand layout existed: a tensor-map gather costs the TMA unit ~3.6 cycles per box
    //  row whatever its width, a warp of 16-byte cp.async ~1600 cycles of issue per K/V stage and 3x the DRAM sectors.)
    const size_t head_elems = (size_t)T * 32;  // one (sequence, q|k|v, head) block
    uint32_t g = 0;
    int it = 0;
    ItemPos ip = first_item();
    for (int item = 0; item < my_items; ++item, ++it, advance(ip)) {
      const int s = ip.s, h = ip.h, qt = ip.qt;
      const __nv_bfloat16* qsrc = qkv + ((size_t)s * 3 * H + h) * head_elems;
      const __nv_bfloat16* ksrc = qsrc + (size_t)H * head_elems;
      const __nv_bfloat16* vsrc = ksrc + (size_t)H * head_elems;
      const int qb = it & 1;
      ptx::mbar_wait(&q_empty[qb], (uint32_t)((it >> 1) & 1) ^ 1, 10, 200u);
      if (ptx::elect_one()) {
        const uint32_t bytes = (uint32_t)min(128, T - qt * 128) * 64u;
        ptx::mbar_expect_tx(&q_full[qb], bytes);
        ptx::bulk_load_hint(sQ + qb * kAttnTileBytes, qsrc + (size_t)qt * 128 * 32, bytes, &q_full[qb], ptx::kL2EvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < nqt; ++j, ++g) {
        const int st = (int)(g % NS);
        AVB_TRACE(0, g, 0);
        ptx::mbar_wait(&kv_empty[st], (uint32_t)((g / NS) & 1) ^ 1, 11, 200u);
        AVB_TRACE(0, g, 1);
        if (ptx::elect_one()) {
          uint8_t* sk = sKV + st * 2 * kAttnTileBytes;
          const uint32_t bytes = (uint32_t)min(128, T - j * 128) * 64u;
          ptx::mbar_expect_tx(&kv_full[st], 2 * bytes);
          // every query tile of the (sequence, head) streams the same K / V: keep them in L2 (Q is read once)
          const uint64_t pol = AVB_ATTN_KV_POLICY;
          ptx::bulk_load_hint(sk, ksrc + (size_t)j * 128 * 32, bytes, &kv_full[st], pol);
          ptx::bulk_load_hint(sk + kAttnTileBytes, vsrc + (size_t)j * 128 * 32, bytes, &kv_full[st], pol);
        }
        __syncwarp();
        AVB_TRACE(0, g, 2);
      }
    }
  } else if (warp < 4) {
    // ======================= MMA issuer of stream `sid`
    // One thread issues every tcgen05.mma of the stream, so per-MMA scalar work is on the critical path:
    // all shared-memory descriptors are built once and advanced by adding (byte offset >> 4) to them.
    if (my_items > 0) {  // all 32 lanes run the control flow; one elected lane issues the tcgen05 instructions
      // number of key columns the MMAs of kv tile j touch (multiple of 16)
      auto kcols = [&](int j) { int v = T - j * 128; v = v > 128 ? 128 : v; return (v + 15) & ~15; };
      const uint64_t dQ = ptx::make_smem_desc(ptx::smem_u32(sQ), 16, 512, ptx::kSwz64);
      const uint64_t dK = ptx::make_smem_desc(ptx::smem_u32(sKV), 16, 512, ptx::kSwz64);
      const uint64_t dV = ptx::make_smem_desc(ptx::smem_u32(sKV) + kAttnTileBytes, 16, 512, ptx::kSwz64);
      // ones tile: K-major, no swizzle, 2 x 2 core matrices of 128 B (every element is 1.0, so any layout works)
      const uint64_t d_ones = ptx::make_smem_desc(ptx::smem_u32(s_ones), 128, 256, ptx::kSwzNone);
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
      constexpr uint32_t idesc_l = ptx::make_idesc_bf16(128, 16, 0, 0);
      constexpr uint64_t kStageStep = (2 * kAttnTileBytes) >> 4;
      auto issue_S = [&](long long g, int it, int j) {
        if (j == 0) { ptx::mbar_wait(&q_full[it & 1], (uint32_t)((it >> 1) & 1), 20, 40u); }
        const int st = (int)(g % NS);
        ptx::mbar_wait(&kv_full[st], (uint32_t)((g / NS) & 1), 21, 40u);
        ptx::tc_fence_after();  // bulk copies and the tensor core both work through the async proxy: no proxy fence
        const int kc = kcols(j);
        const uint32_t idesc = kc == 128 ? idesc_s : ptx::make_idesc_bf16(128, kc, 0, 0);
        const uint64_t dk = dK + (uint64_t)st * kStageStep;
        const uint64_t dq = dQ + (uint64_t)((it & 1) * (kAttnTileBytes >> 4));
        if (ptx::elect_one()) {
          ptx::umma_bf16(tmem_S, dq, dk, idesc, 0);
          ptx::umma_bf16(tmem_S, dq + 2, dk + 2, idesc, 1);  // +32 bytes: second K=16 slice of the 64-byte rows
          ptx::umma_commit(s_full);
          if (j == nqt - 1) ptx::umma_commit(&q_empty[it & 1]);  // last S of the item: this Q buffer may be overwritten
        }
        __syncwarp();
      };
      long long g = 0;
      int it = 0;
      int item = 0;
      issue_S(0, 0, 0);
      while (item < my_items) {
        for (int j = 0; j < nqt; ++j, ++g) {
          // S of the next step goes to the tensor core as soon as this step's softmax has read S
          const bool more = (j + 1 < nqt) || (item + 1 < my_items);
          AVB_TRACE(1, g, 0);
          if (more) {
            ptx::mbar_wait(s_free, (uint32_t)(g & 1), 24, 40u);
            AVB_TRACE(1, g, 1);
            ptx::tc_fence_after();
            if (j + 1 < nqt) issue_S(g + 1, it, j + 1);
            else issue_S(g + 1, it + 1, 0);
          }
          AVB_TRACE(1, g, 2);
          ptx::mbar_wait(p_full, (uint32_t)(g & 1), 23, 40u);
          AVB_TRACE(1, g, 3);
          ptx::tc_fence_after();
          const int st = (int)(g % NS);
          const uint64_t dv = dV + (uint64_t)st * kStageStep;
          const int ksteps = kcols(j) >> 4;
          // P: 8 TMEM columns (16 bf16) per K=16 slice; V: [keys x 32], 16 keys (1 KB) per slice
          if (ptx::elect_one()) {
            if (ksteps == 8) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                ptx::umma_bf16_ts(tmem_O, tmem_P + k * 8, dv + (uint64_t)(k * 64), idesc_pv, (j | k) != 0);
                ptx::umma_bf16_ts(tmem_O + 32, tmem_P + k * 8, d_ones, idesc_l, (j | k) != 0);
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                ptx::umma_bf16_ts(tmem_O, tmem_P + k * 8, dv + (uint64_t)(k * 64), idesc_pv, (j | k) != 0);
                ptx::umma_bf16_ts(tmem_O + 32, tmem_P + k * 8, d_ones, idesc_l, (j | k) != 0);
              }
            }
            ptx::umma_commit(&kv_empty[st]);
            ptx::umma_commit(pv_done);
          }
          __syncwarp();
          AVB_TRACE(1, g, 4);
        }
        ++item;
        ++it;
      }
    }
  } else {
    // ======================= softmax / rescale / epilogue warps of stream `sid`
    // 8 warps: warp (q, half) owns TMEM lanes 32q..32q+31 (query rows) and key columns [64*half, 64*half+64).
    const int w8 = (warp - 4) % kAttnSoftmaxWarps;
    const int q = w8 & 3, half = w8 >> 2;
    const int r = q * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int bar_id = 1 + sid * 4 + q;  // named barrier shared by the two warps of a row quarter
    const uint32_t tS = tmem_S + half * 64 + lane_off;   // this warp's 64 S columns
    const uint32_t tP = tmem_P + half * 32 + lane_off;   // ... and the 32 packed P columns they become
    const uint32_t tO = tmem_O + lane_off;

    long long g = 0;
    int it = 0;
    // The epilogue of an item (O / L -> bf16 -> out) is deferred into the first step of the next item, after the
    // exponentials: by then its last PV MMA has long completed (no exposed MMA latency), and since the next PV is only
    // issued after every softmax thread has arrived on p_full, O needs no second buffer.
    // out is the out-projection's A-operand layout (outproj_tc_kernel): row m = s*T + position goes to
    // out[m >> 7][head][m & 127][32 ch] with the 16-byte chunks XOR-swizzled by (m >> 1) & 3.
    bool pend = false;
    int pend_m0 = 0, pend_h = 0, pend_t0 = 0;  // sequence-major index of tile row 0, head, first position of the tile
    auto epilogue = [&]() {  // each half writes 16 of the 32 output channels of its rows
      uint32_t o[16];
      ptx::tmem_ld16(tO + half * 16, o);
      const float lsum = __uint_as_float(ptx::tmem_ld1(tO + 32));
      ptx::tmem_wait_ld();
      if (pend_t0 + r < T) {
        if (kFast && !(lsum < 1.2676506e30f && lsum > 7.8886091e-31f)) *guard = 1;  // outside (2^-100, 2^100), inf or NaN
        const unsigned m = (unsigned)(pend_m0 + r);
        const float inv = 1.0f / lsum;
        __nv_bfloat16* orow = out + (((size_t)(m >> 7) * H + pend_h) * 128 + (m & 127)) * 32;
        const unsigned swz = (m >> 1) & 3;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          uint4 pk;
          pk.x = ptx::pack_bf16(__uint_as_float(o[e * 8 + 0]) * inv, __uint_as_float(o[e * 8 + 1]) * inv);
          pk.y = ptx::pack_bf16(__uint_as_float(o[e * 8 + 2]) * inv, __uint_as_float(o[e * 8 + 3]) * inv);
          pk.z = ptx::pack_bf16(__uint_as_float(o[e * 8 + 4]) * inv, __uint_as_float(o[e * 8 + 5]) * inv);
          pk.w = ptx::pack_bf16(__uint_as_float(o[e * 8 + 6]) * inv, __uint_as_float(o[e * 8 + 7]) * inv);
#ifndef AVB_ATTN_NOSTORE  // timing experiment: epilogue without its global stores
          *reinterpret_cast<uint4*>(orow + (((unsigned)(half * 2 + e) ^ swz) << 3)) = pk;
#else
          if (pk.x == 0x12345678u) *reinterpret_cast<uint4*>(orow + (((unsigned)(half * 2 + e) ^ swz) << 3)) = pk;
#endif
        }
      }
    };
    if constexpr (kFast) {
      // ---- single-pass softmax WITHOUT a running maximum.
      // softmax is invariant to a per-row shift of the scores; implementations subtract the running row maximum only
      // to keep exp2 in range.  bf16 P and the fp32 accumulators have the full 8-bit exponent, so for scores within
      // +-100 (log2 units, i.e. +-69 nats - far beyond what LayerNorm-ed activations produce) exp2(s) itself is in
      // range and the normalisation O / L cancels the missing shift exactly, with no loss of precision.  That removes
      // everything the exact variant below pays for: the second read of S, the row-max reduction, the shared-memory
      // exchange + barrier between the two warps of a row quarter on every tile, the subtraction, and the O rescales -
      // a dependent chain as long as the exponentials themselves.  S is read once and handed back to the tensor core
      // at once.  Out-of-range scores cannot go unnoticed: on the MUFU lanes they give P = inf or a row sum of 0, which
      // the epilogue checks (2^-100 < L < 2^100); the polynomial lanes (whose exponent arithmetic would wrap) are
      // range-checked on the raw scores.  Either raises `guard`, and the exact variant recomputes the launch.
      const uint32_t a_sfull = ptx::smem_u32(s_full), a_sfree = ptx::smem_u32(s_free);
      const uint32_t a_pfull = ptx::smem_u32(p_full), a_pvdone = ptx::smem_u32(pv_done);
      uint32_t gp = 0;  // parity of the current step (g & 1)
      float pmax = -INFINITY;
      const bool bounded = norms != nullptr && norms[0] * norms[1] <= 1.0e4f;  // false also for NaN
      ItemPos ip = first_item();
      // -DAVB_ATTN_PIPE_LD=1 (A/B build, slower - see the macro): the S tile of the NEXT step is requested from TMEM right
      // behind this step's P store, so that its load latency overlaps the store's completion and the p_full hand-off.
      uint32_t sa[32], sb[32];
      bool have_s = false;
      for (int item = 0; item < my_items; ++item, ++it) {
        const int s = ip.s, h = ip.h, qt = ip.qt;
        for (int j = 0; j < nqt; ++j, ++g, gp ^= 1u) {
          AVB_TRACE(trole, g, 0);
          if (!have_s) {
            ptx::mbar_wait_s(a_sfull, gp);
            AVB_TRACE(trole, g, 1);
            ptx::tc_fence_after();
            ptx::tmem_ld32(tS, sa);
            ptx::tmem_ld32(tS + 32, sb);
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive_s(a_sfree);  // the only read of S: the tensor core may start the next S right away
          }
          AVB_TRACE(trole, g, 2);
          const int nv = T - j * 128 - half * 64;  // valid key columns in this warp's half (may be <= 0)
          // ---- P = exp2(S) -> packed bf16 pairs
          uint32_t pk[32];
          if (nv < 64 || !bounded) {  // ragged tile (masked keys are -inf) or no score bound: clamped polynomial lanes
            if (nv < 64) {  // keys beyond the sequence end: score -inf, probability 0
              if (nv < 32) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                  if (e >= nv) sa[e] = 0xff800000u;
              }
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (32 + e >= nv) sb[e] = 0xff800000u;
            }
            exp_chunk32<true>(sa, pk, pmax);
            exp_chunk32<true>(sb, pk + 16, pmax);
          } else {
            exp_chunk32<false>(sa, pk, pmax);
            exp_chunk32<false>(sb, pk + 16, pmax);
          }
          AVB_TRACE(trole, g, 4);
          if (g >= 1) {  // P is free and O complete once the PV / L MMAs of the previous step have finished
            ptx::mbar_wait_s(a_pvdone, gp ^ 1u);
            ptx::tc_fence_after();
          }
          if (j == 0 && pend) { epilogue(); pend = false; }
          AVB_TRACE(trole, g, 5);
          ptx::tmem_st32(tP, pk);
#if AVB_ATTN_PIPE_LD
          // non-blocking test (a phase may complete between two lanes' tests, so the warp takes lane 0's answer); if the
          // tile is not there yet the next step fetches it the ordinary way
          have_s = ((j + 1 < nqt) || (item + 1 < my_items)) && ptx::mbar_test_s(a_sfull, gp ^ 1u);
          have_s = __shfl_sync(0xffffffffu, (int)have_s, 0) != 0;
          if (have_s) {
            ptx::tc_fence_after();
            ptx::tmem_ld32(tS, sa);
            ptx::tmem_ld32(tS + 32, sb);
          }
#endif
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          ptx::mbar_arrive_s(a_pfull);
#if AVB_ATTN_PIPE_LD
          if (have_s) {
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive_s(a_sfree);
          }
#endif
          AVB_TRACE(trole, g, 7);
        }
        pend = true;
        pend_h = h;
        pend_t0 = qt * 128;
        pend_m0 = s * T + qt * 128;
        advance(ip);
      }
      if (pmax > kAttnOverflowGuard) *guard = 1;  // a polynomial lane left the safe exponent range
    } else {
    ItemPos ip = first_item();
    for (int item = 0; item < my_items; ++item, ++it) {
      const int s = ip.s, h = ip.h, qt = ip.qt;
      float m = -INFINITY;
      for (int j = 0; j < nqt; ++j, ++g) {
        AVB_TRACE(trole, g, 0);
        ptx::mbar_wait(s_full, (uint32_t)(g & 1), 30);
        AVB_TRACE(trole, g, 1);
        ptx::tc_fence_after();
        const int nv = T - j * 128 - half * 64;  // valid key columns in this warp's half (may be <= 0)
        // ---- pass 1: row maximum over this half (S is re-read in pass 2: TMEM reads are cheap, registers are not:
        //      640 threads leave 96 per thread), then exchange with the partner warp
        const bool full_half = nv >= 64;
        float mx;
        if (full_half) {
          uint32_t sa[32], sb[32];
          ptx::tmem_ld32(tS, sa);
          ptx::tmem_ld32(tS + 32, sb);
          ptx::tmem_wait_ld();
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[e]), __uint_as_float(sa[e + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sb[e]), __uint_as_float(sb[e + 1])));
          }
          mx = fmaxf(mx0, mx1);
        } else {
          mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < nv) {
              uint32_t sv[32];
              ptx::tmem_ld32(tS + c * 32, sv);
              ptx::tmem_wait_ld();
              const int rem = nv - c * 32;
#pragma unroll
              for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (e < rem) ? __uint_as_float(sv[e]) : -INFINITY);
            }
          }
        }
        AVB_TRACE(trole, g, 2);
        float* xc = s_xchg + (((g & 1) * kAttnStreams + sid) * 2) * 128;
        xc[half * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        mx = fmaxf(mx, xc[(half ^ 1) * 128 + r]);
        // ---- lazy rescale: keep the stale reference max unless it grew by more than 2^8 (warp-uniform vote;
        //      both warps of a row quarter see the same values and take the same decision)
        const bool need = mx > m + kAttnRescaleThreshold;  // true on the first tile (m = -inf)
        const bool any_need = __any_sync(0xffffffffu, need);
        AVB_TRACE(trole, g, 3);
        if (g >= 1) {  // P is free and O complete once the PV / L MMAs of the previous step have finished
          ptx::mbar_wait(pv_done, (uint32_t)((g - 1) & 1), 31);
          ptx::tc_fence_after();
        }
        if (j == 0 && pend) { epilogue(); pend = false; }
        AVB_TRACE(trole, g, 4);
        if (any_need) {
          const float m_new = need ? mx : m;
          if (j > 0 && half == 0) {  // half 0 rescales O and the denominator column in TMEM
            const float alpha = need ? fast_exp2(m - m_new) : 1.0f;
            uint32_t o[32];
            ptx::tmem_ld32(tO, o);
            uint32_t lv = ptx::tmem_ld1(tO + 32);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            lv = __float_as_uint(__uint_as_float(lv) * alpha);
            ptx::tmem_st32(tO, o);
            ptx::tmem_st1(tO + 32, lv);
            ptx::tmem_wait_st();
          }
          m = m_new;
        }
        // ---- pass 2: P = exp2(S - m) -> packed bf16 pairs -> TMEM (the row sum is done by the L MMA)
        {
          uint32_t pk[32];
          if (full_half) {
            uint32_t sa[32], sb[32];
            ptx::tmem_ld32(tS, sa);
            ptx::tmem_ld32(tS + 32, sb);
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive(s_free);  // last read of S: hand the buffer back to the tensor core
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if ((e % kAttnPolyEvery) == kAttnPolyEvery - 1) {  // this pair goes through the FMA-pipe polynomial
                pk[e] = ptx::pack_bf16(poly_exp2(__uint_as_float(sa[2 * e]) - m), poly_exp2(__uint_as_float(sa[2 * e + 1]) - m));
                pk[16 + e] = ptx::pack_bf16(poly_exp2(__uint_as_float(sb[2 * e]) - m), poly_exp2(__uint_as_float(sb[2 * e + 1]) - m));
              } else {
                pk[e] = ptx::pack_bf16(fast_exp2(__uint_as_float(sa[2 * e]) - m), fast_exp2(__uint_as_float(sa[2 * e + 1]) - m));
                pk[16 + e] = ptx::pack_bf16(fast_exp2(__uint_as_float(sb[2 * e]) - m), fast_exp2(__uint_as_float(sb[2 * e + 1]) - m));
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {  // static indices only: pk[] must stay in registers
              if (c * 32 < nv) {
                uint32_t sv[32];
                ptx::tmem_ld32(tS + c * 32, sv);
                ptx::tmem_wait_ld();
                const int rem = nv - c * 32;  // keys beyond the sequence end get probability 0
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float p0 = (2 * e < rem) ? fast_exp2(__uint_as_float(sv[2 * e]) - m) : 0.f;
                  const float p1 = (2 * e + 1 < rem) ? fast_exp2(__uint_as_float(sv[2 * e + 1]) - m) : 0.f;
                  pk[c * 16 + e] = ptx::pack_bf16(p0, p1);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) pk[c * 16 + e] = 0u;
              }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(s_free);
          }
          AVB_TRACE(trole, g, 7);
          ptx::tmem_st32(tP, pk);
          ptx::tmem_wait_st();
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_full);
      }
      pend = true;
      pend_h = h;
      pend_t0 = qt * 128;
      pend_m0 = s * T + qt * 128;
      advance(ip);
    }
    }
    if (pend) {
      ptx::mbar_wait(pv_done, (uint32_t)((g - 1) & 1), 33);
      ptx::tc_fence_after();
      epilogue();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// Test helper: the score bound {max |q_h|^2, max |k_h|^2} of a qkv buffer in the attention operand layout
// [seq][q,k,v][head][pos][32 ch] (what ln_gemm2_tc_kernel records in the product path); rows = S * T * H (pos, head) rows
// per operand.  The chunk swizzle permutes within a 64-byte row, which a sum of squares does not see.
__global__ void qk_norms_kernel(const __nv_bfloat16* __restrict__ qkv, size_t rows, int H, int T, float* __restrict__ norms) {
  float mq = 0.f, mk = 0.f;
  const size_t per_seq = (size_t)H * T;  // (head, pos) rows of one operand of one sequence
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
    const size_t sq = i / per_seq, rem = i - sq * per_seq;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const uint4* p = reinterpret_cast<const uint4*>(qkv + ((sq * 3 + w) * per_seq + rem) * 32);
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 v = p[c];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float lo = __uint_as_float(u[t] << 16), hi = __uint_as_float(u[t] & 0xffff0000u);
          ss = fmaf(lo, lo, fmaf(hi, hi, ss));
        }
      }
      if (w == 0) mq = fmaxf(mq, ss); else mk = fmaxf(mk, ss);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
    mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<int*>(norms), __float_as_int(mq));
    atomicMax(reinterpret_cast<int*>(norms) + 1, __float_as_int(mk));
  }
}

// Test helper: token-major qkv[tok][3*H*32] -> the attention operand layout of ln_gemm_tc_kernel<2> (one thread per
// 16-byte chunk).  The product path never runs this: its QKV projection writes the operand layout directly.
__global__ void repack_qkv_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int n, int d,
                                  int H, int axis) {
  const size_t chunks = (size_t)B * n * d * 3 * H * 4;
  const int T = axis == 0 ? d : n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx & 3);
    size_t r = idx >> 2;
    const int h = (int)(r % H); r /= H;
    const int w = (int)(r % 3); r /= 3;
    const size_t tok = r;
    const int j = (int)(tok % d), i = (int)((tok / d) % n), b = (int)(tok / ((size_t)n * d));
    const size_t seq = axis == 0 ? (size_t)b * n + i : (size_t)b * d + j;
    const int pos = axis == 0 ? j : i;
    const uint4 v = *reinterpret_cast<const uint4*>(in + tok * (size_t)(3 * H * 32) + (size_t)(w * H + h) * 32 + c4 * 8);
    *reinterpret_cast<uint4*>(out + (((seq * 3 + w) * H + h) * T + pos) * 32 + ((c4 ^ ((pos >> 1) & 3)) << 3)) = v;
  }
}

// Test helper: attention output in the out-projection's A-operand layout -> token-major out[tok][H*32].
__global__ void unpack_attn_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int n, int d,
                                   int H, int axis) {
  const size_t chunks = (size_t)B * n * d * H * 4;
  const RowMap rm{n, d, axis};
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (size_t)gridDim.x * blockDim.x) {
    const unsigned c4 = (unsigned)(idx & 3);
    size_t r = idx >> 2;
    const int h = (int)(r % H);
    const unsigned m = (unsigned)(r / H);
    const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)(m >> 7) * H + h) * 128 + (m & 127)) * 32 + ((c4 ^ ((m >> 1) & 3)) << 3));
    *reinterpret_cast<uint4*>(out + (size_t)rm.token((int)m) * (H * 32) + h * 32 + c4 * 8) = v;
  }
}

// Test helpers (inverse directions of the two above).
__global__ void unpack_qkv_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int n, int d,
                                  int H, int axis) {
  const size_t chunks = (size_t)B * n * d * 3 * H * 4;
  const int T = axis == 0 ? d : n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx & 3);
    size_t r = idx >> 2;
    const int h = (int)(r % H); r /= H;
    const int w = (int)(r % 3); r /= 3;
    const size_t tok = r;
    const int j = (int)(tok % d), i = (int)((tok / d) % n), b = (int)(tok / ((size_t)n * d));
    const size_t seq = axis == 0 ? (size_t)b * n + i : (size_t)b * d + j;
    const int pos = axis == 0 ? j : i;
    const uint4 v = *reinterpret_cast<const uint4*>(in + (((seq * 3 + w) * H + h) * T + pos) * 32 + ((c4 ^ ((pos >> 1) & 3)) << 3));
    *reinterpret_cast<uint4*>(out + tok * (size_t)(3 * H * 32) + (size_t)(w * H + h) * 32 + c4 * 8) = v;
  }
}
__global__ void pack_attn_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int n, int d,
                                 int H, int axis) {
  const size_t chunks = (size_t)B * n * d * H * 4;
  const RowMap rm{n, d, axis};
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (size_t)gridDim.x * blockDim.x) {
    const unsigned c4 = (unsigned)(idx & 3);
    size_t r = idx >> 2;
    const int h = (int)(r % H);
    const unsigned m = (unsigned)(r / H);
    const uint4 v = *reinterpret_cast<const uint4*>(in + (size_t)rm.token((int)m) * (H * 32) + h * 32 + c4 * 8);
    *reinterpret_cast<uint4*>(out + (((size_t)(m >> 7) * H + h) * 128 + (m & 127)) * 32 + ((c4 ^ ((m >> 1) & 3)) << 3)) = v;
  }
}

}  // namespace avb
