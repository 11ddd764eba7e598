// Fused attention sub-layer for blocks that attend over the variable axis d (even blocks, model.py:56-65) when a
// whole sequence fits one 128-row tile (T = d <= 128):
//
//     delta[tokens of (b, i)] = MHA(LN_q(z), LN_k(z), LN_v(z)) . Wo + bo        (bf16 rows, added to z by the FFN kernel)
//
// One CTA owns a whole sequence (b, i): its T tokens are T consecutive 512-byte rows of z.  Unfused, this sub-layer
// writes and re-reads qkv (1536 B/token each way) and the attention output (512 B/token) through HBM - three
// launches that each sit at 70-84 % of the HBM peak (profiles/r1g-r1j).  Here neither exists in HBM: per token the
// kernel reads 512 B of z and writes 256 B of delta.
//
// A sequence is processed head by head (a "unit" u = sequence * 8 + head):
//   LN      : 4 warps normalise the T fp32 rows into a bf16 [128 x 128] A tile (128B-swizzled, double-buffered by
//             sequence; rows >= T are zero)
//   G_u     : [Q_h | K_h | V_h] = xhat . Wqkv_h^T          M=128, N=96, K=128  -> TMEM (fp32)
//   conv_u  : + bias -> bf16 -> Q / K / V tiles in shared memory (the 64B-swizzled images the attention MMAs read)
//   S_u     : S = Q K^T                                    M=128, N=kc, K=32   -> TMEM          (kc = T rounded up to 16)
//   softmax : P = exp2(S) without a running maximum, bf16 pairs -> TMEM (same scheme and guard as attention_tc_kernel)
//   PV_u    : O = P V, L = P 1                             M=128, N=32|16, K=kc -> TMEM
//   Oepi_u  : O / L -> bf16 -> TMEM A tile [128 x 32] (tcgen05.st; shared memory / HBM when the out-projection is not fused)
//   OP_u    : DELTA += O_h . Wo_h^T                        M=128, N=128, K=32  -> TMEM, accumulated over the 8 heads
//   Depi    : DELTA + bo' -> bf16 -> 256-byte rows of delta (token order), staged in the finished xhat tile, TMA store
// kFuseOut = false stops after Oepi and writes the attention output in outproj_tc_kernel's A-operand layout instead
// (the out-projection then runs as its own launch); kept as the staging step of this kernel and as its A/B reference.
//
// Weights: the per-head operand images are packed once at load time (avb_load_weights): Wqkv_h as two K blocks of
// [96 rows x 64] bf16 in the 128B-swizzle image (12 KB each), Wo_h as [128 rows x 32] in the 64B-swizzle image
// (8 KB).  Wo (64 KB) stays resident in shared memory; the Wqkv K blocks stream from L2 through a ring of 1-D bulk
// copies (192 KB per sequence; the L2 -> SM path, not HBM, carries them).
//
// TMEM (512 columns): SBUF0 | SBUF1 (128 each: the QKV accumulator of a unit, then its S tile) | P 64 | O 32 | L 16 |
// bf16 O tile 16 (A operand of OP) | DELTA 128.  Software pipeline, while the 8 softmax warps (the MUFU-bound stage)
// work on unit u:
//     attention issuer:  PV_{u-1}, S_{u+1}          projection issuer:  G_{u+2}, OP_{u-1}
// Two issuing warps because each MMA group waits on a different producer (softmax, conv, LayerNorm + weight ring,
// Oepi): one thread blocking on them in turn serialised everything (r2b trace: 3550 clocks per unit).  What the
// tensor pipe's in-order execution used to guarantee between the two groups is now explicit: conv_u waits for
// S_{u-1} before it overwrites the single-buffered Q / K tiles, Oepi_u waits for OP_{u-1} before it rewrites the O tile.
// Measured cost of an M = 128, K = 16 tcgen05.mma on B200 (scripts/microbench_mma.cu): ~42 + N/2 clocks with A in
// shared memory, ~10 + N/2 with A in TMEM - per unit G 8 x 89, S 2 x 97, PV 7 x 26, L 7 x 18, OP 2 x 74 = 1360 clocks.
//   warp 0: weight producer   warp 1: attention MMA issuer   warp 2: TMEM alloc   warp 3: projection MMA issuer
//   warps 4-7: LayerNorm slices of the next sequence, Oepi, Depi (TMEM lane quarter = warp & 3; 120 registers)
//   warps 8-11: conv only (88 registers)      warps 12-19: softmax (lane quarter x column half; 104 registers)
// DESIGN.md 3.2 has the two rounds of trace / ncu findings this structure comes from (6.64 -> 3.3 ms per launch).
#pragma once
#include "kernels_tc.cuh"

namespace avb {

constexpr int kDaThreads = 640;
constexpr int kDaXBytes = 128 * 128 * 2;    // xhat A tile: two 128B-swizzled K blocks of [128 x 64]
constexpr int kDaWkbBytes = 96 * 128;       // one K block [96 rows x 64] of a head's Wqkv image
constexpr int kDaWoHeadBytes = 128 * 64;    // one head's [128 rows x 32] slice of Wo
constexpr int kDaTileBytes = 128 * 32 * 2;  // Q / K / V / O tile
constexpr int kDaHeads = 8;
constexpr int kDaVBufs = 3;
#ifndef AVB_DATTN_QKBUFS  // Q / K tile buffers: 2 = conv of the next unit does not wait for S of this one (W ring 3 stages instead of 4)
#define AVB_DATTN_QKBUFS 1  // (r2g: 3.29-3.34 ms either way, and with OP on its own issuing warp - kept at the simpler setting)
#endif
constexpr int kDaQKBufs = AVB_DATTN_QKBUFS;

struct DaBias { float bo[128]; };  // bo' = bo + bv Wo as a kernel parameter: Depi adds it straight from the constant bank

template <bool kFuseOut>
struct DaCfg {
  static constexpr int kWStages = kFuseOut ? (kDaQKBufs == 2 ? 3 : 4) : 6;
  static constexpr int kWoBytes = kFuseOut ? kDaHeads * kDaWoHeadBytes : 0;
  static constexpr int kOBufs = 0;  // the bf16 O tile of the fused out-projection lives in TMEM (A operand of OP)
  static constexpr int kDataBytes =
      2 * kDaXBytes + kWStages * kDaWkbBytes + kWoBytes + (2 * kDaQKBufs + kDaVBufs + kOBufs) * kDaTileBytes;
  static constexpr int kBiasBytes = (kDaHeads * 32 + (kFuseOut ? 0 : 256)) * 4;  // bq (| bv when the out-projection is not fused) as fp32; bo' is a kernel parameter
  static constexpr int kStageBytes = 0;  // Depi stages its boxes in the xhat tile of the sequence it finishes
  // (no alignment slack: the dynamic shared memory window starts 1024-aligned - the kernel traps if it ever does not)
  static constexpr int kSmemBytes = kDataBytes + 1024 /*barriers, TMEM slot, ones tile*/ + kBiasBytes + kStageBytes;
};
static_assert(DaCfg<true>::kSmemBytes <= 232448, "dattn_tc_kernel: shared memory over the sm_100 limit");

// Bounded (trapping) mbarrier waits in the softmax warps as well: a protocol bug must fail the launch, never hang the
// GPU box.  0 selects the lean hot-loop waits once a build has been validated on hardware.
#ifndef AVB_DATTN_G_GROUP  // projection MMAs issued per pacing group (1, 2 or 4 = unpaced).  Measured (r2d): 4.20 / 4.62 /
#define AVB_DATTN_G_GROUP 4  // 5.49 ms for 4 / 2 / 1 - the projection's own latency (s_free -> G -> conv -> S) is as critical
#endif                       // as the PV MMAs it delays, so pacing it loses more than it frees
#ifndef AVB_DATTN_WAIT_NS  // nanosleep between polls of the two long waits of the conv / epilogue warps (0 = tight polling)
#define AVB_DATTN_WAIT_NS 0
#endif
#ifndef AVB_DATTN_BOUNDED_WAITS
#define AVB_DATTN_BOUNDED_WAITS 1
#endif
#ifndef AVB_DATTN_POLY_EVERY  // every n-th pair of exponentials on the FMA pipe (17 = none).  This kernel is bound by issue slots
#define AVB_DATTN_POLY_EVERY 17  // and role latencies, not by the MUFU (XU pipe 42 %): a polynomial pair costs 11 slots, a MUFU pair 3.
#endif                           // r2g, ms per launch: none 3.24 | 1/6 3.25 | 1/4 3.30 | 1/3 3.33 (the n-axis kernel's share) | 1/2 3.44
#ifndef AVB_DATTN_OP_WARP  // 1: OP has its own issuing warp (2) instead of sharing warp 3 with G
#define AVB_DATTN_OP_WARP 0
#endif
#ifndef AVB_DATTN_LN_H0  // first of the four units of a sequence whose epilogue carries a LayerNorm slice of the next sequence
#define AVB_DATTN_LN_H0 1  // (r2g, ms per launch: 3.47 / 3.40 / 3.51 for 0 / 1 / 2 - the unit behind a Depi stays light)
#endif
#if AVB_DATTN_BOUNDED_WAITS
#define AVB_DA_WAIT(bar_ptr, parity, tag) ptx::mbar_wait((bar_ptr), (parity), (tag))
#else
#define AVB_DA_WAIT(bar_ptr, parity, tag) ptx::mbar_wait_s(ptx::smem_u32(bar_ptr), (parity))
#endif

#ifdef AVB_DATTN_DEBUG
// Debug builds only: CTA 0 dumps the intermediates of its first sequence (see scripts/debug_dattn.py)
//   [0, 8*128*96)            Q|K|V rows after bias, as float, per head
//   + 8*128*128              raw scores read by the softmax warps, per head
//   + 8*128*32               O / L per head
//   + 128*128                DELTA + bo
__device__ float* g_dattn_dbg = nullptr;
#endif

// kPack (with kFuseOut): sequences of T <= 64 tokens are packed g = 128 / T to a tile - the tile's g T rows are g T
// consecutive tokens, every per-row stage is unchanged, and the softmax masks the scores between different sequences
// (block-diagonal P), so a tile costs what it always costs and a sequence 1 / g of it.  "Sequence" below then reads
// "tile"; tm_delta_last serves a last tile that holds fewer than g sequences.
template <bool kFuseOut, bool kPack = false>
__global__ void __launch_bounds__(kDaThreads, 1)
dattn_tc_kernel(const float* __restrict__ z, const uint8_t* __restrict__ wqkv_img, const uint8_t* __restrict__ wo_img,
                const float* __restrict__ bqkv, const float* __restrict__ bo, __nv_bfloat16* __restrict__ out, int S, int T,
                int* __restrict__ guard, const __grid_constant__ CUtensorMap tm_delta, const __grid_constant__ DaBias cb,
                const __grid_constant__ CUtensorMap tm_delta_last) {
  using Cfg = DaCfg<kFuseOut>;
  constexpr int NS = Cfg::kWStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();  // the swizzled operand images need 1024-byte alignment
  uint8_t* sX = smem;                               // [2][32 KB]
  uint8_t* sW = sX + 2 * kDaXBytes;                 // [NS][12 KB]
  uint8_t* sWo = sW + NS * kDaWkbBytes;             // [8][8 KB] (kFuseOut)
  uint8_t* sQ = sWo + Cfg::kWoBytes;                // [kDaQKBufs] x { Q 8 KB | K 8 KB }
  uint8_t* sK = sQ + kDaTileBytes;
  uint8_t* sV = sQ + kDaQKBufs * 2 * kDaTileBytes;  // [3][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* x_full = bars;                 // [2] 128 LN threads
  uint64_t* x_empty = bars + 2;            // [2] commit
  uint64_t* w_full = bars + 4;             // [NS] tx
  uint64_t* w_empty = w_full + NS;         // [NS] commit
  uint64_t* g_full = w_empty + NS;         // [2] commit: QKV accumulator of unit u in SBUF[u & 1]
  uint64_t* qkv_ready = g_full + 2;        // [2] 128 conv threads
  uint64_t* s_full = qkv_ready + 2;        // [2] commit
  uint64_t* s_free = s_full + 2;           // [2] 256 softmax threads
  uint64_t* o_ready = s_free + 2;          // [2] 128 conv threads: O read (and its bf16 tile written)
  uint64_t* p_full = o_ready + 2;          // 256 softmax threads
  uint64_t* pv_done = p_full + 1;          // commit
  uint64_t* d_full = pv_done + 1;          // commit
  uint64_t* d_free = d_full + 1;           // 128 conv threads
  uint64_t* wo_full = d_free + 1;          // tx
  uint64_t* op_done = wo_full + 1;         // commit: OP of a unit has read the bf16 O tile in TMEM (kFuseOut)
  uint64_t* v_free = op_done + 1;          // [kDaVBufs] commit: PV of a unit has read its V tile
  uint64_t* o_free = v_free + kDaVBufs;    // [2] 128 epilogue threads: the O / L accumulators are in registers
  uint64_t* g_pace = o_free + 2;           // commit: a group of projection MMAs has completed (issue pacing)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_pace + 1);
  uint32_t* s_ones = reinterpret_cast<uint32_t*>(smem + Cfg::kDataBytes + 512);  // 512 B of bf16 1.0
  // Biases live in shared memory: a __ldg from the epilogue warps queues in the L1 pipeline behind the LayerNorm warps'
  // row loads (L2 latency each), which ncu showed as ~500 clocks of long-scoreboard stall per 32-column chunk.
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::kDataBytes + 1024);       // [256] bq | [128] bo' = bo + bv Wo

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = kPack ? 128 / T : 1;       // sequences per tile
  const int Tt = g * T;                    // rows of a full tile
  const int ntiles = (S + g - 1) / g;
  const int nseq = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;  // tiles of this CTA
  const int U = nseq * kDaHeads;
  const int kc = (Tt + 15) & ~15;  // key columns the MMAs touch
  // valid rows of the CTA's tile number si (the last tile of the launch may hold fewer than g sequences)
  auto vrows_of = [&](int si) {
    const int tile = (int)blockIdx.x + si * (int)gridDim.x;
    return kPack ? min(g, S - tile * g) * T : T;
  };
  // trace builds (-DAVB_TRACE_ENABLE, scripts/trace_dattn.py): CTA 0 stamps clock64() per role and unit
  long long* trace = (blockIdx.x == 0 && lane == 0 && (warp <= 4 || warp == 8 || warp == 12)) ? g_attn_trace : nullptr;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&x_full[i], 128); ptx::mbar_init(&x_empty[i], 1);
      ptx::mbar_init(&g_full[i], 1);   ptx::mbar_init(&qkv_ready[i], 128);
      ptx::mbar_init(&s_full[i], 1);   ptx::mbar_init(&s_free[i], 256);
      ptx::mbar_init(&o_ready[i], 128);
    }
    for (int i = 0; i < NS; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    ptx::mbar_init(p_full, 256); ptx::mbar_init(pv_done, 1);
    ptx::mbar_init(d_full, 1);   ptx::mbar_init(d_free, 128);
    ptx::mbar_init(wo_full, 1);
    ptx::mbar_init(op_done, 1);
    for (int i = 0; i < kDaVBufs; ++i) ptx::mbar_init(&v_free[i], 1);
    ptx::mbar_init(&o_free[0], 128); ptx::mbar_init(&o_free[1], 128);
    ptx::mbar_init(g_pace, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x >= 128 && threadIdx.x < 256) s_ones[threadIdx.x - 128] = 0x3F803F80u;
  for (int i = threadIdx.x; i < kDaHeads * 32 + (kFuseOut ? 0 : 256); i += kDaThreads)
    s_bias[i] = i < kDaHeads * 32 ? bqkv[i] : bqkv[i + kDaHeads * 32];  // bv = bqkv + 512
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tP = tmem_base + 256, tO = tmem_base + 320, tL = tmem_base + 352, tOb = tmem_base + 368, tD = tmem_base + 384;

  // Register budget per warpgroup (the pool holds what the CTA was launched with, 640 x 96 = 128 x (64 + 120 + 88) +
  // 256 x 104).  Each setmaxnreg sits at the top of its own role branch: ptxas sizes the register allocation of a
  // region by the setmaxnreg that dominates it.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ======================= weight producer: Wo once, then the Wqkv K blocks of every unit (head = unit & 7)
    if (kFuseOut && nseq > 0) {
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(wo_full, kDaHeads * kDaWoHeadBytes);
        for (int hh = 0; hh < kDaHeads; ++hh)
          ptx::bulk_load_hint(sWo + hh * kDaWoHeadBytes, wo_img + hh * kDaWoHeadBytes, kDaWoHeadBytes, wo_full, ptx::kL2EvictLast);
      }
      __syncwarp();
    }
    for (int kbi = 0; kbi < 2 * U; ++kbi) {
      const int st = kbi % NS;
      AVB_TRACE(0, kbi, 0);
      ptx::mbar_wait(&w_empty[st], (uint32_t)((kbi / NS) & 1) ^ 1, 80, 100u);
      AVB_TRACE(0, kbi, 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&w_full[st], kDaWkbBytes);
        ptx::bulk_load_hint(sW + st * kDaWkbBytes, wqkv_img + (size_t)((kbi >> 1) & 7) * (2 * kDaWkbBytes) + (kbi & 1) * kDaWkbBytes,
                            kDaWkbBytes, &w_full[st], ptx::kL2EvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ======================= attention MMA issuer: scores S_u = Q K^T, then O = P V and L = P 1
    // (The projection MMAs G / OP have their own issuing warp: a single issuer blocked on one of its four barrier
    //  kinds held back the other three - trace r2b: 3550 clocks per unit with the tensor pipe idle most of the time.)
    if (U > 0) {
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
      constexpr uint32_t idesc_l = ptx::make_idesc_bf16(128, 16, 0, 0);
      const uint32_t idesc_s = ptx::make_idesc_bf16(128, kc, 0, 0);
      const int ksteps = kc >> 4;
      const uint64_t dQ = ptx::make_smem_desc(ptx::smem_u32(sQ), 16, 512, ptx::kSwz64);
      const uint64_t dK = ptx::make_smem_desc(ptx::smem_u32(sK), 16, 512, ptx::kSwz64);
      const uint64_t dV0 = ptx::make_smem_desc(ptx::smem_u32(sV), 16, 512, ptx::kSwz64);
      const uint64_t d_ones = ptx::make_smem_desc(ptx::smem_u32(s_ones), 128, 256, ptx::kSwzNone);
      auto issue_S = [&](int u) {  // scores of unit u into SBUF[u & 1] (over the QKV accumulator conv_u has consumed)
        ptx::mbar_wait(&qkv_ready[u & 1], (uint32_t)((u >> 1) & 1), 84, 40u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(u & 1) * 128u;
          const uint64_t qk = (uint64_t)((u % kDaQKBufs) * ((2 * kDaTileBytes) >> 4));
          ptx::umma_bf16(d_tmem, dQ + qk, dK + qk, idesc_s, 0);
          ptx::umma_bf16(d_tmem, dQ + qk + 2, dK + qk + 2, idesc_s, 1);
          ptx::umma_commit(&s_full[u & 1]);
        }
        __syncwarp();
      };
      auto issue_PV = [&](int u) {
        ptx::mbar_wait(p_full, (uint32_t)(u & 1), 85, 40u);
        if (u >= 1) ptx::mbar_wait(&o_free[(u - 1) & 1], (uint32_t)(((u - 1) >> 1) & 1), 86, 40u);  // O of unit u-1 has been read
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dv = dV0 + (uint64_t)((u % kDaVBufs) * (kDaTileBytes >> 4));
          // Straight-line issue: as a rolled loop this was ~20 uniform-datapath instructions per pair of MMAs, ~1000
          // clocks per unit for the issuing thread - long enough for G of three units ahead (ready ~400 clocks after P)
          // to get into the in-order tensor pipe between the PV MMAs, 89 clocks each (r2g trace: p_full -> pv_done 1450
          // clocks, the softmax warps of the next unit waiting ~500 of them to store their P tile).
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < ksteps) {
              ptx::umma_bf16_ts(tO, tP + k * 8, dv + (uint64_t)(k * 64), idesc_pv, k != 0);
              ptx::umma_bf16_ts(tL, tP + k * 8, d_ones, idesc_l, k != 0);
            }
          }
          ptx::umma_commit(pv_done);
          ptx::umma_commit(&v_free[u % kDaVBufs]);
        }
        __syncwarp();
      };
      // (Tried, r2e: non-blocking tests of both groups' barriers, whichever is ready first - 4.42 vs 4.00 ms: the polling warp
      //  competes with the softmax warps of its SM sub-partition for issue slots.)
      issue_S(0);
      for (int u = 0; u < U; ++u) {
        AVB_TRACE(1, u, 0);
        if (u >= 1) issue_PV(u - 1);
        AVB_TRACE(1, u, 1);
        if (u + 1 < U) issue_S(u + 1);
        AVB_TRACE(1, u, 2);
      }
      issue_PV(U - 1);
    }
  } else if (warp == 3 || (AVB_DATTN_OP_WARP && warp == 2)) {
    // ======================= projection MMA issuer: G_u = xhat Wqkv_h^T two units ahead, DELTA += O_h Wo_h^T one behind
    if (U > 0) {
      constexpr uint32_t idesc_g = ptx::make_idesc_bf16(128, 96, 0, 0);
      constexpr uint32_t idesc_op = ptx::make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dX0 = ptx::make_smem_desc(ptx::smem_u32(sX), 16, 1024, ptx::kSwz128);
      const uint64_t dW0 = ptx::make_smem_desc(ptx::smem_u32(sW), 16, 1024, ptx::kSwz128);
      const uint64_t dWo0 = ptx::make_smem_desc(ptx::smem_u32(sWo), 16, 512, ptx::kSwz64);
      uint32_t pace_phase = 0;
      auto issue_G = [&](int u) {  // QKV projection of unit u into SBUF[u & 1]
        const int si = u >> 3, h = u & 7, buf = si & 1;
        if (h == 0) ptx::mbar_wait(&x_full[buf], (uint32_t)((si >> 1) & 1), 81, 40u);
        if (u >= 2) ptx::mbar_wait(&s_free[u & 1], (uint32_t)(((u >> 1) - 1) & 1), 82, 40u);  // softmax of unit u-2 has read S
        AVB_TRACE(5, (u >= 2 ? u - 2 : 499), 5);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(u & 1) * 128u;
        for (int kb = 0; kb < 2; ++kb) {
          const int kbi = 2 * u + kb, st = kbi % NS;
          ptx::mbar_wait(&w_full[st], (uint32_t)((kbi / NS) & 1), 83, 40u);
          ptx::tc_fence_after();
          // The tensor pipe executes in issue order and a blocked issue holds the issuing thread: eight projection MMAs
          // queued back to back (712 clocks, needed two units from now) held up the PV MMAs of the attention issuer
          // (needed now) - r2d trace: PV issue 1450 clocks.  The projection is paced in groups of AVB_DATTN_G_GROUP MMAs,
          // each waited for, so the attention MMAs slip in between.
          const uint64_t da = dX0 + (uint64_t)(buf * (kDaXBytes >> 4) + kb * ((128 * 128) >> 4));
          const uint64_t db = dW0 + (uint64_t)(st * (kDaWkbBytes >> 4));
#pragma unroll 1
          for (int k0 = 0; k0 < 4; k0 += AVB_DATTN_G_GROUP) {
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = k0; k < k0 + AVB_DATTN_G_GROUP; ++k)
                ptx::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_g, (kb | k) != 0);
              if (AVB_DATTN_G_GROUP < 4) ptx::umma_commit(g_pace);
              if (k0 + AVB_DATTN_G_GROUP == 4) {
                ptx::umma_commit(&w_empty[st]);
                if (kb == 1) {
                  ptx::umma_commit(&g_full[u & 1]);
                  if (h == 7) ptx::umma_commit(&x_empty[buf]);
                }
              }
            }
            __syncwarp();
            if (AVB_DATTN_G_GROUP < 4) { ptx::mbar_wait(g_pace, pace_phase, 99, 20u); pace_phase ^= 1u; }
          }
        }
      };
      auto issue_OP = [&](int u) {  // DELTA (+)= O_h . Wo_h^T
        const int si = u >> 3, h = u & 7;
        ptx::mbar_wait(&o_ready[u & 1], (uint32_t)((u >> 1) & 1), 87, 40u);
        if (h == 0 && si >= 1) ptx::mbar_wait(d_free, (uint32_t)((si - 1) & 1), 88, 40u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t db = dWo0 + (uint64_t)(h * (kDaWoHeadBytes >> 4));
          ptx::umma_bf16_ts(tD, tOb, db, idesc_op, h != 0);  // A = the bf16 O tile in TMEM: 8 columns per K = 16 slice
          ptx::umma_bf16_ts(tD, tOb + 8, db + 2, idesc_op, 1);
          ptx::umma_commit(op_done);  // the O tile may be overwritten (Oepi of the next unit)
          if (h == 7) ptx::umma_commit(d_full);
        }
        __syncwarp();
      };
      if (AVB_DATTN_OP_WARP && warp == 2) {
        if (kFuseOut) {
          ptx::mbar_wait(wo_full, 0, 89, 40u);
          for (int u = 0; u < U; ++u) issue_OP(u);
        }
      } else {
        if (kFuseOut && !AVB_DATTN_OP_WARP) ptx::mbar_wait(wo_full, 0, 89, 40u);
        issue_G(0);
        if (U > 1) issue_G(1);
        for (int u = 0; u < U; ++u) {
          AVB_TRACE(5, u, 0);
          if (u + 2 < U) issue_G(u + 2);
          AVB_TRACE(5, u, 1);
          if (kFuseOut && !AVB_DATTN_OP_WARP && u >= 1) issue_OP(u - 1);
          AVB_TRACE(5, u, 2);
        }
        if (kFuseOut && !AVB_DATTN_OP_WARP) issue_OP(U - 1);
      }
    }
  }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    // ======================= LayerNorm + epilogue warps (Oepi, Depi): thread = row r of the tile = TMEM lane for the
    // epilogues.  The LayerNorm of the NEXT sequence is cut into four 8-row slices per warp, one per unit h = 0..3, with
    // the slice's loads in flight across that unit's Oepi (which has no async-proxy fence any more - a fence is
    // MEMBAR.ALL.CTA and would drain them): a whole 32-row LayerNorm in one piece would hold up the O hand-back once per
    // sequence.  These warps wait ~1200 clocks per unit for PV otherwise (r2d trace); the conv warps have no slack.
    const int w = warp - 4, q = w, r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int rs = lane >> 3, j8 = lane & 7;
    float4 ln[2][4];
    auto ln8_load = [&](const float* zs, int r0, int vr) {  // rows 32w + r0 .. + 7 of the sequence (zeros past its end)
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const int row = w * 32 + r0 + uu * 4 + rs;
        const float4* zp = reinterpret_cast<const float4*>(zs + (size_t)row * 128) + j8;
#pragma unroll
        for (int k = 0; k < 4; ++k) ln[uu][k] = row < vr ? __ldcg(zp + k * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto ln8_finish = [&](uint32_t a_tile, int r0) {  // same layout as ln_warp_rows_to_umma_tile
      // Sum and sum of squares in one pass (as the FFN kernel's LayerNorm does), the four reductions of the slice's two
      // row groups interleaved: three shuffle rounds on the epilogue warps' critical path instead of twelve dependent ones.
      float sum[2], sq[2];
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s0 += ln[uu][k].x + ln[uu][k].y; s1 += ln[uu][k].z + ln[uu][k].w;
          q0 = fmaf(ln[uu][k].x, ln[uu][k].x, q0); q1 = fmaf(ln[uu][k].y, ln[uu][k].y, q1);
          q0 = fmaf(ln[uu][k].z, ln[uu][k].z, q0); q1 = fmaf(ln[uu][k].w, ln[uu][k].w, q1);
        }
        sum[uu] = s0 + s1; sq[uu] = q0 + q1;
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          sum[uu] += __shfl_xor_sync(0xffffffffu, sum[uu], o);
          sq[uu] += __shfl_xor_sync(0xffffffffu, sq[uu], o);
        }
      }
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const float mean = sum[uu] * (1.0f / 128.0f);
        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, sq[uu] * (1.0f / 128.0f)), 0.f) + kLnEps);
        const float shift = -mean * rstd;
        const int rr = w * 32 + r0 + uu * 4 + rs;
        const uint32_t rbase = a_tile + (uint32_t)((rr >> 3) * 1024 + (rr & 7) * 128 + (j8 & 1) * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int chunk = (j8 + 8 * (k & 1)) >> 1;
          const uint32_t addr = rbase + (uint32_t)((k >> 1) * (128 * 128) + ((chunk ^ (rr & 7)) << 4));
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr),
                       "r"(ptx::pack_bf16(fmaf(ln[uu][k].x, rstd, shift), fmaf(ln[uu][k].y, rstd, shift))),
                       "r"(ptx::pack_bf16(fmaf(ln[uu][k].z, rstd, shift), fmaf(ln[uu][k].w, rstd, shift)))
                       : "memory");
        }
      }
    };
    // Oepi in two parts, so that the accumulator hand-back (part A) never waits behind the Depi of the previous sequence:
    //   A: O, L -> registers, o_free (PV of the next unit may start), O / L -> packed bf16
    //   B: the bf16 tile -> TMEM as the A operand of OP (or the attention output in HBM), o_ready
    auto oepi_a = [&](int u, uint4 (&pk)[4]) {
      const int h = u & 7;
      ptx::mbar_wait(pv_done, (uint32_t)(u & 1), 92, AVB_DATTN_WAIT_NS);
      AVB_TRACE(4, u, 3);
      AVB_TRACE(7, u, 0);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld32(tO + lane_off, o);
      const float lsum = __uint_as_float(ptx::tmem_ld1(tL + lane_off));
      ptx::tmem_wait_ld();
      // the accumulators are in registers: PV of the next unit may start (the O-accumulator hand-back is the unit
      // pipeline's critical cycle - r2c trace - so it does not wait for the arithmetic below)
      ptx::tc_fence_before();
      ptx::mbar_arrive(&o_free[u & 1]);
      AVB_TRACE(7, u, 1);
      if (r < vrows_of(u >> 3) && !(lsum < 1.2676506e30f && lsum > 7.8886091e-31f)) *guard = 1;  // outside (2^-100, 2^100), inf or NaN
      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(lsum));  // 1 ulp: far below the bf16 rounding of O
#ifdef AVB_DATTN_DEBUG
      if (g_dattn_dbg && blockIdx.x == 0 && u < 8)
        for (int t = 0; t < 32; ++t) g_dattn_dbg[(size_t)8 * 128 * 96 + 8 * 128 * 128 + ((size_t)u * 128 + r) * 32 + t] = __uint_as_float(o[t]) * inv;
#endif
      if constexpr (kFuseOut) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pk[e].x = ptx::pack_bf16(__uint_as_float(o[e * 8 + 0]) * inv, __uint_as_float(o[e * 8 + 1]) * inv);
          pk[e].y = ptx::pack_bf16(__uint_as_float(o[e * 8 + 2]) * inv, __uint_as_float(o[e * 8 + 3]) * inv);
          pk[e].z = ptx::pack_bf16(__uint_as_float(o[e * 8 + 4]) * inv, __uint_as_float(o[e * 8 + 5]) * inv);
          pk[e].w = ptx::pack_bf16(__uint_as_float(o[e * 8 + 6]) * inv, __uint_as_float(o[e * 8 + 7]) * inv);
        }
      } else {
        // (the separate out-projection launch uses the plain bo, so the value bias is added here)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b0 = ptx::ld_shared_v4f(ptx::smem_u32(s_bias) + (uint32_t)((kDaHeads * 32 + h * 32 + e * 8) * 4));
          const float4 b1 = ptx::ld_shared_v4f(ptx::smem_u32(s_bias) + (uint32_t)((kDaHeads * 32 + h * 32 + e * 8 + 4) * 4));
          pk[e].x = ptx::pack_bf16(__uint_as_float(o[e * 8 + 0]) * inv + b0.x, __uint_as_float(o[e * 8 + 1]) * inv + b0.y);
          pk[e].y = ptx::pack_bf16(__uint_as_float(o[e * 8 + 2]) * inv + b0.z, __uint_as_float(o[e * 8 + 3]) * inv + b0.w);
          pk[e].z = ptx::pack_bf16(__uint_as_float(o[e * 8 + 4]) * inv + b1.x, __uint_as_float(o[e * 8 + 5]) * inv + b1.y);
          pk[e].w = ptx::pack_bf16(__uint_as_float(o[e * 8 + 6]) * inv + b1.z, __uint_as_float(o[e * 8 + 7]) * inv + b1.w);
        }
      }
      AVB_TRACE(7, u, 2);
    };
    auto oepi_b = [&](int u, const uint4 (&pk)[4]) {
      if constexpr (kFuseOut) {
        // straight back into TMEM as the A operand of OP: no shared-memory round trip, and no async-proxy fence
        // (MEMBAR.ALL.CTA: it would drain this thread's LayerNorm loads in flight)
        if (u >= 1) ptx::mbar_wait(op_done, (uint32_t)((u - 1) & 1), 97);  // OP of unit u-1 has read the tile
        AVB_TRACE(7, u, 3);
        ptx::tc_fence_after();
        uint32_t ob[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) { ob[e * 4 + 0] = pk[e].x; ob[e * 4 + 1] = pk[e].y; ob[e * 4 + 2] = pk[e].z; ob[e * 4 + 3] = pk[e].w; }
        ptx::tmem_st16(tOb + lane_off, ob);
        ptx::tmem_wait_st();
        AVB_TRACE(7, u, 4);
      } else {
        if (r < T) {  // (never packed) outproj_tc_kernel's A-operand layout: [m >> 7][head][m & 127][32 ch], chunks at c ^ ((m >> 1) & 3)
          const int h = u & 7;
          const int s = (int)blockIdx.x + (u >> 3) * (int)gridDim.x;
          const unsigned m = (unsigned)s * (unsigned)T + (unsigned)r;
          __nv_bfloat16* orow = out + (((size_t)(m >> 7) * kDaHeads + h) * 128 + (m & 127)) * 32;
          const unsigned mswz = (m >> 1) & 3;
#pragma unroll
          for (int e = 0; e < 4; ++e) *reinterpret_cast<uint4*>(orow + (((unsigned)e ^ mswz) << 3)) = pk[e];
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&o_ready[u & 1]);
    };
    auto depi = [&](int si) {  // DELTA + bo' -> bf16 -> the sequence's 256-byte delta rows (token order)
      // Row-per-thread 16-byte global stores cost ~1800 LSU data wavefronts per sequence and the next chunk's TMEM load
      // waited for the stores to release their source registers (ncu r2g: 3100 clocks per sequence, with the whole unit
      // pipeline stalled behind it once per sequence).  Each warp now stages [32 rows x 32 channels] boxes in the 64B-
      // swizzle image (conflict-free 16-byte shared stores) and one lane hands them to the TMA: a 3-D tensor map
      // {channel, row, sequence} clips the rows past the end of the sequence.  The stage is the warp's own 8 KB of this
      // sequence's xhat tile: G of the last head has read it (d_full is causally behind g_full of that unit), and the
      // next writer is this same warp's LayerNorm slice two sequences on, behind the wait_group.read in the unit loop.
      // The bias comes from the constant bank (kernel parameter): no shared-memory loads in front of the adds.
      AVB_TRACE(6, si, 0);
      ptx::mbar_wait(d_full, (uint32_t)(si & 1), 93);
      AVB_TRACE(6, si, 1);
      ptx::tc_fence_after();
      const int s = (int)blockIdx.x + si * (int)gridDim.x;
      const uint32_t stage = ptx::smem_u32(sX + (si & 1) * kDaXBytes) + (uint32_t)(w * 4096);  // + 16 KB for chunks 2, 3
      const uint32_t swz = (uint32_t)((lane >> 1) & 3);
#pragma unroll
      for (int cp = 0; cp < 4; cp += 2) {
        uint32_t v[2][32];
        ptx::tmem_ld32(tD + cp * 32 + lane_off, v[0]);
        ptx::tmem_ld32(tD + cp * 32 + 32 + lane_off, v[1]);
        ptx::tmem_wait_ld();
        if (cp == 2) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(d_free);  // accumulator read: the next sequence's first OP may overwrite it
        }
        if (cp == 0) AVB_TRACE(6, si, 2);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cc = cp + c2;
          const uint32_t my = stage + (uint32_t)((cc >> 1) * 16384 + (cc & 1) * 2048 + lane * 64);
#ifdef AVB_DATTN_DEBUG
          if (g_dattn_dbg && blockIdx.x == 0 && si == 0)
            for (int t = 0; t < 32; ++t)
              g_dattn_dbg[(size_t)8 * 128 * 96 + 8 * 128 * 128 + 8 * 128 * 32 + (size_t)r * 128 + cc * 32 + t] =
                  __uint_as_float(v[c2][t]) + cb.bo[cc * 32 + t];
#endif
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float* b = cb.bo + cc * 32 + e * 8;
            const uint32_t p0 = ptx::pack_bf16(__uint_as_float(v[c2][e * 8 + 0]) + b[0], __uint_as_float(v[c2][e * 8 + 1]) + b[1]);
            const uint32_t p1 = ptx::pack_bf16(__uint_as_float(v[c2][e * 8 + 2]) + b[2], __uint_as_float(v[c2][e * 8 + 3]) + b[3]);
            const uint32_t p2 = ptx::pack_bf16(__uint_as_float(v[c2][e * 8 + 4]) + b[4], __uint_as_float(v[c2][e * 8 + 5]) + b[5]);
            const uint32_t p3 = ptx::pack_bf16(__uint_as_float(v[c2][e * 8 + 6]) + b[6], __uint_as_float(v[c2][e * 8 + 7]) + b[7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + (((uint32_t)e ^ swz) << 4)), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                         : "memory");
          }
        }
        if (cp == 0) AVB_TRACE(6, si, 3);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      AVB_TRACE(6, si, 4);
      const int vr = vrows_of(si);
      if (lane < 4 && w * 32 < vr) {  // one box per lane: a single issue slot instead of four dependent ones
        if (kPack && vr < Tt)         // the launch's last tile, fewer sequences: its own map clips at its own row count
          ptx::tma_store_3d(&tm_delta_last, stage + (uint32_t)((lane >> 1) * 16384 + (lane & 1) * 2048), lane * 32, w * 32, 0);
        else
          ptx::tma_store_3d(&tm_delta, stage + (uint32_t)((lane >> 1) * 16384 + (lane & 1) * 2048), lane * 32, w * 32, s);
        ptx::bulk_commit();
      }
      AVB_TRACE(6, si, 5);
    };

    if (nseq > 0) {  // the first sequence's LayerNorm, slice by slice
      const float* zs = z + (size_t)blockIdx.x * Tt * 128;
      if (nseq > 1) prefetch_rows_l2(z + (size_t)((int)blockIdx.x + (int)gridDim.x) * Tt * 128, vrows_of(1), w * 32, lane);
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 8) {
        ln8_load(zs, r0, vrows_of(0));
        ln8_finish(ptx::smem_u32(sX), r0);
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&x_full[0]);
    }
    for (int si = 0; si < nseq; ++si) {
      const int s_next = (int)blockIdx.x + (si + 1) * (int)gridDim.x, nbuf = (si + 1) & 1;
      const bool has_next = si + 1 < nseq;
      if (si + 2 < nseq) prefetch_rows_l2(z + (size_t)(s_next + (int)gridDim.x) * Tt * 128, vrows_of(si + 2), w * 32, lane);
      for (int h = 0; h < kDaHeads; ++h) {
        const int u = si * kDaHeads + h;
        const bool slice = has_next && h >= AVB_DATTN_LN_H0 && h < AVB_DATTN_LN_H0 + 4;
        const int hs = h - AVB_DATTN_LN_H0;  // slice index
        AVB_TRACE(4, u, 0);
        uint4 pk[4];
        if (slice) {
          if (hs == 0) ptx::mbar_wait(&x_empty[nbuf], (uint32_t)(((si + 1) >> 1) & 1) ^ 1, 90);
          ln8_load(z + (size_t)s_next * Tt * 128, hs * 8, vrows_of(si + 1));  // (issued a unit earlier, behind the previous slice: 3.50 vs 3.36 ms)
        }
        // (Tried, r2g: the second half of the previous sequence's Depi behind this unit's accumulator hand-back, 3.72 vs
        //  3.57 ms - the hand-back of the unit after that waits instead.)
        oepi_a(u, pk);
        oepi_b(u, pk);
        AVB_TRACE(4, u, 2);
        if (slice) {
          if (kFuseOut && hs == 0 && si > 0) {  // Depi of the previous sequence staged its boxes in this buffer
            if (lane < 4) ptx::bulk_wait_read<0>();
            __syncwarp();
          }
          ln8_finish(ptx::smem_u32(sX + nbuf * kDaXBytes), hs * 8);
          if (hs == 3) {
            ptx::fence_proxy_async();
            ptx::mbar_arrive(&x_full[nbuf]);
          }
        }
        if (kFuseOut && h == kDaHeads - 1) depi(si);
        AVB_TRACE(4, u, 5);
      }
    }
    if (kFuseOut && lane < 4) ptx::bulk_wait_all();  // the last boxes must have left shared memory before the CTA exits
  } else if (warp < 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // ======================= conv warps: thread = row r of the tile = TMEM lane.  Nothing else runs here: qkv_ready ->
    // S_u -> softmax_u is the unit pipeline's critical chain.
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)(r * 64), swz = (uint32_t)((r >> 1) & 3);
    const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);
    auto conv_wait = [&](int u) {
      ptx::mbar_wait(&g_full[u & 1], (uint32_t)((u >> 1) & 1), 91, AVB_DATTN_WAIT_NS);
      // single-buffered Q / K tiles: S of the previous unit (issued by another warp than G) must have read them.  With two
      // buffers the reader was S of unit u - 2, which the softmax had loaded before G of this unit could be issued.
      if (kDaQKBufs == 1 && u >= 1) ptx::mbar_wait(&s_full[(u - 1) & 1], (uint32_t)(((u - 1) >> 1) & 1), 96);
      // ... and the V buffer (three of them) is last read by PV of unit u - 3
      if (u >= kDaVBufs) ptx::mbar_wait(&v_free[u % kDaVBufs], (uint32_t)((u / kDaVBufs - 1) & 1), 98);
      AVB_TRACE(2, u, 2);
    };
    auto conv = [&](int u) {  // QKV accumulator -> + bias -> bf16 -> Q / K / V tiles
      const int h = u & 7;
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(u & 1) * 128u + lane_off;
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        uint32_t v[32];
        ptx::tmem_ld32(tacc + w * 32, v);
        // Only Q carries its bias here.  The key bias shifts every score of a query row by the same q . bk, which the
        // softmax cancels exactly; the value bias passes through the convex combination unchanged (sum_j p_j = 1) and
        // is folded into the out-projection bias at load time (bo + bv Wo, avb_load_weights) - 64 FADD and 16 shared
        // loads less per row and unit on warps whose instruction count is what bounds the unit period.
        float4 bb[8];
        if (w == 0) {
          const uint32_t bp = ptx::smem_u32(s_bias) + (uint32_t)(h * 32 * 4);  // shared-memory broadcast (explicit ld.shared)
#pragma unroll
          for (int e = 0; e < 8; ++e) bb[e] = ptx::ld_shared_v4f(bp + e * 16);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) bb[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ptx::tmem_wait_ld();
        const uint32_t qk = (uint32_t)((u % kDaQKBufs) * 2 * kDaTileBytes);
        const uint32_t tile = w == 0 ? aQ + qk : (w == 1 ? aK + qk : aV + (uint32_t)((u % kDaVBufs) * kDaTileBytes));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[e * 8 + t]);
          if (w == 0) {
            f[0] += bb[e * 2].x;     f[1] += bb[e * 2].y;     f[2] += bb[e * 2].z;     f[3] += bb[e * 2].w;
            f[4] += bb[e * 2 + 1].x; f[5] += bb[e * 2 + 1].y; f[6] += bb[e * 2 + 1].z; f[7] += bb[e * 2 + 1].w;
          }
#ifdef AVB_DATTN_DEBUG
          if (g_dattn_dbg && blockIdx.x == 0 && u < 8)
            for (int t = 0; t < 8; ++t) g_dattn_dbg[((size_t)u * 128 + r) * 96 + w * 32 + e * 8 + t] = f[t];
#endif
          uint4 pk;
          pk.x = ptx::pack_bf16(f[0], f[1]); pk.y = ptx::pack_bf16(f[2], f[3]);
          pk.z = ptx::pack_bf16(f[4], f[5]); pk.w = ptx::pack_bf16(f[6], f[7]);
          ptx::st_shared_v4(tile + row_off + (((uint32_t)e ^ swz) << 4), pk);
        }
        AVB_TRACE(2, u, 4 + w);
      }
      ptx::fence_proxy_async();  // tile writes (generic proxy) -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      ptx::mbar_arrive(&qkv_ready[u & 1]);
    };

    for (int u = 0; u < U; ++u) {
      AVB_TRACE(2, u, 0);
      conv_wait(u);
      conv(u);
      AVB_TRACE(2, u, 3);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ======================= softmax warps: warp (q, half) owns TMEM lanes 32q.. and key columns [64 half, 64 half + 64)
    const int q = warp & 3, half = (warp - 12) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a_pfull = ptx::smem_u32(p_full);
    const uint32_t tPw = tP + half * 32 + lane_off;
    const int nv = T - half * 64;  // valid key columns in this warp's half (may be <= 0)
    // kPack: the keys of a row are those of its own sequence, columns [lo, hi) of the tile - per thread one 32-bit mask
    // for each of its two 32-column chunks, the same for every unit (rows past the tile's sequences: empty window)
    uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu;
    bool any0 = true, any1 = true, all0 = true, all1 = true;
    if constexpr (kPack) {
      const int r = q * 32 + lane, sq = r / T;
      const int lo = sq * T, hi = sq < g ? lo + T : lo;
      auto chunk_mask = [&](int c0) {
        const int a = lo - c0, b = hi - c0;
        const uint32_t mb = b >= 32 ? 0xffffffffu : (b <= 0 ? 0u : (1u << b) - 1u);
        const uint32_t ma = a <= 0 ? 0xffffffffu : (a >= 32 ? 0u : ~((1u << a) - 1u));
        return ma & mb;
      };
      m0 = chunk_mask(half * 64); m1 = chunk_mask(half * 64 + 32);
      any0 = __any_sync(0xffffffffu, m0 != 0u); all0 = __all_sync(0xffffffffu, m0 == 0xffffffffu);
      any1 = __any_sync(0xffffffffu, m1 != 0u); all1 = __all_sync(0xffffffffu, m1 == 0xffffffffu);
    }
    float pmax = -INFINITY;
    for (int u = 0; u < U; ++u) {
      AVB_TRACE(3, u, 0);
      AVB_DA_WAIT(&s_full[u & 1], (uint32_t)((u >> 1) & 1), 94);
      AVB_TRACE(3, u, 1);
      ptx::tc_fence_after();
      const uint32_t tS = tmem_base + (uint32_t)(u & 1) * 128u + half * 64 + lane_off;
      uint32_t sa[32], sb[32];
      ptx::tmem_ld32(tS, sa);
      ptx::tmem_ld32(tS + 32, sb);
      ptx::tmem_wait_ld();
      ptx::tc_fence_before();
      ptx::mbar_arrive_s(ptx::smem_u32(&s_free[u & 1]));  // the only read of S: the buffer may take the next-but-one unit's projection
      AVB_TRACE(3, u, 2);
      if constexpr (kPack) {  // scores against other sequences' keys: -inf, probability 0
        if (any0 && !all0) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((m0 >> e) & 1u)) sa[e] = 0xff800000u;
        }
        if (any1 && !all1) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((m1 >> e) & 1u)) sb[e] = 0xff800000u;
        }
      } else if (nv < 64) {  // keys beyond the sequence end: score -inf, probability 0
        if (nv < 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e >= nv) sa[e] = 0xff800000u;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (32 + e >= nv) sb[e] = 0xff800000u;
      }
#ifdef AVB_DATTN_DEBUG
      if (g_dattn_dbg && blockIdx.x == 0 && u < 8) {
        const int r = q * 32 + lane;
        for (int t = 0; t < 32; ++t) {
          g_dattn_dbg[(size_t)8 * 128 * 96 + ((size_t)u * 128 + r) * 128 + half * 64 + t] = __uint_as_float(sa[t]);
          g_dattn_dbg[(size_t)8 * 128 * 96 + ((size_t)u * 128 + r) * 128 + half * 64 + 32 + t] = __uint_as_float(sb[t]);
        }
      }
#endif
      uint32_t pk[32];
      if (kPack ? any0 : nv > 0) {  // (a 32-column chunk past the end of the sequence is all zeros: no exponentials for it)
        exp_chunk32<true, AVB_DATTN_POLY_EVERY>(sa, pk, pmax);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      }
      if (kPack ? any1 : nv > 32) {
        exp_chunk32<true, AVB_DATTN_POLY_EVERY>(sb, pk + 16, pmax);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[16 + e] = 0u;
      }
      AVB_TRACE(3, u, 3);
      if (u >= 1) {  // P is free once the PV / L MMAs of the previous unit have finished
        AVB_DA_WAIT(pv_done, (uint32_t)((u - 1) & 1), 95);
        ptx::tc_fence_after();
      }
      AVB_TRACE(3, u, 4);
      ptx::tmem_st32(tPw, pk);
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive_s(a_pfull);
      AVB_TRACE(3, u, 5);
    }
    if (pmax > kAttnOverflowGuard) *guard = 1;  // a polynomial lane left the safe exponent range
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace avb
