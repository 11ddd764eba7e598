// C-ABI of the B200-native AVICI inference forward (see include/avici_b200.h).
//
// Host-side driver: weight packing (LayerNorm / softmax-scale folding for the bf16 path), the
// per-block kernel schedule and workspace carving.  The device work lives in kernels_f32.cuh
// (prologue / epilogue kernels + the true-fp32 parity path) and kernels_tc.cuh (bf16 tcgen05 path).
#include "../../include/avici_b200.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels_f32.cuh"
#include "kernels_tc.cuh"
#include "kernels_dattn.cuh"
#include "kernels_split.cuh"
#include "kernels_metrics.cuh"

namespace {

using bf16 = __nv_bfloat16;

thread_local std::string g_create_error;

struct Block {
  // exact fp32 parameters (both modes keep them; 17 MB in total)
  const float *ln_g[4], *ln_b[4];  // q, k, v, ffn LayerNorms
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *w1, *b1, *w2, *b2;
  // bf16 tensor-core copies: K-major [N,K], LayerNorm affine and softmax scale folded in
  const bf16 *wqkv_t, *wo_t, *w1_t, *w2_t;
  const float *bqkv_f, *b1_f;
  const float *bo_d;  // dattn_tc_kernel: bo + bv' Wo (the value bias passes through the softmax average unchanged)
  // fp32 mode, split-precision tensor-core GEMMs (kernels_split.cuh): hi / lo bf16 halves of the K-major weights with the
  // LayerNorm affine folded in (no softmax scale: the fp32 attention kernel applies it), their [128 x 64] tensor maps
  CUtensorMap tm_s_qkv[2], tm_s_o[2], tm_s_1[2], tm_s_2[2];
  avb::DaBias dattn_bias;  // host copy of bo_d for dattn_tc_kernel's parameter
  avb::FfnBias ffn_bias;  // host copy of b1' (LayerNorm-folded) and b2 for the resident FFN kernel's parameter (F = 512)
  const float *zero_k;
  // per-head operand images of dattn_tc_kernel: Wqkv_h as 2 K blocks of [96 x 64] (128B swizzle), Wo_h as [128 x 32] (64B swizzle)
  const uint8_t *wd_qkv, *wd_o;
  CUtensorMap tm_wqkv_h, tm_wo, tm_w1q, tm_w2q;  // wqkv: [128 x 64] boxes (CTA-pair QKV), wo: [128 x 64], w1 / w2: [64 x 64] (CTA-pair FFN)
};

}  // namespace

// ------------------------------------------------------------------ NCCL (resolved at run time)
// The library carries no link-time dependency on NCCL: the handful of entry points the axially sharded forward needs
// are looked up in the libnccl.so.2 that is already loaded in the process (PyTorch bundles one) or, failing that,
// in the system's.  The declarations below restate the stable public ABI of nccl.h (2.x).
extern "C" {
typedef struct ncclComm* avb_nccl_comm_t;
typedef struct { char internal[128]; } avb_nccl_unique_id;
}
namespace {
struct NcclApi {
  int (*GetUniqueId)(avb_nccl_unique_id*) = nullptr;
  int (*CommInitRank)(avb_nccl_comm_t*, int, avb_nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(avb_nccl_comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, avb_nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, avb_nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, avb_nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
constexpr int kNcclFloat32 = 7;  // ncclDataType_t::ncclFloat32

const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy PyTorch has loaded, if any
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
    auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p && api.why.empty()) api.why = std::string("missing NCCL symbol ") + name; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.why.empty();
  });
  return api;
}
}  // namespace

struct avb_model_s {
  avb_config cfg{};
  std::string err;
  int device = 0;
  int num_sms = 148;
  bool loaded = false;
  void* arena = nullptr;  // all weights
  std::vector<Block> blocks;
  const float *emb_w = nullptr, *emb_b = nullptr;
  const float *lnf_g = nullptr, *lnf_b = nullptr;  // final LN
  const float *lnu_g = nullptr, *lnu_b = nullptr, *wu = nullptr, *bu = nullptr;
  const float *lnv_g = nullptr, *lnv_b = nullptr, *wv = nullptr, *bv = nullptr;
  const float *temp = nullptr, *bias = nullptr;
  void* host_ws = nullptr;  // internal workspace of avb_forward_host (one host-buffer call at a time per handle)
  size_t host_ws_bytes = 0;
  cudaStream_t host_stream = nullptr;
  std::mutex host_mu;
  long long launches = 0;
  int dattn_mode = 1;  // d-axis blocks with d <= 128: 1 = fully fused kernel, 2 = fused up to the attention output, 0 = unfused
  int dattn_pack = 1;  // fused d-axis kernel: pack 128 / d sequences per tile when d <= 64 (0: one sequence per tile)
  int head_tc = 1;  // bf16 mode: edge-logit contraction of the head on the tensor cores (0: SIMT head_logits_kernel)
  int f32_tc = 1;  // fp32 mode: projections and FFN as split-precision tensor-core GEMMs (0: SIMT sgemm_kernel)
  avb_nccl_comm_t comm = nullptr;  // avb_comm_init: communicator of the axially sharded forward (one rank per handle)
  int comm_rank = 0, comm_world = 1;
  // optional per-kernel-category CUDA-event timing (avb_profile_*)
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  double prof_ms[AVB_PROF_CATEGORIES] = {0};
  long long prof_n[AVB_PROF_CATEGORIES] = {0};
};

namespace {

int fail(avb_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

// Records a CUDA-event pair around the launches issued while it is alive (only when profiling is on).
struct ProfScope {
  avb_handle h; cudaStream_t st; int cat; cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t take(avb_handle h) {
    if (h->ev_used == h->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
    return h->ev_pool[h->ev_used++];
  }
  ProfScope(avb_handle h_, int cat_, cudaStream_t st_) : h(h_), st(st_), cat(cat_) {
    if (h && h->prof) { a = take(h); b = take(h); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (a) { cudaEventRecord(b, st); h->prof_recs.push_back({cat, a, b}); }
  }
};

// Every entry point that takes a handle runs on the handle's device (one process may hold one handle per GPU):
// switch to it for the duration of the call and restore the caller's current device afterwards.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define AVB_CUDA(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail((h), AVB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define AVB_LAUNCH_CHECK(h)                                                                        \
  do {                                                                                             \
    (h)->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return fail((h), AVB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ Haiku parameter names
std::string sfx(int i) { return i == 0 ? std::string() : "_" + std::to_string(i); }

// ------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, cols] bf16 row-major, box [box_rows x 64 cols], 128B swizzle
// fp32 [rows x cols] matrix in [box_rows x 32 cols] boxes (128-byte rows, 128B swizzle): the z tiles the FFN stores
bool make_tmap_2d_f32(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_tmap_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ kernel launchers
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: every launcher remembers which devices
// it has prepared (one process may hold one handle per GPU).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <class F> cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// Debug / A-B switches read from the environment exist only in -DAVB_AB_BUILD builds (scripts/ab_run.sh); the
// production library never calls getenv.
inline const char* ab_env(const char* name) {
#ifdef AVB_AB_BUILD
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// LayerNorm + QKV projection, CTA pairs (cluster of 2, cta_group::2); tm_b must have [128 x 64] boxes.
// run_if: see ln_gemm2_tc_kernel.
int launch_ln_gemm2_tc(avb_handle h, const float* z, const CUtensorMap& tm_b, const float* bias, void* out, int ldo,
                       int M, int N, cudaStream_t st, int num_sms, long long* launches, int seq_n, int seq_d, int seq_axis,
                       const int* run_if = nullptr, int seq_dg = 0, float* norms = nullptr) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    return cudaFuncSetAttribute(avb::ln_gemm2_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kLnGemmSmemBytes);
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(ln_gemm2_tc): %s", cudaGetErrorString(e));
  const int num_pairs = ((M + 127) / 128 + 1) / 2;
  const int grid = 2 * std::min(num_pairs, num_sms / 2);
  avb::ln_gemm2_tc_kernel<2><<<grid, avb::kLnGemmThreads, avb::kLnGemmSmemBytes, st>>>(
      z, tm_b, bias, reinterpret_cast<bf16*>(out), ldo, M, N, seq_n, seq_d, seq_axis, seq_dg, run_if, norms);
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "ln_gemm2_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// delta[token][128] (bf16) = attn . Wo^T + bo
int launch_outproj_tc(avb_handle h, const void* attn, const CUtensorMap& tm_w, const float* bias, void* delta, int M,
                      int n, int d, int axis, cudaStream_t st, int num_sms, long long* launches, const int* run_if = nullptr,
                      int seq_dg = 0) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    return cudaFuncSetAttribute(avb::outproj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kOutProjSmemBytes);
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(outproj_tc): %s", cudaGetErrorString(e));
  const int grid = std::min((M + 127) / 128, num_sms);
  avb::outproj_tc_kernel<<<grid, avb::kOutProjThreads, avb::kOutProjSmemBytes, st>>>(
      reinterpret_cast<const bf16*>(attn), tm_w, bias, reinterpret_cast<bf16*>(delta), M, n, d, axis, seq_dg, run_if);
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "outproj_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// CTA-pair (cluster of 2, cta_group::2) fused FFN; tensor maps with [64 x 64] boxes.  The resident-weight variant
// serves the published width (F = 512, four resident chunks), the weight-ring variant every other multiple of 256.
int launch_ffn2_tc(avb_handle h, float* z, const CUtensorMap& tm_w1q, const CUtensorMap& tm_w2q, const float* b1,
                   const float* b2, int M, int F, cudaStream_t st, int num_sms, long long* launches,
                   const void* delta = nullptr, int n = 0, int d = 0, int axis = 0, const avb::FfnBias* host_bias = nullptr) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    cudaError_t r = cudaFuncSetAttribute(avb::ffn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kFfnSmemBytes);
    if (r == cudaSuccess)
      r = cudaFuncSetAttribute(avb::ffn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kFfnSmemBytes);
    return r;
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(ffn2_tc): %s", cudaGetErrorString(e));
  bool use_resident = F == 512;
#if !AVB_DELTA_TOKENS
  use_resident = false;  // the resident version reads delta next to z: token order only
#endif
  if (const char* r = ab_env("AVB_FFN_RESIDENT")) use_resident = use_resident && r[0] != '0';
  const int num_pairs = ((M + 127) / 128 + 1) / 2;
  const int grid = 2 * std::min(num_pairs, num_sms / 2);
  CUtensorMap tm_z;  // the residual stream of this call, for the final epilogue's tensor-map stores
  if (!make_tmap_2d_f32(&tm_z, z, (uint64_t)M, 128, 32)) return fail(h, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for z");
  static const avb::FfnBias zero_bias{};
  avb::FfnBias fetched;
  if (use_resident && !host_bias) {  // test hook: the biases live on the device only
    if (cudaMemcpy(fetched.b1, b1, sizeof fetched.b1, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(fetched.b2, b2, sizeof fetched.b2, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(h, AVB_ERR_CUDA, "ffn2_tc: fetching the biases failed");
    host_bias = &fetched;
  }
  if (use_resident)
    avb::ffn2_tc_kernel<true><<<grid, avb::kFfnThreads, avb::kFfnSmemBytes, st>>>(z, reinterpret_cast<const bf16*>(delta), n, d, axis,
                                                                                   tm_z, tm_w1q, tm_w2q, b1, b2, M, F, *host_bias);
  else
    avb::ffn2_tc_kernel<false><<<grid, avb::kFfnThreads, avb::kFfnSmemBytes, st>>>(z, reinterpret_cast<const bf16*>(delta), n, d, axis,
                                                                                    tm_z, tm_w1q, tm_w2q, b1, b2, M, F, zero_bias);
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "ffn2_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// Axial attention = two launches on the same stream: the fast kernel (no running maximum) and the exact kernel, which
// exits at once unless the fast one raised `guard` (scores outside +-100 log2 units; never seen on LayerNorm-ed inputs,
// but the result must not depend on that).  guard: 4 bytes of device memory owned by the caller's workspace, so
// concurrent forwards on different streams / workspaces stay independent.
// mode 0: fast + guarded exact (reset_guard: clear the guard first); 1: the exact kernel alone, unconditionally;
// 3: the exact kernel alone, only if the guard is already up (fallback behind dattn_tc_kernel).
int launch_attention_tc(avb_handle h, const void* qkv, void* out, int B, int n, int d, int H, int axis,
                        cudaStream_t st, int num_sms, long long* launches, int* guard, int mode = 0,
                        const float* norms = nullptr) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    cudaError_t r = cudaFuncSetAttribute(avb::attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kAttnSmemBytes);
    if (r == cudaSuccess)
      r = cudaFuncSetAttribute(avb::attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kAttnSmemBytes);
    return r;
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e));
  if (mode == 0) {
    if (const char* m = ab_env("AVB_ATTN_MODE")) mode = !strcmp(m, "exact") ? 1 : (!strcmp(m, "fast") ? 2 : 0);
  }
  const int T = axis == 0 ? d : n;
  const long long S = axis == 0 ? (long long)B * n : (long long)B * d;
  const long long items = S * H * ((T + 127) / 128);
  if (items >= (1LL << 31)) return fail(h, AVB_ERR_INVALID, "attention: too many (sequence, head, tile) items in one call (%lld)", items);
  // a stream owns whole (sequence, head) pairs and walks their query tiles itself; with fewer than ~12 pairs per stream
  // the pairs are cut into qsplit parts of query tiles (a divisor of their number) so that the last round fills up
  const int nqt = (T + 127) / 128;
  int qsplit = 1;
  for (int cand = 2; cand <= nqt && S * H * qsplit < 12LL * avb::kAttnStreams * num_sms; ++cand)
    if (nqt % cand == 0) qsplit = cand;
  const int grid = (int)std::min<long long>((S * H * qsplit + avb::kAttnStreams - 1) / avb::kAttnStreams, num_sms);
  if (mode == 0 || mode == 2) {
    if (cudaMemsetAsync(guard, 0, sizeof(int), st) != cudaSuccess) return fail(h, AVB_ERR_CUDA, "attention guard reset failed");
    avb::attention_tc_kernel<true><<<grid, avb::kAttnThreads, avb::kAttnSmemBytes, st>>>(
        reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, n, d, H, axis, guard, norms, qsplit);
    if (launches) ++*launches;
  }
  if (mode != 2) {
    avb::attention_tc_kernel<false><<<grid, avb::kAttnThreads, avb::kAttnSmemBytes, st>>>(
        reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, n, d, H, axis, mode == 1 ? nullptr : guard, nullptr, qsplit);
    if (launches) ++*launches;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "attention_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// Fused d-axis attention sub-layer (dattn_tc_kernel): one CTA per sequence of T = d <= 128 tokens.
// fuse_out: delta rows (bf16, token order) come out of the kernel; otherwise the attention output in the
// out-projection's A-operand layout.  `guard` must have been cleared by the caller.
int launch_dattn_tc(avb_handle h, const float* z, const Block& W, void* out, int S, int T, bool fuse_out, cudaStream_t st,
                    int num_sms, long long* launches, int* guard) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    cudaError_t r = cudaFuncSetAttribute(avb::dattn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         avb::DaCfg<true>::kSmemBytes);
    if (r == cudaSuccess)
      r = cudaFuncSetAttribute(avb::dattn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::DaCfg<true>::kSmemBytes);
    if (r == cudaSuccess)
      r = cudaFuncSetAttribute(avb::dattn_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::DaCfg<false>::kSmemBytes);
    return r;
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(dattn_tc): %s", cudaGetErrorString(e));
  if (T < 1 || T > 128 || S < 1) return fail(h, AVB_ERR_INVALID, "dattn_tc: sequence length %d not in 1..128", T);
  // sequences of T <= 64 tokens are packed g = 128 / T to a tile (kPack): the delta rows are then {128 channels, g T rows,
  // tiles}, and a last tile with fewer sequences gets a map of its own that clips at its own row count
  const bool pack = fuse_out && T <= 64 && h->dattn_pack;
  const int g = pack ? 128 / T : 1, Tt = g * T;
  const int ntiles = (S + g - 1) / g, nfull = S / g, rem = S - nfull * g;
  const int grid = std::min(ntiles, num_sms);
  CUtensorMap tm_delta, tm_last;
  memset(&tm_delta, 0, sizeof tm_delta);
  memset(&tm_last, 0, sizeof tm_last);
  if (fuse_out) {
    // delta rows as {128 channels, rows of a tile, tiles} bf16, stored in [32 ch x 32 rows] boxes (64B swizzle): the row
    // bound of the map clips what a warp's box holds past the end of its tile
    EncodeTiledFn fn = encode_tiled_fn();
    auto encode = [&](CUtensorMap* tm, void* base, int rows, int tiles) {
      cuuint64_t gdim[3] = {128, (cuuint64_t)rows, (cuuint64_t)tiles};
      cuuint64_t gstr[2] = {256, (cuuint64_t)Tt * 256};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      return fn && fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = encode(&tm_delta, out, Tt, std::max(nfull, 1));
    if (ok && rem > 0) ok = encode(&tm_last, reinterpret_cast<bf16*>(out) + (size_t)nfull * Tt * 128, rem * T, 1);
    if (!ok) return fail(h, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the delta rows");
    if (pack)
      avb::dattn_tc_kernel<true, true><<<grid, avb::kDaThreads, avb::DaCfg<true>::kSmemBytes, st>>>(
          z, W.wd_qkv, W.wd_o, W.bqkv_f, W.bo_d, reinterpret_cast<bf16*>(out), S, T, guard, tm_delta, W.dattn_bias, tm_last);
    else
      avb::dattn_tc_kernel<true, false><<<grid, avb::kDaThreads, avb::DaCfg<true>::kSmemBytes, st>>>(
          z, W.wd_qkv, W.wd_o, W.bqkv_f, W.bo_d, reinterpret_cast<bf16*>(out), S, T, guard, tm_delta, W.dattn_bias, tm_last);
  } else {
    avb::dattn_tc_kernel<false, false><<<grid, avb::kDaThreads, avb::DaCfg<false>::kSmemBytes, st>>>(
        z, W.wd_qkv, W.wd_o, W.bqkv_f, W.bo_d, reinterpret_cast<bf16*>(out), S, T, guard, tm_delta, W.dattn_bias, tm_last);
  }
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "dattn_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// C = epi(A' W + bias) with fp32 operands split into bf16 pairs (kernels_split.cuh); K = 64 * KB columns of A starting at A,
// weight columns [k_off, k_off + K) of the pre-split W^T behind tm[0] (hi) / tm[1] (lo).
template <int KB, bool LN, int EPI>
int launch_gemm_split(avb_handle h, const float* A, int lda, const CUtensorMap* tm, int k_off, const float* bias, const float* R,
                      int ldr, float* C, int ldc, int M, int N, cudaStream_t st, avb::SplitSeq sq = avb::SplitSeq{0, 0, 0, 0, 0},
                      float* norms = nullptr) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    return cudaFuncSetAttribute(avb::gemm_split_tc_kernel<KB, LN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                avb::SplitCfg<KB>::kSmemBytes);
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(gemm_split_tc): %s", cudaGetErrorString(e));
  const int grid = std::min((M + 127) / 128, h->num_sms);
  avb::gemm_split_tc_kernel<KB, LN, EPI><<<grid, avb::kSplitThreads, avb::SplitCfg<KB>::kSmemBytes, st>>>(
      A, lda, tm[0], tm[1], k_off, bias, R, ldr, C, ldc, M, N, sq, norms);
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "gemm_split_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

// fp32-mode attention on the tensor cores (attention_split_tc_kernel): operands from the split QKV projection
int launch_attention_split(avb_handle h, const void* qkvs, float* attn, int B, int n, int d, int H, int axis, int dg, cudaStream_t st,
                           const float* norms) {
  static PerDeviceOnce once;
  cudaError_t e = once.run([] {
    return cudaFuncSetAttribute(avb::attention_split_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kSplitAttnSmemBytes);
  });
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(attention_split_tc): %s", cudaGetErrorString(e));
  const int T = axis == 0 ? d : n;
  const long long S = axis == 0 ? (long long)B * n : (long long)B * d;
  const long long items = S * H * ((T + 127) / 128);
  if (items >= (1LL << 31)) return fail(h, AVB_ERR_INVALID, "attention: too many (sequence, head, tile) items in one call (%lld)", items);
  const int grid = (int)std::min<long long>(items, 2LL * h->num_sms);
  avb::attention_split_tc_kernel<<<grid, avb::kSplitAttnThreads, avb::kSplitAttnSmemBytes, st>>>(
      reinterpret_cast<const bf16*>(qkvs), attn, B, n, d, H, axis, dg, norms);
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "attention_split_tc launch: %s", cudaGetErrorString(e));
  return AVB_OK;
}

template <int EPI, bool LN>
void launch_sgemm(const float* A, int lda, const float* W, int ldw, const float* bias, const float* g,
                  const float* b, const float* R, int ldr, float* C, int ldc, size_t M, int N, int K,
                  cudaStream_t st) {
  dim3 grid((unsigned)((M + 63) / 64), N / 64);
  avb::sgemm_kernel<EPI, LN><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, g, b, R, ldr, C, ldc, M, N, K);
}

// ------------------------------------------------------------------ workspace carving
struct BlockWs {
  // bf16 mode: big (qkv bf16[N,768] aliased with h bf16[N,512]) | attn bf16[N,256]
  // f32 mode:  big (qkv f32[N,768] aliased with h f32[N,512]) | attn f32[N,256]
  size_t xhat_off, big_off, attn_off, guard_off, total;
};
BlockWs block_ws_layout(const avb_config& c, size_t ntok) {
  BlockWs w{};
  const size_t HD = (size_t)c.num_heads * c.key_size, F = (size_t)c.widening_factor * c.dim;
  size_t off = 0;
  if (c.compute_dtype == AVB_DTYPE_BF16) {
    w.xhat_off = off;  // (the LayerNorm output lives only in shared memory)
    w.big_off = off;  off = align_up(off + ntok * std::max(3 * HD, F) * 2, 1024);
    w.attn_off = off; off = align_up(off + align_up(ntok, 128) * HD * 2, 1024);  // whole 128-row tiles (A-operand layout)
    w.guard_off = off; off += 1024;  // overflow guard of the fast attention kernel (launch_attention_tc)
  } else {
    w.big_off = off;  off = align_up(off + ntok * std::max(3 * HD, F) * 4, 1024);
    w.attn_off = off; off = align_up(off + ntok * HD * 4, 1024);
    w.guard_off = off; off += 1024;  // score bound of the split-precision attention
  }
  w.total = off;
  return w;
}

struct FwdWs {
  size_t xs_off, xtmp_off, iv_off, stat_off, pooled_off, pooled_c_off, uv_off, z_off, blk_off, total;
  int Bc;  // datasets (after chunk expansion) processed per pass
};
constexpr size_t kWsBudget = (size_t)40 << 30;  // cap on the per-pass activation working set

FwdWs fwd_ws_layout(const avb_config& c, int B, int n, int d, int chunks) {
  FwdWs w{};
  const int Be = B * chunks, ne = n / chunks;
  const size_t cells = (size_t)B * n * d;
  size_t off = 0;
  w.xs_off = off; off = align_up(off + cells * 4, 1024);
  w.xtmp_off = off; off = align_up(off + (chunks > 1 ? cells * 4 : 0), 1024);  // standardizer output before the chunk gather
  w.iv_off = off; off = align_up(off + (chunks > 1 ? cells * 4 : 0), 1024);
  w.stat_off = off; off = align_up(off + (size_t)B * (n + d + 8) * 8, 1024);
  w.pooled_off = off; off = align_up(off + (size_t)B * d * c.dim * 4, 1024);
  w.pooled_c_off = off; off = align_up(off + (chunks > 1 ? (size_t)Be * d * c.dim * 4 : 0), 1024);
  w.uv_off = off; off = align_up(off + (size_t)2 * B * d * c.out_dim * 4, 1024);
  const size_t tok1 = (size_t)ne * d;
  const size_t per_ds = tok1 * c.dim * 4 + block_ws_layout(c, tok1).total + 4096;
  int Bc = (int)std::max<size_t>(1, std::min<size_t>((size_t)Be, kWsBudget / per_ds));
  w.Bc = Bc;
  w.z_off = off; off = align_up(off + (size_t)Bc * tok1 * c.dim * 4, 1024);
  w.blk_off = off; off += block_ws_layout(c, (size_t)Bc * tok1).total;
  w.total = off;
  return w;
}

// rows i -> (chunk i % chunks, position i / chunks)   (model.py:105-106)
__global__ void chunk_gather_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int n, int d,
                                    int chunks) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * n * d;
  if (idx >= total) return;
  const int j = (int)(idx % d);
  const size_t bi = idx / d;
  const int i = (int)(bi % n);
  const size_t b = bi / n;
  const int cs = n / chunks;
  const int c = i % chunks, r = i / chunks;
  dst[(((b * chunks + c) * cs) + r) * d + j] = src[idx];
}
__global__ void chunk_max_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t per_ds, int chunks,
                                 size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t b = idx / per_ds, e = idx % per_ds;
  float m = -INFINITY;
  for (int c = 0; c < chunks; ++c) m = fmaxf(m, src[(b * chunks + c) * per_ds + e]);
  dst[idx] = m;
}

bool shape_ok(avb_handle h, int B, int n, int d) { return h && B >= 1 && n >= 1 && d >= 1; }

}  // namespace

// ====================================================================================== C ABI
extern "C" {

const char* avb_version(void) { return "avici_b200 0.2.0 sm_100a"; }

const char* avb_last_error(avb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int avb_create(avb_handle* out, const avb_config* cfg) {
  if (!out || !cfg) return fail(nullptr, AVB_ERR_INVALID, "avb_create: null argument");
  *out = nullptr;
  if (cfg->dim != 128) return fail(nullptr, AVB_ERR_INVALID, "unsupported dim=%d (kernels are built for dim=128)", cfg->dim);
  if (cfg->key_size != 32) return fail(nullptr, AVB_ERR_INVALID, "unsupported key_size=%d (kernels are built for 32)", cfg->key_size);
  if (cfg->num_heads < 1 || cfg->num_heads > 16 || (cfg->num_heads * cfg->key_size) % 64 != 0)
    return fail(nullptr, AVB_ERR_INVALID, "unsupported num_heads=%d", cfg->num_heads);
  if (cfg->widening_factor < 1 || (cfg->widening_factor * cfg->dim) % 256 != 0)
    return fail(nullptr, AVB_ERR_INVALID, "unsupported widening_factor=%d (FFN width must be a multiple of 256)", cfg->widening_factor);
  if (cfg->layers < 1) return fail(nullptr, AVB_ERR_INVALID, "layers must be >= 1");
  if (cfg->out_dim < 1 || cfg->out_dim > 128) return fail(nullptr, AVB_ERR_INVALID, "unsupported out_dim=%d (1..128)", cfg->out_dim);
  if (cfg->compute_dtype != AVB_DTYPE_F32 && cfg->compute_dtype != AVB_DTYPE_BF16)
    return fail(nullptr, AVB_ERR_INVALID, "unknown compute_dtype=%d", cfg->compute_dtype);
  if (cfg->compute_dtype == AVB_DTYPE_BF16 && cfg->num_heads != 2 * avb::kOutProjKB)
    return fail(nullptr, AVB_ERR_INVALID, "the bf16 tensor-core path is built for num_heads=%d (got %d); use compute_dtype fp32",
                2 * avb::kOutProjKB, cfg->num_heads);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, AVB_ERR_CUDA, "device sm_%d%d is not Blackwell sm_100 (kernels are sm_100a only)", prop.major, prop.minor);
  avb_handle h = new avb_model_s();
  h->cfg = *cfg;
  h->device = dev;
  h->num_sms = prop.multiProcessorCount;
  if (const char* m = ab_env("AVB_DATTN_MODE")) h->dattn_mode = m[0] == '0' ? 0 : (m[0] == '2' ? 2 : 1);
  e = cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return fail(nullptr, AVB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  *out = h;
  return AVB_OK;
}

void avb_destroy(avb_handle h) {
  if (!h) return;
  DeviceGuard dg(h->device);
  if (h->comm) nccl_api().CommDestroy(h->comm);
  if (h->arena) cudaFree(h->arena);
  if (h->host_ws) cudaFree(h->host_ws);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  delete h;
}

int64_t avb_param_count(avb_handle h) {
  if (!h) return -1;
  const avb_config& c = h->cfg;
  const int64_t k = c.dim, HD = (int64_t)c.num_heads * c.key_size, F = (int64_t)c.widening_factor * c.dim;
  const int64_t per_block = 4 * 2 * k + 3 * (k * HD + HD) + (HD * k + k) + (k * F + F) + (F * k + k);
  return 2 * k + k + 2 * (int64_t)c.layers * per_block + 2 * k + 2 * (2 * k + k * c.out_dim + c.out_dim) + 2;
}

int64_t avb_launch_count(avb_handle h) { return h ? h->launches : -1; }

int avb_profile_enable(avb_handle h, int32_t on) {
  if (!h) return AVB_ERR_INVALID;
  h->prof = on != 0;
  return AVB_OK;
}

int avb_profile_read(avb_handle h, double* ms_out, int64_t* count_out, int32_t reset) {
  if (!h) return AVB_ERR_INVALID;
  DeviceGuard dg(h->device);
  AVB_CUDA(h, cudaDeviceSynchronize());
  for (const auto& r : h->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { h->prof_ms[r.cat] += ms; h->prof_n[r.cat]++; }
  }
  h->prof_recs.clear();
  h->ev_used = 0;
  for (int i = 0; i < AVB_PROF_CATEGORIES; ++i) {
    if (ms_out) ms_out[i] = h->prof_ms[i];
    if (count_out) count_out[i] = h->prof_n[i];
    if (reset) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  }
  return AVB_OK;
}

int avb_load_weights(avb_handle h, const avb_tensor_desc* tensors, int32_t count) {
  if (!h || !tensors) return fail(h, AVB_ERR_INVALID, "avb_load_weights: null argument");
  DeviceGuard dg(h->device);
  const avb_config& c = h->cfg;
  const int k = c.dim, HD = c.num_heads * c.key_size, F = c.widening_factor * c.dim, nb = 2 * c.layers;
  std::map<std::string, std::pair<const float*, int64_t>> tab;
  for (int i = 0; i < count; ++i) tab[tensors[i].name] = {tensors[i].data, tensors[i].numel};
  std::string missing;
  auto get = [&](const std::string& name, int64_t numel) -> const float* {
    auto it = tab.find(name);
    if (it == tab.end() || it->second.second != numel || !it->second.first) {
      if (missing.empty()) missing = name + (it == tab.end() ? " (absent)" : " (wrong size)");
      return nullptr;
    }
    return it->second.first;
  };

  // host staging: one fp32 arena and one bf16 arena, then a single upload each
  std::vector<float> f32;
  std::vector<bf16> b16;
  f32.reserve((size_t)avb_param_count(h) + (size_t)nb * (3 * HD + F) + 1024);
  auto push_f32 = [&](const float* src, size_t nel) -> size_t {
    size_t off = align_up(f32.size(), 64);
    f32.resize(off + nel);
    if (src) memcpy(f32.data() + off, src, nel * sizeof(float));
    return off;
  };
  auto push_b16 = [&](size_t nel) -> size_t {
    size_t off = align_up(b16.size(), 512);
    b16.resize(off + nel);
    return off;
  };
  struct BlockOff {
    size_t ln_g[4], ln_b[4], wq, bq, wk, bk, wv, bv, wo, bo, w1, b1, w2, b2, bqkv_f, b1_f, bo_d, zero_k;
    size_t s_qkv[2], s_o[2], s_1[2], s_2[2];
    size_t wqkv_t, wo_t, w1_t, w2_t, wd_qkv, wd_o;
  };
  std::vector<BlockOff> offs(nb);
  const std::string R = "BaseModel/";
  const size_t emb_w = push_f32(get(R + "linear/w", 2 * k), 2 * k);
  const size_t emb_b = push_f32(get(R + "linear/b", k), k);
  const double qscale = 1.4426950408889634 / std::sqrt((double)c.key_size);  // log2(e)/sqrt(key_size)
  for (int i = 0; i < nb; ++i) {
    BlockOff& o = offs[i];
    const float *g[4], *b[4];
    for (int t = 0; t < 4; ++t) {
      const std::string ln = R + "layer_norm" + sfx(4 * i + t);
      g[t] = get(ln + "/scale", k); b[t] = get(ln + "/offset", k);
      o.ln_g[t] = push_f32(g[t], k); o.ln_b[t] = push_f32(b[t], k);
    }
    const std::string mha = R + "multi_head_attention" + sfx(i);
    const float* wq = get(mha + "/query/w", (int64_t)k * HD); const float* bq = get(mha + "/query/b", HD);
    const float* wk = get(mha + "/key/w", (int64_t)k * HD);   const float* bk = get(mha + "/key/b", HD);
    const float* wv = get(mha + "/value/w", (int64_t)k * HD); const float* bv = get(mha + "/value/b", HD);
    const float* wo = get(mha + "/linear/w", (int64_t)HD * k); const float* bo = get(mha + "/linear/b", k);
    const float* w1 = get(R + "linear" + sfx(2 * i + 1) + "/w", (int64_t)k * F);
    const float* b1 = get(R + "linear" + sfx(2 * i + 1) + "/b", F);
    const float* w2 = get(R + "linear" + sfx(2 * i + 2) + "/w", (int64_t)F * k);
    const float* b2 = get(R + "linear" + sfx(2 * i + 2) + "/b", k);
    o.wq = push_f32(wq, (size_t)k * HD); o.bq = push_f32(bq, HD);
    o.wk = push_f32(wk, (size_t)k * HD); o.bk = push_f32(bk, HD);
    o.wv = push_f32(wv, (size_t)k * HD); o.bv = push_f32(bv, HD);
    o.wo = push_f32(wo, (size_t)HD * k); o.bo = push_f32(bo, k);
    o.w1 = push_f32(w1, (size_t)k * F);  o.b1 = push_f32(b1, F);
    o.w2 = push_f32(w2, (size_t)F * k);  o.b2 = push_f32(b2, k);
    o.bqkv_f = push_f32(nullptr, 3 * HD);
    o.b1_f = push_f32(nullptr, F);
    o.bo_d = push_f32(nullptr, k);
    o.zero_k = push_f32(nullptr, k);
    if (!missing.empty()) continue;
    if (c.compute_dtype == AVB_DTYPE_F32) {
      // split-precision operands of gemm_split_tc_kernel: W' = diag(gamma) W in fp64 -> fp32 -> (hi, lo) bf16, K-major [N, K]
      auto split_store = [&](size_t hi_off, size_t lo_off, size_t idx, double w) {
        const float wf = (float)w;
        const bf16 hi = __float2bfloat16(wf);
        b16[hi_off + idx] = hi;
        b16[lo_off + idx] = __float2bfloat16(wf - __bfloat162float(hi));
      };
      for (int p = 0; p < 2; ++p) {
        o.s_qkv[p] = push_b16((size_t)3 * HD * k); o.s_o[p] = push_b16((size_t)k * HD);
        o.s_1[p] = push_b16((size_t)F * k);        o.s_2[p] = push_b16((size_t)k * F);
      }
      const float* ws[3] = {wq, wk, wv}; const float* bs[3] = {bq, bk, bv};
      for (int t = 0; t < 3; ++t) {
        const double sc = t == 0 ? qscale : 1.0;  // q carries log2e / sqrt(key): the split attention kernel uses exp2
        for (int col = 0; col < HD; ++col) {
          double acc = bs[t][col];
          for (int r = 0; r < k; ++r) {
            acc += (double)b[t][r] * ws[t][(size_t)r * HD + col];
            split_store(o.s_qkv[0], o.s_qkv[1], ((size_t)t * HD + col) * k + r, (double)g[t][r] * ws[t][(size_t)r * HD + col] * sc);
          }
          f32[o.bqkv_f + (size_t)t * HD + col] = (float)(acc * sc);
        }
      }
      for (int col = 0; col < k; ++col)
        for (int r = 0; r < HD; ++r) split_store(o.s_o[0], o.s_o[1], (size_t)col * HD + r, wo[(size_t)r * k + col]);
      for (int col = 0; col < F; ++col) {
        double acc = b1[col];
        for (int r = 0; r < k; ++r) {
          acc += (double)b[3][r] * w1[(size_t)r * F + col];
          split_store(o.s_1[0], o.s_1[1], (size_t)col * k + r, (double)g[3][r] * w1[(size_t)r * F + col]);
        }
        f32[o.b1_f + col] = (float)acc;
      }
      for (int col = 0; col < k; ++col)
        for (int r = 0; r < F; ++r) split_store(o.s_2[0], o.s_2[1], (size_t)col * F + r, w2[(size_t)r * k + col]);
    }
    if (c.compute_dtype == AVB_DTYPE_BF16) {
      // LN(x) @ W + b = xhat @ (diag(gamma) W) + (beta @ W + b); q additionally carries log2e/sqrt(key)
      o.wqkv_t = push_b16((size_t)3 * HD * k);
      const float* ws[3] = {wq, wk, wv}; const float* bs[3] = {bq, bk, bv};
      for (int t = 0; t < 3; ++t) {
        const double sc = t == 0 ? qscale : 1.0;
        for (int col = 0; col < HD; ++col) {
          double acc = bs[t][col];
          for (int r = 0; r < k; ++r) {
            acc += (double)b[t][r] * ws[t][(size_t)r * HD + col];
            b16[o.wqkv_t + ((size_t)t * HD + col) * k + r] =
                __float2bfloat16((float)((double)g[t][r] * ws[t][(size_t)r * HD + col] * sc));
          }
          f32[o.bqkv_f + (size_t)t * HD + col] = (float)(acc * sc);
        }
      }
      o.wo_t = push_b16((size_t)k * HD);
      for (int col = 0; col < k; ++col) {
        double acc = bo[col];
        for (int r = 0; r < HD; ++r) {
          const bf16 wb = __float2bfloat16(wo[(size_t)r * k + col]);
          b16[o.wo_t + (size_t)col * HD + r] = wb;
          acc += (double)f32[o.bqkv_f + (size_t)2 * HD + r] * (double)__bfloat162float(wb);
        }
        f32[o.bo_d + col] = (float)acc;
      }
      o.w1_t = push_b16((size_t)F * k);
      for (int col = 0; col < F; ++col) {
        double acc = b1[col];
        for (int r = 0; r < k; ++r) {
          acc += (double)b[3][r] * w1[(size_t)r * F + col];
          b16[o.w1_t + (size_t)col * k + r] = __float2bfloat16((float)((double)g[3][r] * w1[(size_t)r * F + col]));
        }
        f32[o.b1_f + col] = (float)acc;
      }
      o.w2_t = push_b16((size_t)k * F);
      for (int col = 0; col < k; ++col)
        for (int r = 0; r < F; ++r) b16[o.w2_t + (size_t)col * F + r] = __float2bfloat16(w2[(size_t)r * k + col]);
      // dattn_tc_kernel operand images (byte-exact shared-memory layouts, fetched with 1-D bulk copies)
      if (c.num_heads == avb::kDaHeads) {
        const size_t img_qkv = (size_t)avb::kDaHeads * 2 * avb::kDaWkbBytes / 2, img_o = (size_t)avb::kDaHeads * avb::kDaWoHeadBytes / 2;
        o.wd_qkv = push_b16(img_qkv);
        o.wd_o = push_b16(img_o);
        for (int hh = 0; hh < avb::kDaHeads; ++hh) {
          for (int kb = 0; kb < 2; ++kb)
            for (int r = 0; r < 96; ++r) {  // row r = (q|k|v, channel jj of the head)
              const int w = r >> 5, jj = r & 31;
              const bf16* src = &b16[o.wqkv_t + ((size_t)w * HD + hh * 32 + jj) * k + kb * 64];
              for (int ch = 0; ch < 8; ++ch) {  // 16-byte chunk ch -> slot ch ^ (r & 7) of the 128-byte row
                const size_t dst = o.wd_qkv + ((size_t)(hh * 2 + kb) * avb::kDaWkbBytes + (size_t)(r >> 3) * 1024 + (r & 7) * 128 +
                                               ((ch ^ (r & 7)) << 4)) / 2;
                for (int e = 0; e < 8; ++e) b16[dst + e] = src[ch * 8 + e];
              }
            }
          for (int r = 0; r < 128; ++r) {  // row r = output channel; 32 head channels = 64 bytes, chunk ch -> ch ^ ((r >> 1) & 3)
            const bf16* src = &b16[o.wo_t + (size_t)r * HD + hh * 32];
            for (int ch = 0; ch < 4; ++ch) {
              const size_t dst = o.wd_o + ((size_t)hh * avb::kDaWoHeadBytes + (size_t)r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)) / 2;
              for (int e = 0; e < 8; ++e) b16[dst + e] = src[ch * 8 + e];
            }
          }
        }
      }
    }
  }
  const size_t lnf_g = push_f32(get(R + "layer_norm" + sfx(4 * nb) + "/scale", k), k);
  const size_t lnf_b = push_f32(get(R + "layer_norm" + sfx(4 * nb) + "/offset", k), k);
  const size_t lnu_g = push_f32(get(R + "layer_norm" + sfx(4 * nb + 1) + "/scale", k), k);
  const size_t lnu_b = push_f32(get(R + "layer_norm" + sfx(4 * nb + 1) + "/offset", k), k);
  const size_t wu = push_f32(get(R + "linear" + sfx(2 * nb + 1) + "/w", (int64_t)k * c.out_dim), (size_t)k * c.out_dim);
  const size_t bu = push_f32(get(R + "linear" + sfx(2 * nb + 1) + "/b", c.out_dim), c.out_dim);
  const size_t lnv_g = push_f32(get(R + "layer_norm" + sfx(4 * nb + 2) + "/scale", k), k);
  const size_t lnv_b = push_f32(get(R + "layer_norm" + sfx(4 * nb + 2) + "/offset", k), k);
  const size_t wv = push_f32(get(R + "linear" + sfx(2 * nb + 2) + "/w", (int64_t)k * c.out_dim), (size_t)k * c.out_dim);
  const size_t bv = push_f32(get(R + "linear" + sfx(2 * nb + 2) + "/b", c.out_dim), c.out_dim);
  const size_t temp = push_f32(get("BaseModel/learned_temp", 1), 1);
  const size_t bias = push_f32(get("BaseModel/final_matrix_bias", 1), 1);
  if (!missing.empty()) return fail(h, AVB_ERR_MISSING, "weight tensor missing or mis-sized: %s", missing.c_str());

  const size_t f32_bytes = align_up(f32.size() * 4, 1024), b16_bytes = align_up(b16.size() * 2, 1024);
  if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
  AVB_CUDA(h, cudaMalloc(&h->arena, f32_bytes + b16_bytes + 1024));
  float* df = reinterpret_cast<float*>(h->arena);
  bf16* db = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(h->arena) + f32_bytes);
  AVB_CUDA(h, cudaMemcpy(df, f32.data(), f32.size() * 4, cudaMemcpyHostToDevice));
  if (!b16.empty()) AVB_CUDA(h, cudaMemcpy(db, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));

  h->emb_w = df + emb_w; h->emb_b = df + emb_b;
  h->lnf_g = df + lnf_g; h->lnf_b = df + lnf_b;
  h->lnu_g = df + lnu_g; h->lnu_b = df + lnu_b; h->wu = df + wu; h->bu = df + bu;
  h->lnv_g = df + lnv_g; h->lnv_b = df + lnv_b; h->wv = df + wv; h->bv = df + bv;
  h->temp = df + temp; h->bias = df + bias;
  h->blocks.assign(nb, Block{});
  for (int i = 0; i < nb; ++i) {
    Block& B = h->blocks[i]; const BlockOff& o = offs[i];
    for (int t = 0; t < 4; ++t) { B.ln_g[t] = df + o.ln_g[t]; B.ln_b[t] = df + o.ln_b[t]; }
    B.wq = df + o.wq; B.bq = df + o.bq; B.wk = df + o.wk; B.bk = df + o.bk; B.wv = df + o.wv; B.bv = df + o.bv;
    B.wo = df + o.wo; B.bo = df + o.bo; B.w1 = df + o.w1; B.b1 = df + o.b1; B.w2 = df + o.w2; B.b2 = df + o.b2;
    B.bqkv_f = df + o.bqkv_f; B.b1_f = df + o.b1_f; B.bo_d = df + o.bo_d; B.zero_k = df + o.zero_k;
    if (k == 128) memcpy(B.dattn_bias.bo, f32.data() + o.bo_d, sizeof B.dattn_bias.bo);
    if (c.compute_dtype == AVB_DTYPE_BF16 && F == 512) {
      memcpy(B.ffn_bias.b1, f32.data() + o.b1_f, sizeof B.ffn_bias.b1);
      memcpy(B.ffn_bias.b2, f32.data() + o.b2, sizeof B.ffn_bias.b2);
    }
    if (c.compute_dtype == AVB_DTYPE_F32) {
      bool ok = true;
      for (int p = 0; p < 2; ++p)
        ok = ok && make_tmap_2d(&B.tm_s_qkv[p], db + o.s_qkv[p], 3 * HD, k, 128) && make_tmap_2d(&B.tm_s_o[p], db + o.s_o[p], k, HD, 128) &&
             make_tmap_2d(&B.tm_s_1[p], db + o.s_1[p], F, k, 128) && make_tmap_2d(&B.tm_s_2[p], db + o.s_2[p], k, F, 128);
      if (!ok) return fail(h, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for block %d split weights", i);
    }
    if (c.compute_dtype == AVB_DTYPE_BF16) {
      B.wqkv_t = db + o.wqkv_t; B.wo_t = db + o.wo_t; B.w1_t = db + o.w1_t; B.w2_t = db + o.w2_t;
      B.wd_qkv = reinterpret_cast<const uint8_t*>(db + o.wd_qkv); B.wd_o = reinterpret_cast<const uint8_t*>(db + o.wd_o);
      bool ok = make_tmap_2d(&B.tm_wqkv_h, B.wqkv_t, 3 * HD, k, 128) && make_tmap_2d(&B.tm_wo, B.wo_t, k, HD, 128) &&
                make_tmap_2d(&B.tm_w1q, B.w1_t, F, k, 64) && make_tmap_2d(&B.tm_w2q, B.w2_t, k, F, 64);
      if (!ok) return fail(h, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for block %d weights", i);
    }
  }
  h->loaded = true;
  return AVB_OK;
}

// ----------------------------------------------------------------------------------- stages
int avb_standardize(avb_handle h, const float* x, int32_t B, int32_t n, int32_t d, int32_t mode, float* xs,
                    void* scratch, size_t scratch_bytes, void* stream) {
  if (!shape_ok(h, B, n, d) || !x || !xs) return fail(h, AVB_ERR_INVALID, "avb_standardize: bad argument");
  DeviceGuard dg(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)B * n * d;
  ProfScope ps(h, AVB_PROF_PROLOGUE, st);
  if (mode == AVB_STANDARDIZE_NONE) {
    if (xs != x) AVB_CUDA(h, cudaMemcpyAsync(xs, x, total * 4, cudaMemcpyDeviceToDevice, st));
    return AVB_OK;
  }
  if (scratch_bytes < (size_t)B * (n + d + 8) * 8 || !scratch) return fail(h, AVB_ERR_WORKSPACE, "avb_standardize: scratch too small");
  float* sc = reinterpret_cast<float*>(scratch);
  if (mode == AVB_STANDARDIZE_DEFAULT) {
    float* mean = sc; float* stdv = sc + (size_t)B * d;
    avb::zscore_stats_kernel<<<dim3((d + 31) / 32, B), dim3(32, 8), 0, st>>>(x, n, d, mean, stdv);
    AVB_LAUNCH_CHECK(h);
    avb::zscore_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, mean, stdv, n, d, total, xs);
    AVB_LAUNCH_CHECK(h);
  } else if (mode == AVB_STANDARDIZE_COUNT) {
    const size_t rows = (size_t)B * n;
    avb::cpm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, d, xs);
    AVB_LAUNCH_CHECK(h);
    avb::cpm_stats_kernel<<<B, 1024, 0, st>>>(xs, (size_t)n * d, sc);
    AVB_LAUNCH_CHECK(h);
    avb::cpm_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xs, sc, (size_t)n * d, total);
    AVB_LAUNCH_CHECK(h);
  } else {
    return fail(h, AVB_ERR_INVALID, "unknown standardize mode %d", mode);
  }
  return AVB_OK;
}

int avb_embed(avb_handle h, const float* xs, const float* interv, int64_t ntok, float* z, void* stream) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_embed: weights not loaded");
  if (!xs || !z || ntok < 1) return fail(h, AVB_ERR_INVALID, "avb_embed: bad argument");
  DeviceGuard dg(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>((ntok + 7) / 8, (int64_t)h->num_sms * 16);
  ProfScope ps(h, AVB_PROF_PROLOGUE, st);
  avb::embed_kernel<<<grid, 256, 0, st>>>(xs, interv, h->emb_w, h->emb_b, (size_t)ntok, z);
  AVB_LAUNCH_CHECK(h);
  return AVB_OK;
}

size_t avb_block_workspace_bytes(avb_handle h, int32_t B, int32_t n, int32_t d) {
  if (!shape_ok(h, B, n, d)) return 0;
  return block_ws_layout(h->cfg, (size_t)B * n * d).total;
}

// dg > 0: the tokens of the (single-dataset, n-sharded) block are stored in d / dg column blocks (RowMap); axis d only.
static int block_impl(avb_handle h, int32_t bi, float* z, int32_t B, int32_t n, int32_t d, int32_t axis, int32_t dg, void* ws,
                      size_t ws_bytes, void* stream) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_block: weights not loaded");
  if (!shape_ok(h, B, n, d) || !z || (axis != 0 && axis != 1) || bi < 0 || bi >= (int)h->blocks.size())
    return fail(h, AVB_ERR_INVALID, "avb_block: bad argument");
  if (dg != 0 && (dg < 0 || axis != AVB_AXIS_D || B != 1 || d % dg != 0))
    return fail(h, AVB_ERR_INVALID, "avb_block_sharded: column blocks need axis d, one dataset and dg | d (dg=%d, d=%d)", dg, d);
#if !AVB_DELTA_TOKENS
  if (dg != 0) return fail(h, AVB_ERR_INVALID, "avb_block_sharded: not available in -DAVB_DELTA_TOKENS=0 builds");
#endif
  DeviceGuard dev_guard(h->device);
  const avb_config& c = h->cfg;
  const size_t ntok = (size_t)B * n * d;
  if (ntok > (size_t)0x7fffff00) return fail(h, AVB_ERR_INVALID, "avb_block: too many tokens per call (%zu)", ntok);
  const BlockWs L = block_ws_layout(c, ntok);
  if (!ws || ws_bytes < L.total) return fail(h, AVB_ERR_WORKSPACE, "avb_block: workspace %zu < %zu", ws_bytes, L.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  const Block& W = h->blocks[bi];
  const int k = c.dim, H = c.num_heads, HD = H * c.key_size, F = c.widening_factor * c.dim;
  const int M = (int)ntok;

  if (c.compute_dtype == AVB_DTYPE_BF16) {
    bf16* qkv = reinterpret_cast<bf16*>(w8 + L.big_off);
    bf16* attn = reinterpret_cast<bf16*>(w8 + L.attn_off);
    int rc;
    int* guard = reinterpret_cast<int*>(w8 + L.guard_off);
    // attention half: LN -> QKV -> attention -> out-projection, leaving the sub-layer's update as bf16 `delta` rows in
    // the qkv buffer (dead once the attention has run); the FFN's LayerNorm warps add it to z.
    const long long S_d = (long long)B * n;
    if (axis == AVB_AXIS_D && d <= 128 && H == avb::kDaHeads && h->dattn_mode != 0 && dg == 0) {
      // a whole sequence fits one tile: the sub-layer is ONE kernel, qkv and the attention output never reach HBM.
      // The guarded exact fallback (unfused kernels, each returning at once while the guard is down) recomputes delta
      // if a score left the fast softmax's range.
      const bool fuse_out = h->dattn_mode == 1;
      { ProfScope ps(h, AVB_PROF_DATTN, st);
        if (cudaMemsetAsync(guard, 0, sizeof(int), st) != cudaSuccess) return fail(h, AVB_ERR_CUDA, "attention guard reset failed");
        if ((rc = launch_dattn_tc(h, z, W, fuse_out ? (void*)qkv : (void*)attn, (int)S_d, d, fuse_out, st, h->num_sms, &h->launches, guard)))
          return rc;
        if ((rc = launch_ln_gemm2_tc(h, z, W.tm_wqkv_h, W.bqkv_f, qkv, 3 * HD, M, 3 * HD, st, h->num_sms, &h->launches, n, d, axis, guard)))
          return rc;
        if ((rc = launch_attention_tc(h, qkv, attn, B, n, d, H, axis, st, h->num_sms, &h->launches, guard, 3))) return rc;
        if ((rc = launch_outproj_tc(h, attn, W.tm_wo, W.bo, qkv, M, n, d, axis, st, h->num_sms, &h->launches,
                                    fuse_out ? guard : nullptr))) return rc; }
    } else {
    float* norms = reinterpret_cast<float*>(guard) + 1;  // {max |q_h|^2, max |k_h|^2} of this block's projection
    { ProfScope ps(h, AVB_PROF_QKV, st);  // LayerNorm fused into the projection's prologue;
      // rows in sequence-major order of `axis`, output in the attention kernel's operand layout
      if (cudaMemsetAsync(norms, 0, 2 * sizeof(float), st) != cudaSuccess) return fail(h, AVB_ERR_CUDA, "score-bound reset failed");
      if ((rc = launch_ln_gemm2_tc(h, z, W.tm_wqkv_h, W.bqkv_f, qkv, 3 * HD, M, 3 * HD, st, h->num_sms, &h->launches,
                                   n, d, axis, nullptr, dg, norms))) return rc; }
    { ProfScope ps(h, axis == 0 ? AVB_PROF_ATTN_D : AVB_PROF_ATTN_N, st);
      if ((rc = launch_attention_tc(h, qkv, attn, B, n, d, H, axis, st, h->num_sms, &h->launches, guard, 0, norms))) return rc; }
    { ProfScope ps(h, AVB_PROF_OUT, st);
      if ((rc = launch_outproj_tc(h, attn, W.tm_wo, W.bo, qkv, M, n, d, axis, st, h->num_sms, &h->launches, nullptr, dg)))
        return rc; }
    }
    // FFN half: z += delta; LN -> Linear + ReLU -> Linear + residual, one kernel
    { ProfScope ps(h, AVB_PROF_FFN, st);
      if ((rc = launch_ffn2_tc(h, z, W.tm_w1q, W.tm_w2q, W.b1_f, W.b2, M, F, st, h->num_sms, &h->launches, qkv, n, d, axis,
                               F == 512 ? &W.ffn_bias : nullptr)))
        return rc; }
  } else {
    float* qkv = reinterpret_cast<float*>(w8 + L.big_off);
    float* hid = qkv;
    float* attn = reinterpret_cast<float*>(w8 + L.attn_off);
    // projections and FFN on the tensor cores as split-precision products (kernels_split.cuh) when the shape fits
    const bool tc = h->f32_tc && k == 128 && HD == 256 && F % 256 == 0 && F <= 768 && ntok <= (size_t)0x7fffff00;
    int rc;
    float* snorms = reinterpret_cast<float*>(w8 + L.guard_off);  // {max |q_h|^2, max |k_h|^2} of this block's projection
    if (tc) {
      ProfScope ps(h, AVB_PROF_QKV, st);
      if (cudaMemsetAsync(snorms, 0, 2 * sizeof(float), st) != cudaSuccess) return fail(h, AVB_ERR_CUDA, "score-bound reset failed");
      if ((rc = launch_gemm_split<2, true, avb::kSplitEpiQkvOperands>(h, z, k, W.tm_s_qkv, 0, W.bqkv_f, nullptr, 0, qkv, 3 * HD, M, 3 * HD,
                                                                       st, avb::SplitSeq{n, d, axis, dg, H}, snorms))) return rc;
    } else
    { ProfScope ps(h, AVB_PROF_QKV, st);
      launch_sgemm<0, true>(z, k, W.wq, HD, W.bq, W.ln_g[0], W.ln_b[0], nullptr, 0, qkv, 3 * HD, ntok, HD, k, st);
      AVB_LAUNCH_CHECK(h);
      launch_sgemm<0, true>(z, k, W.wk, HD, W.bk, W.ln_g[1], W.ln_b[1], nullptr, 0, qkv + HD, 3 * HD, ntok, HD, k, st);
      AVB_LAUNCH_CHECK(h);
      launch_sgemm<0, true>(z, k, W.wv, HD, W.bv, W.ln_g[2], W.ln_b[2], nullptr, 0, qkv + 2 * HD, 3 * HD, ntok, HD, k, st);
      AVB_LAUNCH_CHECK(h); }
    const int T = axis == 0 ? d : n;
    const long long S = axis == 0 ? (long long)B * n : (long long)B * d;
    const long long inner = axis == 0 ? 1 : d;
    const long long outer_stride = axis == 0 ? d : (long long)n * d;
    const long long inner_stride = axis == 0 ? 0 : 1;
    const long long tok_stride = axis == 0 ? 1 : d;
    if (tc) {
      ProfScope ps(h, axis == 0 ? AVB_PROF_ATTN_D : AVB_PROF_ATTN_N, st);
      if ((rc = launch_attention_split(h, qkv, attn, B, n, d, H, axis, dg, st, snorms))) return rc;
    } else
    { ProfScope ps(h, axis == 0 ? AVB_PROF_ATTN_D : AVB_PROF_ATTN_N, st);
      for (long long s0 = 0; s0 < S; s0 += 65535) {  // grid.z is limited to 65535
        const int sz = (int)std::min<long long>(65535, S - s0);
        avb::attention_f32_kernel<<<dim3((T + 127) / 128, H, sz), 128, 0, st>>>(
            qkv, attn, T, H, s0, inner, outer_stride, inner_stride, tok_stride, 1.0f / sqrtf((float)c.key_size), dg,
            (long long)n * dg);
        h->launches++;
      } }
    { cudaError_t e__ = cudaGetLastError();
      if (e__ != cudaSuccess) return fail(h, AVB_ERR_CUDA, "attention_f32 launch failed: %s", cudaGetErrorString(e__)); }
    if (tc) {
      { ProfScope ps(h, AVB_PROF_OUT, st);
        if ((rc = launch_gemm_split<4, false, avb::kSplitEpiResidual>(h, attn, HD, W.tm_s_o, 0, W.bo, z, k, z, k, M, k, st))) return rc; }
      { ProfScope ps(h, AVB_PROF_FFN, st);
        if ((rc = launch_gemm_split<2, true, avb::kSplitEpiRelu>(h, z, k, W.tm_s_1, 0, W.b1_f, nullptr, 0, hid, F, M, F, st))) return rc;
        for (int k0 = 0; k0 < F; k0 += 256)  // K = 256 per launch; the later passes accumulate through the residual epilogue
          if ((rc = launch_gemm_split<4, false, avb::kSplitEpiResidual>(h, hid + k0, F, W.tm_s_2, k0, k0 == 0 ? W.b2 : W.zero_k, z, k, z,
                                                                       k, M, k, st))) return rc; }
    } else {
    { ProfScope ps(h, AVB_PROF_OUT, st);
      launch_sgemm<2, false>(attn, HD, W.wo, k, W.bo, nullptr, nullptr, z, k, z, k, ntok, k, HD, st);
      AVB_LAUNCH_CHECK(h); }
    { ProfScope ps(h, AVB_PROF_FFN, st);
      launch_sgemm<1, true>(z, k, W.w1, F, W.b1, W.ln_g[3], W.ln_b[3], nullptr, 0, hid, F, ntok, F, k, st);
      AVB_LAUNCH_CHECK(h);
      launch_sgemm<2, false>(hid, F, W.w2, k, W.b2, nullptr, nullptr, z, k, z, k, ntok, k, F, st);
      AVB_LAUNCH_CHECK(h); }
    }
  }
  return AVB_OK;
}

int avb_block(avb_handle h, int32_t bi, float* z, int32_t B, int32_t n, int32_t d, int32_t axis, void* ws,
              size_t ws_bytes, void* stream) {
  return block_impl(h, bi, z, B, n, d, axis, 0, ws, ws_bytes, stream);
}

int avb_block_sharded(avb_handle h, int32_t bi, float* z, int32_t rows, int32_t d, int32_t dg, void* ws, size_t ws_bytes,
                      void* stream) {
  return block_impl(h, bi, z, 1, rows, d, AVB_AXIS_D, dg, ws, ws_bytes, stream);
}

int avb_pool(avb_handle h, const float* z, int32_t B, int32_t n, int32_t d, float* pooled, int32_t accumulate,
             void* stream) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_pool: weights not loaded");
  if (!shape_ok(h, B, n, d) || !z || !pooled) return fail(h, AVB_ERR_INVALID, "avb_pool: bad argument");
  DeviceGuard dg(h->device);
  if (B > 65535) return fail(h, AVB_ERR_INVALID, "avb_pool: B > 65535");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope ps(h, AVB_PROF_POOL_HEAD, st);
  avb::ln_max_kernel<<<dim3(d, B), 256, 0, st>>>(z, h->lnf_g, h->lnf_b, n, d, pooled, accumulate);
  AVB_LAUNCH_CHECK(h);
  return AVB_OK;
}

static int head_impl(avb_handle h, const float* pooled, int32_t B, int32_t d, float* probs, void* scratch,
                     size_t scratch_bytes, void* stream, int log_probs) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_head: weights not loaded");
  if (B < 1 || d < 1 || !pooled || !probs) return fail(h, AVB_ERR_INVALID, "avb_head: bad argument");
  DeviceGuard dg(h->device);
  const avb_config& c = h->cfg;
  const size_t need = (size_t)2 * B * d * c.out_dim * 4;
  if (!scratch || scratch_bytes < need) return fail(h, AVB_ERR_WORKSPACE, "avb_head: scratch %zu < %zu", scratch_bytes, need);
  if (B > 65535) return fail(h, AVB_ERR_INVALID, "avb_head: B > 65535");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* u = reinterpret_cast<float*>(scratch);
  float* v = u + (size_t)B * d * c.out_dim;
  ProfScope ps(h, AVB_PROF_POOL_HEAD, st);
  avb::head_uv_kernel<<<(unsigned)((size_t)B * d), 128, 0, st>>>(pooled, h->lnu_g, h->lnu_b, h->wu, h->bu, h->lnv_g,
                                                               h->lnv_b, h->wv, h->bv, c.out_dim, u, v);
  AVB_LAUNCH_CHECK(h);
  if (c.compute_dtype == AVB_DTYPE_BF16 && c.out_dim == 128 && h->head_tc) {
    // the u . v^T contraction on the tensor cores (split-precision operands, kernels_split.cuh)
    static PerDeviceOnce once;
    cudaError_t e = once.run([] {
      return cudaFuncSetAttribute(avb::head_logits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, avb::kHeadTcSmemBytes);
    });
    if (e != cudaSuccess) return fail(h, AVB_ERR_CUDA, "cudaFuncSetAttribute(head_logits_tc): %s", cudaGetErrorString(e));
    avb::head_logits_tc_kernel<<<dim3((d + 127) / 128, (d + 127) / 128, B), avb::kHeadTcThreads, avb::kHeadTcSmemBytes, st>>>(
        u, v, d, h->temp, h->bias, c.mask_diag, log_probs, probs);
  } else {
    avb::head_logits_kernel<<<dim3((d + 15) / 16, (d + 15) / 16, B), dim3(16, 16), 0, st>>>(
        u, v, d, c.out_dim, h->temp, h->bias, c.mask_diag, log_probs, probs);
  }
  AVB_LAUNCH_CHECK(h);
  return AVB_OK;
}

int avb_head(avb_handle h, const float* pooled, int32_t B, int32_t d, float* probs, void* scratch, size_t scratch_bytes,
             void* stream) {
  return head_impl(h, pooled, B, d, probs, scratch, scratch_bytes, stream, 0);
}

int avb_head_logprobs(avb_handle h, const float* pooled, int32_t B, int32_t d, float* logprobs, void* scratch,
                      size_t scratch_bytes, void* stream) {
  return head_impl(h, pooled, B, d, logprobs, scratch, scratch_bytes, stream, 1);
}

// ----------------------------------------------------------------------------------- forward
static int resolve_chunks(avb_handle h, int n, int chunk_size, int* chunks) {
  *chunks = 1;
  if (chunk_size > 0 && chunk_size < n) {  // model.py:100
    if (n % chunk_size) return fail(h, AVB_ERR_INVALID, "observations `n` must be divisible by `experimental_chunk_size`");
    *chunks = n / chunk_size;
  }
  return AVB_OK;
}

size_t avb_workspace_bytes(avb_handle h, int32_t B, int32_t n, int32_t d, int32_t chunk_size) {
  if (!shape_ok(h, B, n, d)) return 0;
  int chunks = 1;
  if (resolve_chunks(h, n, chunk_size, &chunks)) return 0;
  return fwd_ws_layout(h->cfg, B, n, d, chunks).total;
}

static int forward_impl(avb_handle h, const float* x, const float* interv, int32_t B, int32_t n, int32_t d,
                        int32_t standardize, int32_t chunk_size, float* probs, void* ws, size_t ws_bytes, void* stream,
                        int log_probs) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_forward: weights not loaded");
  if (!shape_ok(h, B, n, d) || !x || !probs) return fail(h, AVB_ERR_INVALID, "avb_forward: bad argument");
  DeviceGuard dg(h->device);
  int chunks = 1, rc;
  if ((rc = resolve_chunks(h, n, chunk_size, &chunks))) return rc;
  const avb_config& c = h->cfg;
  const FwdWs L = fwd_ws_layout(c, B, n, d, chunks);
  if (!ws || ws_bytes < L.total) return fail(h, AVB_ERR_WORKSPACE, "avb_forward: workspace %zu < %zu", ws_bytes, L.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  float* xs = reinterpret_cast<float*>(w8 + L.xs_off);
  float* pooled = reinterpret_cast<float*>(w8 + L.pooled_off);
  float* z = reinterpret_cast<float*>(w8 + L.z_off);
  const size_t cells = (size_t)B * n * d;

  // with chunking the standardizer writes to its own staging buffer and the gather below produces xs
  float* xstd = chunks > 1 ? reinterpret_cast<float*>(w8 + L.xtmp_off) : xs;
  if ((rc = avb_standardize(h, x, B, n, d, standardize, xstd, w8 + L.stat_off, (size_t)B * (n + d + 8) * 8, stream))) return rc;
  const float* iv = interv;
  float* pool_dst = pooled;
  int Be = B, ne = n;
  if (chunks > 1) {
    // chunk c of dataset b becomes dataset b*chunks + c with n/chunks rows (model.py:105-106)
    chunk_gather_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(xstd, xs, B, n, d, chunks);
    AVB_LAUNCH_CHECK(h);
    if (interv) {
      float* ivp = reinterpret_cast<float*>(w8 + L.iv_off);
      chunk_gather_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(interv, ivp, B, n, d, chunks);
      AVB_LAUNCH_CHECK(h);
      iv = ivp;
    }
    Be = B * chunks; ne = n / chunks;
    pool_dst = reinterpret_cast<float*>(w8 + L.pooled_c_off);
  }
  const int nb = (int)h->blocks.size();
  const size_t tok1 = (size_t)ne * d;
  for (int b0 = 0; b0 < Be; b0 += L.Bc) {
    const int bc = std::min(L.Bc, Be - b0);
    const size_t ntok = (size_t)bc * tok1;
    if ((rc = avb_embed(h, xs + (size_t)b0 * tok1, iv ? iv + (size_t)b0 * tok1 : nullptr, (int64_t)ntok, z, stream))) return rc;
    const size_t bws = block_ws_layout(c, ntok).total;
    for (int i = 0; i < nb; ++i)
      if ((rc = avb_block(h, i, z, bc, ne, d, (i & 1) ? AVB_AXIS_N : AVB_AXIS_D, w8 + L.blk_off, bws, stream))) return rc;
    if ((rc = avb_pool(h, z, bc, ne, d, pool_dst + (size_t)b0 * d * c.dim, 0, stream))) return rc;
  }
  if (chunks > 1) {
    const size_t per_ds = (size_t)d * c.dim, total = (size_t)B * per_ds;
    chunk_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pool_dst, pooled, per_ds, chunks, total);
    AVB_LAUNCH_CHECK(h);
  }
  return head_impl(h, pooled, B, d, probs, w8 + L.uv_off, (size_t)2 * B * d * c.out_dim * 4, stream, log_probs);
}

int avb_forward(avb_handle h, const float* x, const float* interv, int32_t B, int32_t n, int32_t d, int32_t standardize,
                int32_t chunk_size, float* probs, void* ws, size_t ws_bytes, void* stream) {
  return forward_impl(h, x, interv, B, n, d, standardize, chunk_size, probs, ws, ws_bytes, stream, 0);
}

int avb_forward_logprobs(avb_handle h, const float* x, const float* interv, int32_t B, int32_t n, int32_t d,
                         int32_t standardize, int32_t chunk_size, float* logprobs, void* ws, size_t ws_bytes, void* stream) {
  return forward_impl(h, x, interv, B, n, d, standardize, chunk_size, logprobs, ws, ws_bytes, stream, 1);
}

int avb_forward_host(avb_handle h, const float* x_host, const float* interv_host, int32_t B, int32_t n, int32_t d,
                     int32_t standardize, int32_t chunk_size, float* probs_host) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_forward_host: weights not loaded");
  if (!shape_ok(h, B, n, d) || !x_host || !probs_host) return fail(h, AVB_ERR_INVALID, "avb_forward_host: bad argument");
  DeviceGuard dg(h->device);
  std::lock_guard<std::mutex> lock(h->host_mu);  // the internal workspace and stream are per handle
  int chunks = 1, rc;
  if ((rc = resolve_chunks(h, n, chunk_size, &chunks))) return rc;
  const size_t cells = (size_t)B * n * d, pbytes = (size_t)B * d * d * 4;
  const size_t io = align_up(cells * 4, 1024) * 2 + align_up(pbytes, 1024);
  const size_t need = io + fwd_ws_layout(h->cfg, B, n, d, chunks).total;
  if (h->host_ws_bytes < need) {
    if (h->host_ws) cudaFree(h->host_ws);
    h->host_ws = nullptr; h->host_ws_bytes = 0;
    AVB_CUDA(h, cudaMalloc(&h->host_ws, need));
    h->host_ws_bytes = need;
  }
  uint8_t* w8 = reinterpret_cast<uint8_t*>(h->host_ws);
  float* dx = reinterpret_cast<float*>(w8);
  float* di = reinterpret_cast<float*>(w8 + align_up(cells * 4, 1024));
  float* dp = reinterpret_cast<float*>(w8 + 2 * align_up(cells * 4, 1024));
  cudaStream_t st = h->host_stream;  // the handle's own non-blocking stream (not the legacy default stream)
  AVB_CUDA(h, cudaMemcpyAsync(dx, x_host, cells * 4, cudaMemcpyHostToDevice, st));
  if (interv_host) AVB_CUDA(h, cudaMemcpyAsync(di, interv_host, cells * 4, cudaMemcpyHostToDevice, st));
  if ((rc = avb_forward(h, dx, interv_host ? di : nullptr, B, n, d, standardize, chunk_size, dp, w8 + io,
                        h->host_ws_bytes - io, st)))
    return rc;
  AVB_CUDA(h, cudaMemcpyAsync(probs_host, dp, pbytes, cudaMemcpyDeviceToHost, st));
  AVB_CUDA(h, cudaStreamSynchronize(st));
  return AVB_OK;
}

// ----------------------------------------------------------------------------------- axially sharded forward (config 4)
namespace {
struct ShardWs {
  size_t xs_off, stat_off, xsb_off, ivb_off, za_off, zb_off, blk_off, pl_off, pooled_off, uv_off, total;
};
ShardWs shard_ws_layout(const avb_config& c, int n, int d, int world) {
  ShardWs w{};
  const size_t ng = (size_t)n / world, ntok = ng * d;
  size_t off = 0;
  w.xs_off = off; off = align_up(off + (size_t)n * d * 4, 1024);
  w.stat_off = off; off = align_up(off + (size_t)(n + d + 8) * 8, 1024);
  w.xsb_off = off; off = align_up(off + ntok * 4, 1024);
  w.ivb_off = off; off = align_up(off + ntok * 4, 1024);
  w.za_off = off; off = align_up(off + ntok * c.dim * 4, 1024);
  w.zb_off = off; off = align_up(off + ntok * c.dim * 4, 1024);
  w.blk_off = off; off += block_ws_layout(c, ntok).total;
  w.pl_off = off; off = align_up(off + (size_t)(d / world) * c.dim * 4, 1024);
  w.pooled_off = off; off = align_up(off + (size_t)d * c.dim * 4, 1024);
  w.uv_off = off; off = align_up(off + (size_t)2 * d * c.out_dim * 4, 1024);
  w.total = off;
  return w;
}

// rows [r0, r0 + ng) of src[n, d] -> column-blocked [G][ng][dg] (RowMap with dg > 0)
__global__ void shard_gather_kernel(const float* __restrict__ src, float* __restrict__ dst, int r0, int ng, int d, int dg) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)ng * d) return;
  const size_t blk = (size_t)ng * dg;
  const int q = (int)(idx / blk);
  const size_t rem = idx - (size_t)q * blk;
  const int i = (int)(rem / dg), jj = (int)(rem - (size_t)i * dg);
  dst[idx] = src[(size_t)(r0 + i) * d + (size_t)q * dg + jj];
}
}  // namespace

int avb_comm_unique_id(void* id_out) {
  if (!id_out) return fail(nullptr, AVB_ERR_INVALID, "avb_comm_unique_id: null argument");
  const NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(nullptr, AVB_ERR_CUDA, "NCCL unavailable: %s", nc.why.c_str());
  avb_nccl_unique_id id;
  const int rc = nc.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, AVB_ERR_CUDA, "ncclGetUniqueId: %s", nc.GetErrorString(rc));
  memcpy(id_out, &id, sizeof id);
  return AVB_OK;
}

int avb_comm_init(avb_handle h, const void* id128, int32_t rank, int32_t world) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(h, AVB_ERR_INVALID, "avb_comm_init: bad argument");
  const NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(h, AVB_ERR_CUDA, "NCCL unavailable: %s", nc.why.c_str());
  DeviceGuard dg(h->device);
  if (h->comm) { nc.CommDestroy(h->comm); h->comm = nullptr; }
  avb_nccl_unique_id id;
  memcpy(&id, id128, sizeof id);
  const int rc = nc.CommInitRank(&h->comm, world, id, rank);
  if (rc != 0) { h->comm = nullptr; return fail(h, AVB_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", rank, world, nc.GetErrorString(rc)); }
  h->comm_rank = rank; h->comm_world = world;
  return AVB_OK;
}

int avb_comm_destroy(avb_handle h) {
  if (!h) return AVB_ERR_INVALID;
  if (h->comm) {
    DeviceGuard dg(h->device);
    nccl_api().CommDestroy(h->comm);
    h->comm = nullptr; h->comm_rank = 0; h->comm_world = 1;
  }
  return AVB_OK;
}

size_t avb_sharded_workspace_bytes(avb_handle h, int32_t n, int32_t d, int32_t world) {
  if (!shape_ok(h, 1, n, d) || world < 1 || n % world || d % world) return 0;
  return shard_ws_layout(h->cfg, n, d, world).total;
}

int avb_forward_sharded(avb_handle h, const float* x, const float* interv, int32_t n, int32_t d, int32_t standardize,
                        float* probs, void* ws, size_t ws_bytes, void* stream, void* comm, int32_t rank, int32_t world) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_forward_sharded: weights not loaded");
  if (!shape_ok(h, 1, n, d) || !x || !probs) return fail(h, AVB_ERR_INVALID, "avb_forward_sharded: bad argument");
  avb_nccl_comm_t cm = comm ? reinterpret_cast<avb_nccl_comm_t>(comm) : h->comm;
  if (!comm && h->comm) { rank = h->comm_rank; world = h->comm_world; }
  if (world < 1 || rank < 0 || rank >= world) return fail(h, AVB_ERR_INVALID, "avb_forward_sharded: bad rank %d / world %d", rank, world);
  if (n % world) return fail(h, AVB_ERR_INVALID, "observations `n` must be divisible by `device_count` if sharding is enabled. ");
  if (d % world) return fail(h, AVB_ERR_INVALID, "variables `d` must be divisible by `device_count` for the axial (n-shard <-> d-shard) sharding. ");
  if (world > 1 && !cm) return fail(h, AVB_ERR_INVALID, "avb_forward_sharded: no communicator (avb_comm_init)");
  const NcclApi& nc = nccl_api();
  if (world > 1 && !nc.ok) return fail(h, AVB_ERR_CUDA, "NCCL unavailable: %s", nc.why.c_str());
  DeviceGuard dev_guard(h->device);
  const avb_config& c = h->cfg;
  const ShardWs L = shard_ws_layout(c, n, d, world);
  if (!ws || ws_bytes < L.total) return fail(h, AVB_ERR_WORKSPACE, "avb_forward_sharded: workspace %zu < %zu", ws_bytes, L.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  const int ng = n / world, dg = d / world;
  const size_t ntok = (size_t)ng * d, chunk = (size_t)ng * dg * c.dim;
  float* xs = reinterpret_cast<float*>(w8 + L.xs_off);
  float* xsb = reinterpret_cast<float*>(w8 + L.xsb_off);
  float* ivb = reinterpret_cast<float*>(w8 + L.ivb_off);
  float* zc = reinterpret_cast<float*>(w8 + L.za_off);   // current residual stream
  float* zn = reinterpret_cast<float*>(w8 + L.zb_off);   // target of the next exchange
  float* pooled_loc = reinterpret_cast<float*>(w8 + L.pl_off);
  float* pooled = reinterpret_cast<float*>(w8 + L.pooled_off);
  int rc;
  // every rank standardizes the whole [n, d] matrix (the z-score needs all n rows of a column; 20 MB at n=5000, d=1000),
  // then keeps its n/G rows in column-blocked order
  if ((rc = avb_standardize(h, x, 1, n, d, standardize, xs, w8 + L.stat_off, (size_t)(n + d + 8) * 8, stream))) return rc;
  { ProfScope ps(h, AVB_PROF_PROLOGUE, st);
    const unsigned gb = (unsigned)((ntok + 255) / 256);
    shard_gather_kernel<<<gb, 256, 0, st>>>(xs, xsb, rank * ng, ng, d, dg);
    AVB_LAUNCH_CHECK(h);
    if (interv) { shard_gather_kernel<<<gb, 256, 0, st>>>(interv, ivb, rank * ng, ng, d, dg); AVB_LAUNCH_CHECK(h); } }
  if ((rc = avb_embed(h, xsb, interv ? ivb : nullptr, (int64_t)ntok, zc, stream))) return rc;
  const int nb = (int)h->blocks.size();
  const size_t bws = block_ws_layout(c, ntok).total;
  for (int i = 0; i < nb; ++i) {
    if ((i & 1) == 0) {  // attend over d on the n-shard [G][n/G][d/G] (column blocks)
      if ((rc = block_impl(h, i, zc, 1, ng, d, AVB_AXIS_D, world > 1 ? dg : 0, w8 + L.blk_off, bws, stream))) return rc;
    } else {             // attend over n on the d-shard [n][d/G]
      if ((rc = block_impl(h, i, zc, 1, n, dg, AVB_AXIS_N, 0, w8 + L.blk_off, bws, stream))) return rc;
    }
    if (i + 1 < nb && world > 1) {
      // n-shard <-> d-shard: both layouts are G chunks of [n/G][d/G][dim]; chunk p goes to rank p and the chunk received
      // from rank p lands at position p - one grouped send/recv, no pack or unpack pass, fp32 (bit-identical to one GPU)
      ProfScope ps(h, AVB_PROF_COMM, st);
      int e = nc.GroupStart();
      for (int p = 0; p < world && e == 0; ++p) {
        e = nc.Send(zc + (size_t)p * chunk, chunk, kNcclFloat32, p, cm, st);
        if (e == 0) e = nc.Recv(zn + (size_t)p * chunk, chunk, kNcclFloat32, p, cm, st);
      }
      const int e2 = nc.GroupEnd();
      if (e == 0) e = e2;
      if (e != 0) return fail(h, AVB_ERR_CUDA, "NCCL exchange after block %d: %s", i, nc.GetErrorString(e));
      std::swap(zc, zn);
    }
  }
  // the last block attends over n: z is d-sharded [n][d/G], so the final LayerNorm + max over n is local (model.py:86-89)
  if ((rc = avb_pool(h, zc, 1, n, dg, world > 1 ? pooled_loc : pooled, 0, stream))) return rc;
  if (world > 1) {
    ProfScope ps(h, AVB_PROF_COMM, st);
    const int e = nc.AllGather(pooled_loc, pooled, (size_t)dg * c.dim, kNcclFloat32, cm, st);
    if (e != 0) return fail(h, AVB_ERR_CUDA, "ncclAllGather(pooled): %s", nc.GetErrorString(e));
  }
  return head_impl(h, pooled, 1, d, probs, w8 + L.uv_off, (size_t)2 * d * c.out_dim * 4, stream, 0);
}

// ----------------------------------------------------------------------------------- test hooks
namespace {
// scratch of the layout test hooks (grown on demand, process lifetime)
void* test_scratch(size_t bytes) {
  static void* buf = nullptr;
  static size_t cap = 0;
  if (cap < bytes) {
    if (buf) cudaFree(buf);
    buf = nullptr; cap = 0;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return nullptr;
    cap = bytes;
  }
  return buf;
}
}  // namespace

int avb_test_qkv_layout_bf16(const float* z, const void* wt, const float* bias, void* out, int32_t B, int32_t n, int32_t d,
                             int32_t H, int32_t axis, void* stream) {
  if (!z || !wt || !bias || !out || B < 1 || n < 1 || d < 1 || H != 8 || (axis != 0 && axis != 1))
    return fail(nullptr, AVB_ERR_INVALID, "avb_test_qkv_layout_bf16: bad argument (H must be 8)");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "no CUDA device");
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t ntok = (size_t)B * n * d;
  const int N = 3 * H * 32;
  void* packed = test_scratch(ntok * N * sizeof(bf16));
  if (!packed) return fail(nullptr, AVB_ERR_CUDA, "cudaMalloc(test scratch) failed");
  CUtensorMap tb;
  if (!make_tmap_2d(&tb, wt, N, 128, 128)) return fail(nullptr, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rc = launch_ln_gemm2_tc(nullptr, z, tb, bias, packed, N, (int)ntok, N, st, sms, nullptr, n, d, axis);
  if (rc != AVB_OK) return rc;
  avb::unpack_qkv_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(packed), reinterpret_cast<bf16*>(out), B, n, d, H, axis);
  return cudaGetLastError() == cudaSuccess ? AVB_OK : fail(nullptr, AVB_ERR_CUDA, "unpack_qkv launch failed");
}

int avb_test_outproj_bf16(const void* attn, const void* wot, const float* bias, void* delta, int32_t B, int32_t n, int32_t d,
                          int32_t axis, void* stream) {
  if (!attn || !wot || !bias || !delta || B < 1 || n < 1 || d < 1 || (axis != 0 && axis != 1))
    return fail(nullptr, AVB_ERR_INVALID, "avb_test_outproj_bf16: bad argument");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "no CUDA device");
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t ntok = (size_t)B * n * d;
  const int H = 2 * avb::kOutProjKB, HD = H * 32;
  void* packed = test_scratch(((ntok + 127) / 128) * 128 * HD * sizeof(bf16));
  if (!packed) return fail(nullptr, AVB_ERR_CUDA, "cudaMalloc(test scratch) failed");
  CUtensorMap tw;
  if (!make_tmap_2d(&tw, wot, 128, HD, 128)) return fail(nullptr, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  avb::pack_attn_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(attn), reinterpret_cast<bf16*>(packed), B, n, d, H, axis);
  return launch_outproj_tc(nullptr, packed, tw, bias, delta, (int)ntok, n, d, axis, st, sms, nullptr);
}

int avb_test_ffn_bf16(float* z, const void* w1t, const void* w2t, const float* b1, const float* b2, int32_t M, int32_t F,
                      void* stream) {
  if (!z || !w1t || !w2t || !b1 || !b2 || M < 1 || F % 256) return fail(nullptr, AVB_ERR_INVALID, "avb_test_ffn_bf16: bad shape");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "no CUDA device");
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CUtensorMap t1, t2;
  if (!make_tmap_2d(&t1, w1t, F, 128, 64) || !make_tmap_2d(&t2, w2t, 128, F, 64))
    return fail(nullptr, AVB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return launch_ffn2_tc(nullptr, z, t1, t2, b1, b2, M, F, reinterpret_cast<cudaStream_t>(stream), sms, nullptr);
}

// ----------------------------------------------------------------------------------- graph metrics (8(f) row 4)
size_t avb_graph_metrics_workspace_bytes(int32_t B, int32_t d) {
  if (B < 1 || d < 1) return 0;
  return (size_t)4 * B * d * d * sizeof(double);
}

int avb_graph_metrics(const float* probs, const int32_t* g_true, float threshold, int32_t B, int32_t d, int64_t* out,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!probs || !out || B < 1 || d < 1) return fail(nullptr, AVB_ERR_INVALID, "avb_graph_metrics: bad argument");
  if ((size_t)d * d > (size_t)1 << 30) return fail(nullptr, AVB_ERR_INVALID, "avb_graph_metrics: d too large");
  const size_t need = avb_graph_metrics_workspace_bytes(B, d);
  if (!workspace || workspace_bytes < need)
    return fail(nullptr, AVB_ERR_WORKSPACE, "avb_graph_metrics: workspace %zu < %zu", workspace_bytes, need);
  if (B > 65535) return fail(nullptr, AVB_ERR_INVALID, "avb_graph_metrics: B > 65535");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long* o = reinterpret_cast<long long*>(out);
  avb::graph_counts_kernel<<<B, 256, 0, st>>>(probs, g_true, threshold, d, o);
  const size_t per = (size_t)B * d * d;
  double* zb[2] = {reinterpret_cast<double*>(workspace), reinterpret_cast<double*>(workspace) + per};  // z ping/pong
  double* rb[2] = {zb[1] + per, zb[1] + 2 * per};                                                      // result ping/pong
  avb::acyc_init_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(probs, threshold, d, per, zb[0]);  // mat = I + G/d
  const dim3 grid((d + 15) / 16, (d + 15) / 16, B);
  auto mm = [&](const double* x, const double* y, double* c) { avb::bmm_f64_kernel<<<grid, 256, 0, st>>>(x, y, c, d); };
  // numpy.linalg.matrix_power(mat, d): the multiplication order decides the fp64 rounding, so it is numpy's
  const double* result = nullptr;
  if (d == 1) {
    result = zb[0];
  } else if (d == 2) {
    mm(zb[0], zb[0], rb[0]); result = rb[0];
  } else if (d == 3) {
    mm(zb[0], zb[0], zb[1]); mm(zb[1], zb[0], rb[0]); result = rb[0];
  } else {
    // binary decomposition: z = a, a^2, a^4, ...; result = product of the z whose bit is set, low bit first
    int zi = 0, ri = 0, n = d;
    bool first_z = true;
    while (n > 0) {
      if (first_z) first_z = false;
      else { mm(zb[zi], zb[zi], zb[zi ^ 1]); zi ^= 1; }
      const int bit = n & 1;
      n >>= 1;
      if (bit) {
        if (!result) {
          if (cudaMemcpyAsync(rb[ri], zb[zi], per * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return fail(nullptr, AVB_ERR_CUDA, "avb_graph_metrics: copy failed");
        } else {
          mm(rb[ri], zb[zi], rb[ri ^ 1]); ri ^= 1;
        }
        result = rb[ri];
      }
    }
  }
  avb::acyc_trace_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(result, d, B, o);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "avb_graph_metrics: launch failed: %s", cudaGetErrorString(e));
  return AVB_OK;
}

int avb_threshold_metrics(const float* probs, const int32_t* g_true, int32_t B, int32_t d, double* out, void* stream) {
  if (!probs || !g_true || !out || B < 1 || d < 1) return fail(nullptr, AVB_ERR_INVALID, "avb_threshold_metrics: bad argument");
  if ((size_t)d * d > (size_t)1 << 30) return fail(nullptr, AVB_ERR_INVALID, "avb_threshold_metrics: d too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = d * d >= 1024 ? 1024 : 256;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int bc = std::min(65535, B - b0);
    const size_t off = (size_t)b0 * d * d;
    avb::threshold_metrics_kernel<<<bc, threads, 0, st>>>(probs + off, g_true + off, d, out + (size_t)b0 * 4);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "avb_threshold_metrics: launch failed: %s", cudaGetErrorString(e));
  return AVB_OK;
}

int avb_test_set_dattn_mode(avb_handle h, int32_t mode) {
  if (!h || mode < 0 || mode > 2) return fail(h, AVB_ERR_INVALID, "avb_test_set_dattn_mode: mode must be 0, 1 or 2");
  h->dattn_mode = mode;
  return AVB_OK;
}

int avb_test_set_f32_tc(avb_handle h, int32_t on) {
  if (!h) return AVB_ERR_INVALID;
  h->f32_tc = on != 0;
  return AVB_OK;
}

int avb_test_set_dattn_pack(avb_handle h, int32_t on) {
  if (!h) return AVB_ERR_INVALID;
  h->dattn_pack = on != 0;
  return AVB_OK;
}

int avb_test_set_head_tc(avb_handle h, int32_t on) {
  if (!h) return AVB_ERR_INVALID;
  h->head_tc = on != 0;
  return AVB_OK;
}

int avb_test_dattn_bf16(avb_handle h, int32_t bi, const float* z, int32_t B, int32_t n, int32_t d, int32_t fuse_out,
                        void* delta, int32_t* guard_host, void* stream) {
  if (!h || !h->loaded) return fail(h, AVB_ERR_INVALID, "avb_test_dattn_bf16: weights not loaded");
  if (h->cfg.compute_dtype != AVB_DTYPE_BF16 || h->cfg.num_heads != avb::kDaHeads)
    return fail(h, AVB_ERR_INVALID, "avb_test_dattn_bf16: bf16 handle with 8 heads required");
  if (!shape_ok(h, B, n, d) || d > 128 || !z || !delta || bi < 0 || bi >= (int)h->blocks.size())
    return fail(h, AVB_ERR_INVALID, "avb_test_dattn_bf16: bad argument (d <= 128)");
  DeviceGuard dg(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Block& W = h->blocks[bi];
  const size_t ntok = (size_t)B * n * d;
  const int HD = h->cfg.num_heads * h->cfg.key_size;
  uint8_t* scratch = reinterpret_cast<uint8_t*>(test_scratch(((ntok + 127) / 128) * 128 * HD * sizeof(bf16) + 1024));
  if (!scratch) return fail(h, AVB_ERR_CUDA, "cudaMalloc(test scratch) failed");
  int* guard = reinterpret_cast<int*>(scratch);
  bf16* attn = reinterpret_cast<bf16*>(scratch + 1024);
  AVB_CUDA(h, cudaMemsetAsync(guard, 0, sizeof(int), st));
  int rc;
  if ((rc = launch_dattn_tc(h, z, W, fuse_out ? delta : (void*)attn, B * n, d, fuse_out != 0, st, h->num_sms, nullptr, guard))) return rc;
  if (!fuse_out && (rc = launch_outproj_tc(h, attn, W.tm_wo, W.bo, delta, (int)ntok, n, d, 0, st, h->num_sms, nullptr))) return rc;
  if (guard_host) {
    AVB_CUDA(h, cudaMemcpyAsync(guard_host, guard, sizeof(int), cudaMemcpyDeviceToHost, st));
    AVB_CUDA(h, cudaStreamSynchronize(st));
  }
  return AVB_OK;
}

int avb_debug_set_dattn_dump(void* dump_dev) {
#ifdef AVB_DATTN_DEBUG
  float* p = reinterpret_cast<float*>(dump_dev);
  return cudaMemcpyToSymbol(avb::g_dattn_dbg, &p, sizeof p) == cudaSuccess ? AVB_OK : AVB_ERR_CUDA;
#else
  (void)dump_dev;
  return fail(nullptr, AVB_ERR_INVALID, "avb_debug_set_dattn_dump: not a -DAVB_DATTN_DEBUG build");
#endif
}

int avb_debug_set_trace(void* trace_dev) {
  long long* p = reinterpret_cast<long long*>(trace_dev);
  return cudaMemcpyToSymbol(avb::g_attn_trace, &p, sizeof p) == cudaSuccess ? AVB_OK : AVB_ERR_CUDA;
}

int avb_test_attention_bf16(const void* qkv, void* out, int32_t B, int32_t n, int32_t d, int32_t H, int32_t axis,
                            void* stream) {
  if (!qkv || !out || B < 1 || n < 1 || d < 1 || H < 1 || (axis != 0 && axis != 1))
    return fail(nullptr, AVB_ERR_INVALID, "avb_test_attention_bf16: bad argument");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "no CUDA device");
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static int* guard = nullptr;  // test entry point: one process-wide guard word (+ score bound) and repack buffer
  static void* packed = nullptr;
  static size_t packed_bytes = 0;
  if (!guard && cudaMalloc(&guard, 4 * sizeof(int)) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "cudaMalloc(guard) failed");
  float* norms = reinterpret_cast<float*>(guard) + 1;
  const size_t need = ((size_t)B * n * d * 3 + (((size_t)B * n * d + 127) / 128) * 128) * H * 32 * sizeof(bf16);
  if (packed_bytes < need) {
    if (packed) cudaFree(packed);
    packed = nullptr; packed_bytes = 0;
    if (cudaMalloc(&packed, need) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "cudaMalloc(repack buffer) failed");
    packed_bytes = need;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool prepacked = ab_env("AVB_ATTN_PREPACKED") != nullptr;  // profiling scripts: time the attention alone
  if (prepacked) {  // the score bound the QKV projection would have recorded (profiling runs: once per process, so that
    static bool have_norms = false;  // the timed launches are the attention kernel alone)
    if (!have_norms) {
      if (cudaMemsetAsync(norms, 0, 2 * sizeof(float), st) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "memset failed");
      avb::qk_norms_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), (size_t)B * n * d * H, H, axis == 0 ? d : n, norms);
      cudaStreamSynchronize(st);
      have_norms = true;
    }
    return launch_attention_tc(nullptr, qkv, out, B, n, d, H, axis, st, sms, nullptr, guard, 0, norms);
  }
  // token-major test tensors <-> the kernels' operand layouts; `out` must hold whole 128-row tiles in the packed form,
  // so the packed output lives behind the packed input in the scratch buffer
  bf16* pin = reinterpret_cast<bf16*>(packed);
  bf16* pout = pin + (size_t)B * n * d * 3 * H * 32;
  avb::repack_qkv_kernel<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), pin, B, n, d, H, axis);
  if (cudaMemsetAsync(norms, 0, 2 * sizeof(float), st) != cudaSuccess) return fail(nullptr, AVB_ERR_CUDA, "memset failed");
  avb::qk_norms_kernel<<<sms * 8, 256, 0, st>>>(pin, (size_t)B * n * d * H, H, axis == 0 ? d : n, norms);
  const int rc = launch_attention_tc(nullptr, pin, pout, B, n, d, H, axis, st, sms, nullptr, guard, 0, norms);
  if (rc != AVB_OK) return rc;
  avb::unpack_attn_kernel<<<sms * 8, 256, 0, st>>>(pout, reinterpret_cast<bf16*>(out), B, n, d, H, axis);
  return cudaGetLastError() == cudaSuccess ? AVB_OK : fail(nullptr, AVB_ERR_CUDA, "unpack_attn launch failed");
}

}  // extern "C"
