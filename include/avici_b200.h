/* avici_b200.h — C-ABI of the B200-native AVICI inference forward.
 *
 * Drop-in boundary.  The reference (larslorch/avici) is pure Python/JAX and has no
 * FFI; its operator seam for this path is
 *
 *   InferenceModel.net.apply(params, rng, x[...,n,d,2], is_training=False, experimental_chunk_size)
 *        -> logits[...,d,d]                                   avici/model.py:217 (class at :93-144)
 *   AVICIModel.__call_main__(x[n,d], interv[n,d])  -> p[d,d]  avici/pretrain.py:53-67
 *
 * Every entry point below replaces one piece of that seam and cites it.  Plain pointers and sizes
 * only; no torch types.  All `*_dev` pointers are CUDA device pointers on the current device, all
 * work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous unless stated.
 * A handle belongs to the device that was current in avb_create; entry points taking a handle switch to
 * that device for the duration of the call (pointers and stream must belong to it).
 * The library never falls back to a CPU path: without a CUDA device every compute call returns
 * AVB_ERR_CUDA.
 */
#ifndef AVICI_B200_H
#define AVICI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVB_OK 0
#define AVB_ERR_INVALID 1     /* bad argument / unsupported shape (reference: AssertionError)      */
#define AVB_ERR_CUDA 2        /* CUDA runtime / driver failure                                      */
#define AVB_ERR_MISSING 3     /* a required weight tensor was not supplied                          */
#define AVB_ERR_WORKSPACE 4   /* caller workspace too small                                         */

#define AVB_DTYPE_F32 0       /* fp32 SIMT kernels, true-fp32 accumulate (parity mode, <=1e-4)      */
#define AVB_DTYPE_BF16 1      /* bf16 tcgen05/TMEM tensor-core kernels, fp32 accumulate (<=1e-2)    */

#define AVB_STANDARDIZE_NONE 0
#define AVB_STANDARDIZE_DEFAULT 1 /* jax_standardize_default_simple, avici/utils/data_jax.py:72-79 */
#define AVB_STANDARDIZE_COUNT 2   /* jax_standardize_count_simple,   avici/utils/data_jax.py:48-59 */

#define AVB_AXIS_D 0 /* attend over the variable axis d (even blocks, model.py:82-84)    */
#define AVB_AXIS_N 1 /* attend over the observation axis n (odd blocks)                  */

typedef struct avb_model_s* avb_handle;

/* Network shape: the kwargs of BaseModel.__init__ (avici/model.py:26-50) plus
 * InferenceModel's mask_diag (model.py:156).  `layers` is the reference's `layers`
 * (layer PAIRS; the network has 2*layers blocks, model.py:42). */
typedef struct {
  int32_t layers;
  int32_t dim;
  int32_t key_size;
  int32_t num_heads;
  int32_t widening_factor;
  int32_t out_dim;
  int32_t mask_diag;
  int32_t compute_dtype; /* AVB_DTYPE_* */
} avb_config;

/* One leaf of the Haiku parameter dict (SURVEY.md §8c naming), e.g.
 * name = "BaseModel/multi_head_attention_3/query/w", host fp32, row-major in Haiku's own
 * shape (Linear w:[in,out]).  Replaces `state.params` (avici/pretrain.py:267). */
typedef struct {
  const char* name;
  const float* data; /* HOST pointer */
  int64_t numel;
} avb_tensor_desc;

/* Library / build info: "avici_b200 <ver> sm_100a". */
const char* avb_version(void);

/* Replaces InferenceModel.__init__ + hk.transform(model_class(**kwargs)) (model.py:150-172).
 * Unsupported shapes (dim!=128, key_size!=32, ...) -> AVB_ERR_INVALID, never a fallback. */
int avb_create(avb_handle* out, const avb_config* cfg);
void avb_destroy(avb_handle h);
const char* avb_last_error(avb_handle h); /* h may be NULL: last error of avb_create */

/* Replaces passing `params` to net.apply (model.py:217): copies, folds LayerNorm affine
 * parameters and the softmax scale into the projections (bf16 mode), packs K-major bf16
 * copies for the tensor-core kernels and uploads.  Synchronous. */
int avb_load_weights(avb_handle h, const avb_tensor_desc* tensors, int32_t count);

/* Number of fp32 parameters the configured network expects (4 269 442 for the defaults). */
int64_t avb_param_count(avb_handle h);

/* Bytes of scratch `avb_forward` needs for a batch of B datasets of shape [n,d] (0 on a bad
 * argument, e.g. a chunk_size that does not divide n). */
size_t avb_workspace_bytes(avb_handle h, int32_t B, int32_t n, int32_t d, int32_t chunk_size);

/* Replaces AVICIModel.__call_main__ (pretrain.py:53-67) on a batch (leading dims as in
 * train.py:61): stack interv channel (NULL = zeros, pretrain.py:56-57) -> standardize ->
 * BaseModel forward -> exp(log_sigmoid(logits)), diag = 0 (model.py:203-240).
 * x_dev/interv_dev: [B,n,d] fp32; probs_dev: [B,d,d] fp32.
 * chunk_size: experimental_chunk_size (model.py:99-116); <=0 or >=n = off; must divide n. */
int avb_forward(avb_handle h, const float* x_dev, const float* interv_dev, int32_t B, int32_t n,
                int32_t d, int32_t standardize, int32_t chunk_size, float* probs_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream);

/* infer_edge_logprobs (model.py:203-222) on the same inputs: log_sigmoid(logits) in fp32, diagonal = -inf when
 * mask_diag.  Unlike log(avb_forward's p) it keeps its relative precision where p underflows. */
int avb_forward_logprobs(avb_handle h, const float* x_dev, const float* interv_dev, int32_t B, int32_t n,
                         int32_t d, int32_t standardize, int32_t chunk_size, float* logprobs_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (what the Python wrapper's numpy path uses): copies x/interv to
 * the device, runs avb_forward on an internal workspace and the handle's own stream, and copies probs back.
 * Synchronous; calls on the same handle are serialised. */
int avb_forward_host(avb_handle h, const float* x_host, const float* interv_host, int32_t B,
                     int32_t n, int32_t d, int32_t standardize, int32_t chunk_size,
                     float* probs_host);

/* ---- stage-level entry points (used by the axially-sharded multi-GPU driver, which puts an
 * all-to-all between blocks; also by the parity tests).  All operate on a residual stream
 * z_dev[B,n,d,dim] fp32, token index t = (b*n + i)*d + j. ---- */

/* The standardizers alone: xs_dev[B,n,d] <- standardize(x_dev)[...,0].  scratch_dev needs
 * >= B*(n+d+8)*8 bytes. */
int avb_standardize(avb_handle h, const float* x_dev, int32_t B, int32_t n, int32_t d, int32_t mode,
                    float* xs_dev, void* scratch_dev, size_t scratch_bytes, void* stream);

/* hk.Linear(dim) embedding of the [.,2] (value, interv) pairs (model.py:95).
 * xs_dev/interv_dev: [ntok]; z_dev: [ntok,dim]. */
int avb_embed(avb_handle h, const float* xs_dev, const float* interv_dev, int64_t ntok,
              float* z_dev, void* stream);

/* One `_block` (model.py:53-75) applied in place; `axis` selects the attended axis (the
 * reference alternates it with swapaxes, model.py:84; here z is never physically transposed). */
size_t avb_block_workspace_bytes(avb_handle h, int32_t B, int32_t n, int32_t d);
int avb_block(avb_handle h, int32_t block_idx, float* z_dev, int32_t B, int32_t n, int32_t d,
              int32_t axis, void* workspace_dev, size_t workspace_bytes, void* stream);

/* The same `_block`, attending over d, on the n-shard of ONE axially sharded dataset whose `rows` x d tokens are stored
 * in d / dg column blocks: token (i, j) at ((j / dg) * rows + i) * dg + j % dg.  Column block q is the contiguous chunk the
 * all-to-all of avb_forward_sharded sends to rank q, so re-sharding needs no pack / unpack pass.  dg = 0: plain layout. */
int avb_block_sharded(avb_handle h, int32_t block_idx, float* z_dev, int32_t rows, int32_t d, int32_t dg,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* Final LayerNorm + max over n (model.py:86-89): pooled_dev[B,d,dim].  When `accumulate` != 0
 * the result is max-combined into pooled_dev (chunk path model.py:116, n-sharded ranks). */
int avb_pool(avb_handle h, const float* z_dev, int32_t B, int32_t n, int32_t d, float* pooled_dev,
             int32_t accumulate, void* stream);

/* u/v heads, cosine logits, temperature, bias, sigmoid, zero diagonal
 * (model.py:121-141, 218-220, 239): probs_dev[B,d,d].  scratch_dev >= 2*B*d*out_dim*4 bytes. */
int avb_head(avb_handle h, const float* pooled_dev, int32_t B, int32_t d, float* probs_dev,
             void* scratch_dev, size_t scratch_bytes, void* stream);

/* Same head, log-probabilities: log_sigmoid(logits), diagonal = -inf (model.py:218-220). */
int avb_head_logprobs(avb_handle h, const float* pooled_dev, int32_t B, int32_t d, float* logprobs_dev,
                      void* scratch_dev, size_t scratch_bytes, void* stream);

/* ---- one large dataset sharded axially over the GPUs of a node (BASELINE config 4; replaces the reference's
 * `shard_if_possible` path, avici/pretrain.py:124-143, which shards n only and leaves the collectives to XLA/GSPMD).
 * One handle and one rank per GPU.  Blocks that attend over d run on the rank's n-shard [n/G, d] (column-blocked, see
 * avb_block_sharded), blocks that attend over n on its d-shard [n, d/G]; between consecutive blocks the fp32 residual
 * stream is re-sharded with ONE grouped ncclSend/ncclRecv all-to-all (chunk p -> rank p), in place of the layout, so the
 * result is bit-identical to the single-GPU forward.  The final LayerNorm + max over n is local to the d-shard; the pooled
 * [d/G, dim] rows are all-gathered and every rank computes the head.  Requires G | n and G | d.
 *   avb_comm_unique_id : rank 0 creates the 128-byte NCCL unique id (broadcast it to the other ranks by any means)
 *   avb_comm_init      : every rank, collectively: ncclCommInitRank on the handle's device; the handle owns the comm
 *   avb_forward_sharded: x_dev / interv_dev are the FULL [n, d] matrices on every rank (20 MB at n=5000, d=1000),
 *                        probs_dev[d, d] is produced on every rank.  `comm`: an ncclComm_t of the caller's, or NULL for
 *                        the handle's own (then rank / world are the handle's).  NCCL is resolved with dlopen at the first
 *                        call (the libnccl.so.2 already loaded in the process, e.g. PyTorch's); world = 1 needs none. */
int avb_comm_unique_id(void* id_out /*128 bytes*/);
int avb_comm_init(avb_handle h, const void* id /*128 bytes*/, int32_t rank, int32_t world);
int avb_comm_destroy(avb_handle h);
size_t avb_sharded_workspace_bytes(avb_handle h, int32_t n, int32_t d, int32_t world);
int avb_forward_sharded(avb_handle h, const float* x_dev, const float* interv_dev, int32_t n, int32_t d,
                        int32_t standardize, float* probs_dev, void* workspace_dev, size_t workspace_bytes, void* stream,
                        void* comm /*ncclComm_t or NULL*/, int32_t rank, int32_t world);

/* ---- graph metrics on the device: replaces the per-dataset numpy calls of the eval loop
 * avici/train.py:125-151 -> avici/metrics.py:5-24 (shd), :27-32 (n_edges), :35-51 (is_acyclic / is_cyclic),
 * :54-87 (classification_metrics: precision / recall / f1 follow from tp, fp, fn).
 * probs_dev[B,d,d] f32; the predicted graph is probs > threshold (train.py:63).  g_true_dev[B,d,d] int32 0/1 or NULL
 * (then shd / n_edges_true / tp / fp / fn are computed against the empty graph).  out_dev[B][AVB_GM_FIELDS] int64.
 * `cyclic` is the REFERENCE's test, trace((I + G/d)^d) - d not isclose 0 in fp64 with numpy's matrix_power
 * multiplication order, not a graph search: cycles too long to reach its 1e-8 tolerance count as acyclic there too.
 * No handle: the call needs no weights; errors are reported through avb_last_error(NULL). ---- */
#define AVB_GM_SHD 0
#define AVB_GM_EDGES_PRED 1
#define AVB_GM_EDGES_TRUE 2
#define AVB_GM_TP 3
#define AVB_GM_FP 4
#define AVB_GM_FN 5
#define AVB_GM_CYCLIC 6
#define AVB_GM_FIELDS 8
size_t avb_graph_metrics_workspace_bytes(int32_t B, int32_t d);
int avb_graph_metrics(const float* probs_dev, const int32_t* g_true_dev, float threshold, int32_t B, int32_t d,
                      int64_t* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Threshold-free metrics per dataset: out_dev[B][4] f64 = {auroc, auprc, average precision, 0}, the values
 * avici/metrics.py:90-127 (`threshold_metrics`) gets from sklearn's roc_curve/auc, precision_recall_curve/auc and
 * average_precision_score on the flattened [d*d] scores, with the same corner cases (no positives and no predicted
 * mass -> 1, 1, 1; exactly one of them -> 0.5, 0, 0).  fp64 sums: agreement with sklearn to ~1e-14, not bit-exact. */
int avb_threshold_metrics(const float* probs_dev, const int32_t* g_true_dev, int32_t B, int32_t d, double* out_dev,
                          void* stream);

/* Kernel launches issued by this handle since creation (bench.py's `gpu_launches`). */
int64_t avb_launch_count(avb_handle h);

/* Optional per-kernel-category timing with CUDA events recorded on the launching stream (bench.py's
 * `roofline`).  avb_profile_read synchronises the device, adds the elapsed times of all scopes
 * recorded since the last read to running totals and copies totals (ms) and scope counts out. */
#define AVB_PROF_PROLOGUE 0  /* standardize + embed                                               */
#define AVB_PROF_QKV 1       /* LayerNorm + Q/K/V projections (unfused blocks)                     */
#define AVB_PROF_ATTN_D 2    /* attention over d (unfused blocks: d > 128, fp32 mode)              */
#define AVB_PROF_ATTN_N 3    /* attention over n                                                   */
#define AVB_PROF_OUT 4       /* output projection (unfused blocks)                                 */
#define AVB_PROF_FFN 5       /* FFN sub-layer (LayerNorm, up-projection, ReLU, down-projection)    */
#define AVB_PROF_POOL_HEAD 6 /* final LN + max-pool + head                                         */
#define AVB_PROF_DATTN 7     /* fused LayerNorm + QKV + attention over d + out-projection (d<=128) */
#define AVB_PROF_COMM 8      /* axially sharded forward: all-to-all exchanges + the pooled all-gather  */
#define AVB_PROF_CATEGORIES 9
int avb_profile_enable(avb_handle h, int32_t on);
int avb_profile_read(avb_handle h, double* ms_out /*[AVB_PROF_CATEGORIES]*/,
                     int64_t* count_out /*[AVB_PROF_CATEGORIES]*/, int32_t reset);

/* ---- unit-test hooks for the tensor-core kernels (bf16 handles only) ---- */
/* z[M,128] fp32 += relu(LayerNorm(z, no affine) @ W1t[F,128]^T + b1) @ W2t[128,F]^T + b2, in place (F % 256 == 0). */
int avb_test_ffn_bf16(float* z_dev, const void* w1t_bf16_dev, const void* w2t_bf16_dev, const float* b1_dev,
                      const float* b2_dev, int32_t M, int32_t F, void* stream);
/* Axial attention on packed qkv[B,n,d,3*H*32] bf16 (q pre-scaled by log2e/sqrt(32)):
 * out[B,n,d,H*32] bf16. */
int avb_test_attention_bf16(const void* qkv_bf16_dev, void* out_bf16_dev, int32_t B, int32_t n,
                            int32_t d, int32_t H, int32_t axis, void* stream);

/* The QKV projection as the block runs it (ln_gemm_tc_kernel<2>): LayerNorm + projection of z[B*n*d,128] with rows taken
 * in the sequence-major order of `axis` and the result written in the attention operand layout; the hook converts
 * that layout back to token-major out[B,n,d,3*H*32] bf16 so a test can compare it with a plain reference. */
int avb_test_qkv_layout_bf16(const float* z_dev, const void* wt_bf16_dev, const float* bias_dev,
                             void* out_tokmajor_bf16_dev, int32_t B, int32_t n, int32_t d, int32_t H,
                             int32_t axis, void* stream);
/* The out-projection as the block runs it (outproj_tc_kernel): delta[B*n*d,128] bf16 (token order) = attn @
 * Wot[128,256]^T + bias, with attn given token-major [B,n,d,256] bf16 and packed by the hook into the kernel's
 * A-operand layout (sequence-major rows of `axis`). */
int avb_test_outproj_bf16(const void* attn_tokmajor_bf16_dev, const void* wot_bf16_dev, const float* bias_dev,
                          void* delta_bf16_dev, int32_t B, int32_t n, int32_t d, int32_t axis, void* stream);

/* The fused d-axis attention sub-layer as the block runs it (dattn_tc_kernel, blocks that attend over d <= 128):
 * delta[B*n*d,128] bf16 (token order) = MHA(LN_q(z), LN_k(z), LN_v(z)) . Wo + bo of block `block_idx` of the loaded
 * weights, z[B*n*d,128] fp32.  fuse_out 0 runs the kernel up to the attention output and the out-projection as its
 * own launch.  *guard_host (may be NULL) receives the fast softmax's range guard (0 = in range); reading it
 * synchronises the stream. */
int avb_test_dattn_bf16(avb_handle h, int32_t block_idx, const float* z_dev, int32_t B, int32_t n, int32_t d,
                        int32_t fuse_out, void* delta_bf16_dev, int32_t* guard_host, void* stream);
/* fp32 handles: 1 (default) = projections and FFN as split-precision (bf16 x 3, fp32 accumulate) tensor-core GEMMs,
 * 0 = the SIMT fp32 GEMM kernels (A/B reference of the parity mode). */
int avb_test_set_f32_tc(avb_handle h, int32_t on);
/* bf16 handles: 1 (default) = the head's u . v^T contraction on the tensor cores (head_logits_tc_kernel), 0 = SIMT. */
int avb_test_set_head_tc(avb_handle h, int32_t on);
/* bf16 handles: 1 (default) = the fused d-axis kernel packs 128 / d sequences per 128-row tile when d <= 64, 0 = one per tile. */
int avb_test_set_dattn_pack(avb_handle h, int32_t on);
/* Which kernels run the attention half of a block that attends over d <= 128 (bf16 mode): 1 (default) the fully
 * fused dattn_tc_kernel, 2 the same kernel up to the attention output + the separate out-projection, 0 the unfused
 * QKV / attention / out-projection launches (what every other block uses).  A/B and test knob. */
int avb_test_set_dattn_mode(avb_handle h, int32_t mode);
/* Debug builds (-DAVB_DATTN_DEBUG) only: device buffer of 278528 floats that CTA 0 of dattn_tc_kernel fills with the
 * intermediates of its first sequence (NULL switches the dump off); AVB_ERR_INVALID in production builds. */
int avb_debug_set_dattn_dump(void* dump_dev);

/* Debug: device buffer of 3*4096 int64 that block 0 of the attention kernel fills with clock64() stamps of
 * its producer / MMA / softmax roles (NULL switches tracing off). */
int avb_debug_set_trace(void* trace_dev);

#ifdef __cplusplus
}
#endif
#endif /* AVICI_B200_H */
