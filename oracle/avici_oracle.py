"""CPU oracle for the AVICI inference forward (TEST INFRASTRUCTURE — not product code).

This file is a plain-numpy restatement of the reference algorithm for the hot path
`AVICIModel.__call__ -> __call_main__ -> InferenceModel.infer_edge_probs -> BaseModel.__call__`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product package `avici_b200` never does.

PARITY UNPINNED: the reference (larslorch/avici) is pure JAX + dm-haiku; neither is
installed here and the reference ships no tests, golden vectors or stored outputs
(SURVEY.md §4, §8c).  The arithmetic lives in un-vendored third-party packages
(`dm-haiku>=0.0.8`, `jax>=0.3.17`, lower bounds only — setup.py:27-36).  Their published
algorithms are restated below, one clearly-named function each, anchored on the
reference's own call sites:

  reference file:line                      restated here
  ---------------------------------------  -----------------------------------
  avici/model.py:15-16   layer_norm        haiku_layer_norm
  avici/model.py:56-65   _block (MHA)      haiku_multi_head_attention, block
  avici/model.py:67-75   _block (FFN)      block
  avici/model.py:78-90   _all_blocks..     all_blocks_and_max
  avici/model.py:93-144  __call__          base_model_forward
  avici/model.py:99-116  chunk path        base_model_forward(chunk_size=…)
  avici/model.py:203-240 infer_edge_*      infer_edge_logprobs / infer_edge_probs
  avici/model.py:175-200 sample_graphs     (probabilities only; RNG is JAX threefry)
  avici/pretrain.py:53-67  __call_main__   call_main
  avici/pretrain.py:106-156 __call__       avici_call
  avici/utils/data_jax.py:72-79            standardize_default_simple
  avici/utils/data_jax.py:10-29,48-59      standardize_count_simple

The oracle is dtype-generic: run it with float64 parameters/inputs for the "gold"
tier, float32 for the "reference precision" tier (XLA:CPU computes true fp32).
"""
import math

import numpy as np

LN_EPS = 1e-5  # hk.LayerNorm default eps


# --------------------------------------------------------------------------------------
# Haiku layer semantics (dm-haiku, not under /root/reference; see header)
# --------------------------------------------------------------------------------------
def haiku_linear(x, p):
    """hk.Linear: y = x @ w + b with w:[in,out] (call sites model.py:71,73,95,124,128)."""
    return x @ p["w"] + p["b"]


def haiku_layer_norm(x, p):
    """hk.LayerNorm(axis=-1, create_scale=True, create_offset=True), eps=1e-5, biased variance
    (model.py:15-16)."""
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)  # ddof=0
    inv = p["scale"] / np.sqrt(var + x.dtype.type(LN_EPS))
    return inv * (x - mean) + p["offset"]


def _softmax_last(a):
    a = a - a.max(axis=-1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(axis=-1, keepdims=True)


def haiku_multi_head_attention(q_in, k_in, v_in, params, prefix, num_heads, key_size,
                               max_logit_elems=1 << 26):
    """hk.MultiHeadAttention(num_heads, key_size, w_init_scale=2.0, model_size) (model.py:59-64).

    The sequence axis is axis -2 of the inputs; all leading axes are batch.  Projections carry
    a bias, heads are laid out head-major in the projected feature axis (column c = h*key+j),
    logits are scaled by 1/sqrt(key_size), softmax over keys, final Linear(model_size).
    The T x T logits are materialised exactly like Haiku does, but in slabs over the flattened
    batch axis so that the (n=1000, d=100) case fits in memory.
    """
    dt = q_in.dtype
    q = haiku_linear(q_in, params[f"{prefix}/query"])
    k = haiku_linear(k_in, params[f"{prefix}/key"])
    v = haiku_linear(v_in, params[f"{prefix}/value"])
    lead, T = q.shape[:-2], q.shape[-2]
    H, D = num_heads, key_size
    q = q.reshape(-1, T, H, D)
    k = k.reshape(-1, T, H, D)
    v = v.reshape(-1, T, H, D)
    S = q.shape[0]
    out = np.empty((S, T, H * D), dtype=dt)
    slab = max(1, int(max_logit_elems // max(1, H * T * T)))
    scale = dt.type(1.0 / math.sqrt(D))
    for s0 in range(0, S, slab):
        s1 = min(S, s0 + slab)
        logits = np.einsum("sthd,sThd->shtT", q[s0:s1], k[s0:s1]) * scale
        w = _softmax_last(logits)
        o = np.einsum("shtT,sThd->sthd", w, v[s0:s1])
        out[s0:s1] = o.reshape(s1 - s0, T, H * D)
    out = out.reshape(*lead, T, H * D)
    return haiku_linear(out, params[f"{prefix}/linear"])


# --------------------------------------------------------------------------------------
# Haiku parameter names: counters in module creation order
# --------------------------------------------------------------------------------------
def _sfx(i):
    return "" if i == 0 else f"_{i}"


def block_param_names(i, root="BaseModel"):
    """Names used by block i (creation order in model.py:56-73)."""
    return dict(
        ln_q=f"{root}/layer_norm{_sfx(4 * i)}",
        ln_k=f"{root}/layer_norm{_sfx(4 * i + 1)}",
        ln_v=f"{root}/layer_norm{_sfx(4 * i + 2)}",
        mha=f"{root}/multi_head_attention{_sfx(i)}",
        ln_ffn=f"{root}/layer_norm{_sfx(4 * i + 3)}",
        ffn1=f"{root}/linear{_sfx(2 * i + 1)}",
        ffn2=f"{root}/linear{_sfx(2 * i + 2)}",
    )


# --------------------------------------------------------------------------------------
# BaseModel (model.py:24-144)
# --------------------------------------------------------------------------------------
def block(z, params, i, cfg):
    """One `_block` (model.py:53-75), dropout rate 0 at inference.  Attends over axis -2."""
    nm = block_param_names(i)
    q_in = haiku_layer_norm(z, params[nm["ln_q"]])
    k_in = haiku_layer_norm(z, params[nm["ln_k"]])
    v_in = haiku_layer_norm(z, params[nm["ln_v"]])
    z = z + haiku_multi_head_attention(q_in, k_in, v_in, params, nm["mha"],
                                       cfg["num_heads"], cfg["key_size"])
    z_in = haiku_layer_norm(z, params[nm["ln_ffn"]])
    h = np.maximum(haiku_linear(z_in, params[nm["ffn1"]]), 0)
    z = z + haiku_linear(h, params[nm["ffn2"]])
    return z


def all_blocks_and_max(z, params, cfg):
    """`_all_blocks_and_max` (model.py:78-90): 2*layers blocks, swapping axes -3/-2 after each,
    final LN, max over axis -3 (the observation axis, since the block count is even)."""
    n_blocks = 2 * cfg["layers"]
    assert n_blocks % 2 == 0
    for i in range(n_blocks):
        z = block(z, params, i, cfg)
        z = np.swapaxes(z, -3, -2)
    z = haiku_layer_norm(z, params[f"BaseModel/layer_norm{_sfx(4 * n_blocks)}"])
    return z.max(axis=-3)


def base_model_forward(params, x, cfg, chunk_size=None):
    """`BaseModel.__call__` (model.py:93-144), is_training=False.  x: [..., n, d, 2] -> logits [..., d, d]."""
    n_blocks = 2 * cfg["layers"]
    z = haiku_linear(x, params["BaseModel/linear"])
    n = z.shape[-3]
    if chunk_size is not None and 0 < chunk_size < n:
        # model.py:99-116: reshape n -> [chunk_size, chunks]; row i goes to chunk (i mod chunks)
        assert n % chunk_size == 0, "observations `n` must be divisible by `experimental_chunk_size` "
        chunks = n // chunk_size
        z = z.reshape(*z.shape[:-3], chunk_size, chunks, *z.shape[-2:])
        z = np.swapaxes(z, -3, 0)
        z = np.stack([all_blocks_and_max(z[c], params, cfg) for c in range(chunks)], axis=0)
        z = z.max(axis=0)
    else:
        z = all_blocks_and_max(z, params, cfg)

    u = haiku_linear(haiku_layer_norm(z, params[f"BaseModel/layer_norm{_sfx(4 * n_blocks + 1)}"]),
                     params[f"BaseModel/linear{_sfx(2 * n_blocks + 1)}"])
    v = haiku_linear(haiku_layer_norm(z, params[f"BaseModel/layer_norm{_sfx(4 * n_blocks + 2)}"]),
                     params[f"BaseModel/linear{_sfx(2 * n_blocks + 2)}"])
    u = u / np.linalg.norm(u, axis=-1, ord=2, keepdims=True)
    v = v / np.linalg.norm(v, axis=-1, ord=2, keepdims=True)
    logit_ij = np.einsum("...id,...jd->...ij", u, v)
    temp = params["BaseModel"]["learned_temp"].squeeze()
    logit_ij = logit_ij * np.exp(temp)
    logit_ij = logit_ij + params["BaseModel"]["final_matrix_bias"].squeeze()
    assert logit_ij.shape[-1] == x.shape[-2] and logit_ij.shape[-2] == x.shape[-2]
    return logit_ij


# --------------------------------------------------------------------------------------
# InferenceModel (model.py:203-240)
# --------------------------------------------------------------------------------------
def log_sigmoid(x):
    """jax.nn.log_sigmoid = -softplus(-x), numerically stable form."""
    return -(np.maximum(-x, 0) + np.log1p(np.exp(-np.abs(x))))


def infer_edge_logprobs(params, x, cfg, mask_diag=True, chunk_size=None):
    logits = base_model_forward(params, x, cfg, chunk_size=chunk_size)
    logp = log_sigmoid(logits)
    if mask_diag:
        d = logp.shape[-1]
        logp = logp.copy()
        logp[..., np.arange(d), np.arange(d)] = -np.inf  # set_diagonal, model.py:19-21
    return logp


def infer_edge_probs(params, x, cfg, mask_diag=True, chunk_size=None):
    with np.errstate(divide="ignore"):
        return np.exp(infer_edge_logprobs(params, x, cfg, mask_diag=mask_diag, chunk_size=chunk_size))


# --------------------------------------------------------------------------------------
# Standardizers (utils/data_jax.py)
# --------------------------------------------------------------------------------------
def standardize_default_simple(x):
    """jax_standardize_default_simple (data_jax.py:72-79): z-score channel 0 over the n axis,
    population std, std==0 -> 1; channel 1 untouched."""
    x = x.copy()
    ref = x[..., 0]
    mean = ref.mean(-2, keepdims=True)
    std = ref.std(-2, keepdims=True)
    x[..., 0] = (ref - mean) / np.where(std == 0.0, x.dtype.type(1.0), std)
    return x


def _cpm_standardizer(x):
    """_jax_cpm_standardizer (data_jax.py:10-17)."""
    lib = x.sum(-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.log2(x / (lib * x.dtype.type(1e-6)))
    return np.where(np.isclose(x, 0.0), np.nan, val)


def _cpm_shift_scale(x, shift, scale):
    """_jax_cpm_shift_scale (data_jax.py:20-29)."""
    with np.errstate(divide="ignore", invalid="ignore"):  # global std == 0: 0/0 -> NaN -> 0 below, as in the reference
        x = (x - np.where(np.isnan(shift), 0.0, shift)) / np.where(np.isnan(scale), 1.0, scale)
    if x.size != 0:
        x = np.where(np.isnan(x), 0.0, x)
    return x


def standardize_count_simple(x):
    """jax_standardize_count_simple (data_jax.py:48-59): log2-CPM, minus global nan-min,
    divided by global nan-std (per dataset: axes -1,-2), NaN -> 0."""
    x = x.copy()
    dt = x.dtype
    cpm = _cpm_standardizer(x[..., 0])
    with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore", RuntimeWarning)
        gmin = np.nanmin(cpm, axis=(-1, -2), keepdims=True)
        gstd = np.nanstd(cpm, axis=(-1, -2), keepdims=True)
    x[..., 0] = _cpm_shift_scale(cpm, gmin, gstd).astype(dt)
    return x


# --------------------------------------------------------------------------------------
# AVICIModel (pretrain.py:53-156)
# --------------------------------------------------------------------------------------
def call_main(params, x, interv, cfg, expects_counts=False, mask_diag=True, chunk_size=None):
    """`__call_main__` (pretrain.py:53-67): stack the intervention channel, standardize, forward.
    Accepts leading batch dims like the jitted reference does."""
    if interv is None:
        x2 = np.stack([x, np.zeros_like(x)], axis=-1)
    else:
        assert x.shape == interv.shape
        x2 = np.stack([x, interv.astype(x.dtype)], axis=-1)
    x2 = standardize_count_simple(x2) if expects_counts else standardize_default_simple(x2)
    return infer_edge_probs(params, x2, cfg, mask_diag=mask_diag, chunk_size=chunk_size)


def avici_call(params, x, cfg, interv=None, return_probs=True, expects_counts=False,
               mask_diag=True, experimental_chunk_size=None, dtype=np.float32):
    """`AVICIModel.__call__` (pretrain.py:70-156) without the device placement."""
    assert x.ndim == 2, "`x` must be a 2D array of shape [n, d]."
    if interv is not None:
        assert x.shape == interv.shape, "`x` and `interv` must have the same shape when provided."
        assert interv.ndim == 2
        assert np.all(np.isclose(interv, 0) | np.isclose(interv, 1)), \
            "Intervention mask `interv` has to be binary."
        interv = np.around(interv, 0)
    p = cast_params(params, dtype)
    out = call_main(p, x.astype(dtype), None if interv is None else interv.astype(dtype), cfg,
                    expects_counts=expects_counts, mask_diag=mask_diag,
                    chunk_size=experimental_chunk_size)
    if not return_probs:
        out = (out > 0.5).astype(int)
    return out


def cast_params(params, dtype):
    return {m: {k: np.asarray(v, dtype=dtype) for k, v in leaves.items()} for m, leaves in params.items()}


DEFAULT_CFG = dict(layers=8, dim=128, key_size=32, num_heads=8, widening_factor=4, out_dim=128)


def flops_per_dataset(n, d, cfg=DEFAULT_CFG):
    """Algorithmic FLOPs per dataset, SURVEY.md §8a (MACs x 2; LN/softmax excluded)."""
    k, H, dk, wf = cfg["dim"], cfg["num_heads"], cfg["key_size"], cfg["widening_factor"]
    nb = 2 * cfg["layers"]
    per_tok_lin = 2 * (3 * k * H * dk + H * dk * k + 2 * k * wf * k)
    attn = 4 * H * dk  # per (token, key): QK^T + PV
    return (n * d * (nb * per_tok_lin + (nb // 2) * attn * (n + d))
            + 2 * 2 * k * n * d + 2 * 2 * d * k * cfg["out_dim"] + 2 * d * d * cfg["out_dim"])
