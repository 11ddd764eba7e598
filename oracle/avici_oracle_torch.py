"""torch-CPU twin of oracle/avici_oracle.py (TEST / BASELINE INFRASTRUCTURE — not product code).

Same algorithm as the reference forward (materialised `softmax(QK^T)` like `hk.MultiHeadAttention`,
unfused FFN, true fp32), restated with torch CPU ops so that it uses all host cores (MKL/oneDNN).
This is what `bench.py` times as the CPU baseline ("restated reference, torch-CPU; JAX is not
installable offline").  PARITY UNPINNED, like the numpy oracle it mirrors; the two
are checked against each other in tests/test_oracle.py.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / --impl reference) may import
this file.  Reference call sites: avici/model.py:53-144,203-240; avici/pretrain.py:53-67;
avici/utils/data_jax.py:72-79.
"""
import math

import torch

from .avici_oracle import block_param_names, _sfx

LN_EPS = 1e-5


def to_torch(params, dtype=torch.float32):
    return {m: {k: torch.as_tensor(v).to(dtype) for k, v in leaves.items()} for m, leaves in params.items()}


def _linear(x, p):
    return x @ p["w"] + p["b"]


def _ln(x, p):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), p["scale"], p["offset"], LN_EPS)


def _mha(q_in, k_in, v_in, P, prefix, H, D, max_logit_elems=1 << 27):
    q = _linear(q_in, P[f"{prefix}/query"])
    k = _linear(k_in, P[f"{prefix}/key"])
    v = _linear(v_in, P[f"{prefix}/value"])
    lead, T = q.shape[:-2], q.shape[-2]
    q = q.reshape(-1, T, H, D).permute(0, 2, 1, 3)
    k = k.reshape(-1, T, H, D).permute(0, 2, 1, 3)
    v = v.reshape(-1, T, H, D).permute(0, 2, 1, 3)
    S = q.shape[0]
    out = torch.empty((S, T, H * D), dtype=q.dtype)
    slab = max(1, int(max_logit_elems // max(1, H * T * T)))
    for s0 in range(0, S, slab):
        s1 = min(S, s0 + slab)
        logits = (q[s0:s1] @ k[s0:s1].transpose(-1, -2)) / math.sqrt(D)
        w = torch.softmax(logits, dim=-1)
        out[s0:s1] = (w @ v[s0:s1]).permute(0, 2, 1, 3).reshape(s1 - s0, T, H * D)
    return _linear(out.reshape(*lead, T, H * D), P[f"{prefix}/linear"])


def block(z, P, i, cfg):
    nm = block_param_names(i)
    z = z + _mha(_ln(z, P[nm["ln_q"]]), _ln(z, P[nm["ln_k"]]), _ln(z, P[nm["ln_v"]]), P, nm["mha"],
                 cfg["num_heads"], cfg["key_size"])
    h = torch.relu(_linear(_ln(z, P[nm["ln_ffn"]]), P[nm["ffn1"]]))
    return z + _linear(h, P[nm["ffn2"]])


def forward_probs(P, x, interv, cfg, mask_diag=True):
    """`__call_main__` with the default standardizer on x[..., n, d] (torch CPU tensors) -> p[..., d, d]."""
    mean = x.mean(-2, keepdim=True)
    std = x.std(-2, keepdim=True, unbiased=False)
    xs = (x - mean) / torch.where(std == 0.0, torch.ones_like(std), std)
    iv = torch.zeros_like(xs) if interv is None else interv.to(xs.dtype)
    return forward_probs_standardized(P, torch.stack([xs, iv], dim=-1), cfg, mask_diag=mask_diag)


def forward_cosines(P, x2, cfg):
    """BaseModel.__call__ (model.py:93-135) on an already standardized x2[..., n, d, 2] up to the cosine
    similarities u.v^T [..., d, d] (before temperature and bias)."""
    nb = 2 * cfg["layers"]
    z = _linear(x2, P["BaseModel/linear"])
    for i in range(nb):
        z = block(z, P, i, cfg).transpose(-3, -2)
    z = _ln(z, P[f"BaseModel/layer_norm{_sfx(4 * nb)}"]).amax(dim=-3)
    u = _linear(_ln(z, P[f"BaseModel/layer_norm{_sfx(4 * nb + 1)}"]), P[f"BaseModel/linear{_sfx(2 * nb + 1)}"])
    v = _linear(_ln(z, P[f"BaseModel/layer_norm{_sfx(4 * nb + 2)}"]), P[f"BaseModel/linear{_sfx(2 * nb + 2)}"])
    u = u / torch.linalg.norm(u, dim=-1, keepdim=True)
    v = v / torch.linalg.norm(v, dim=-1, keepdim=True)
    return torch.einsum("...id,...jd->...ij", u, v)


def probs_from_cosines(cos, temp, bias, mask_diag=True):
    """model.py:136-141 (temperature, bias) + 218-220, 239 (exp(log_sigmoid), diagonal -> 0)."""
    logits = cos * torch.exp(torch.as_tensor(temp, dtype=cos.dtype)) + torch.as_tensor(bias, dtype=cos.dtype)
    p = torch.exp(torch.nn.functional.logsigmoid(logits))
    if mask_diag:
        d = p.shape[-1]
        p = p * (1.0 - torch.eye(d, dtype=p.dtype))
    return p


def forward_probs_standardized(P, x2, cfg, mask_diag=True):
    cos = forward_cosines(P, x2, cfg)
    return probs_from_cosines(cos, P["BaseModel"]["learned_temp"].squeeze(), P["BaseModel"]["final_matrix_bias"].squeeze(),
                              mask_diag=mask_diag)
