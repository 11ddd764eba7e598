"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C-ABI, against the
CPU oracle on identical inputs and weights.

Tolerances are the north star's: max|dp| <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode, thresholded
graphs identical except for edges within that margin of 0.5.  Nothing here reads /root/reference.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

import avici_b200
from avici_b200 import _lib
from avici_b200.params import init_params
from oracle import avici_oracle as O
from oracle import avici_oracle_torch as OT

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}
# Two heads.  REF_INIT is the reference's own random init (cosine_temp_init=2.0 from example-custom/train.py:78,
# logit_bias_init=-3.0 from model.py:34): the north star's "same random-init weights" condition, under which the
# 1e-4 / 1e-2 criteria are asserted.  STRESS (temp 3 -> logits = 20*cos, bias 0) spreads p over (0,1) so that the
# 0.5 threshold and the full sigmoid slope are exercised; it multiplies every cosine error by 20*0.25 = 5 in p.
# fp32 mode must meet 1e-4 under STRESS too; bf16 operand rounding (weights, LN output, qkv, attention output,
# FFN hidden; measured max|dcos| ~ 4e-3 over 16 blocks) is held to 3e-2 there.
REF_INIT = dict(cosine_temp_init=2.0, logit_bias_init=-3.0)
STRESS = dict(cosine_temp_init=3.0, logit_bias_init=0.0)
TOL_STRESS = {"fp32": 1e-4, "bf16": 3e-2}


def oracle_probs(params, x, interv, layers, dtype=torch.float64):
    cfg = dict(O.DEFAULT_CFG, layers=layers)
    P = OT.to_torch(params, dtype)
    xi = None if interv is None else torch.as_tensor(np.asarray(interv)).to(dtype)
    return OT.forward_probs(P, torch.as_tensor(np.asarray(x)).to(dtype), xi, cfg).numpy()


def check_parity(p, ref, tol):
    err = float(np.abs(p - ref).max())
    assert err <= tol, f"max|dp| = {err:.3e} > {tol}"
    margin = np.abs(ref - 0.5) <= tol
    assert np.array_equal((p > 0.5)[~margin], (ref > 0.5)[~margin]), "thresholded graphs differ outside the margin"
    return err


# ------------------------------------------------------------------------------ kernel unit tests
@pytest.mark.parametrize("B,n,d,axis", [(1, 8, 100, 0), (2, 200, 50, 0), (2, 200, 50, 1), (1, 300, 130, 1),
                                        (3, 1000, 7, 1), (1, 1, 1, 0), (1, 5, 3, 1)])
def test_qkv_projection_operand_layout(built_lib, B, n, d, axis):
    """ln_gemm_tc_kernel<2>: rows gathered in sequence-major order of `axis`, output in the attention operand layout
    (the hook converts it back to token-major): must equal the plain LayerNorm + projection of every token."""
    H, N, M = 8, 768, B * n * d
    g = torch.Generator(device="cuda").manual_seed(M + axis)
    z = torch.randn(M, 128, device="cuda", generator=g) * 3 + torch.randn(M, 1, device="cuda", generator=g)
    wt = (torch.randn(N, 128, device="cuda", generator=g) / 128 ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    rc = built_lib.avb_test_qkv_layout_bf16(z.data_ptr(), wt.data_ptr(), bias.data_ptr(), out.data_ptr(), B, n, d, H, axis,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(built_lib, None, rc)
    torch.cuda.synchronize()
    xhat = torch.nn.functional.layer_norm(z, (128,), eps=1e-5).to(torch.bfloat16).float()
    ref = xhat @ wt.float().T + bias
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * max(ref.abs().max().item(), 1.0), f"qkv layout err {err}"


@pytest.mark.parametrize("B,n,d,axis", [(1, 8, 100, 0), (2, 200, 50, 0), (2, 200, 50, 1), (1, 300, 130, 1),
                                        (3, 1000, 7, 1), (1, 1, 1, 0), (1, 5, 3, 1)])
def test_outproj_operand_layout_and_row_map(built_lib, B, n, d, axis):
    """outproj_tc_kernel: A operand in the tile-major head-split layout, delta rows written through the sequence-major
    row map: delta[token] = attn @ Wo^T + bo (bf16) for every token, exactly once."""
    M = B * n * d
    g = torch.Generator(device="cuda").manual_seed(M * 3 + axis)
    attn = torch.randn(M, 256, device="cuda", generator=g).to(torch.bfloat16)
    wot = (torch.randn(128, 256, device="cuda", generator=g) / 16).to(torch.bfloat16)
    bias = torch.randn(128, device="cuda", generator=g)
    delta = torch.full((M, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = built_lib.avb_test_outproj_bf16(attn.data_ptr(), wot.data_ptr(), bias.data_ptr(), delta.data_ptr(), B, n, d, axis,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(built_lib, None, rc)
    torch.cuda.synchronize()
    ref = attn.float() @ wot.float().T + bias
    err = (delta.float() - ref).abs().max().item()
    assert err <= 2 ** -8 * max(ref.abs().max().item(), 1.0) + 1e-6, f"outproj err {err}"  # one bf16 rounding of the update


@pytest.mark.parametrize("M", [128, 1000, 333, 70001])
def test_ffn_tc_against_torch(built_lib, M):
    F = 512
    g = torch.Generator(device="cuda").manual_seed(M)
    z = torch.randn(M, 128, device="cuda", generator=g) * 2 + torch.randn(M, 1, device="cuda", generator=g)
    w1t = (torch.randn(F, 128, device="cuda", generator=g) / 128 ** 0.5).to(torch.bfloat16)
    w2t = (torch.randn(128, F, device="cuda", generator=g) / F ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(F, device="cuda", generator=g) * 0.5
    b2 = torch.randn(128, device="cuda", generator=g) * 0.5
    xhat = torch.nn.functional.layer_norm(z, (128,), eps=1e-5).to(torch.bfloat16).float()
    hid = (xhat @ w1t.float().T + b1).relu().to(torch.bfloat16).float()
    ref = z + hid @ w2t.float().T + b2
    out = z.clone()
    rc = built_lib.avb_test_ffn_bf16(out.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), b1.data_ptr(), b2.data_ptr(), M, F,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(built_lib, None, rc)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err <= 2e-2, f"ffn_tc err {err}"  # an xhat that rounds differently by one bf16 ulp moves a hidden unit by ~1e-2


def attention_reference(qkv, B, n, d, H, axis):
    """fp32 softmax attention over axis `axis` of qkv[B,n,d,3*H*32]; q is pre-scaled by log2e/sqrt(32)."""
    x = qkv.float().reshape(B, n, d, 3, H, 32)
    q, k, v = x[..., 0, :, :], x[..., 1, :, :], x[..., 2, :, :]       # [B,n,d,H,32]
    if axis == 1:
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))              # [B,d,n,H,32]: sequence axis last-but-2
    logits = torch.einsum("bsthe,bsThe->bshtT", q, k) * float(np.log(2.0))
    o = torch.einsum("bshtT,bsThe->bsthe", torch.softmax(logits, -1), v)
    if axis == 1:
        o = o.transpose(1, 2)
    return o.reshape(B, n, d, H * 32)


@pytest.mark.parametrize("B,n,d,H,axis", [(1, 8, 100, 8, 0), (2, 200, 50, 8, 0), (2, 200, 50, 8, 1),
                                          (1, 300, 130, 8, 0), (1, 300, 130, 8, 1), (1, 1000, 16, 8, 1),
                                          (3, 128, 128, 2, 0), (3, 128, 128, 2, 1), (1, 1, 1, 8, 0)])
def test_attention_tc_against_torch(built_lib, B, n, d, H, axis):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + n + d)
    qkv = (torch.randn(B, n, d, 3 * H * 32, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    out = torch.zeros(B, n, d, H * 32, device="cuda", dtype=torch.bfloat16)
    rc = built_lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(built_lib, None, rc)
    torch.cuda.synchronize()
    ref = attention_reference(qkv, B, n, d, H, axis)
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2, f"attention_tc err {err}"  # bf16 P and output rounding on |v| ~ 4


@pytest.mark.gpu
@pytest.mark.parametrize("axis", [0, 1])
def test_attention_tc_out_of_range_scores_fall_back_to_exact(built_lib, axis):
    """The fast attention kernel exponentiates raw scores (no running maximum); scores beyond +-100 (log2 units) must
    raise its guard so that the exact kernel recomputes the launch.  |q.k| reaches several hundred here."""
    B, n, d, H = 1, 300, 130, 8
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, n, d, 3 * H * 32, device="cuda", generator=g)
    qkv[..., :2 * H * 32] *= 4.0  # q and k: score std ~ 90, tails far beyond the guard
    qkv = qkv.to(torch.bfloat16)
    out = torch.zeros(B, n, d, H * 32, device="cuda", dtype=torch.bfloat16)
    rc = built_lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(built_lib, None, rc)
    torch.cuda.synchronize()
    ref = attention_reference(qkv, B, n, d, H, axis)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2, f"attention_tc (guarded) err {err}"


# ------------------------------------------------------------------------------ end-to-end parity
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))))
def test_golden_vectors(path, dtype):
    d = np.load(path)
    layers, counts = int(d["layers"]), bool(d["expects_counts"])
    kw = dict(layers=layers, **STRESS)
    model = avici_b200.random_init_model(kw, seed=int(d["param_seed"]), perturb=True, expects_counts=counts,
                                         compute_dtype=dtype)
    interv = d["interv"].astype(np.float32) if d["interv"].size else None
    chunk = int(d["chunk"]) if int(d["chunk"]) > 0 else None
    p = model(x=d["x"].astype(np.float32), interv=interv, experimental_chunk_size=chunk)
    assert isinstance(p, np.ndarray) and p.shape == d["p"].shape and np.all(np.diag(p) == 0)
    check_parity(p, d["p"], TOL_STRESS[dtype])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_config1_full_depth(dtype):
    """BASELINE config 1: scm-v0-shaped (16 blocks) random-init forward on simulate_data(d=50, n=200, rff-gauss)."""
    g, x, _ = avici_b200.simulate_data(d=50, n=200, seed=0, domain="rff-gauss")
    for head, tols in ((REF_INIT, TOL), (STRESS, TOL_STRESS)):
        kw = dict(layers=8, **head)
        for perturb in (False, True):
            model = avici_b200.random_init_model(kw, seed=0, perturb=perturb, compute_dtype=dtype)
            ref = oracle_probs(model.params, x, None, 8)
            p = model(x=x.astype(np.float32))
            err = check_parity(p, ref, tols[dtype])
            print(f"config1[{dtype}, temp={head['cosine_temp_init']}, perturb={perturb}]: max|dp| = {err:.2e}")
            gh = model(x=x.astype(np.float32), return_probs=False)
            assert gh.dtype.kind == "i" and set(np.unique(gh)) <= {0, 1}


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_batched_interventional_lin_gauss(dtype):
    """Config 3 shape family scaled to oracle size: lin-gauss with the intervention channel, batch of datasets;
    batched call == per-dataset calls == oracle."""
    kw = dict(layers=2, **STRESS)
    data = [avici_b200.simulate_data(d=30, n=160, n_interv=40, seed=s, domain="lin-gauss") for s in range(3)]
    x = np.stack([t[1] for t in data]).astype(np.float32)
    iv = np.stack([t[2] for t in data]).astype(np.float32)
    model = avici_b200.random_init_model(kw, seed=1, perturb=True, compute_dtype=dtype)
    pb = model.infer_batch(x, iv)
    pt = model.infer_batch(torch.as_tensor(x).cuda(), torch.as_tensor(iv).cuda()).cpu().numpy()
    assert np.array_equal(pb, pt)  # host-buffer entry point == device-buffer entry point
    for k in range(3):
        ref = oracle_probs(model.params, x[k].astype(np.float64), iv[k], 2)
        check_parity(pb[k], ref, TOL_STRESS[dtype])
        single = model(x=x[k], interv=iv[k].copy())
        assert np.abs(single - pb[k]).max() <= (1e-6 if dtype == "fp32" else 1e-2)


def test_standardizers_match_reference_outputs():
    """Rows a3 / a4 on the device: avb_standardize against outputs of the reference's own numpy standardizers
    (avici/utils/data.py:10-29,66-78; fixtures from tests/golden/standardize/make_standardize_golden.py), float32 inputs,
    including constant / all-zero columns, all-zero rows, an all-zero dataset and std == 0."""
    blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "standardize", "standardize_cases.npz"))
    eng = avici_b200.random_init_model(dict(layers=1), seed=0, compute_dtype="fp32").engine
    lib, h = eng.lib, eng.handle
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    checked = 0
    for key in blob["names"]:
        key = str(key)
        if not key.endswith("float32"):
            continue
        x, want, counts = blob[f"{key}__x"], blob[f"{key}__y"], bool(blob[f"{key}__counts"])
        n, d = x.shape
        xt = torch.as_tensor(x).cuda().contiguous()
        xs = torch.full_like(xt, float("nan"))
        scratch = torch.empty((n + d + 8) * 8, dtype=torch.uint8, device="cuda")
        mode = _lib.AVB_STANDARDIZE_COUNT if counts else _lib.AVB_STANDARDIZE_DEFAULT
        eng._check(lib.avb_standardize(h, xt.data_ptr(), 1, n, d, mode, xs.data_ptr(), scratch.data_ptr(),
                                       scratch.numel(), st))
        got = xs.cpu().numpy()
        assert np.isfinite(got).all(), key
        err = float(np.abs(got - want).max())
        assert err <= 1e-5 * max(1.0, float(np.abs(want).max())), f"{key}: {err}"
        checked += 1
    assert checked >= 20


def test_count_standardizer_gene_shape_fp32():
    """Config 5 family scaled down: count data + count standardizer, fp32 bit-tolerance check."""
    kw = dict(layers=2, **STRESS)
    g, x, _ = avici_b200.simulate_data(d=40, n=120, seed=2, domain="gene-ecoli")
    model = avici_b200.random_init_model(kw, seed=2, perturb=True, expects_counts=True, compute_dtype="fp32")
    cfg = dict(O.DEFAULT_CFG, layers=2)
    ref = O.avici_call(model.params, x, cfg, expects_counts=True, dtype=np.float64)
    check_parity(model(x=x.astype(np.float32)), ref, 1e-4)
    with pytest.warns(UserWarning, match="count data"):
        model(x=x.astype(np.float32) + 0.5)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_full_size_properties(dtype):
    """n=1000, d=100 (BASELINE configs 2/3 shape): size-independent properties instead of the (hours-long)
    full oracle: row-permutation invariance, variable-permutation equivariance, diag = 0, cosine bounds."""
    layers = 2 if dtype == "fp32" else 8
    kw = dict(layers=layers, **STRESS)
    model = avici_b200.random_init_model(kw, seed=0, perturb=True, compute_dtype=dtype)
    g, x, iv = avici_b200.simulate_data(d=100, n=800, n_interv=200, seed=5, domain="lin-gauss")
    x = x.astype(np.float32); iv = iv.astype(np.float32)
    rng = np.random.default_rng(0)
    pr, pc = rng.permutation(1000), rng.permutation(100)
    xb = np.stack([x, x[pr], x[:, pc]]); ib = np.stack([iv, iv[pr], iv[:, pc]])
    p = model.infer_batch(xb, ib)
    tol = 5e-5 if dtype == "fp32" else 3e-2
    assert np.abs(p[0] - p[1]).max() <= tol
    assert np.abs(p[0][np.ix_(pc, pc)] - p[2]).max() <= tol
    assert np.all(np.diagonal(p, axis1=1, axis2=2) == 0)
    sig = lambda t: 1 / (1 + np.exp(-t))
    off = p[0][~np.eye(100, dtype=bool)]
    assert off.min() >= sig(-np.exp(3.0)) - 1e-6 and off.max() <= sig(np.exp(3.0)) + 1e-6
    assert np.isfinite(p).all()


def test_stage_api_matches_forward_and_axial_layouts():
    """The stage-level C-ABI (what the axially-sharded driver calls) reproduces avb_forward, and running the
    d-axis blocks on row shards / n-axis blocks on column shards (a 2-rank emulation on one GPU, shards processed
    in turn) gives the same result: the multi-GPU path only adds the all-to-all data movement."""
    from avici_b200 import sharded  # noqa: F401
    kw = dict(layers=2, **STRESS)
    model = avici_b200.random_init_model(kw, seed=4, perturb=True, compute_dtype="fp32")
    eng = model.engine
    lib, h = eng.lib, eng.handle
    g, x, iv = avici_b200.simulate_data(d=24, n=96, n_interv=32, seed=9, domain="lin-gauss")
    n, d = x.shape
    xt, it = torch.as_tensor(x, dtype=torch.float32).cuda(), torch.as_tensor(iv, dtype=torch.float32).cuda()
    full = model(x=xt, interv=it.clone())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xs = torch.empty_like(xt)
    scratch = torch.empty((n + d + 8) * 8, dtype=torch.uint8, device="cuda")
    eng._check(lib.avb_standardize(h, xt.data_ptr(), 1, n, d, 1, xs.data_ptr(), scratch.data_ptr(), scratch.numel(), st))
    z = torch.empty(n, d, 128, device="cuda")
    eng._check(lib.avb_embed(h, xs.data_ptr(), it.data_ptr(), n * d, z.data_ptr(), st))
    G = 2
    ws = torch.empty(int(lib.avb_block_workspace_bytes(h, 1, n, d)), dtype=torch.uint8, device="cuda")
    for i in range(4):
        if i % 2 == 0:
            shards = [z[r * n // G:(r + 1) * n // G].contiguous() for r in range(G)]
            for s in shards:
                eng._check(lib.avb_block(h, i, s.data_ptr(), 1, n // G, d, 0, ws.data_ptr(), ws.numel(), st))
            z = torch.cat(shards, 0)
        else:
            shards = [z[:, r * d // G:(r + 1) * d // G].contiguous() for r in range(G)]
            for s in shards:
                eng._check(lib.avb_block(h, i, s.data_ptr(), 1, n, d // G, 1, ws.data_ptr(), ws.numel(), st))
            z = torch.cat(shards, 1)
    pooled = torch.empty(1, d, 128, device="cuda")
    eng._check(lib.avb_pool(h, z.contiguous().data_ptr(), 1, n, d, pooled.data_ptr(), 0, st))
    probs = torch.empty(d, d, device="cuda")
    hs = torch.empty(2 * d * 128 * 4, dtype=torch.uint8, device="cuda")
    eng._check(lib.avb_head(h, pooled.data_ptr(), 1, d, probs.data_ptr(), hs.data_ptr(), hs.numel(), st))
    torch.cuda.synchronize()
    assert (probs - full).abs().max().item() <= 1e-6


def test_edge_shapes_and_errors():
    kw = dict(layers=1, **STRESS)
    for dtype in ("fp32", "bf16"):
        model = avici_b200.random_init_model(kw, seed=0, perturb=True, compute_dtype=dtype)
        for n, d in ((1, 2), (2, 1), (5, 3), (129, 7), (7, 129)):
            rng = np.random.default_rng(n * 100 + d)
            x = rng.normal(size=(n, d))
            ref = oracle_probs(model.params, x, None, 1)
            p = model(x=x.astype(np.float32))
            if np.isfinite(ref).all():
                check_parity(p, ref, TOL_STRESS[dtype])
            else:  # d == 1 etc: the reference itself produces NaN-free output only where defined
                assert p.shape == ref.shape
        with pytest.raises(AssertionError, match="divisible"):
            model(x=np.zeros((10, 4), dtype=np.float32), experimental_chunk_size=3)
    eng = model.engine
    with pytest.raises(KeyError):
        bad = init_params(kw, seed=0)
        del bad["BaseModel/linear_2"]
        eng.load_params(bad)


def test_sample_graphs_and_logprobs():
    kw = dict(layers=1, **STRESS)
    model = avici_b200.random_init_model(kw, seed=0, perturb=True, compute_dtype="fp32")
    g, x, _ = avici_b200.simulate_data(d=12, n=50, seed=1, domain="lin-gauss")
    xs = O.standardize_default_simple(np.stack([x, np.zeros_like(x)], -1)).astype(np.float32)
    im = model._model
    p = im.infer_edge_probs(model.params, xs)
    ref = O.infer_edge_probs(O.cast_params(model.params, np.float64), xs.astype(np.float64), dict(O.DEFAULT_CFG, layers=1))
    check_parity(p, ref, 1e-4)
    lp = im.infer_edge_logprobs(model.params, None, xs)
    assert np.all(np.isneginf(np.diag(lp)))
    s = im.sample_graphs(2000, model.params, 0, xs)
    assert tuple(s.shape) == (2000, 12, 12) and s.diagonal(dim1=-2, dim2=-1).sum().item() == 0
    assert np.abs(s.float().mean(0).cpu().numpy() - p).max() < 0.06
