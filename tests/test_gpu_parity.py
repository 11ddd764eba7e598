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


def _attn_sublayer_reference(params, i, z, H=8, D=32):
    """delta = MHA(LN_q(z), LN_k(z), LN_v(z)) . Wo + bo of block i in fp32 torch (model.py:56-64), attention over axis -2."""
    nm = O.block_param_names(i)
    P = lambda m: {k: torch.as_tensor(np.asarray(v), device="cuda", dtype=torch.float32) for k, v in params[m].items()}  # noqa: E731
    ln = lambda x, p: torch.nn.functional.layer_norm(x, (128,), p["scale"], p["offset"], 1e-5)  # noqa: E731
    lin = lambda x, p: x @ p["w"] + p["b"]  # noqa: E731
    q = lin(ln(z, P(nm["ln_q"])), P(nm["mha"] + "/query"))
    k = lin(ln(z, P(nm["ln_k"])), P(nm["mha"] + "/key"))
    v = lin(ln(z, P(nm["ln_v"])), P(nm["mha"] + "/value"))
    B, n, d, _ = z.shape
    q, k, v = (t.reshape(B, n, d, H, D) for t in (q, k, v))
    w = torch.softmax(torch.einsum("bnthe,bnThe->bnhtT", q, k) / D ** 0.5, -1)
    o = torch.einsum("bnhtT,bnThe->bnthe", w, v).reshape(B, n, d, H * D)
    return lin(o, P(nm["mha"] + "/linear"))


@pytest.mark.parametrize("fuse_out", [1, 0])
@pytest.mark.parametrize("B,n,d", [(1, 8, 100), (2, 37, 50), (1, 300, 128), (1, 5, 3), (1, 200, 16), (1, 2000, 100),
                                   (1, 1, 1), (3, 149, 64), (1, 40, 113)])
def test_dattn_fused_sublayer_against_torch(B, n, d, fuse_out):
    """dattn_tc_kernel (LayerNorm + QKV + attention over d + out-projection in one kernel, one CTA per sequence) through
    its hook, both with the out-projection fused and as a separate launch: delta rows of every token, exactly once."""
    model = avici_b200.random_init_model(dict(layers=1), seed=3, perturb=True, compute_dtype="bf16")
    eng = model.engine
    g = torch.Generator(device="cuda").manual_seed(B * 7919 + n * 31 + d)
    z = torch.randn(B, n, d, 128, device="cuda", generator=g) * 2 + torch.randn(B, n, d, 1, device="cuda", generator=g)
    for bi in (0, 1):  # block 1's weights too (its own LayerNorm / projection parameters)
        delta = torch.full((B * n * d, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
        guard = C.c_int32(-1)
        eng._check(eng.lib.avb_test_dattn_bf16(eng.handle, bi, z.data_ptr(), B, n, d, fuse_out, delta.data_ptr(),
                                               C.byref(guard), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        ref = _attn_sublayer_reference(model.params, bi, z).reshape(B * n * d, 128)
        assert guard.value == 0
        assert torch.isfinite(delta.float()).all(), "rows missing or non-finite"
        err = (delta.float() - ref).abs().max().item()
        assert err <= 3e-2 * max(ref.abs().max().item(), 1.0), f"dattn err {err} (scale {ref.abs().max().item()})"


@pytest.mark.parametrize("B,n,d", [(1, 7, 50), (2, 33, 20), (1, 130, 64), (1, 5, 3), (1, 1, 1), (1, 301, 33), (3, 11, 42)])
def test_dattn_packed_sequences_equal_one_per_tile(B, n, d):
    """d <= 64: the fused d-axis kernel packs 128 // d sequences into one 128-row tile and masks the scores between
    different sequences.  Against the same kernel with one sequence per tile (every stage but the softmax mask is
    row-wise, so the two agree to the bf16 rounding of P) and against the torch reference; sequence counts that do and do
    not fill the last tile."""
    model = avici_b200.random_init_model(dict(layers=1), seed=5, perturb=True, compute_dtype="bf16")
    eng = model.engine
    lib, h = eng.lib, eng.handle
    g = torch.Generator(device="cuda").manual_seed(B * 131 + n * 7 + d)
    z = torch.randn(B, n, d, 128, device="cuda", generator=g) * 2 + torch.randn(B, n, d, 1, device="cuda", generator=g)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for pack in (1, 0):
        eng._check(lib.avb_test_set_dattn_pack(h, pack))
        delta = torch.full((B * n * d + 64, 128), float("nan"), device="cuda", dtype=torch.bfloat16)  # + 64 guard rows
        guard = C.c_int32(-1)
        eng._check(lib.avb_test_dattn_bf16(h, 0, z.data_ptr(), B, n, d, 1, delta.data_ptr(), C.byref(guard), st))
        torch.cuda.synchronize()
        assert guard.value == 0
        assert torch.isnan(delta[B * n * d:].float()).all(), "rows written past the last sequence"
        res[pack] = delta[:B * n * d].float()
    eng._check(lib.avb_test_set_dattn_pack(h, 1))
    ref = _attn_sublayer_reference(model.params, 0, z).reshape(B * n * d, 128)
    scale = max(ref.abs().max().item(), 1.0)
    assert torch.isfinite(res[1]).all(), "rows missing or non-finite"
    assert (res[1] - ref).abs().max().item() <= 3e-2 * scale
    assert (res[1] - res[0]).abs().max().item() <= 2e-2 * scale


def test_dattn_modes_agree_and_guard_falls_back():
    """The same forward with the d-axis sub-layer fully fused (default), fused up to the attention output, and unfused:
    all within bf16 tolerance of each other; with score magnitudes far outside the fast softmax's range the fused
    kernel raises its guard and the exact unfused fallback recomputes the block (finite, equal to the unfused path)."""
    g_, x, iv = avici_b200.simulate_data(d=100, n=150, n_interv=50, seed=11, domain="lin-gauss")
    x = x.astype(np.float32); iv = iv.astype(np.float32)
    for blow_up in (False, True):
        model = avici_b200.random_init_model(dict(layers=2, **STRESS), seed=5, perturb=True, compute_dtype="bf16")
        if blow_up:
            for i in range(4):
                for nm_ in ("query", "key"):
                    model.params[f"BaseModel/multi_head_attention{'' if i == 0 else '_' + str(i)}/{nm_}"]["w"] *= 12.0
        eng = model.engine
        ps = {}
        for mode in (1, 2, 0):
            eng._check(eng.lib.avb_test_set_dattn_mode(eng.handle, mode))
            ps[mode] = model(x=x, interv=iv.copy())
            assert np.isfinite(ps[mode]).all()
        eng._check(eng.lib.avb_test_set_dattn_mode(eng.handle, 1))
        tol = 3e-2
        assert np.abs(ps[1] - ps[0]).max() <= tol and np.abs(ps[2] - ps[0]).max() <= tol, \
            (blow_up, np.abs(ps[1] - ps[0]).max(), np.abs(ps[2] - ps[0]).max())
        if not blow_up:
            ref = oracle_probs(model.params, x, iv, 2)
            check_parity(ps[1], ref, TOL_STRESS["bf16"])


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
    # the same network under the reference's own head init (temp 2, bias -3: the north star's weights): the fixture's
    # STRESS p determines the cosine matrix, hence the REF_INIT p; held to the north-star tolerances 1e-4 / 1e-2
    with np.errstate(divide="ignore", invalid="ignore"):
        cos = (np.log(d["p"]) - np.log1p(-d["p"]) - STRESS["logit_bias_init"]) / np.exp(STRESS["cosine_temp_init"])
    np.fill_diagonal(cos, 0.0)
    model_ref = avici_b200.random_init_model(dict(layers=layers, **REF_INIT), seed=int(d["param_seed"]), perturb=True,
                                             expects_counts=counts, compute_dtype=dtype)
    p_ref = model_ref(x=d["x"].astype(np.float32), interv=interv, experimental_chunk_size=chunk)
    check_parity(p_ref, _probs_from_cos(cos, REF_INIT["cosine_temp_init"], REF_INIT["logit_bias_init"]), TOL[dtype])


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
            tol = tols[dtype]
            if dtype == "fp32" and head is STRESS:
                # The 5x-amplified head on fresh-init weights is ill-conditioned on this dataset (values up to 50): the
                # fp32 restatement of the reference itself sits 1.8e-4 from the fp64 one, and rounding x to fp32 alone
                # moves the fp64 result by 8e-5.  The 1e-4 criterion is the north star's for the reference's own head
                # (REF_INIT, asserted above without slack); here the bound is 1e-4 on top of the measured fp32 noise floor.
                floor = float(np.abs(oracle_probs(model.params, x.astype(np.float32), None, 8, dtype=torch.float32) - ref).max())
                tol = tol + 2.0 * floor
            err = check_parity(p, ref, tol)
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


# ------------------------------------------------------------------------------ BASELINE configs at their stated sizes
FULL = os.path.join(os.path.dirname(__file__), "golden", "full")


def _probs_from_cos(cos, temp, bias):
    """model.py:136-141 + 218-220, 239 applied to the oracle's cosine matrix (fp64)."""
    logits = cos * np.exp(temp) + bias
    p = np.exp(-(np.maximum(-logits, 0) + np.log1p(np.exp(-np.abs(logits)))))
    idx = np.arange(p.shape[-1])
    p[..., idx, idx] = 0.0
    return p


def _full_case(name):
    d = np.load(os.path.join(FULL, name + ".npz"))
    interv = d["interv"].astype(np.float32) if d["interv"].size else None
    return d["x"], interv, d["cos"], bool(d["expects_counts"]), int(d["param_seed"])


def _check_full(name, dtype, batched):
    """The CUDA path at a BASELINE config's stated size against the committed fp64 oracle output (tests/golden/
    make_full_size_golden.py), identical float32 inputs and weights, under the reference's own head init (north-star
    tolerances) and under the STRESS head (every cosine error x5 in p)."""
    x, interv, cos, counts, seed = _full_case(name)
    out = {}
    for tag, head, tols in (("ref-init", REF_INIT, TOL), ("stress", STRESS, TOL_STRESS)):
        model = avici_b200.random_init_model(dict(layers=8, **head), seed=seed, perturb=True, expects_counts=counts,
                                             compute_dtype=dtype)
        if batched:
            p = model.infer_batch(x, interv)
        else:
            p = np.stack([model(x=x[b], interv=None if interv is None else interv[b].copy()) for b in range(x.shape[0])])
        assert p.shape == cos.shape and np.isfinite(p).all()
        ref = _probs_from_cos(cos, head["cosine_temp_init"], head["logit_bias_init"])
        errs = [check_parity(p[b], ref[b], tols[dtype]) for b in range(x.shape[0])]
        out[tag] = max(errs)
        print(f"{name}[{dtype}, {tag}]: max|dp| = {max(errs):.2e} (tol {tols[dtype]})")
    return out


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_config2_stated_size_vs_oracle(dtype):
    """BASELINE config 2 (the bench workload): rff-gauss d=100 n=1000, 16 blocks, a batch of datasets."""
    _check_full("c2_rff_n1000_d100", dtype, batched=True)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_config3_stated_size_vs_oracle(dtype):
    """BASELINE config 3: lin-gauss d=100, 800 observational + 200 interventional rows, intervention channel."""
    _check_full("c3_lin_iv_n1000_d100", dtype, batched=False)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_config5_stated_size_vs_oracle(dtype):
    """BASELINE config 5: count data d=200 n=2000 through the count standardizer; fp32 <= 1e-4 is the config's own
    criterion, the bf16 path (d > 128: the unfused d-axis kernels) is held to its 1e-2 as well."""
    _check_full("c5_gene_n2000_d200", dtype, batched=False)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_config4_twin_axial_shards_vs_oracle(dtype):
    """BASELINE config 4's twin (n=640, d=256; the full n=5000, d=1000 is hours of oracle time): the axially sharded
    schedule - d-axis blocks on row shards, n-axis blocks on column shards, 4 shards processed in turn through the
    stage-level C-ABI in the layouts each rank of `avb_forward_sharded` uses (column-blocked n-shards) - against the
    ORACLE."""
    x, interv, cos, counts, seed = _full_case("c4twin_n640_d256")
    G = 4
    for head, tols in ((REF_INIT, TOL), (STRESS, TOL_STRESS)):
        model = avici_b200.random_init_model(dict(layers=8, **head), seed=seed, perturb=True, compute_dtype=dtype)
        eng = model.engine
        lib, h = eng.lib, eng.handle
        n, d = x.shape[1:]
        xt = torch.as_tensor(x[0]).cuda()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        xs = torch.empty_like(xt)
        scratch = torch.empty((n + d + 8) * 8, dtype=torch.uint8, device="cuda")
        eng._check(lib.avb_standardize(h, xt.data_ptr(), 1, n, d, 1, xs.data_ptr(), scratch.data_ptr(), scratch.numel(), st))
        z = torch.empty(n, d, 128, device="cuda")
        eng._check(lib.avb_embed(h, xs.data_ptr(), None, n * d, z.data_ptr(), st))
        ws = torch.empty(int(max(lib.avb_block_workspace_bytes(h, 1, n // G, d), lib.avb_block_workspace_bytes(h, 1, n, d // G))),
                         dtype=torch.uint8, device="cuda")
        from avici_b200 import sharded
        for i in range(16):
            if i % 2 == 0:  # the n-shard in column blocks [G][n/G][d/G][k], as avb_forward_sharded keeps it
                shards = [sharded.to_column_blocks(z[r * n // G:(r + 1) * n // G], G) for r in range(G)]
                for s_ in shards:
                    eng._check(lib.avb_block_sharded(h, i, s_.data_ptr(), n // G, d, d // G, ws.data_ptr(), ws.numel(), st))
                z = torch.cat([sharded.from_column_blocks(s_) for s_ in shards], 0)
            else:
                shards = [z[:, r * d // G:(r + 1) * d // G].contiguous() for r in range(G)]
                for s_ in shards:
                    eng._check(lib.avb_block(h, i, s_.data_ptr(), 1, n, d // G, 1, ws.data_ptr(), ws.numel(), st))
                z = torch.cat(shards, 1)
        pooled = torch.empty(1, d, 128, device="cuda")
        eng._check(lib.avb_pool(h, z.contiguous().data_ptr(), 1, n, d, pooled.data_ptr(), 0, st))
        probs = torch.empty(d, d, device="cuda")
        hs = torch.empty(2 * d * 128 * 4, dtype=torch.uint8, device="cuda")
        eng._check(lib.avb_head(h, pooled.data_ptr(), 1, d, probs.data_ptr(), hs.data_ptr(), hs.numel(), st))
        torch.cuda.synchronize()
        ref = _probs_from_cos(cos[0], head["cosine_temp_init"], head["logit_bias_init"])
        err = check_parity(probs.cpu().numpy(), ref, tols[dtype])
        # and the unsharded forward of the same dataset
        err1 = check_parity(model(x=x[0]), ref, tols[dtype])
        print(f"c4twin[{dtype}, temp={head['cosine_temp_init']}]: sharded max|dp| = {err:.2e}, unsharded {err1:.2e}")


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


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_block_sharded_column_blocks_bit_identical(dtype):
    """avb_block_sharded on a column-blocked n-shard == avb_block on the same rows in plain order, bit for bit (the
    layout only changes where a token lives).  The plain run is pinned to the unfused kernels, which is what the sharded
    block uses for every d (the fused d <= 128 kernel rounds at different points)."""
    from avici_b200 import sharded
    model = avici_b200.random_init_model(dict(layers=1, **STRESS), seed=2, perturb=True, compute_dtype=dtype)
    eng = model.engine
    lib, h = eng.lib, eng.handle
    eng._check(lib.avb_test_set_dattn_mode(h, 0))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(0)
    for rows, d, G in ((37, 144, 4), (16, 40, 2), (5, 1000, 8)):
        z0 = torch.randn(rows, d, 128, device="cuda", generator=g) * 1.5
        ws = torch.empty(int(lib.avb_block_workspace_bytes(h, 1, rows, d)), dtype=torch.uint8, device="cuda")
        plain = z0.clone()
        eng._check(lib.avb_block(h, 0, plain.data_ptr(), 1, rows, d, 0, ws.data_ptr(), ws.numel(), st))
        blk = sharded.to_column_blocks(z0, G)
        eng._check(lib.avb_block_sharded(h, 0, blk.data_ptr(), rows, d, d // G, ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        assert torch.equal(sharded.from_column_blocks(blk), plain), (rows, d, G)
    with pytest.raises(AssertionError, match="column blocks"):
        eng._check(lib.avb_block_sharded(h, 0, z0.data_ptr(), 5, 1000, 7, ws.data_ptr(), ws.numel(), st))


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_forward_sharded_world1_equals_forward(dtype):
    """avb_forward_sharded with one rank (no communicator needed) is avb_forward."""
    model = avici_b200.random_init_model(dict(layers=2, **STRESS), seed=5, perturb=True, compute_dtype=dtype)
    eng = model.engine
    lib, h = eng.lib, eng.handle
    g, x, iv = avici_b200.simulate_data(d=24, n=96, n_interv=32, seed=9, domain="lin-gauss")
    n, d = x.shape  # 128 rows: 96 observational + 32 interventional
    xt, it = torch.as_tensor(x, dtype=torch.float32).cuda(), torch.as_tensor(iv, dtype=torch.float32).cuda()
    full = model(x=xt, interv=it.clone())
    ws = torch.empty(int(lib.avb_sharded_workspace_bytes(h, n, d, 1)), dtype=torch.uint8, device="cuda")
    probs = torch.empty(d, d, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    eng._check(lib.avb_forward_sharded(h, xt.data_ptr(), it.data_ptr(), n, d, 1, probs.data_ptr(), ws.data_ptr(), ws.numel(),
                                       st, None, 0, 1))
    torch.cuda.synchronize()
    assert torch.equal(probs, full)
    assert lib.avb_sharded_workspace_bytes(h, n, d, 5) == 0  # 5 does not divide n
    with pytest.raises(AssertionError, match="divisible"):
        eng._check(lib.avb_forward_sharded(h, xt.data_ptr(), None, n, d, 1, probs.data_ptr(), ws.data_ptr(), ws.numel(), st,
                                           None, 0, 5))


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_axial_sharding_over_two_gpus(dtype):
    """One process, two GPUs (the reference's `devices=` + `shard_if_possible=True`): the axially sharded forward with
    its NCCL exchanges is bit-identical to the single-GPU forward; `infer_batch(devices=...)` spreads a batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    model = avici_b200.random_init_model(dict(layers=3, **STRESS), seed=6, perturb=True, compute_dtype=dtype)
    model.SHARD_MIN_CELLS = 0
    g, x, iv = avici_b200.simulate_data(d=136, n=320, n_interv=64, seed=3, domain="lin-gauss")
    x = x.astype(np.float32); iv = iv.astype(np.float32)
    one = model(x=x, interv=iv.copy(), devices=[0])
    two = model(x=x, interv=iv.copy(), devices=[0, 1])
    assert np.array_equal(one, two)
    xb = np.stack([x[:200, :50], x[100:300, 50:100], x[120:320, 86:136]])
    assert np.array_equal(model.infer_batch(xb, devices=[0, 1]), model.infer_batch(xb))


def test_f32_split_tensor_core_gemms_vs_simt_and_oracle():
    """fp32 mode: the split-precision (bf16 x 3) tensor-core projections / FFN against the SIMT fp32 GEMMs of the same
    handle and against the fp64 oracle, ragged token counts included (tiles of 128 rows)."""
    kw = dict(layers=3, **STRESS)
    model = avici_b200.random_init_model(kw, seed=11, perturb=True, compute_dtype="fp32")
    eng = model.engine
    for n, d, seed in ((150, 37, 0), (257, 64, 1), (64, 130, 2)):
        g, x, iv = avici_b200.simulate_data(d=d, n=n, n_interv=n // 4, seed=seed, domain="lin-gauss")
        x = x.astype(np.float32); iv = iv.astype(np.float32)
        ref = oracle_probs(model.params, x.astype(np.float64), iv, 3)
        eng._check(eng.lib.avb_test_set_f32_tc(eng.handle, 1))
        p_tc = model(x=x, interv=iv.copy())
        eng._check(eng.lib.avb_test_set_f32_tc(eng.handle, 0))
        p_simt = model(x=x, interv=iv.copy())
        eng._check(eng.lib.avb_test_set_f32_tc(eng.handle, 1))
        e_tc, e_simt = float(np.abs(p_tc - ref).max()), float(np.abs(p_simt - ref).max())
        print(f"fp32 n={n + n // 4} d={d}: split tensor-core max|dp| = {e_tc:.2e}, SIMT {e_simt:.2e}")
        check_parity(p_tc, ref, TOL_STRESS["fp32"])
        check_parity(p_simt, ref, TOL_STRESS["fp32"])
    # scores far beyond the +-100 bound the fast path relies on: the projection's recorded norms must send the attention
    # kernel down its exact online-softmax path (row maximum, rescaled accumulators)
    params = {k: dict(v) for k, v in model.params.items()}
    for i in range(6):
        sfx = "" if i == 0 else f"_{i}"
        for nm in ("query", "key"):
            mod = f"BaseModel/multi_head_attention{sfx}/{nm}"
            params[mod] = {"w": params[mod]["w"] * 3.5, "b": params[mod]["b"] * 3.5}
    big = avici_b200.AVICIModel(params=params, model=model._model, expects_counts=False)
    g, x, iv = avici_b200.simulate_data(d=37, n=150, n_interv=37, seed=0, domain="lin-gauss")
    x = x.astype(np.float32); iv = iv.astype(np.float32)
    ref = oracle_probs(params, x.astype(np.float64), iv, 3)
    p_big = big(x=x, interv=iv.copy())
    eng_big = big.engine
    eng_big._check(eng_big.lib.avb_test_set_f32_tc(eng_big.handle, 0))
    p_big_simt = big(x=x, interv=iv.copy())
    e_tc, e_simt = float(np.abs(p_big - ref).max()), float(np.abs(p_big_simt - ref).max())
    print(f"fp32 split, scores x12: max|dp| = {e_tc:.2e}, SIMT {e_simt:.2e}")
    # sharply peaked softmax: the fp32 restatement of the reference itself drifts from the fp64 one on this model, so both
    # kernels are held to 1e-4 on top of that measured noise floor
    floor = float(np.abs(oracle_probs(params, x, iv, 3, dtype=torch.float32) - ref).max())
    print(f"  fp32 oracle vs fp64 oracle on this model: {floor:.2e}")
    check_parity(p_big_simt, ref, 1e-4 + 2 * floor)
    # the split products carry ~1e-5 relative error per score against 1e-7 in true fp32: with |s| in the hundreds that is
    # the same size as the fp32 noise of the whole forward, hence twice the allowance
    check_parity(p_big, ref, 1e-4 + 4 * floor)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_forward_captures_as_a_cuda_graph(dtype):
    """avb_forward only enqueues on the caller's stream: a CUDA graph of it replays bit-identically on refilled inputs."""
    model = avici_b200.random_init_model(dict(layers=2, **STRESS), seed=8, perturb=True, compute_dtype=dtype)
    eng = model.engine
    data = [avici_b200.simulate_data(d=40, n=120, n_interv=40, seed=s, domain="lin-gauss") for s in range(2)]
    xs = [torch.as_tensor(t[1], dtype=torch.float32).cuda()[None] for t in data]
    ivs = [torch.as_tensor(t[2], dtype=torch.float32).cuda()[None] for t in data]
    x, iv = xs[0].clone(), ivs[0].clone()
    graph, out = eng.capture(x, iv)
    for k in (1, 0):
        x.copy_(xs[k]); iv.copy_(ivs[k])
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eng.forward(xs[k], ivs[k]))


def test_shape_sweep_bf16_vs_fp32_paths():
    """Tile-boundary shapes through BOTH kernel families (bf16: fused d-axis kernel for d <= 128 / unfused above, two-stream
    attention, CTA-pair QKV and FFN; fp32: split-precision GEMMs and attention): the two share no kernel, so agreement to
    the bf16 tolerance on every shape cross-checks the layouts of each (ragged tiles, d = 112 / 113 / 127 / 128 / 129,
    n around multiples of 128, batches)."""
    kw = dict(layers=2, **REF_INIT)
    m16 = avici_b200.random_init_model(kw, seed=21, perturb=True, compute_dtype="bf16")
    m32 = avici_b200.random_init_model(kw, seed=21, perturb=True, compute_dtype="fp32")
    rng = np.random.default_rng(5)
    shapes = [(1, 127, 112), (1, 129, 113), (2, 256, 127), (1, 255, 128), (1, 257, 129), (3, 64, 16), (1, 1, 5), (2, 385, 33),
              (1, 640, 130), (1, 100, 257), (4, 130, 100), (1, 1000, 7)]
    for B, n, d in shapes:
        x = rng.normal(size=(B, n, d)).astype(np.float32) * rng.uniform(0.5, 3.0, size=(1, 1, d)).astype(np.float32)
        iv = (rng.random(size=(B, n, d)) < 0.1).astype(np.float32)
        p16 = m16.infer_batch(x, iv)
        p32 = m32.infer_batch(x, iv)
        assert np.isfinite(p16).all() and np.isfinite(p32).all(), (B, n, d)
        err = float(np.abs(p16 - p32).max())
        assert err <= TOL["bf16"], f"shape {(B, n, d)}: bf16 vs fp32 path max|dp| = {err:.3e}"
        assert np.all(np.diagonal(p16, axis1=1, axis2=2) == 0)


def test_head_contraction_on_tensor_cores_matches_simt():
    """bf16 handles run the head's u . v^T on the tensor cores (split-precision operands): against the SIMT head of the same
    handle on identical pooled features, probabilities and log-probabilities, d below / at / above the 128 tile."""
    model = avici_b200.random_init_model(dict(layers=1, **STRESS), seed=13, perturb=True, compute_dtype="bf16")
    eng = model.engine
    lib, h = eng.lib, eng.handle
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(3)
    for B, d in ((3, 100), (1, 128), (2, 300), (1, 1)):
        pooled = torch.randn(B, d, 128, device="cuda", generator=g)
        hs = torch.empty(2 * B * d * 128 * 4, dtype=torch.uint8, device="cuda")
        res = {}
        for tc in (1, 0):
            eng._check(lib.avb_test_set_head_tc(h, tc))
            p = torch.full((B, d, d), float("nan"), device="cuda")
            lp = torch.full((B, d, d), float("nan"), device="cuda")
            eng._check(lib.avb_head(h, pooled.data_ptr(), B, d, p.data_ptr(), hs.data_ptr(), hs.numel(), st))
            eng._check(lib.avb_head_logprobs(h, pooled.data_ptr(), B, d, lp.data_ptr(), hs.data_ptr(), hs.numel(), st))
            torch.cuda.synchronize()
            res[tc] = (p, lp)
        eng._check(lib.avb_test_set_head_tc(h, 1))
        assert (res[1][0] - res[0][0]).abs().max().item() <= 2e-4, (B, d)  # 20 x cos: 1e-5 of split precision
        fin = torch.isfinite(res[0][1])
        assert torch.equal(fin, torch.isfinite(res[1][1]))
        assert (res[1][1][fin] - res[0][1][fin]).abs().max().item() <= 1e-3 if fin.any() else True
        assert torch.all(torch.diagonal(res[1][0], dim1=1, dim2=2) == 0)


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
    ref_lp = O.infer_edge_logprobs(O.cast_params(model.params, np.float64), xs.astype(np.float64), dict(O.DEFAULT_CFG, layers=1))
    off = ~np.eye(12, dtype=bool)
    assert np.abs(lp[off] - ref_lp[off]).max() <= 2e-3  # log_sigmoid(logits), model.py:218; logits within ~1e-3 in fp32
    # a head that saturates: p underflows in fp32 but log_sigmoid keeps its relative precision (reference semantics)
    sat = avici_b200.random_init_model(dict(layers=1, cosine_temp_init=6.0, logit_bias_init=-200.0), seed=0, perturb=True,
                                       compute_dtype="fp32")
    lps = sat._model.infer_edge_logprobs(sat.params, None, xs)
    ref_s = O.infer_edge_logprobs(O.cast_params(sat.params, np.float64), xs.astype(np.float64), dict(O.DEFAULT_CFG, layers=1))
    assert np.isfinite(lps[off]).all() and np.all(lps[off] < -80)  # exp() of these is 0 in fp32: log(p) would be -inf
    assert np.abs(lps[off] / ref_s[off] - 1).max() <= 1e-3
    s = im.sample_graphs(2000, model.params, 0, xs)
    assert tuple(s.shape) == (2000, 12, 12) and s.diagonal(dim1=-2, dim2=-1).sum().item() == 0
    assert np.abs(s.float().mean(0).cpu().numpy() - p).max() < 0.06
