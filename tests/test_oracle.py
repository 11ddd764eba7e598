"""CPU tests of the oracle (oracle/avici_oracle.py): the self-checks SURVEY.md §8c derives from the
reference source (nothing else pins it: the reference ships no tests or vectors), the torch twin, and
the committed golden vectors."""
import glob
import os

import numpy as np
import pytest

from avici_b200.params import init_params, param_count
from avici_b200.synthetic import simulate_data
from oracle import avici_oracle as O

KW = dict(layers=1, cosine_temp_init=3.0, logit_bias_init=0.0)
CFG = dict(O.DEFAULT_CFG, layers=1)


@pytest.fixture(scope="module")
def setup():
    params = init_params(KW, seed=0, perturb=True, dtype=np.float64)
    g, x, interv = simulate_data(d=9, n=24, n_interv=8, seed=1, domain="lin-gauss")
    return params, x, interv


def test_param_count_matches_reference_architecture():
    # (vii) 4 269 442 fp32 parameters for the published shape
    assert param_count() == 4269442
    shapes = init_params(dict(layers=1), seed=0)
    assert "BaseModel/multi_head_attention_1/query" in shapes and "BaseModel/linear_4" in shapes
    assert shapes["BaseModel"]["learned_temp"].shape == (1, 1, 1)


def test_row_permutation_invariance(setup):
    params, x, interv = setup
    p = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    perm = np.random.default_rng(0).permutation(x.shape[0])
    p2 = O.avici_call(params, x[perm], CFG, interv=interv[perm].copy(), dtype=np.float64)
    assert np.abs(p - p2).max() < 1e-12


def test_variable_permutation_equivariance(setup):
    params, x, interv = setup
    p = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    perm = np.random.default_rng(1).permutation(x.shape[1])
    p2 = O.avici_call(params, x[:, perm], CFG, interv=interv[:, perm].copy(), dtype=np.float64)
    assert np.abs(p[np.ix_(perm, perm)] - p2).max() < 1e-12


def test_diag_zero_and_bounds(setup):
    params, x, interv = setup
    p = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    assert np.all(np.diag(p) == 0)
    off = p[~np.eye(p.shape[0], dtype=bool)]
    sig = lambda t: 1 / (1 + np.exp(-t))
    assert off.min() >= sig(0.0 - np.exp(3.0)) - 1e-12 and off.max() <= sig(0.0 + np.exp(3.0)) + 1e-12
    g = O.avici_call(params, x, CFG, interv=interv.copy(), return_probs=False, dtype=np.float64)
    assert g.dtype.kind == "i" and np.array_equal(g, (p > 0.5).astype(int))


def test_chunk_identity_and_semantics(setup):
    params, x, interv = setup
    p = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    for cs in (None, 0, x.shape[0], x.shape[0] + 5):  # (v): outside 0 < cs < n == unchunked
        assert np.array_equal(p, O.avici_call(params, x, CFG, interv=interv.copy(), experimental_chunk_size=cs,
                                              dtype=np.float64))
    pc = O.avici_call(params, x, CFG, interv=interv.copy(), experimental_chunk_size=8, dtype=np.float64)
    assert np.abs(pc - p).max() > 1e-6  # chunking changes the function (README.md:155-162)
    with pytest.raises(AssertionError):
        O.avici_call(params, x, CFG, experimental_chunk_size=5, dtype=np.float64)


def test_batched_equals_stacked(setup):
    params, x, interv = setup
    g2, x2, iv2 = simulate_data(d=9, n=24, n_interv=8, seed=2, domain="lin-gauss")
    xb, ib = np.stack([x, x2]), np.stack([interv, iv2])
    pb = O.call_main(O.cast_params(params, np.float64), xb, ib, CFG)
    for k, (xx, ii) in enumerate(((x, interv), (x2, iv2))):
        assert np.abs(pb[k] - O.avici_call(params, xx, CFG, interv=ii.copy(), dtype=np.float64)).max() < 1e-12


def test_default_interv_is_zero_channel(setup):
    params, x, _ = setup
    a = O.avici_call(params, x, CFG, dtype=np.float64)
    b = O.avici_call(params, x, CFG, interv=np.zeros_like(x), dtype=np.float64)
    assert np.array_equal(a, b)


def test_standardizers_against_numpy_formulas():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 20, 5)) * 4 + 1
    x[:, :, 2] = 7.0  # constant column: std 0 -> divide by 1
    xs = O.standardize_default_simple(np.stack([x, np.ones_like(x)], -1))
    assert np.allclose(xs[..., 0].mean(-2), 0, atol=1e-12)
    assert np.allclose(xs[..., 0][:, :, [0, 1, 3, 4]].std(-2), 1)
    assert np.all(xs[..., 0][:, :, 2] == 0) and np.all(xs[..., 1] == 1)
    c = rng.poisson(3.0, size=(2, 30, 6)).astype(float)
    c[0, 4] = 0  # all-zero row -> NaN -> 0
    cs = O.standardize_count_simple(np.stack([c, np.zeros_like(c)], -1))[..., 0]
    assert np.isfinite(cs).all() and cs.min() == 0.0 and np.all(cs[0, 4] == 0)
    lib = c.sum(-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = np.where(c == 0, np.nan, np.log2(c / (lib * 1e-6)))
    ref = (ref - np.nanmin(ref, axis=(1, 2), keepdims=True)) / np.nanstd(ref, axis=(1, 2), keepdims=True)
    assert np.allclose(cs, np.nan_to_num(ref, nan=0.0))


def test_fp32_tier_noise_floor(setup):
    params, x, interv = setup
    p64 = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    p32 = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float32)
    assert np.abs(p64 - p32).max() < 2e-5  # the fp32 restatement sits well inside the 1e-4 criterion


def test_torch_twin_matches_numpy_oracle(setup):
    import torch
    from oracle import avici_oracle_torch as OT
    params, x, interv = setup
    p = O.avici_call(params, x, CFG, interv=interv.copy(), dtype=np.float64)
    pt = OT.forward_probs(OT.to_torch(params, torch.float64), torch.as_tensor(x), torch.as_tensor(interv), CFG)
    assert np.abs(pt.numpy() - p).max() < 1e-10
    xb = torch.as_tensor(np.stack([x, x[::-1]]))
    pb = OT.forward_probs(OT.to_torch(params, torch.float64), xb, None, CFG)
    assert np.abs(pb[0].numpy() - pb[1].numpy()).max() < 1e-10  # batched + row-permutation invariant


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))))
def test_golden_vectors(path):
    d = np.load(path)
    params = init_params(dict(layers=int(d["layers"]), cosine_temp_init=3.0, logit_bias_init=0.0),
                         seed=int(d["param_seed"]), perturb=True, dtype=np.float64)
    cfg = dict(O.DEFAULT_CFG, layers=int(d["layers"]))
    interv = d["interv"] if d["interv"].size else None
    chunk = int(d["chunk"]) if int(d["chunk"]) > 0 else None
    p = O.avici_call(params, d["x"], cfg, interv=None if interv is None else interv.copy(),
                     expects_counts=bool(d["expects_counts"]), experimental_chunk_size=chunk, dtype=np.float64)
    assert np.abs(p - d["p"]).max() < 1e-12


# ---------------------------------------------------------------------------- graph metrics
def _metrics_cases():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics", "metrics_cases.npz")
    z = np.load(path)
    return [(str(n), z[f"{n}__probs"], z[f"{n}__true"], z[f"{n}__ref"]) for n in z["names"]]


def test_threshold_metrics_oracle_matches_reference_outputs():
    """auroc / auprc / ap: the sort-based numpy restatement of sklearn's curves against the reference's own
    threshold_metrics outputs (metrics.py:90-127, sklearn 1.9.0), corner cases included."""
    import os
    from oracle import metrics_oracle as MO
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics", "metrics_cases.npz"))
    corner = set()
    for n in z["names"]:
        n = str(n)
        m = MO.threshold_metrics(z[f"{n}__true"], z[f"{n}__probs"])
        got, want = np.array([m["auroc"], m["auprc"], m["ap"]]), z[f"{n}__thr"]
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13, err_msg=n)
        if tuple(want) in ((1.0, 1.0, 1.0), (0.5, 0.0, 0.0)):
            corner.add(tuple(want))
    assert len(corner) == 2  # both corner-case branches of metrics.py:112-127 are in the golden set


def test_metrics_oracle_matches_reference_outputs():
    """PINNED: tests/golden/metrics/metrics_cases.npz holds outputs of the reference's own avici/metrics.py
    (tests/golden/metrics/make_metrics_golden.py); the numpy restatement must reproduce every one of them."""
    from oracle import metrics_oracle as MO
    cases = _metrics_cases()
    assert len(cases) >= 70
    seen_missed_cycle = False
    for name, probs, true, ref in cases:
        pred = (probs > 0.5).astype(np.int32)
        assert int(MO.shd(true, pred)) == int(ref[0]), name
        assert int(MO.n_edges(pred)) == int(ref[1]), name
        assert int(MO.is_cyclic(pred)) == int(ref[2]), name
        cm = MO.classification_metrics(true, pred)
        assert (cm["precision"], cm["recall"], cm["f1"]) == pytest.approx(tuple(ref[3:6]), abs=1e-15), name
        if name.startswith("cycle_d100_L13"):
            seen_missed_cycle = int(ref[2]) == 0
    # the reference's matrix-power test does not see a lone 13-cycle at d = 100: the oracle (and the CUDA path) must
    # return the reference's answer, not the graph-theoretic one
    assert seen_missed_cycle


def test_metrics_oracle_shd_properties():
    from oracle import metrics_oracle as MO
    rng = np.random.default_rng(0)
    g = (rng.random((5, 12, 12)) < 0.2).astype(np.int32)
    h = (rng.random((5, 12, 12)) < 0.2).astype(np.int32)
    assert (MO.shd(g, g) == 0).all()
    assert (MO.shd(g, h) == MO.shd(h, g)).all()
    # a reversed edge counts once (metrics.py:8-9)
    a = np.zeros((4, 4), np.int32); a[0, 1] = 1
    assert MO.shd(a, a.T) == 1


def _standardize_cases():
    path = os.path.join(os.path.dirname(__file__), "golden", "standardize", "standardize_cases.npz")
    blob = np.load(path)
    return [(str(k), blob[f"{k}__x"], blob[f"{k}__y"], bool(blob[f"{k}__counts"])) for k in blob["names"]]


def test_standardizer_oracles_match_reference_outputs():
    """Rows a3 / a4 PINNED: the oracle's standardizers against outputs of the reference's own numpy twins
    (avici/utils/data.py:10-29,66-78, generated by tests/golden/standardize/make_standardize_golden.py), including
    constant / all-zero columns, all-zero rows (library size 0), an all-zero dataset and std == 0."""
    cases = _standardize_cases()
    assert len(cases) >= 40
    for name, x, want, counts in cases:
        x2 = np.stack([x, np.ones_like(x)], -1)  # channel 1 (here all ones) must come back untouched
        got = (O.standardize_count_simple if counts else O.standardize_default_simple)(x2)
        assert got.dtype == x.dtype and np.array_equal(got[..., 1], x2[..., 1]), name
        assert np.isfinite(got).all(), name
        tol = 1e-12 if x.dtype == np.float64 else 2e-6
        err = np.abs(got[..., 0] - want).max() if want.size else 0.0
        assert err <= tol * max(1.0, float(np.abs(want).max())), f"{name}: {err}"
