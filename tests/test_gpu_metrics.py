"""GPU parity of the on-device graph metrics against (1) outputs of the reference's own
avici/metrics.py committed under tests/golden/metrics/ and (2) the numpy oracle on seeded batches.  Integer outputs
must be bit-exact; precision / recall / f1 are ratios of those integers (<= 1 ulp); `cyclic` must be the reference's
answer, including the long cycles its fp64 matrix-power test does not see."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from avici_b200 import _lib
from avici_b200 import metrics as GM
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics", "metrics_cases.npz")


def _cases():
    z = np.load(GOLDEN)
    return [(str(n), z[f"{n}__probs"], z[f"{n}__true"], z[f"{n}__ref"]) for n in z["names"]]


def test_reference_golden_vectors(built_lib):
    checked = 0
    for name, probs, true, ref in _cases():
        m = GM.graph_metrics(probs, true, threshold=0.5)
        assert int(m["shd"]) == int(ref[0]), name
        assert int(m["edges"]) == int(ref[1]), name
        assert int(m["cyclic"]) == int(ref[2]), name
        assert float(m["precision"]) == pytest.approx(ref[3], abs=1e-15), name
        assert float(m["recall"]) == pytest.approx(ref[4], abs=1e-15), name
        assert float(m["f1"]) == pytest.approx(ref[5], abs=1e-15), name
        checked += 1
    assert checked >= 70


@pytest.mark.parametrize("B,d,dens", [(64, 100, 0.03), (7, 33, 0.2), (3, 1, 0.5), (130, 17, 0.08), (2, 257, 0.01)])
def test_batch_matches_oracle(built_lib, B, d, dens):
    rng = np.random.default_rng(B * 1000 + d)
    probs = rng.random((B, d, d)).astype(np.float32)
    probs = np.where(rng.random((B, d, d)) < dens, 0.5 + 0.5 * probs, 0.5 * probs).astype(np.float32)
    # half of the batch acyclic (upper-triangular support under a permutation), so both answers occur
    for b in range(0, B, 2):
        perm = rng.permutation(d)
        mask = np.triu(np.ones((d, d), bool), 1)[np.ix_(perm, perm)]
        probs[b] = np.where(mask, probs[b], 0.25 * probs[b])
    true = (rng.random((B, d, d)) < 2.0 / max(d, 2)).astype(np.int32)
    want = MO.graph_metrics(probs, true, 0.5)
    got = GM.graph_metrics(torch.from_numpy(probs).cuda(), torch.from_numpy(true).cuda(), 0.5)
    for k in ("shd", "edges", "cyclic"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("precision", "recall", "f1"):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-15, err_msg=k)
    if B >= 7 and d > 1:
        assert 0 < want["cyclic"].sum() < B
    # size-independent properties: shd(g, g) = 0; counts consistent
    same = GM.graph_metrics((probs > 0.5).astype(np.float32), (probs > 0.5).astype(np.int32))
    assert (same["shd"] == 0).all() and (same["fp"] == 0).all() and (same["fn"] == 0).all()
    assert np.array_equal(same["tp"], got["edges"])


def test_other_thresholds_and_reference_named_functions(built_lib):
    rng = np.random.default_rng(5)
    d = 24
    probs = rng.random((d, d)).astype(np.float32)
    true = (rng.random((d, d)) < 0.1).astype(np.int32)
    for thr in (0.1, 0.5, 0.9):
        pred = (probs > thr).astype(np.int32)
        m = GM.graph_metrics(probs, true, threshold=thr)
        assert int(m["shd"]) == int(MO.shd(true, pred)) and int(m["edges"]) == int(pred.sum())
        assert int(m["cyclic"]) == int(MO.is_cyclic(pred))
    pred = (probs > 0.8).astype(np.int32)
    assert int(GM.shd(true, pred)) == int(MO.shd(true, pred))
    assert int(GM.n_edges(pred)) == int(pred.sum())
    assert GM.is_cyclic(pred) == MO.is_cyclic(pred) and GM.is_acyclic(pred) == MO.is_acyclic(pred)
    cm, want = GM.classification_metrics(true, pred), MO.classification_metrics(true, pred)
    assert cm == pytest.approx(want, abs=1e-15)
    # batched shd keeps the leading shape (metrics.py:12-14)
    g = (rng.random((2, 3, 9, 9)) < 0.3).astype(np.int32)
    h = (rng.random((2, 3, 9, 9)) < 0.3).astype(np.int32)
    assert np.array_equal(GM.shd(g, h), MO.shd(g, h))


def test_threshold_metrics_reference_golden_and_oracle(built_lib):
    """auroc / auprc / ap (metrics.py:90-127): fp64 sums, tolerance 1e-12 against the reference's outputs."""
    z = np.load(GOLDEN)
    for n in z["names"]:
        n = str(n)
        m = GM.threshold_metrics(z[f"{n}__true"], z[f"{n}__probs"])
        np.testing.assert_allclose([m["auroc"], m["auprc"], m["ap"]], z[f"{n}__thr"], rtol=0, atol=1e-12, err_msg=n)
    rng = np.random.default_rng(11)
    for B, d, grid in ((64, 100, 0), (5, 33, 7), (3, 2, 3), (2, 130, 0)):
        true = (rng.random((B, d, d)) < 3.0 / max(d, 3)).astype(np.int32)
        probs = (0.3 * true + 0.7 * rng.random((B, d, d))).astype(np.float32)
        if grid:
            probs = (np.round(probs * grid) / grid).astype(np.float32)  # heavy ties
        true[0] = 0                       # no positives, predicted mass > 0 -> 0.5, 0, 0
        if B > 2:
            true[1] = 0; probs[1] = 0.0   # neither -> 1, 1, 1
        got = GM.threshold_metrics_batch(torch.from_numpy(true).cuda(), torch.from_numpy(probs).cuda())
        for b in range(B):
            want = MO.threshold_metrics(true[b], probs[b])
            for k in ("auroc", "auprc", "ap"):  # an all-positive truth has no ROC curve: nan on both sides (sklearn warns)
                np.testing.assert_allclose(got[k][b], want[k], rtol=0, atol=1e-12, equal_nan=True, err_msg=f"{B} {d} {b} {k}")
    # a perfect ranking: every positive above every negative
    true = (rng.random((20, 20)) < 0.1).astype(np.int32)
    m = GM.threshold_metrics(true, np.where(true, 0.9, 0.1).astype(np.float32))
    assert m == pytest.approx({"auroc": 1.0, "auprc": 1.0, "ap": 1.0}, abs=1e-15)


def test_c_abi_errors(built_lib):
    lib = built_lib
    p = torch.zeros((2, 4, 4), device="cuda")
    out = torch.zeros((2, 8), dtype=torch.int64, device="cuda")
    ws = torch.zeros(lib.avb_graph_metrics_workspace_bytes(2, 4), dtype=torch.uint8, device="cuda")
    assert lib.avb_graph_metrics(None, None, C.c_float(0.5), 2, 4, out.data_ptr(), ws.data_ptr(), ws.numel(), None) == _lib.AVB_ERR_INVALID
    assert lib.avb_graph_metrics(p.data_ptr(), None, C.c_float(0.5), 2, 4, out.data_ptr(), ws.data_ptr(), 16, None) == _lib.AVB_ERR_WORKSPACE
    assert b"workspace" in lib.avb_last_error(None)
    assert lib.avb_graph_metrics(p.data_ptr(), None, C.c_float(0.5), 2, 4, out.data_ptr(), ws.data_ptr(), ws.numel(), None) == _lib.AVB_OK
    torch.cuda.synchronize()
    assert out[:, :7].abs().sum().item() == 0  # empty prediction, no truth: every count 0, acyclic
    tout = torch.zeros((2, 4), dtype=torch.float64, device="cuda")
    assert lib.avb_threshold_metrics(p.data_ptr(), None, 2, 4, tout.data_ptr(), None) == _lib.AVB_ERR_INVALID
