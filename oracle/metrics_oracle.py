"""TEST INFRASTRUCTURE ONLY - numpy restatement of the thresholded graph metrics of `avici/metrics.py`.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(avici_b200/metrics.py -> avb_graph_metrics) never does.

PINNED: unlike the forward (whose jax/haiku dependencies are absent), `avici/metrics.py` needs only numpy and sklearn and
imports in this container, so tests/golden/metrics/metrics_cases.npz holds outputs of THE REFERENCE ITSELF
(tests/golden/metrics/make_metrics_golden.py, run here against /root/reference) and tests/test_oracle.py checks this restatement
against them.
"""
import numpy as onp


def shd(g, h):
    """avici/metrics.py:5-24: |g - h| + transpose, double edges count once, upper triangle (with diagonal) summed."""
    g, h = onp.asarray(g), onp.asarray(h)
    assert g.ndim == h.ndim
    diff = onp.abs(g.astype(onp.int64) - h.astype(onp.int64))
    both = diff + onp.swapaxes(diff, -1, -2)
    return onp.triu(onp.minimum(both, 1)).sum(axis=(-1, -2))


def n_edges(g):
    """avici/metrics.py:27-32"""
    return onp.asarray(g).sum(axis=(-1, -2))


def is_acyclic(g):
    """avici/metrics.py:35-43: trace((I + g/d)^d) - d isclose 0, fp64, numpy's matrix_power."""
    g = onp.asarray(g)
    d = g.shape[-1]
    powd = onp.linalg.matrix_power(onp.eye(d) + g / d, d)
    return bool(onp.isclose(onp.trace(powd) - d, 0.0))


def is_cyclic(g):
    """avici/metrics.py:46-51"""
    return not is_acyclic(g)


def classification_metrics(true, pred):
    """avici/metrics.py:54-87 with sklearn's binary precision / recall / f1 written out from the counts."""
    t = onp.asarray(true).reshape(-1) != 0
    p = onp.asarray(pred).reshape(-1) != 0
    if p.sum() > 0 and t.sum() > 0:
        tp = float((t & p).sum()); fp = float((~t & p).sum()); fn = float((t & ~p).sum())
        prec, rec = tp / (tp + fp), tp / (tp + fn)
        f1 = 2 * tp / (2 * tp + fp + fn) if tp > 0 else 0.0
        return {"precision": prec, "recall": rec, "f1": f1}
    if p.sum() == 0 and t.sum() == 0:
        return {"precision": 1.0, "recall": 1.0, "f1": 1.0}
    return {"precision": 0.0, "recall": 0.0, "f1": 0.0}


def graph_metrics(probs, true, threshold=0.5):
    """What the eval loop accumulates per dataset (avici/train.py:125-151) for a [B, d, d] batch."""
    probs, true = onp.asarray(probs), onp.asarray(true)
    pred = (probs > threshold).astype(onp.int32)
    out = {k: [] for k in ("shd", "edges", "cyclic", "precision", "recall", "f1")}
    for j in range(pred.shape[0]):
        out["shd"].append(int(shd(true[j], pred[j])))
        out["edges"].append(int(n_edges(pred[j])))
        out["cyclic"].append(int(is_cyclic(pred[j])))
        m = classification_metrics(true[j], pred[j])
        for k in ("precision", "recall", "f1"):
            out[k].append(m[k])
    return {k: onp.asarray(v) for k, v in out.items()}
