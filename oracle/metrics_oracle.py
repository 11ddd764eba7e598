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


def _binary_clf_curve(t, s):
    """sklearn.metrics._ranking._binary_clf_curve (scikit-learn 1.9.0, the version in this image; un-vendored
    dependency of avici/metrics.py:101-104): stable descending sort, one point per distinct score."""
    order = onp.argsort(-s, kind="stable")  # sklearn: argsort(kind="mergesort")[::-1]; ties inside a score do not matter
    s, t = s[order], t[order]
    distinct = onp.where(onp.diff(s))[0]
    idx = onp.r_[distinct, t.size - 1]
    tps = onp.cumsum(t, dtype=onp.float64)[idx]
    fps = 1.0 + idx - tps
    return fps, tps


def threshold_metrics(true, pred):
    """avici/metrics.py:90-127: auroc = auc(roc_curve), auprc = auc(recall, precision), ap = average_precision_score."""
    t = (onp.asarray(true).reshape(-1) != 0).astype(onp.float64)
    s = onp.asarray(pred).reshape(-1).astype(onp.float64)
    if s.sum() > 0 and t.sum() > 0:
        fps, tps = _binary_clf_curve(t, s)
        # roc_curve: origin prepended; dropping collinear points (drop_intermediate) does not change the trapezoid area
        fpr = onp.r_[0.0, fps] / fps[-1] if fps[-1] > 0 else onp.full(fps.size + 1, onp.nan)
        tpr = onp.r_[0.0, tps] / tps[-1]
        auroc = float(onp.trapezoid(tpr, fpr))
        # precision_recall_curve: precision = tps / (tps + fps), recall = tps / P, reversed, (recall 0, precision 1) appended
        precision = onp.r_[(tps / (tps + fps))[::-1], 1.0]
        recall = onp.r_[(tps / tps[-1])[::-1], 0.0]
        auprc = float(-onp.trapezoid(precision, recall))  # auc() on a decreasing x
        ap = float(-onp.sum(onp.diff(recall) * precision[:-1]))
        return {"auroc": auroc, "auprc": auprc, "ap": ap}
    if s.sum() == 0 and t.sum() == 0:
        return {"auroc": 1.0, "auprc": 1.0, "ap": 1.0}
    return {"auroc": 0.5, "auprc": 0.0, "ap": 0.0}
