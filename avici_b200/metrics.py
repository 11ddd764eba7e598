"""Graph metrics over batches of predictions that are already on the GPU.

Host-side mirror of `avici/metrics.py` for the metrics the eval loop computes per dataset at a decision threshold
(`avici/train.py:125-151`): `shd` (metrics.py:5-24), `n_edges` (:27-32), `is_acyclic` / `is_cyclic` (:35-51),
`classification_metrics` (:54-87).  Same names, argument meaning and corner-case values as the reference, but one
call (`graph_metrics`) covers the whole `[B, d, d]` batch with four kinds of kernel launches through the C-ABI entry point
`avb_graph_metrics`; nothing is computed on the host except turning tp / fp / fn counts into precision / recall / f1.

`threshold_metrics` (auroc / auprc / ap, metrics.py:90-127: sklearn's roc_curve, precision_recall_curve, auc,
average_precision_score) runs on the device too (`avb_threshold_metrics`), written as sort-free sums over the positive
entries; it agrees with sklearn to ~1e-14 (fp64 sums), not bit for bit.  Not here: the calibration statistics (:130-220).

There is no CPU fallback: every function needs the in-tree CUDA extension and an sm_100 device.
"""
import ctypes as C

import numpy as onp

from . import _lib

FIELDS = ("shd", "edges", "edges_true", "tp", "fp", "fn", "cyclic")


def _torch():
    import torch
    return torch


def _as_cuda(a, dtype):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("avici_b200.metrics: no CUDA device; there is no CPU fallback")
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(onp.ascontiguousarray(a))
    return t.to(device="cuda", dtype=dtype).contiguous()


def graph_metrics(pred, true=None, threshold=0.5):
    """Per-dataset metrics of `pred > threshold` against `true`, for `[..., d, d]` batches.

    pred: probabilities (or a 0/1 graph) as numpy array or torch tensor; true: 0/1 graph of the same shape or None.
    Returns a dict of numpy int64 arrays of shape `[...]` with keys FIELDS (`edges` = n_edges of the prediction,
    `cyclic` in {0, 1}), plus float64 `precision`, `recall`, `f1` with the reference's corner cases
    (metrics.py:71-87: both empty -> 1.0, exactly one empty -> 0.0).
    """
    torch = _torch()
    lib = _lib.load()
    p = _as_cuda(pred, torch.float32)
    assert p.ndim >= 2 and p.shape[-1] == p.shape[-2], "pred must be [..., d, d]"
    lead, d = tuple(p.shape[:-2]), int(p.shape[-1])
    B = int(onp.prod(lead)) if lead else 1
    g = None
    if true is not None:
        g = _as_cuda(true, torch.int32)
        assert tuple(g.shape) == tuple(p.shape), "true and pred must have the same shape"  # metrics.py:16
    out = torch.empty((max(B, 1), 8), dtype=torch.int64, device="cuda")
    res = {}
    if B > 0:
        # the C-ABI takes at most 65535 datasets per call
        step = 65535
        p2 = p.reshape(B, d, d)
        g2 = g.reshape(B, d, d) if g is not None else None
        for b0 in range(0, B, step):
            bc = min(step, B - b0)
            nbytes = lib.avb_graph_metrics_workspace_bytes(bc, d)
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            rc = lib.avb_graph_metrics(p2[b0:].data_ptr(), g2[b0:].data_ptr() if g2 is not None else None,
                                       C.c_float(float(threshold)), bc, d, out[b0:].data_ptr(), ws.data_ptr(), nbytes,
                                       torch.cuda.current_stream().cuda_stream)
            _lib.check(lib, None, rc)
    host = out[:B].cpu().numpy()
    for k, name in enumerate(FIELDS):
        res[name] = host[:, k].reshape(lead).copy()
    tp, fp, fn = (res[k].astype(onp.float64) for k in ("tp", "fp", "fn"))
    n_pred, n_true = tp + fp, tp + fn
    both = (n_pred > 0) & (n_true > 0)
    none = (n_pred == 0) & (n_true == 0)
    with onp.errstate(divide="ignore", invalid="ignore"):
        prec = onp.where(both, tp / n_pred, onp.where(none, 1.0, 0.0))
        rec = onp.where(both, tp / n_true, onp.where(none, 1.0, 0.0))
        # sklearn's binary f1 = 2 tp / (2 tp + fp + fn), 0 when tp = 0
        f1 = onp.where(both, onp.where(tp > 0, 2 * tp / (2 * tp + fp + fn), 0.0), onp.where(none, 1.0, 0.0))
    res.update(precision=prec, recall=rec, f1=f1)
    return res


def shd(g, h):
    """Structural Hamming distance of `[..., d, d]` 0/1 graphs (metrics.py:5-24)."""
    assert onp.ndim(g) == onp.ndim(h)
    return graph_metrics(h, g)["shd"][()]


def n_edges(g):
    """Number of edges (metrics.py:27-32)."""
    return graph_metrics(g)["edges"][()]


def is_acyclic(g):
    """The reference's trace-of-matrix-power test for one `[d, d]` graph (metrics.py:35-43)."""
    return not bool(graph_metrics(g)["cyclic"][()])


def is_cyclic(g):
    """metrics.py:46-51"""
    return not is_acyclic(g)


def classification_metrics(true, pred):
    """precision / recall / f1 of ONE flattened prediction (metrics.py:54-87)."""
    t = onp.asarray(true.cpu() if hasattr(true, "cpu") else true)
    d2 = int(t.size)
    # any shape is flattened by the reference; lay it out as a [1, d2] x ... matrix is unnecessary: counts are
    # elementwise, so a square zero-padded container gives the same tp / fp / fn
    side = int(onp.ceil(onp.sqrt(max(d2, 1))))
    tt = onp.zeros(side * side, dtype=onp.int32)
    pp = onp.zeros(side * side, dtype=onp.float32)
    tt[:d2] = t.reshape(-1)
    pp[:d2] = onp.asarray(pred.cpu() if hasattr(pred, "cpu") else pred).reshape(-1)
    m = graph_metrics(pp.reshape(side, side), tt.reshape(side, side))
    return {"precision": float(m["precision"]), "recall": float(m["recall"]), "f1": float(m["f1"])}


def threshold_metrics_batch(true, pred):
    """auroc / auprc / ap per dataset for `[..., d, d]` batches: dict of float64 arrays of shape `[...]`."""
    torch = _torch()
    lib = _lib.load()
    p = _as_cuda(pred, torch.float32)
    g = _as_cuda(true, torch.int32)
    assert p.ndim >= 2 and p.shape[-1] == p.shape[-2], "pred must be [..., d, d]"
    assert tuple(g.shape) == tuple(p.shape), "true and pred must have the same shape"
    lead, d = tuple(p.shape[:-2]), int(p.shape[-1])
    B = int(onp.prod(lead)) if lead else 1
    out = torch.empty((max(B, 1), 4), dtype=torch.float64, device="cuda")
    if B > 0:
        rc = lib.avb_threshold_metrics(p.data_ptr(), g.data_ptr(), B, d, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
        _lib.check(lib, None, rc)
    host = out[:B].cpu().numpy()
    return {"auroc": host[:, 0].reshape(lead).copy(), "auprc": host[:, 1].reshape(lead).copy(),
            "ap": host[:, 2].reshape(lead).copy()}


def threshold_metrics(true, pred):
    """auroc / auprc / ap of ONE flattened prediction (metrics.py:90-127)."""
    t = onp.asarray(true.cpu() if hasattr(true, "cpu") else true).reshape(-1)
    s = onp.asarray(pred.cpu() if hasattr(pred, "cpu") else pred).reshape(-1)
    n = int(t.size)
    side = int(onp.ceil(onp.sqrt(max(n, 1))))
    if side * side != n:
        # the C-ABI takes square [d, d] inputs; a flattened array of another length is padded with entries that are
        # neither positives nor able to outrank anything: label 0, score -inf.  They would still count as negatives,
        # so only exact squares are accepted instead of silently changing the metric.
        raise AssertionError("threshold_metrics: input must have d*d entries")
    m = threshold_metrics_batch(t.reshape(side, side), s.reshape(side, side))
    return {"auroc": float(m["auroc"]), "auprc": float(m["auprc"]), "ap": float(m["ap"])}
