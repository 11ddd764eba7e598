"""Batched evaluation step on the device: forward of a batch of datasets, then the eval loop's metrics.

Mirror of the reference's `eval_batch_helper` (avici/train.py:51-69: edge probabilities and their thresholded graphs)
followed by the per-dataset metric loop of `evaluate` (avici/train.py:125-151), with the reference's scalar names
(`shd_0_5`, `edges_0_5`, `cyclic_0_5`, `precision_0_5`, `recall_0_5`, `f1_0_5`, `auroc`, `auprc`, `ap`; a threshold t
is spelled `f"{t}".replace(".", "_")`) and its averaging (mean over the batch).  The predictions never leave the GPU
between the forward and the metrics when `x` is a CUDA tensor.  The training loss the reference logs alongside
(`model.loss`, train.py:55-56) belongs to the training path and is not computed here.
"""
import numpy as onp

from . import metrics as _metrics


def threshold_key(thres):
    """train.py:63: f"{thres}".replace(".", "_")"""
    return f"{thres}".replace(".", "_")


def eval_batch(model, x, g, interv=None, decision_thresholds=(0.5,), experimental_chunk_size=None):
    """`x[B, n, d]`, true graphs `g[B, d, d]` -> (dict of batch-mean scalars, edge probabilities `[B, d, d]`).

    `model` is an `avici_b200.AVICIModel`; `x` / `interv` / `g` numpy arrays or torch tensors (CUDA tensors stay on
    the device end to end)."""
    assert x.ndim == 3 and g.ndim == 3, "`x` must be [B, n, d] and `g` [B, d, d]"  # train.py:91
    assert x.shape[0] == g.shape[0] and x.shape[2] == g.shape[1] == g.shape[2]
    probs = model.infer_batch(x, interv=interv, return_probs=True, experimental_chunk_size=experimental_chunk_size)
    scalars = {}
    for thres in decision_thresholds:
        key = threshold_key(thres)
        m = _metrics.graph_metrics(probs, g, threshold=thres)
        scalars[f"shd_{key}"] = float(onp.mean(m["shd"]))
        scalars[f"edges_{key}"] = float(onp.mean(m["edges"]))
        scalars[f"cyclic_{key}"] = float(onp.mean(m["cyclic"]))
        for k in ("precision", "recall", "f1"):
            scalars[f"{k}_{key}"] = float(onp.mean(m[k]))
    tm = _metrics.threshold_metrics_batch(g, probs)
    for k in ("auroc", "auprc", "ap"):
        scalars[k] = float(onp.mean(tm[k]))
    return scalars, probs
