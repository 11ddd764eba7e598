"""Generates tests/golden/metrics/metrics_cases.npz with outputs of THE REFERENCE's own `avici/metrics.py`.

That module needs only numpy and sklearn, both present in this container, so it is loaded straight from the
read-only reference tree by file path (importing the `avici` package would pull jax).  The vectors pin
oracle/metrics_oracle.py and, through it, the CUDA path (avb_graph_metrics).  The reference tree does not exist on the
GPU box, hence the committed file.  Run from the repo root:  python tests/golden/metrics/make_metrics_golden.py
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/avici/metrics.py"


def load_reference_metrics():
    spec = importlib.util.spec_from_file_location("ref_avici_metrics", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def random_dag(rng, d, p):
    perm = rng.permutation(d)
    g = np.triu((rng.random((d, d)) < p).astype(np.int32), 1)
    return g[np.ix_(perm, perm)][:, :].copy()


def single_cycle(d, length, rng, extra_dag_p=0.0):
    """A DAG (optionally) plus exactly one directed cycle through `length` nodes that keeps the rest acyclic."""
    g = np.zeros((d, d), np.int32)
    if extra_dag_p > 0:
        g = np.triu((rng.random((d, d)) < extra_dag_p).astype(np.int32), 1)
        g[:length, :] = 0
        g[:, :length] = 0
    for k in range(length):
        g[k, (k + 1) % length] = 1
    return g


def cases():
    rng = np.random.default_rng(20240501)
    out = []
    for d in (1, 2, 3, 4, 5, 7, 20, 50, 100, 130):
        for dens in (0.02, 0.1, 0.4):
            true = random_dag(rng, d, min(1.0, 2.0 / max(d, 2)))
            probs = np.where(rng.random((d, d)) < dens, 0.5 + 0.5 * rng.random((d, d)), 0.5 * rng.random((d, d))).astype(np.float32)
            np.fill_diagonal(probs, 0.0)
            out.append((f"rand_d{d}_p{dens}", probs, true))
    for d in (20, 100):
        true = random_dag(rng, d, 2.0 / d)
        # an acyclic prediction: a noisy copy of another DAG
        pred = random_dag(rng, d, 3.0 / d)
        out.append((f"dag_d{d}", pred.astype(np.float32), true))
        out.append((f"exact_d{d}", true.astype(np.float32), true))
        out.append((f"reversed_d{d}", true.T.astype(np.float32).copy(), true))
        out.append((f"empty_pred_d{d}", np.zeros((d, d), np.float32), true))
        out.append((f"empty_true_d{d}", pred.astype(np.float32), np.zeros((d, d), np.int32)))
        out.append((f"both_empty_d{d}", np.zeros((d, d), np.float32), np.zeros((d, d), np.int32)))
    # one directed cycle of a given length: the reference's fp64 test stops seeing it past ~12 edges at d = 100
    for d, lengths in ((100, (2, 3, 8, 11, 12, 13, 14, 20, 100)), (30, (2, 10, 20, 30)), (130, (12, 16, 17, 18, 40))):
        for L in lengths:
            out.append((f"cycle_d{d}_L{L}", single_cycle(d, L, rng).astype(np.float32), random_dag(rng, d, 2.0 / d)))
            out.append((f"cycle_dag_d{d}_L{L}", single_cycle(d, L, rng, extra_dag_p=0.05).astype(np.float32),
                        random_dag(rng, d, 2.0 / d)))
    # heavy score ties (probabilities on a coarse grid): the curve has one point per DISTINCT score
    for d in (12, 40):
        true = random_dag(rng, d, 3.0 / d)
        probs = (np.round(rng.random((d, d)) * 6) / 6).astype(np.float32)
        out.append((f"ties_d{d}", probs, true))
        out.append((f"ties_informative_d{d}", np.clip(np.round((0.5 * true + 0.6 * rng.random((d, d))) * 5) / 5, 0, 1).astype(np.float32), true))
    # self loops and bidirected edges (shd's diagonal / double-edge rules)
    d = 6
    g = np.zeros((d, d), np.int32); g[0, 1] = g[1, 0] = 1; g[2, 2] = 1; g[3, 4] = 1
    h = np.zeros((d, d), np.int32); h[0, 1] = 1; h[4, 3] = 1; h[5, 5] = 1
    out.append(("loops_bidirected", h.astype(np.float32), g))
    return out


def main():
    ref = load_reference_metrics()
    store = {}
    names = []
    for name, probs, true in cases():
        pred = (probs > 0.5).astype(np.int32)  # avici/train.py:63
        cm = ref.classification_metrics(true, pred)
        store[f"{name}__probs"] = probs
        store[f"{name}__true"] = true.astype(np.int32)
        store[f"{name}__ref"] = np.array([float(ref.shd(true, pred)), float(ref.n_edges(pred)), float(ref.is_cyclic(pred)),
                                          cm["precision"], cm["recall"], cm["f1"]], dtype=np.float64)
        tm = ref.threshold_metrics(true, probs)  # avici/metrics.py:90-127 (sklearn curves)
        store[f"{name}__thr"] = np.array([tm["auroc"], tm["auprc"], tm["ap"]], dtype=np.float64)
        names.append(name)
    store["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_cases.npz")
    np.savez_compressed(path, **store)
    cyc = {n: int(store[f"{n}__ref"][2]) for n in names if n.startswith("cycle_d")}
    print(f"{len(names)} cases -> {path} ({os.path.getsize(path)} bytes)")
    print("reference is_cyclic on single cycles:", cyc)


if __name__ == "__main__":
    main()
