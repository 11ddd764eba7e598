"""numpy restatement of the reference's example data simulator (benchmark / test input generator).

Follows `avici.utils.example.simulate_data` (avici/utils/example.py:10-73) and the pieces of
`avici/synthetic/*` it reaches for the example domains (config/examples/{lin-gauss,rff-gauss,
gene-ecoli}.yaml): a random DAG with ~2 edges per variable (Erdos-Renyi graph.py:208-222 or a
preferential-attachment scale-free graph restating graph.py:240-246 without igraph), linear or
random-Fourier-feature additive-noise mechanisms (linear.py:45-82, rff.py:60-114, utils.py:10-31),
ancestral sampling with single-node interventions (utils.py:35-151) and the final row permutation
(example.py:63-64).  The RNG stream is NOT bit-identical to the reference (igraph / call order);
the value distributions are.  `gene-ecoli` is a count-data surrogate (negative-binomial expression
with dropout and library-size variation) standing in for the vendored SERGIO simulator.
"""
import math

import numpy as onp

DOMAINS = ("lin-gauss", "rff-gauss", "gene-ecoli")


def _signed_uniform(rng, low, high, shape):
    sign = rng.choice([-1.0, 1.0], size=shape)
    return sign * rng.uniform(low, high, size=shape)


def erdos_renyi(rng, n_vars, edges_per_var=2):
    """graph.py:208-222."""
    n_edges = edges_per_var * n_vars
    p = min(n_edges / ((n_vars * (n_vars - 1)) / 2), 0.99)
    mat = rng.binomial(n=1, p=p, size=(n_vars, n_vars)).astype(int)
    dag = onp.tril(mat, k=-1)
    perm = rng.permutation(onp.eye(n_vars).astype(int))
    return perm.T @ dag @ perm


def scale_free(rng, n_vars, edges_per_var=2, power=1.0):
    """Directed Barabasi-Albert: node t attaches `edges_per_var` edges to earlier nodes with
    probability ~ (in-degree^power + 1); then vertices are permuted (graph.py:240-246)."""
    mat = onp.zeros((n_vars, n_vars), dtype=int)
    indeg = onp.zeros(n_vars)
    for t in range(1, n_vars):
        m = min(edges_per_var, t)
        w = indeg[:t] ** power + 1.0
        tgt = rng.choice(t, size=m, replace=False, p=w / w.sum())
        mat[t, tgt] = 1
        indeg[tgt] += 1
    perm = rng.permutation(n_vars)
    out = onp.zeros_like(mat)
    out[onp.ix_(perm, perm)] = mat
    return out


def _toporder(g):
    """Kahn's algorithm on adjacency g[i,j] = 1 iff i -> j (utils/graph.py:14-17)."""
    g = g.copy()
    indeg = g.sum(0)
    order, ready = [], [int(j) for j in onp.where(indeg == 0)[0]]
    while ready:
        i = ready.pop()
        order.append(i)
        for j in onp.where(g[i])[0]:
            indeg[j] -= 1
            if indeg[j] == 0:
                ready.append(int(j))
    assert len(order) == g.shape[0], "graph is not a DAG"
    return order


def _sample_scm(rng, g, n_obs, n_int, mech, noise_scale, interv_low=1.0, interv_high=5.0):
    """Ancestral sampling, observational rows then one block of interventional rows per target
    node, interleaved like utils.py:118-134; returns x[n,d], mask[n,d]."""
    d = g.shape[0]
    topo = _toporder(g)
    targets = [None] * (n_obs > 0)
    if n_int > 0:
        targets += sorted(rng.choice(d, size=d, replace=False).tolist())  # n_interv_vars = -1: all nodes
    xs, ms = {"obs": [], "int": []}, {"obs": [], "int": []}
    for tgt in targets:
        rows = n_obs if tgt is None else math.ceil(n_int / d)
        x = onp.zeros((rows, d))
        for j in topo:
            z = rng.normal(size=rows) * noise_scale[j]
            if tgt is not None and j == tgt:
                x[:, j] = _signed_uniform(rng, interv_low, interv_high, rows)
            else:
                x[:, j] = mech(j, x) + z
        mask = onp.zeros((rows, d), dtype=onp.float32)
        if tgt is not None:
            mask[:, tgt] = 1.0
        key = "obs" if tgt is None else "int"
        xs[key].append(x)
        ms[key].append(mask)
    parts_x, parts_m = [], []
    if xs["obs"]:
        parts_x.append(xs["obs"][0][:n_obs]); parts_m.append(ms["obs"][0][:n_obs])
    if xs["int"]:
        xi = onp.stack(xs["int"]).reshape(-1, d, order="F")[:n_int]
        mi = onp.stack(ms["int"]).reshape(-1, d, order="F")[:n_int]
        parts_x.append(xi); parts_m.append(mi)
    return onp.concatenate(parts_x, 0), onp.concatenate(parts_m, 0)


def simulate_data(d, n, *, n_interv=0, seed=0, domain="rff-gauss", graph="auto"):
    """Same signature and return convention as the reference helper: `(g[d,d] int, x[n+n_interv,d],
    interv[n+n_interv,d] or None)`; `x` is raw (un-standardized), rows permuted."""
    if domain not in DOMAINS:
        raise KeyError(f"unknown domain `{domain}`; implemented: {DOMAINS}")
    rng = onp.random.default_rng(onp.random.SeedSequence(entropy=seed))
    kind = graph if graph != "auto" else ("scale-free" if rng.uniform() < 0.5 else "erdos-renyi")
    if kind == "scale-free":
        g = scale_free(rng, d, 2, power=float(rng.choice([0.5, 1.0, 1.5])))
        if rng.uniform() < 0.5:
            g = g.T.copy()  # ScaleFreeTranspose
    else:
        g = erdos_renyi(rng, d, 2)
    noise_scale = rng.uniform(0.2, 2.0, size=d)
    bias = rng.uniform(-3.0, 3.0, size=d)

    if domain == "lin-gauss":
        w = _signed_uniform(rng, 1.0, 3.0, (d, d)) * g  # w[i,j]: weight of parent i in node j

        def mech(j, x):
            return x @ w[:, j] + bias[j]
    elif domain == "rff-gauss":
        n_rff = 100
        par = [onp.where(g[:, j])[0] for j in range(d)]
        ls = rng.uniform(7.0, 10.0, size=d)
        c = rng.uniform(10.0, 20.0, size=d)
        omega = [rng.normal(0.0, 1.0 / ls[j], size=(len(par[j]), n_rff)) for j in range(d)]
        b = rng.uniform(0, 2 * onp.pi, size=(d, n_rff))
        wr = rng.normal(size=(d, n_rff))

        def mech(j, x):
            phi = onp.cos(x[:, par[j]] @ omega[j] + b[j])
            return math.sqrt(2.0) * c[j] * (phi @ wr[j]) / math.sqrt(n_rff) + bias[j]
    else:  # gene-ecoli surrogate: counts
        w = rng.uniform(0.5, 1.5, size=(d, d)) * g * rng.choice([-1.0, 1.0], size=(d, d))
        base = rng.uniform(0.5, 3.0, size=d)
        topo = _toporder(g)
        n_tot = n + n_interv
        logexp = onp.zeros((n_tot, d))
        mask = onp.zeros((n_tot, d), dtype=onp.float32)
        if n_interv:
            tg = rng.integers(0, d, size=n_interv)
            mask[n + onp.arange(n_interv), tg] = 1.0
        for j in topo:
            logexp[:, j] = base[j] + 0.3 * onp.tanh(logexp @ w[:, j]) + 0.3 * rng.normal(size=n_tot)
        logexp = onp.where(mask > 0, -2.0, logexp)  # knock-down interventions
        lib = rng.lognormal(0.0, 0.3, size=(n_tot, 1))
        mean = onp.exp(logexp) * lib * 5.0
        counts = rng.poisson(mean * rng.gamma(2.0, 0.5, size=mean.shape))
        dropout = rng.uniform(size=mean.shape) < 1.0 / (1.0 + onp.exp(1.5 * (onp.log1p(mean) - 1.0)))
        x = onp.where(dropout, 0, counts).astype(onp.float64)
        perm = rng.permutation(n_tot)
        return g.astype(int), x[perm], (mask[perm] if n_interv else None)

    x, mask = _sample_scm(rng, g, n, n_interv, mech, noise_scale)
    perm = rng.permutation(x.shape[0])
    x, mask = x[perm], mask[perm]
    return g.astype(int), x, (mask if n_interv else None)
