"""numpy restatement of the reference's example data simulator (benchmark / test input generator).

`simulate_data(d, n, n_interv=0, seed=0, domain=..., path=...)` mirrors `avici.utils.example.simulate_data`
(avici/utils/example.py:10-73) for all seven example domains and for user YAML domain files in the reference's
configuration DSL (avici/utils/parse.py:15-103, avici/config/examples/*.yaml):

    reference                                              here
    -----------------------------------------------------  -----------------------------------------
    synthetic/distributions.py:4-63                        Gaussian, Laplace, Cauchy, Uniform, SignedUniform, RandInt, Beta
    synthetic/graph.py:198-222 / 225-262 (igraph)           ErdosRenyi / ScaleFree, ScaleFreeTranspose
    synthetic/graph.py:265-305 (igraph)                     WattsStrogatz
    synthetic/graph.py:308-353 / 356-379 (igraph)           SBM / GRG
    synthetic/noise_scale.py:6-71                           SimpleNoise, HeteroscedasticRFFNoise
    synthetic/linear.py:6-82, rff.py:9-114, utils.py:10-31  LinearAdditive, RFFAdditive
    synthetic/utils.py:35-151                               sample_recursive_scm
    utils/parse.py:15-41, 56-103, buffer.py:40-72           cartesian_dict, parse_config_tree, generate_data
    config/examples/*.yaml                                  EXAMPLE_DOMAINS (the same trees as Python literals)

igraph is not available here, so the three igraph-backed graph priors are restated with numpy (Barabasi-Albert
preferential attachment, the Watts-Strogatz lattice + rewiring, the geometric random graph): same graph FAMILIES and
parameters, not the same RNG stream.  `gene-ecoli` (Ecoli subgraph + the vendored SERGIO simulator, gene.py:12-241)
is a count-data surrogate (negative-binomial expression with dropout and library-size variation): parity and timing
only need identical inputs on both sides.  Nothing here runs on the hot path.
"""
import itertools
import math

import numpy as onp


# ------------------------------------------------------------------------------------------ distributions
class Gaussian:
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, rng, shape=None):
        return self.scale * rng.normal(size=shape)


class Laplace:
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, rng, shape=None):
        return self.scale * rng.laplace(size=shape)


class Cauchy:
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, rng, shape=None):
        return self.scale * rng.standard_cauchy(size=shape)


class Uniform:
    def __init__(self, low, high):
        self.low, self.high = low, high

    def __call__(self, rng, shape=None):
        return rng.uniform(size=shape, low=self.low, high=self.high)


class SignedUniform:
    def __init__(self, low, high):
        self.low, self.high = low, high

    def __call__(self, rng, shape=None):
        sgn = rng.choice([-1, 1], size=shape)
        return sgn * rng.uniform(size=shape, low=self.low, high=self.high)


class RandInt:
    def __init__(self, low, high, endpoint=True):
        self.low, self.high, self.endpoint = low, high, endpoint

    def __call__(self, rng, shape=None):
        return rng.integers(size=shape, low=self.low, high=self.high, endpoint=self.endpoint)


class Beta:
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __call__(self, rng, shape=None):
        return rng.beta(self.a, self.b, size=shape)


# ------------------------------------------------------------------------------------------ graph priors
def _random_relabel(rng, mat):
    """`p.T @ mat @ p` with p a random permutation matrix (graph.py:219-221)."""
    perm = rng.permutation(mat.shape[0])
    out = onp.zeros_like(mat)
    out[onp.ix_(perm, perm)] = mat
    return out


def _toporder(g):
    """Kahn's algorithm on adjacency g[i,j] = 1 iff i -> j (utils/graph.py:14-17)."""
    indeg = g.sum(0).astype(int)
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


class ErdosRenyi:
    """graph.py:198-222."""

    def __init__(self, edges_per_var):
        self.edges_per_var = edges_per_var

    def __call__(self, rng, n_vars):
        n_edges = self.edges_per_var * n_vars
        p = min(n_edges / max((n_vars * (n_vars - 1)) / 2, 1), 0.99)
        mat = rng.binomial(n=1, p=p, size=(n_vars, n_vars)).astype(int)
        return _random_relabel(rng, onp.tril(mat, k=-1))


class ScaleFree:
    """Directed Barabasi-Albert graph (graph.py:225-246 through igraph's `Graph.Barabasi(n, m, directed=True,
    power)`): node t attaches `edges_per_var` edges to distinct earlier nodes with probability proportional to
    in-degree^power + 1, then the vertices are relabelled at random.  Power-law in-degree."""

    def __init__(self, edges_per_var, power=1.0):
        self.edges_per_var, self.power = edges_per_var, power

    def __call__(self, rng, n_vars):
        mat = onp.zeros((n_vars, n_vars), dtype=int)
        indeg = onp.zeros(n_vars)
        for t in range(1, n_vars):
            m = min(int(self.edges_per_var), t)
            w = indeg[:t] ** self.power + 1.0
            tgt = rng.choice(t, size=m, replace=False, p=w / w.sum())
            mat[t, tgt] = 1
            indeg[tgt] += 1
        return _random_relabel(rng, mat)


class ScaleFreeTranspose(ScaleFree):
    """Power-law out-degree (graph.py:249-262)."""

    def __call__(self, rng, n_vars):
        return super().__call__(rng, n_vars).T.copy()


class WattsStrogatz:
    """Small-world graph (graph.py:265-305 through igraph's `Graph.Watts_Strogatz(dim, size, nei, p)`): a `dim`-
    dimensional periodic lattice of side ceil(n_vars^(1/dim)) whose vertices are linked within `nei` steps, every edge
    endpoint rewired with probability p (no loops, no multi-edges); excess vertices dropped, edges oriented by a random
    vertex order."""

    def __init__(self, dim=2, nei=2, p=0.1):
        self.dim, self.nei, self.p = dim, nei, p

    def __call__(self, rng, n_vars):
        size = max(math.ceil(n_vars ** (1.0 / self.dim)), 1)
        while size ** self.dim < n_vars:  # guard against floating-point roots just below an integer
            size += 1
        n_lat = size ** self.dim
        coords = onp.array(list(itertools.product(range(size), repeat=self.dim)))
        edges = set()
        for a in range(n_lat):  # lattice neighbours within `nei` steps along each axis (periodic)
            for axis in range(self.dim):
                for step in range(1, self.nei + 1):
                    c = coords[a].copy()
                    c[axis] = (c[axis] + step) % size
                    b = int(onp.ravel_multi_index(tuple(c), (size,) * self.dim))
                    if a != b:
                        edges.add((min(a, b), max(a, b)))
        edges = sorted(edges)
        present = set(edges)
        for (a, b) in edges:  # rewire the second endpoint with probability p
            if rng.uniform() < self.p:
                for _ in range(16):
                    c = int(rng.integers(n_lat))
                    e = (min(a, c), max(a, c))
                    if c != a and e not in present:
                        present.discard((a, b))
                        present.add(e)
                        break
        keep = onp.sort(rng.choice(n_lat, size=n_vars, replace=False))
        index = {int(v): i for i, v in enumerate(keep)}
        mat = onp.zeros((n_vars, n_vars), dtype=int)
        for (a, b) in present:
            if a in index and b in index:
                mat[index[a], index[b]] = 1  # a < b: upper triangle = acyclic
        return _random_relabel(rng, onp.triu(mat | mat.T, k=1))


class SBM:
    """Stochastic block model (graph.py:308-353)."""

    def __init__(self, edges_per_var, n_blocks, damp=0.2):
        self.edges_per_var, self.n_blocks, self.damp = edges_per_var, n_blocks, damp

    def __call__(self, rng, n_vars):
        n_blocks = min(n_vars, self.n_blocks)
        splits = onp.sort(rng.choice(n_vars, size=n_blocks - 1, replace=False))
        blocks = onp.split(rng.permutation(n_vars), splits)
        sizes = onp.array([b.shape[0] for b in blocks])
        block_edges = (onp.outer(sizes, sizes) - onp.diag(sizes)) / 2
        rel = onp.eye(n_blocks) + self.damp * (1 - onp.eye(n_blocks))
        n_edges = self.edges_per_var * n_vars
        p = min(0.99, n_edges / max(onp.sum(block_edges * rel), 1e-12))
        intra = rng.binomial(n=1, p=p, size=(n_vars, n_vars)).astype(int)
        inter = rng.binomial(n=1, p=self.damp * p, size=(n_vars, n_vars)).astype(int)
        mat = onp.zeros((n_vars, n_vars), dtype=int)
        for i, bi in enumerate(blocks):
            for j, bj in enumerate(blocks):
                mat[onp.ix_(bi, bj)] = (intra if i == j else inter)[onp.ix_(bi, bj)]
        return _random_relabel(rng, onp.triu(mat, k=1))


class GRG:
    """Geometric random graph (graph.py:356-379 through igraph's `Graph._GRG(n, radius)`): points uniform in the unit
    square, linked when closer than `radius`; edges oriented by a random vertex order."""

    def __init__(self, radius=0.1):
        self.radius = radius

    def __call__(self, rng, n_vars):
        pts = rng.uniform(size=(n_vars, 2))
        dist = onp.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
        mat = (dist < self.radius).astype(int)
        return _random_relabel(rng, onp.triu(mat, k=1))


# ------------------------------------------------------------------------------------------ noise models
def draw_rff_params(*, rng, d, length_scale, output_scale, n_rff):
    """synthetic/utils.py:10-31."""
    ls = length_scale(rng, shape=(1,)).item() if callable(length_scale) else length_scale
    c = output_scale(rng, shape=(1,)).item() if callable(output_scale) else output_scale
    return dict(c=c, omega=rng.normal(loc=0, scale=1.0 / ls, size=(d, n_rff)), b=rng.uniform(0, 2 * onp.pi, size=(n_rff,)),
                w=rng.normal(loc=0, scale=1.0, size=(n_rff,)), n_rff=n_rff)


def _rff_function(x_parents, *, omega, b, w, c, n_rff):
    phi = onp.cos(x_parents @ omega + b)
    return onp.sqrt(2.0) * c * (phi @ w) / onp.sqrt(n_rff)


class SimpleNoise:
    """Constant-scale noise (noise_scale.py:6-21)."""

    def __init__(self, dist, scale):
        self.dist, self.scale = dist, scale

    def __call__(self, rng, x, is_parent):
        return self.scale * self.dist(rng, shape=(x.shape[0],))


class HeteroscedasticRFFNoise:
    """Input-dependent noise scale, scale^2 = log(1 + exp(h(x_parents))) with h a random-Fourier-feature draw
    (noise_scale.py:24-52)."""

    def __init__(self, dist, *, rng, d, length_scale, output_scale, n_rff=100):
        self.dist = dist
        self.param = draw_rff_params(rng=rng, d=d, length_scale=length_scale, output_scale=output_scale, n_rff=n_rff)

    def __call__(self, rng, x, is_parent):
        f_x = _rff_function(x[:, onp.where(is_parent)[0]], **self.param)
        scale = onp.sqrt(onp.log(1.0 + onp.exp(f_x)))
        return scale * self.dist(rng, shape=scale.shape)


def init_noise_dist(*, rng, dim, dist, noise_scale, noise_scale_heteroscedastic):
    """noise_scale.py:55-71."""
    if noise_scale is not None:
        assert noise_scale_heteroscedastic is None
        return SimpleNoise(dist, noise_scale(rng))
    if noise_scale_heteroscedastic is not None:
        assert "rff" in noise_scale_heteroscedastic
        return HeteroscedasticRFFNoise(dist, rng=rng, d=int(dim), length_scale=noise_scale_heteroscedastic["length_scale"],
                                       output_scale=noise_scale_heteroscedastic["output_scale"], n_rff=100)
    raise KeyError("neither `noise_scale` nor `noise_scale_heteroscedastic` are given")


# ------------------------------------------------------------------------------------------ SCM sampling
def sample_recursive_scm(*, rng, n_observations_obs, n_observations_int, g, f, nse, interv_dist, n_interv_vars=0):
    """Ancestral sampling over a DAG with single-node interventions (synthetic/utils.py:35-151).
    Returns dict(x_obs[n_obs, d, 2], x_int[n_int, d, 2], is_count_data=False); channel 1 is the intervention mask."""
    n_vars = g.shape[-1]
    toporder = _toporder(g)
    targets = []
    if n_observations_obs > 0:
        targets.append(None)
    if n_observations_int > 0:
        assert n_interv_vars != 0, "Need n_interv_vars != 0 to sample interventional data"
        if n_interv_vars == -1 or n_interv_vars == 1.0:
            n_interv_vars = n_vars
        elif not float(n_interv_vars).is_integer():
            n_interv_vars = math.ceil(n_interv_vars * n_vars)
        n_interv_vars = int(n_interv_vars)
        targets += sorted(rng.choice(n_vars, size=min(n_vars, n_interv_vars), replace=False).tolist())
    xs, ms = {"obs": [], "int": []}, {"obs": [], "int": []}
    for target in targets:
        kind = "obs" if target is None else "int"
        rows = n_observations_obs if target is None else math.ceil(n_observations_int / n_interv_vars)
        x = onp.zeros((rows, n_vars))
        for j in toporder:
            z_j = nse[j](rng=rng, x=x, is_parent=g[:, j])
            if target is not None and j == target:
                x[:, j] = interv_dist(rng, shape=z_j.shape)
            else:
                x[:, j] = f[j](x=x, z=z_j, is_parent=g[:, j])
        mask = onp.zeros((rows, n_vars), dtype=onp.float32)
        if target is not None:
            mask[:, target] = 1.0
        xs[kind].append(x)
        ms[kind].append(mask)

    def interleave(parts):  # rows of the targets interleaved so that clipping the end keeps them balanced (utils.py:118-134)
        return onp.stack(parts).reshape(-1, n_vars, order="F") if parts else onp.zeros((0, n_vars))
    x_obs = onp.stack([interleave(xs["obs"]), interleave(ms["obs"])], -1)[:n_observations_obs]
    x_int = onp.stack([interleave(xs["int"]), interleave(ms["int"])], -1)[:n_observations_int]
    assert x_obs.size != 0 or x_int.size != 0, "Need to sample at least some observations"
    return dict(x_obs=x_obs, x_int=x_int, is_count_data=False)


class LinearAdditive:
    """Linear mechanisms with additive noise (linear.py:6-82)."""

    def __init__(self, param, bias, noise, noise_scale=None, noise_scale_heteroscedastic=None, n_interv_vars=0,
                 interv_dist=None):
        assert interv_dist is not None or n_interv_vars == 0
        self.param, self.bias, self.noise = param, bias, noise
        self.noise_scale, self.noise_scale_heteroscedastic = noise_scale, noise_scale_heteroscedastic
        self.n_interv_vars, self.interv_dist = n_interv_vars, interv_dist

    def __call__(self, rng, g, n_observations_obs, n_observations_int):
        n_vars = g.shape[-1]
        f = []
        for _ in range(n_vars):
            w = self.param(rng, shape=(n_vars,))
            b = self.bias(rng, shape=(1,))
            f.append(lambda x, is_parent, z, theta=w, bias=b: (x @ (theta * is_parent)) + bias + z)
        nse = [init_noise_dist(rng=rng, dim=int(g[:, j].sum()), dist=self.noise, noise_scale=self.noise_scale,
                               noise_scale_heteroscedastic=self.noise_scale_heteroscedastic) for j in range(n_vars)]
        return sample_recursive_scm(rng=rng, n_observations_obs=n_observations_obs, n_observations_int=n_observations_int,
                                    g=g, f=f, nse=nse, interv_dist=self.interv_dist, n_interv_vars=self.n_interv_vars)


class RFFAdditive:
    """Random-Fourier-feature (GP-draw) mechanisms with additive noise (rff.py:9-114)."""

    def __init__(self, length_scale, output_scale, bias, noise, noise_scale=None, noise_scale_heteroscedastic=None,
                 n_interv_vars=0, interv_dist=None):
        assert interv_dist is not None or n_interv_vars == 0
        self.length_scale, self.output_scale, self.bias, self.noise = length_scale, output_scale, bias, noise
        self.noise_scale, self.noise_scale_heteroscedastic = noise_scale, noise_scale_heteroscedastic
        self.n_interv_vars, self.interv_dist = n_interv_vars, interv_dist

    def __call__(self, rng, g, n_observations_obs, n_observations_int):
        n_vars = g.shape[-1]
        f = []
        for j in range(n_vars):
            par = draw_rff_params(rng=rng, d=int(g[:, j].sum()), length_scale=self.length_scale,
                                  output_scale=self.output_scale, n_rff=100)
            b = self.bias(rng, shape=(1,))
            f.append(lambda x, is_parent, z, par=par, bias=b: _rff_function(x[:, onp.where(is_parent)[0]], **par) + bias + z)
        nse = [init_noise_dist(rng=rng, dim=int(g[:, j].sum()), dist=self.noise, noise_scale=self.noise_scale,
                               noise_scale_heteroscedastic=self.noise_scale_heteroscedastic) for j in range(n_vars)]
        return sample_recursive_scm(rng=rng, n_observations_obs=n_observations_obs, n_observations_int=n_observations_int,
                                    g=g, f=f, nse=nse, interv_dist=self.interv_dist, n_interv_vars=self.n_interv_vars)


class GeneCountSurrogate:
    """Stand-in for `GRNSergio` (gene.py:12-241; the vendored SERGIO chemical-Langevin simulator is not restated):
    non-negative integer counts with dropout zeros and library-size variation, knock-down interventions; returns
    `is_count_data=True` so that the count standardizer applies."""

    def __init__(self, n_interv_vars=-1, **unused):
        self.n_interv_vars = n_interv_vars

    def __call__(self, rng, g, n_observations_obs, n_observations_int):
        d = g.shape[-1]
        w = rng.uniform(0.5, 1.5, size=(d, d)) * g * rng.choice([-1.0, 1.0], size=(d, d))
        base = rng.uniform(0.5, 3.0, size=d)
        n_tot = n_observations_obs + n_observations_int
        logexp = onp.zeros((n_tot, d))
        mask = onp.zeros((n_tot, d), dtype=onp.float32)
        if n_observations_int:
            tg = rng.integers(0, d, size=n_observations_int)
            mask[n_observations_obs + onp.arange(n_observations_int), tg] = 1.0
        for j in _toporder(g):
            logexp[:, j] = base[j] + 0.3 * onp.tanh(logexp @ w[:, j]) + 0.3 * rng.normal(size=n_tot)
        logexp = onp.where(mask > 0, -2.0, logexp)  # knock-down
        lib = rng.lognormal(0.0, 0.3, size=(n_tot, 1))
        mean = onp.exp(logexp) * lib * 5.0
        counts = rng.poisson(mean * rng.gamma(2.0, 0.5, size=mean.shape))
        dropout = rng.uniform(size=mean.shape) < 1.0 / (1.0 + onp.exp(1.5 * (onp.log1p(mean) - 1.0)))
        x = onp.stack([onp.where(dropout, 0, counts).astype(onp.float64), mask.astype(onp.float64)], -1)
        return dict(x_obs=x[:n_observations_obs], x_int=x[n_observations_obs:], is_count_data=True)


# ------------------------------------------------------------------------------------------ domain configuration DSL
YAML_CLASS = "__class__"  # avici/definitions.py
REGISTRY = {c.__name__: c for c in (Gaussian, Laplace, Cauchy, Uniform, SignedUniform, RandInt, Beta, ErdosRenyi, ScaleFree,
                                    ScaleFreeTranspose, WattsStrogatz, SBM, GRG, LinearAdditive, RFFAdditive)}
REGISTRY.update(GRNSergio=GeneCountSurrogate, Ecoli=lambda **kw: ScaleFree(edges_per_var=2, power=1.0),
                Yeast=lambda **kw: ScaleFree(edges_per_var=2, power=1.0))  # surrogates, see the module docstring


def cartesian_dict(d):
    """Cartesian product of a nested dict of lists: every list is a set of alternatives (parse.py:15-41)."""
    if isinstance(d, dict):
        keys = list(d.keys())
        for combo in itertools.product(*(list(cartesian_dict(v)) for v in d.values())):
            yield dict(zip(keys, combo))
    elif isinstance(d, list):
        for c in d:
            yield from cartesian_dict(c)
    else:
        yield d


def parse_config_tree(config, force_kwargs=None, registry=None):
    """Instantiate every `{__class__: Name, **kwargs}` node bottom-up; `force_kwargs` overrides matching keys of plain
    dict nodes (e.g. the observation counts) (parse.py:56-103)."""
    registry = REGISTRY if registry is None else registry
    if isinstance(config, dict):
        if YAML_CLASS in config:
            name = config[YAML_CLASS]
            kwargs = {k: parse_config_tree(v, force_kwargs, registry) for k, v in config.items() if k != YAML_CLASS}
            if name not in registry:
                raise SyntaxError(f"{YAML_CLASS} `{name}` of data config is not defined in existing or registered modules. "
                                  f"\nSpelled correctly and specified additional module paths?")
            return registry[name](**kwargs)
        new = {k: parse_config_tree(v, force_kwargs, registry) for k, v in config.items()}
        for k in (force_kwargs or {}):
            if k in new:
                new[k] = force_kwargs[k]
        return new
    if isinstance(config, list):
        return [parse_config_tree(v, force_kwargs, registry) for v in config]
    assert isinstance(config, (int, float, bool, str)) or config is None, f"Unknown yaml entry `{config}`"
    return config


def _load_custom_modules(module_paths):
    """Classes of user modules referenced by a YAML domain file (parse.py:106-154): every public class becomes a
    registry entry."""
    import importlib.util
    import inspect
    from pathlib import Path
    extra = {}
    for p in ([module_paths] if isinstance(module_paths, (str, Path)) else list(module_paths or [])):
        p = Path(p).resolve(strict=True)
        spec = importlib.util.spec_from_file_location(f"avici_b200._custom_.{p.stem}", p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        extra.update({n: c for n, c in inspect.getmembers(mod, inspect.isclass) if not n.startswith("_")})
    return extra


def _graphs(*extra):
    return [{YAML_CLASS: "ScaleFree", "edges_per_var": [2], "power": [0.5, 1.0, 1.5]},
            {YAML_CLASS: "ScaleFreeTranspose", "edges_per_var": [2], "power": [0.5, 1.0, 1.5]},
            {YAML_CLASS: "WattsStrogatz", "dim": [2, 3], "nei": [1], "p": [0.3]},
            {YAML_CLASS: "SBM", "edges_per_var": [2], "n_blocks": [5, 10], "damp": [0.1]},
            {YAML_CLASS: "GRG", "radius": [0.1]}, *extra]


_INTERV = {"n_interv_vars": -1, "interv_dist": [{YAML_CLASS: "SignedUniform", "low": 1.0, "high": 5.0}]}
_BIAS = [{YAML_CLASS: "Uniform", "low": -3.0, "high": 3.0}]
_HOMOSKED = {"noise_scale": [{YAML_CLASS: "Uniform", "low": 0.2, "high": 2.0}]}
_HETEROSKED = {"noise_scale_heteroscedastic": [{"rff": None, "length_scale": 10.0, "output_scale": 2.0}]}
_LIN = {YAML_CLASS: "LinearAdditive", "param": [{YAML_CLASS: "SignedUniform", "low": 1.0, "high": 3.0}], "bias": _BIAS}
_RFF = {YAML_CLASS: "RFFAdditive", "length_scale": [{YAML_CLASS: "Uniform", "low": 7.0, "high": 10.0}],
        "output_scale": [{YAML_CLASS: "Uniform", "low": 10.0, "high": 20.0}], "bias": _BIAS}
_GAUSS = {"noise": [{YAML_CLASS: "Gaussian"}]}
_HEAVY = {"noise": [{YAML_CLASS: "Cauchy"}, {YAML_CLASS: "Laplace"}]}


def _domain(mech, *parts):
    m = dict(mech)
    for p in parts:
        m.update(p)
    m.update(_INTERV)
    return {"data": [{"n_observations_obs": None, "n_observations_int": None, "graph": _graphs(), "mechanism": [m]}]}


# The trees `yaml.safe_load` returns for avici/config/examples/<domain>.yaml (values: lin-gauss.yaml:1-53,
# lin-gauss-heterosked.yaml:1-55, lin-laplace-cauchy.yaml, rff-gauss.yaml:1-57, rff-gauss-heterosked.yaml,
# rff-laplace-cauchy.yaml:1-58, gene-ecoli.yaml), written as Python literals.
EXAMPLE_DOMAINS = {
    "lin-gauss": _domain(_LIN, _GAUSS, _HOMOSKED),
    "lin-gauss-heterosked": _domain(_LIN, _GAUSS, _HETEROSKED),
    "lin-laplace-cauchy": _domain(_LIN, _HEAVY, _HOMOSKED),
    "rff-gauss": _domain(_RFF, _GAUSS, _HOMOSKED),
    "rff-gauss-heterosked": _domain(_RFF, _GAUSS, _HETEROSKED),
    "rff-laplace-cauchy": _domain(_RFF, _HEAVY, _HOMOSKED),
    "gene-ecoli": {"data": [{"n_observations_obs": None, "n_observations_int": None,
                             "graph": [{YAML_CLASS: "Ecoli"}], "mechanism": [{YAML_CLASS: "GRNSergio", "n_interv_vars": -1}]}]},
}
DOMAINS = tuple(EXAMPLE_DOMAINS)


def load_data_config(config, force_kwargs=None, module_paths=None):
    """Expand a domain configuration tree into its list of generative-model specs (parse.py:157-206): the cartesian
    product over every option list, with classes instantiated; each spec is a dict with `graph`, `mechanism`,
    `n_observations_obs`, `n_observations_int`."""
    registry = dict(REGISTRY)
    paths = list(config.get("additional_modules") or []) if isinstance(config.get("additional_modules"), list) \
        else ([config["additional_modules"]] if config.get("additional_modules") else [])
    if module_paths is not None:
        paths += module_paths if isinstance(module_paths, list) else [module_paths]
    registry.update(_load_custom_modules(paths))
    specs = parse_config_tree(list(cartesian_dict(config["data"])), force_kwargs=force_kwargs, registry=registry)
    return specs


def generate_data(rng, *, n_vars, spec_list):
    """`Sampler.generate_data` (buffer.py:40-72): random spec -> graph -> mechanism -> data."""
    spec = spec_list[rng.choice(len(spec_list))]
    g = spec["graph"](rng=rng, n_vars=n_vars)
    data = spec["mechanism"](rng=rng, g=g, n_observations_obs=spec["n_observations_obs"],
                             n_observations_int=spec["n_observations_int"])
    return dict(g=g, n_vars=n_vars, **data)


def simulate_data(d, n, *, n_interv=0, seed=0, domain=None, path=None, module_paths=None):
    """Same signature and return convention as the reference helper (utils/example.py:10-73): `(g[d,d] int,
    x[n+n_interv,d], interv[n+n_interv,d] or None)`; `x` is raw (un-standardized), rows permuted.

    domain: one of `DOMAINS` (the reference's `avici/config/examples/*.yaml`); path: a YAML domain file in the same DSL
    (needs `pyyaml`); module_paths: extra Python modules whose classes the YAML file may name."""
    rng = onp.random.default_rng(onp.random.SeedSequence(entropy=seed))
    if domain is not None:
        assert path is None, "Only specify one of `domain` and `path`"
        if domain not in EXAMPLE_DOMAINS:
            raise KeyError(f"`{domain}` does not exist; implemented example domains: {DOMAINS}")
        config = EXAMPLE_DOMAINS[domain]
    elif path is not None:
        from pathlib import Path
        if not Path(path).is_file():
            raise KeyError(f"`{path}` does not exist.")
        import yaml
        with open(path) as f:
            config = yaml.safe_load(f)
        if config is None:
            raise SyntaxError("`config` is None; make sure there are no false tabs and indents in the .yaml file")
    else:
        raise KeyError("Specify either an an `avici.config.examples` domain (YAML name) or a path to a YAML config.")
    specs = load_data_config(config, force_kwargs=dict(n_observations_obs=n, n_observations_int=n_interv),
                             module_paths=module_paths)
    data = generate_data(rng, n_vars=d, spec_list=specs)
    x = onp.concatenate([data["x_obs"], data["x_int"]], axis=-3)
    x = rng.permutation(x)
    g = onp.asarray(data["g"]).astype(int)
    return (g, x[..., 0], x[..., 1].astype(onp.float32)) if n_interv else (g, x[..., 0], None)
