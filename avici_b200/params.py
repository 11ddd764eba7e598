"""Haiku parameter naming and random initialisation of the AVICI `BaseModel` (host side, numpy).

The reference keeps its weights as a Haiku parameter dict `{module_path: {leaf: array}}`
(`state.params`, avici/pretrain.py:267; module creation order in avici/model.py:53-144).  This
module restates that naming and draws random-init weights from the same
distributions Haiku would use, so that benchmarks and parity tests have "scm-v0-shaped" networks
without the (un-downloadable) checkpoints.
"""
import math

import numpy as np

DEFAULT_MODEL_KWARGS = dict(layers=8, dim=128, key_size=32, num_heads=8, widening_factor=4,
                            dropout=0.1, out_dim=None, logit_bias_init=-3.0, cosine_temp_init=0.0,
                            ln_axis=-1)


def _sfx(i):
    return "" if i == 0 else f"_{i}"


def resolve_model_kwargs(model_kwargs=None):
    """`BaseModel.__init__` defaults (avici/model.py:26-50); unknown keys are dropped like
    `InferenceModel.__init__` does (model.py:165-169)."""
    kw = dict(DEFAULT_MODEL_KWARGS)
    for k, v in (model_kwargs or {}).items():
        if k in kw:
            kw[k] = v
    if kw["out_dim"] is None:
        kw["out_dim"] = kw["dim"]
    if kw["ln_axis"] != -1:
        raise ValueError("only ln_axis=-1 (the reference default and all published models) is supported")
    return kw


def param_shapes(model_kwargs=None):
    """Ordered {module_path: {leaf: shape}} in Haiku creation order (model.py:56-73,95,122-141)."""
    kw = resolve_model_kwargs(model_kwargs)
    k, H, dk, F, od = kw["dim"], kw["num_heads"], kw["key_size"], kw["widening_factor"] * kw["dim"], kw["out_dim"]
    nb = 2 * kw["layers"]
    R = "BaseModel"
    shapes = {f"{R}/linear": {"w": (2, k), "b": (k,)}}
    for i in range(nb):
        for t in range(3):
            shapes[f"{R}/layer_norm{_sfx(4 * i + t)}"] = {"scale": (k,), "offset": (k,)}
        mha = f"{R}/multi_head_attention{_sfx(i)}"
        for name in ("query", "key", "value"):
            shapes[f"{mha}/{name}"] = {"w": (k, H * dk), "b": (H * dk,)}
        shapes[f"{mha}/linear"] = {"w": (H * dk, k), "b": (k,)}
        shapes[f"{R}/layer_norm{_sfx(4 * i + 3)}"] = {"scale": (k,), "offset": (k,)}
        shapes[f"{R}/linear{_sfx(2 * i + 1)}"] = {"w": (k, F), "b": (F,)}
        shapes[f"{R}/linear{_sfx(2 * i + 2)}"] = {"w": (F, k), "b": (k,)}
    shapes[f"{R}/layer_norm{_sfx(4 * nb)}"] = {"scale": (k,), "offset": (k,)}
    shapes[f"{R}/layer_norm{_sfx(4 * nb + 1)}"] = {"scale": (k,), "offset": (k,)}
    shapes[f"{R}/linear{_sfx(2 * nb + 1)}"] = {"w": (k, od), "b": (od,)}
    shapes[f"{R}/layer_norm{_sfx(4 * nb + 2)}"] = {"scale": (k,), "offset": (k,)}
    shapes[f"{R}/linear{_sfx(2 * nb + 2)}"] = {"w": (k, od), "b": (od,)}
    shapes[R] = {"learned_temp": (1, 1, 1), "final_matrix_bias": (1, 1, 1)}
    return shapes


def param_count(model_kwargs=None):
    return sum(int(np.prod(s)) for leaves in param_shapes(model_kwargs).values() for s in leaves.values())


def _trunc_normal(rng, shape, std):
    """hk.initializers.TruncatedNormal: normal truncated to +-2 sigma (rejection sampling)."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * std


def init_params(model_kwargs=None, seed=0, perturb=False, dtype=np.float32):
    """Random-init parameters following Haiku's initialisers (NOT bit-identical to Haiku: its draws
    come from JAX's threefry PRNG).

    * `hk.Linear` default: TruncatedNormal(1/sqrt(fan_in)), b = 0            (embed, model.py:95)
    * `hk.MultiHeadAttention(w_init_scale=2.0)`: VarianceScaling(2.0) = fan-in truncated normal with
      std sqrt(2/fan_in)/0.87962566 for query/key/value/linear                (model.py:59-64)
    * `VarianceScaling(2.0, "fan_in", "uniform")`: U(+-sqrt(6/fan_in))        (model.py:50,71,73,124,128)
    * LayerNorm scale 1 / offset 0; `learned_temp`, `final_matrix_bias` constants (model.py:136-141)

    perturb=True additionally jitters LayerNorm affine parameters and all biases (what a trained
    checkpoint looks like), so that LayerNorm folding / bias handling cannot hide behind identities.
    """
    kw = resolve_model_kwargs(model_kwargs)
    rng = np.random.default_rng(seed)
    params = {}
    for mod, leaves in param_shapes(kw).items():
        p = {}
        if mod == "BaseModel":
            p["learned_temp"] = np.full((1, 1, 1), kw["cosine_temp_init"])
            p["final_matrix_bias"] = np.full((1, 1, 1), kw["logit_bias_init"])
        elif "layer_norm" in mod:
            p["scale"] = np.ones(leaves["scale"])
            p["offset"] = np.zeros(leaves["offset"])
            if perturb:
                p["scale"] = p["scale"] + 0.1 * rng.standard_normal(leaves["scale"])
                p["offset"] = p["offset"] + 0.1 * rng.standard_normal(leaves["offset"])
        else:
            fan_in = leaves["w"][0]
            if "multi_head_attention" in mod:
                w = _trunc_normal(rng, leaves["w"], math.sqrt(2.0 / fan_in) / 0.87962566103423978)
            elif mod == "BaseModel/linear":
                w = _trunc_normal(rng, leaves["w"], 1.0 / math.sqrt(fan_in))
            else:
                lim = math.sqrt(6.0 / fan_in)
                w = rng.uniform(-lim, lim, size=leaves["w"])
            p["w"] = w
            p["b"] = np.zeros(leaves["b"])
            if perturb:
                p["b"] = p["b"] + 0.02 * rng.standard_normal(leaves["b"])
        params[mod] = {k: np.ascontiguousarray(v, dtype=dtype) for k, v in p.items()}
    return params


def flatten_params(params):
    """{module: {leaf: array}} -> [(\"module/leaf\", contiguous float32 array)], the C-ABI's view."""
    flat = []
    for mod, leaves in params.items():
        for leaf, arr in leaves.items():
            flat.append((f"{mod}/{leaf}", np.ascontiguousarray(np.asarray(arr), dtype=np.float32)))
    return flat
