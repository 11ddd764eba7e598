"""`load_pretrained` and the checkpoint loader (drop-in for avici/pretrain.py:159-273 and
avici/utils/load.py:32-48), without jax / haiku / optax in the process.

On-disk format of a reference checkpoint folder: `checkpoint_XXXXXXX.pkl` = pickle of the
`ModelState` NamedTuple (avici/backbone.py:56-63: step, rng, opt_state, params, dual,
dual_penalty_polyak, ave_params) whose leaves are numpy arrays (written at backbone.py:283-292), plus
`kwargs.json` with `neural_net_kwargs` (possibly nested under "model_kwargs", pretrain.py:249-252)
and `inference_model_kwargs`.  Only `.params` is used (pretrain.py:267).
"""
import inspect
import io
import json
import pickle
from collections import namedtuple
from pathlib import Path

import numpy as onp

from .model import AVICIModel, BaseModel, InferenceModel
from .params import init_params, resolve_model_kwargs

CHECKPOINT_KWARGS = "kwargs.json"          # avici/definitions.py:23
HUGGINGFACE_REPO = "larslorch/avici"       # avici/definitions.py:27
PRETRAINED = {                             # avici/definitions.py:28-44, pretrain.py:182-194
    "neurips-linear": False,
    "neurips-rff": False,
    "neurips-grn": True,
    "scm-v0": False,
}

ModelState = namedtuple("ModelState", ["step", "rng", "opt_state", "params", "dual", "dual_penalty_polyak",
                                       "ave_params"])


class _Opaque:
    """Placeholder for classes we do not need (optax optimizer states, jax array wrappers, pytree defs)."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def __setstate__(self, state):
        self.state = state


# the only builtins / collections a checkpoint needs; `builtins.eval`, `exec`, `getattr`, `__import__` ... stay out
_SAFE_BUILTINS = {"dict", "list", "tuple", "set", "frozenset", "slice", "complex", "int", "float", "bool", "str",
                  "bytes", "bytearray", "object", "range"}
_SAFE_COLLECTIONS = {"OrderedDict", "defaultdict", "deque"}


class _CheckpointUnpickler(pickle.Unpickler):
    """Restricted unpickler: resolves `avici.backbone.ModelState` (and the legacy `amortibs.*` alias,
    utils/load.py:3-4) to a local namedtuple, lets numpy and an explicit list of container builtins through, and
    turns every other class (haiku's FlatMapping / FlatComponents, optax states, jax pytree defs) into an inert
    `_Opaque` stub that only records its constructor arguments; `_plain_params` then rebuilds the parameter dict."""

    def find_class(self, module, name):
        if name == "ModelState":  # avici.backbone / legacy amortibs.backbone / our own writer
            return ModelState
        root = module.split(".")[0]
        if root == "numpy":
            return super().find_class(module, name)
        if module == "builtins" and name in _SAFE_BUILTINS:
            return super().find_class(module, name)
        if module == "collections" and name in _SAFE_COLLECTIONS:
            return super().find_class(module, name)
        return type(name, (_Opaque,), {"__module__": module})


def _plain_params(obj, neural_net_kwargs=None):
    """`state.params` as a plain `{module: {leaf: ndarray}}` dict.

    haiku >= 0.0.10 pickles params as nested dicts.  Older versions (the published checkpoints date from 2022,
    requirements.txt:12 `dm-haiku>=0.0.8`) wrap every level in `haiku._src.data_structures.FlatMapping`, which pickles
    either as `FlatMapping(dict)` or as `FlatMapping(FlatComponents(leaves, structure))` with `structure` a jax pytree
    definition.  The first form is unwrapped recursively; for the second the leaves (jax flattens dicts in sorted key
    order) are matched by position and shape against the sorted parameter names of the configured architecture."""
    if isinstance(obj, dict) or hasattr(obj, "items") and not isinstance(obj, _Opaque):
        return {k: _plain_params(v, neural_net_kwargs) for k, v in obj.items()}
    if isinstance(obj, _Opaque):
        args = getattr(obj, "args", ())
        state = getattr(obj, "state", None)
        cand = args[0] if len(args) >= 1 else state
        if isinstance(cand, dict) or (hasattr(cand, "items") and not isinstance(cand, _Opaque)):
            return _plain_params(cand, neural_net_kwargs)
        comp = cand if isinstance(cand, _Opaque) else None  # FlatComponents(leaves, structure)
        leaves = None
        if comp is not None and len(getattr(comp, "args", ())) >= 1:
            leaves = comp.args[0]
        elif isinstance(cand, (tuple, list)) and len(cand) == 2 and isinstance(cand[0], (tuple, list)):
            leaves = cand[0]
        if leaves is not None:
            from .params import param_shapes
            names = sorted((m, l, shp) for m, ls in param_shapes(neural_net_kwargs).items() for l, shp in ls.items())
            if len(names) != len(leaves):
                raise ValueError(f"checkpoint holds {len(leaves)} parameter leaves, the architecture in kwargs.json has "
                                 f"{len(names)}")
            out = {}
            for (m, l, shp), arr in zip(names, leaves):
                arr = onp.asarray(arr)
                if tuple(arr.shape) != tuple(shp):
                    raise ValueError(f"checkpoint leaf for {m}/{l} has shape {arr.shape}, expected {shp}")
                out.setdefault(m, {})[l] = arr
            return out
        raise ValueError(f"cannot recover parameters from pickled object of class {type(obj).__name__}")
    return onp.asarray(obj)


def load_checkpoint(folder):
    """avici/utils/load.py:32-48: newest `*.pkl` in `folder` + `kwargs.json` (acyclicity forced off)."""
    folder = Path(folder)
    last_checkpoint = sorted([p for p in folder.iterdir() if ".pkl" in p.name])[-1]
    with open(last_checkpoint, "rb") as file:
        state = _CheckpointUnpickler(io.BytesIO(file.read())).load()
    with open(folder / CHECKPOINT_KWARGS, "r") as file:
        kwargs = json.load(file)
        kwargs["inference_model_kwargs"]["acyclicity"] = None
    return state, kwargs


def save_checkpoint(folder, params, neural_net_kwargs, inference_model_kwargs=None, step=0):
    """Write a folder `load_pretrained(checkpoint_dir=...)` (and the reference's loader) can read."""
    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    state = ModelState(step=step, rng=None, opt_state=None,
                       params={m: {k: onp.asarray(v) for k, v in l.items()} for m, l in params.items()},
                       dual=None, dual_penalty_polyak=None, ave_params=None)
    with open(folder / f"checkpoint_{step:07d}.pkl", "wb") as f:
        pickle.dump(state, f)
    with open(folder / CHECKPOINT_KWARGS, "w") as f:
        json.dump({"neural_net_kwargs": dict(neural_net_kwargs),
                   "inference_model_kwargs": dict(inference_model_kwargs or {"mask_diag": True})}, f)
    return folder


def load_pretrained(download=None, force_download=False, checkpoint_dir=None, cache_path="./", expects_counts=None,
                    verbose=True, compute_dtype="bf16"):
    """
    Loads an AVICI model (same arguments as the reference, plus `compute_dtype`).

    Args:
        download (str, optional): one of `neurips-linear`, `neurips-rff`, `neurips-grn`, `scm-v0`.  The
            checkpoint is looked up in `cache_path` (layout `<cache_path>/<download>/`); fetching from the
            HuggingFace hub needs `huggingface_hub` and network access.
        force_download (bool, optional): forwarded to the hub download.
        checkpoint_dir (str, optional): local folder with `<name>.pkl` and `kwargs.json`.
        cache_path (str, optional): cache directory for downloads.
        expects_counts (bool, optional): required with `checkpoint_dir`.
        verbose (bool, optional)
        compute_dtype (str): "bf16" (tensor-core path) or "fp32" (parity path).
    Returns:
        `avici_b200.AVICIModel`
    """
    assert (download is None and checkpoint_dir is not None) \
        or (download is not None and checkpoint_dir is None), "Specify exactly one of `download` or `checkpoint_dir`"

    if download is not None:
        if download not in PRETRAINED:
            raise ValueError(f"Unknown download specified: `{download}`")
        expects_counts = PRETRAINED[download]
        if cache_path == "./":
            cache_path = Path(cache_path).resolve() / "cache"
            if verbose:
                print(f"Using default cache_path: `{cache_path}`")
        root_path = Path(cache_path) / download
        have = root_path.is_dir() and (root_path / CHECKPOINT_KWARGS).exists() and \
            any(".pkl" in p.name for p in root_path.iterdir())
        if not have or force_download:
            try:
                from huggingface_hub import hf_hub_download
                paths = [hf_hub_download(repo_id=HUGGINGFACE_REPO, subfolder=download, filename=fn,
                                         cache_dir=cache_path, force_download=force_download)
                         for fn in (CHECKPOINT_KWARGS, "checkpoint_0300000.pkl")]
                root_path = Path(paths[0]).parent
            except Exception as e:  # offline box: say exactly what is missing
                raise FileNotFoundError(
                    f"Pretrained checkpoint `{download}` is not available offline (looked in `{root_path}`; "
                    f"hub download failed: {type(e).__name__}: {e}). Place `kwargs.json` and "
                    f"`checkpoint_*.pkl` there or pass `checkpoint_dir=`.") from e
    else:
        assert expects_counts is not None, "When loading custom model checkpoint, need to specify `expects_counts` as "\
                                           "`True`/`False` to ensure proper data standardization. Specifying " \
                                           "`False` should be the default and implies the data is z-standardized."
        root_path = Path(checkpoint_dir)
        assert root_path.exists(), f"The provided checkpoint_dir `{checkpoint_dir}` does not exist. " \
                                   f"In case you provided a relative path, try the absolute path instead."
        assert root_path.is_dir(), "The provided checkpoint_dir is not a directory. You should specify the path to " \
                                   f"a folder that contains the model `.pkl` (and `{CHECKPOINT_KWARGS}`), " \
                                   "not the path to the `.pkl` alone."
        # the reference asserts on the PARENT folder (pretrain.py:236) but reads the file from the folder
        # itself (utils/load.py:42); accept either location, read from the folder like the reference
        assert (root_path / CHECKPOINT_KWARGS).exists() or (root_path.parent / CHECKPOINT_KWARGS).exists(), \
            f"The provided checkpoint_dir `{checkpoint_dir}` does not contain the file `{CHECKPOINT_KWARGS}`, " \
            f"which is required. "
        if verbose:
            print(f"Loading local checkpoint from `{root_path}`")

    state, loaded_config = load_checkpoint(root_path)
    inference_model_kwargs = dict(loaded_config["inference_model_kwargs"])
    try:
        neural_net_kwargs = loaded_config["neural_net_kwargs"]["model_kwargs"]  # legacy compatibility
    except KeyError:
        neural_net_kwargs = loaded_config["neural_net_kwargs"]

    sig = inspect.signature(InferenceModel).parameters
    for k in [k for k in inference_model_kwargs if k not in sig]:
        del inference_model_kwargs[k]

    inference_model = InferenceModel(**inference_model_kwargs, model_class=BaseModel,
                                     model_kwargs=neural_net_kwargs, compute_dtype=compute_dtype)
    params = state.params if hasattr(state, "params") else state["params"]
    params = _plain_params(params, neural_net_kwargs)  # unwraps haiku < 0.0.10 FlatMapping containers
    return AVICIModel(params=params, model=inference_model, expects_counts=expects_counts)


def random_init_model(model_kwargs=None, seed=0, perturb=True, expects_counts=False, compute_dtype="bf16",
                      mask_diag=True):
    """An `AVICIModel` with random-init weights of the published architecture (the train-script defaults,
    example-custom/train.py:66-80: `cosine_temp_init=2.0`) for benchmarks and parity tests."""
    kw = dict(cosine_temp_init=2.0)
    kw.update(model_kwargs or {})
    kw = resolve_model_kwargs(kw)
    params = init_params(kw, seed=seed, perturb=perturb)
    model = InferenceModel(model_class=BaseModel, model_kwargs=kw, mask_diag=mask_diag, compute_dtype=compute_dtype)
    return AVICIModel(params=params, model=model, expects_counts=expects_counts)
