"""avici_b200 — B200-native implementation of AVICI's inference hot path.

Public surface mirrors `avici/__init__.py:1-4` for the path in scope:
`load_pretrained`, `AVICIModel`, `simulate_data` (+ `InferenceModel`, `BaseModel`).
Importing the package does not need a GPU; any compute call needs the in-tree CUDA extension
(`avici_b200/csrc/libavici_b200.so`, built by `__graft_entry__.build()`) and an sm_100 device.
"""
from .model import AVICIModel, BaseModel, InferenceModel
from .params import init_params, param_count
from .pretrain import load_checkpoint, load_pretrained, random_init_model, save_checkpoint
from .synthetic import simulate_data
from . import metrics  # noqa: F401  (graph metrics on the device, avici/metrics.py mirror)
from .evaluate import eval_batch  # noqa: F401  (train.py:51-69,125-151 on the device)

__all__ = ["load_pretrained", "AVICIModel", "simulate_data", "InferenceModel", "BaseModel", "init_params",
           "param_count", "load_checkpoint", "save_checkpoint", "random_init_model"]
__version__ = "0.2.0"
