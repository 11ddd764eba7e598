"""Times BASELINE config 5 (count data d=200 n=2000, 16 blocks) in the fp32 parity mode and in bf16 mode on one GPU
(the fp32 mode had never been timed).  Prints one JSON line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import avici_b200
from avici_b200 import _lib

d = np.load("tests/golden/full/c5_gene_n2000_d200.npz")
x = torch.as_tensor(d["x"]).cuda()
F = 2000 * 200 * (16 * 524288 + 8 * 1024 * (2000 + 200)) + 512 * 2000 * 200
out = {}
for dtype in ("fp32", "bf16"):
    model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, expects_counts=True, compute_dtype=dtype)
    eng = model.engine
    step = lambda: eng.forward(x, None, standardize=_lib.AVB_STANDARDIZE_COUNT)  # noqa: E731
    step(); torch.cuda.synchronize()
    reps = 3 if dtype == "fp32" else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    eng.profile(True); eng.profile_read(reset=True); step(); prof = eng.profile_read(reset=True); eng.profile(False)
    out[dtype] = {"ms_per_dataset": ms, "tflops": F / ms / 1e9, "kernel_ms": {k: round(v[0], 3) for k, v in prof.items()}}
print(json.dumps({"config5_n2000_d200": out, "flops_per_dataset": F}))
