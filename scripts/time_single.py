"""Latency of ONE dataset through model(x) (the reference's primary call) and small batches, bf16 and fp32 modes."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import avici_b200
from avici_b200 import _lib
for dtype in ("bf16", "fp32"):
    model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, compute_dtype=dtype)
    eng = model.engine
    for (B, n, d) in ((1, 200, 50), (1, 1000, 100), (8, 1000, 100), (1, 2000, 200)):
        x = torch.randn(B, n, d, device="cuda")
        f = lambda: eng.forward(x, None, standardize=_lib.AVB_STANDARDIZE_DEFAULT)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        xn = x[0].cpu().numpy()
        t0 = time.perf_counter(); model(x=xn); t1 = time.perf_counter()
        print(f"{dtype} B={B} n={n} d={d}: {ms:.3f} ms per forward on device ({B / ms * 1e3:.1f} datasets/s); model(x) numpy->numpy one dataset {1e3 * (t1 - t0):.2f} ms")
