"""Forward throughput at small d (bf16), packed vs one sequence per tile in the fused d-axis kernel."""
import sys
import torch
sys.path.insert(0, ".")
import avici_b200
from avici_b200 import _lib
model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, compute_dtype="bf16")
eng = model.engine
for (B, n, d) in ((1, 200, 50), (64, 200, 50), (64, 1000, 20), (32, 1000, 50), (16, 1000, 64)):
    x = torch.randn(B, n, d, device="cuda")
    out = {}
    for pack in (0, 1):
        eng._check(eng.lib.avb_test_set_dattn_pack(eng.handle, pack))
        f = lambda: eng.forward(x, None, standardize=_lib.AVB_STANDARDIZE_DEFAULT)
        for _ in range(3): p = f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        out[pack] = (e0.elapsed_time(e1) / reps, p)
    dp = (out[0][1] - out[1][1]).abs().max().item()
    print(f"B={B} n={n} d={d}: one per tile {out[0][0]:.3f} ms, packed {out[1][0]:.3f} ms ({out[0][0] / out[1][0]:.2f}x), max|dp| between them {dp:.2e}")
