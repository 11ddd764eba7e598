"""Host-side cost of AVICIModel.__call__ on a small dataset (the reference's typical use): cProfile of 200 calls."""
import cProfile, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import avici_b200
model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, compute_dtype="bf16")
g, x, iv = avici_b200.simulate_data(d=50, n=200, seed=0, domain="rff-gauss")
x = x.astype(np.float32)
for _ in range(5): model(x=x)
t = time.perf_counter()
for _ in range(200): model(x=x)
print(f"model(x) n=200 d=50: {(time.perf_counter() - t) / 200 * 1e3:.3f} ms per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): model(x=x)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
