"""Does a CUDA graph of avb_forward help small datasets?  (capture of the C-ABI forward on a side stream, replay)"""
import sys
import torch
sys.path.insert(0, ".")
import avici_b200
from avici_b200 import _lib
model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, compute_dtype="bf16")
eng = model.engine
for (B, n, d) in ((1, 200, 50), (1, 1000, 100), (4, 200, 50)):
    x = torch.randn(B, n, d, device="cuda")
    out = torch.empty(B, d, d, device="cuda")
    f = lambda: eng.forward(x, None, standardize=_lib.AVB_STANDARDIZE_DEFAULT, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    ref = out.clone()
    def timeit(fn, reps=50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_plain = timeit(f)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        f()
    out.zero_()
    g.replay(); torch.cuda.synchronize()
    ok = torch.equal(out, ref)
    t_graph = timeit(g.replay)
    print(f"B={B} n={n} d={d}: plain {t_plain:.3f} ms, graph {t_graph:.3f} ms, identical {ok}")
