"""One fused-FFN launch (and one QKV projection, one out-projection) on a bench-shaped problem, for ncu captures."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from avici_b200 import _lib  # noqa: E402

lib = _lib.load()
M, F = int(sys.argv[1]) if len(sys.argv) > 1 else 1600000, 512
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(M, 128, device="cuda", generator=g)
w1t = (torch.randn(F, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
w2t = (torch.randn(128, F, device="cuda", generator=g) / 22).to(torch.bfloat16)
b1 = torch.randn(F, device="cuda", generator=g)
b2 = torch.randn(128, device="cuda", generator=g)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.avb_test_ffn_bf16(z.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), b1.data_ptr(), b2.data_ptr(), M, F, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"ffn M={M} rc={rc}: {ms:.3f} ms  {M * 1024 / ms / 1e6:.0f} GB/s algorithmic  {M * 262144 / ms / 1e9:.0f} TFLOP/s")
