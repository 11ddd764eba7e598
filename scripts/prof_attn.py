"""One attention launch per axis on a bench-shaped problem (for ncu --set full captures)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from avici_b200 import _lib  # noqa: E402

lib = _lib.load()
B, n, d, H = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1000, 100, 8
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, n, d, 3 * H * 32, device="cuda", generator=g).to(torch.bfloat16)
out = torch.zeros(B, n, d, H * 32, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for axis in (1, 0, 1, 0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis, st)
    e1.record()
    torch.cuda.synchronize()
    T = n if axis == 1 else d
    fl = B * n * d * 4 * H * 32 * T
    print(f"axis={axis} rc={rc} {e0.elapsed_time(e1):.3f} ms  {fl / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s")
