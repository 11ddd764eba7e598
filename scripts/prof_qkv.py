"""One QKV projection (CTA-pair LayerNorm + GEMM, operand-layout output) on a bench-shaped problem, for ncu captures."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from avici_b200 import _lib  # noqa: E402

lib = _lib.load()
B, n, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 16), 1000, 100
M, N = B * n * d, 768
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(M, 128, device="cuda", generator=g)
wt = (torch.randn(N, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
bias = torch.randn(N, device="cuda", generator=g)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for axis in (0, 1, 0, 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.avb_test_qkv_layout_bf16(z.data_ptr(), wt.data_ptr(), bias.data_ptr(), out.data_ptr(), B, n, d, 8, axis, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"qkv+unpack axis={axis} rc={rc}: {e0.elapsed_time(e1):.3f} ms (includes the test hook's unpack kernel)")
