"""One fused LN+GEMM launch per shape on a bench-sized problem (for timing and ncu captures)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from avici_b200 import _lib
lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1600000
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(M, 128, device="cuda", generator=g)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for N, relu in ((768, 0), (512, 1), (768, 0), (512, 1)):
    wt = (torch.randn(N, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.avb_test_ln_gemm_bf16(z.data_ptr(), wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, relu, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gb = M * (512 + N * 2) / 1e9
    print(f"N={N} relu={relu} rc={rc} {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic  {2 * M * N * 128 / ms / 1e9:.0f} TFLOP/s")
