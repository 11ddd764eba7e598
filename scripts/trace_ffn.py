"""Timeline of the fused FFN kernel's roles (block 0) from in-kernel clock64 stamps."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from avici_b200 import _lib
lib = _lib.load()
M, F = 1600000, 512
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(M, 128, device="cuda", generator=g)
w1t = (torch.randn(F, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
w2t = (torch.randn(128, F, device="cuda", generator=g) / 22).to(torch.bfloat16)
b1 = torch.randn(F, device="cuda", generator=g)
b2 = torch.randn(128, device="cuda", generator=g)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rep == 1:
        lib.avb_debug_set_trace(C.c_void_p(trace.data_ptr()))
    e0.record()
    lib.avb_test_ffn_bf16(z.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), b1.data_ptr(), b2.data_ptr(), M, F, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"ffn M={M}: {e0.elapsed_time(e1):.3f} ms  ({M * 1024 / e0.elapsed_time(e1) / 1e6:.0f} GB/s algorithmic)")
lib.avb_debug_set_trace(None)
t = trace.cpu().numpy().reshape(3, 512, 8)
t0 = t[t > 0].min()
names = {0: ["top", "d2_empty", "arrived", "pass1_done", "a_empty"], 1: ["top", "G1next", "w2_full", "h_full", "G2_issued"],
         2: ["top", "d1_full", "math", "h_empty", "h_full_arr", "E2:d2_full", "E2_done", "E2:first_tmem_ld"]}
for role, rn in ((0, "ln"), (1, "mma"), (2, "epilogue")):
    print(f"== {rn}: {names[role]}")
    for step in range(16, 28):
        row = t[role, step, :len(names[role])].astype(np.int64)
        if row.max() == 0:
            continue
        rel = np.where(row > 0, row - t0, -1)
        print(f" step {step:3d} ", rel.tolist(), " next_in", int(t[role, step + 1, 0] - row[0]))
