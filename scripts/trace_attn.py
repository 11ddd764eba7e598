"""Timeline of the attention kernel's roles (block 0, stream 0) from in-kernel clock64 stamps."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from avici_b200 import _lib
lib = _lib.load()
B, n, d, H = 16, 1000, 100, 8
axis = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, n, d, 3 * H * 32, device="cuda", generator=g).to(torch.bfloat16)
out = torch.zeros(B, n, d, H * 32, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis, st)  # warm
torch.cuda.synchronize()
lib.avb_debug_set_trace(C.c_void_p(trace.data_ptr()))
lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis, st)
torch.cuda.synchronize()
lib.avb_debug_set_trace(None)
t = trace.cpu().numpy().reshape(4, 512, 8)
t0 = t[t > 0].min()
names = {0: ["wait_kv_empty", "got", "issued"], 1: ["top", "s_free", "S_issued/wait_p", "p_full", "PV_issued"],
         2: ["top", "s_full", "S_loaded+s_free", "-", "exps_done", "pv_done(+epilogue)", "-", "P_stored+p_full"]}
names[3] = names[2]
for role, rn in ((0, "producer"), (1, "mma"), (2, "softmax s0"), (3, "softmax s1")):
    print(f"== {rn}: events {names[role]} (cycles since kernel start; deltas within a step)")
    for step in list(range(40, 50)):
        row = t[role, step, :len(names[role])]
        if row.max() == 0:
            continue
        rel = row - t0
        print(f" step {step:3d} start {rel[0]:8d} deltas {np.diff(rel).tolist()}  next_step_in {t[role, step + 1, 0] - row[0]}")
print("== exp phases (S loaded -> exps done), absolute cycles: stream 0 | stream 1")
for step in range(40, 50):
    a, b = t[2, step] - t0, t[3, step] - t0
    print(f" step {step}: s0 [{a[2]}, {a[4]}] len {a[4] - a[2]} | s1 [{b[2]}, {b[4]}] len {b[4] - b[2]}")
for role, rn in ((1, "mma"), (2, "softmax s0"), (3, "softmax s1")):
    tops = t[role, 16:400, 0]
    tops = tops[tops > 0]
    print(f"{rn}: mean step period {(tops[-1] - tops[0]) / (len(tops) - 1):.0f} clocks")
sm = t[2, 16:400].astype(np.int64)
for a, b, nm in ((0, 1, "wait s_full"), (1, 2, "tmem ld"), (2, 4, "exps"), (4, 5, "wait pv_done"), (5, 7, "tmem st")):
    print(f"softmax s0 {nm}: mean {(sm[:, b] - sm[:, a]).mean():.0f}")
mm = t[1, 16:400].astype(np.int64)
for a, b, nm in ((0, 1, "wait s_free"), (1, 2, "issue S"), (2, 3, "wait p_full"), (3, 4, "issue PV+L")):
    print(f"mma s0 {nm}: mean {(mm[:, b] - mm[:, a]).mean():.0f}")
