"""SM clock and board power while ONE hot kernel runs back to back for ~2 s each (bench-shaped problems, batch 64):
which kernels the 1000 W cap throttles, and how far.  Uses the library's test hooks; prints one line per kernel."""
import ctypes as C
import statistics
import sys
import time

import torch

sys.path.insert(0, ".")
import avici_b200  # noqa: E402
from bench import ClockSampler  # noqa: E402

B, n, d, H = 64, 1000, 100, 8
model = avici_b200.random_init_model(dict(layers=1), seed=3, perturb=True, compute_dtype="bf16")
eng = model.engine
lib = eng.lib
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(0)
M = B * n * d
z = torch.randn(M, 128, device="cuda", generator=g)
qkv = torch.randn(M, 3 * H * 32, device="cuda", generator=g).to(torch.bfloat16) * 0.5
attn = torch.zeros(M, H * 32, device="cuda", dtype=torch.bfloat16)
delta = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
wt = (torch.randn(768, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
bias = torch.randn(768, device="cuda", generator=g)
w1t = (torch.randn(512, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
w2t = (torch.randn(128, 512, device="cuda", generator=g) / 22).to(torch.bfloat16)
b1 = torch.randn(512, device="cuda", generator=g)
b2 = torch.randn(128, device="cuda", generator=g)
qkv_out = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)

kernels = {
    "attention over n": lambda: lib.avb_test_attention_bf16(qkv.data_ptr(), attn.data_ptr(), B, n, d, H, 1, st),
    "dattn (fused d-axis)": lambda: lib.avb_test_dattn_bf16(eng.handle, 0, z.data_ptr(), B, n, d, 1, delta.data_ptr(), None, st),
    "ffn": lambda: lib.avb_test_ffn_bf16(z.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), b1.data_ptr(), b2.data_ptr(), M, 512, st),
    "qkv (n-axis layout)": lambda: lib.avb_test_qkv_layout_bf16(z.data_ptr(), wt.data_ptr(), bias.data_ptr(), qkv_out.data_ptr(), B, n, d, 8, 1, st),
}
for name, fn in kernels.items():
    for _ in range(3):
        assert fn() == 0, name
    torch.cuda.synchronize()
    s = ClockSampler(torch.cuda.current_device())
    s.start()
    t0, launches = time.time(), 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 2.5:
        for _ in range(10):
            fn()
        launches += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / launches
    rows = s.rows[:]
    clk = s.stop()
    pw = [float(r[2]) for r in rows[3:] if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
    print(f"{name:24s} {ms:7.3f} ms per call (hook)   sm {clk['sm_mhz']} MHz of {clk['sm_max_mhz']}   power {statistics.mean(pw) if pw else float('nan'):.0f} W "
          f"(max {max(pw) if pw else float('nan'):.0f})   {clk['reasons']}")
    time.sleep(1.0)
