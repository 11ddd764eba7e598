"""One-process A/B of builds of the FFN pair kernel through the avb_test_ffn_bf16 hook: the default library against
experimental ones (paths in argv[1:]).  All must produce bit-identical z (same arithmetic, different schedule / store
path); prints the time of each."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from avici_b200 import _lib

libs = {"default": _lib.load()}
res, args = _lib.SIGNATURES["avb_test_ffn_bf16"]
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.avb_test_ffn_bf16.restype = res
    lib.avb_test_ffn_bf16.argtypes = args
    libs[path.split("/")[-1]] = lib
M, F = 1600077, 512
g = torch.Generator(device="cuda").manual_seed(0)
z0 = torch.randn(M, 128, device="cuda", generator=g)
w1t = (torch.randn(F, 128, device="cuda", generator=g) / 11).to(torch.bfloat16)
w2t = (torch.randn(128, F, device="cuda", generator=g) / 22).to(torch.bfloat16)
b1 = torch.randn(F, device="cuda", generator=g)
b2 = torch.randn(128, device="cuda", generator=g)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
outs = {}
for rnd in range(2):  # two rounds over all libraries: clocks drift, the second round is the one to read
    for name, lib in libs.items():
        times = []
        for rep in range(4):
            z = z0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.avb_test_ffn_bf16(z.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), b1.data_ptr(), b2.data_ptr(), M, F, st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            times.append(e0.elapsed_time(e1))
        outs[name] = z
        print(f"round {rnd} {name}: {min(times[1:]):.3f} ms (runs {[round(t, 3) for t in times]})", flush=True)
for name in libs:
    same = torch.equal(outs["default"], outs[name])
    print(name, "bit-identical:", same, " max|diff|", float((outs["default"] - outs[name]).abs().max()),
          " finite:", bool(torch.isfinite(outs[name]).all()))
