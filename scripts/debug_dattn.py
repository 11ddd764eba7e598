"""Stage-by-stage check of dattn_tc_kernel with a -DAVB_DATTN_DEBUG build (AVB_LIB=<that build>): CTA 0 dumps the
intermediates of its first sequence (Q|K|V after bias, the raw scores, O / L per head, DELTA + bo); each is compared
with an fp32 torch computation from the same z and the block's parameters.  Prints max |err| per stage and head."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import avici_b200
from oracle import avici_oracle as O

B, n, d = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 100
fuse_out = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = avici_b200.random_init_model(dict(layers=1), seed=3, perturb=True, compute_dtype="bf16")
eng = model.engine
lib = eng.lib
g = torch.Generator(device="cuda").manual_seed(1)
z = torch.randn(B, n, d, 128, device="cuda", generator=g) * 2 + torch.randn(B, n, d, 1, device="cuda", generator=g)
dump = torch.full((278528,), float("nan"), device="cuda")
rc = lib.avb_debug_set_dattn_dump(C.c_void_p(dump.data_ptr()))
print("set dump rc", rc)
delta = torch.full((B * n * d, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
guard = C.c_int32(-1)
rc = lib.avb_test_dattn_bf16(eng.handle, 0, z.data_ptr(), B, n, d, fuse_out, delta.data_ptr(), C.byref(guard),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
print("rc", rc, lib.avb_last_error(eng.handle), "guard", guard.value)
torch.cuda.synchronize()

nm = O.block_param_names(0)
P = lambda m: {k: torch.as_tensor(np.asarray(v), device="cuda", dtype=torch.float32) for k, v in model.params[m].items()}  # noqa: E731
ln = lambda x, p: torch.nn.functional.layer_norm(x, (128,), p["scale"], p["offset"], 1e-5)  # noqa: E731
lin = lambda x, p: x @ p["w"] + p["b"]  # noqa: E731
zs = z[0, 0]  # CTA 0's first sequence = sequence 0: [d, 128]
q = lin(ln(zs, P(nm["ln_q"])), P(nm["mha"] + "/query")).reshape(d, 8, 32)
k = lin(ln(zs, P(nm["ln_k"])), P(nm["mha"] + "/key")).reshape(d, 8, 32)
v = lin(ln(zs, P(nm["ln_v"])), P(nm["mha"] + "/value")).reshape(d, 8, 32)
scale = 1.4426950408889634 / 32 ** 0.5
dq = dump[:8 * 128 * 96].reshape(8, 128, 96)
ds = dump[8 * 128 * 96:8 * 128 * 96 + 8 * 128 * 128].reshape(8, 128, 128)
do = dump[8 * 128 * 96 + 8 * 128 * 128:8 * 128 * 96 + 8 * 128 * 128 + 8 * 128 * 32].reshape(8, 128, 32)
dd = dump[8 * 128 * 96 + 8 * 128 * 128 + 8 * 128 * 32:].reshape(128, 128)
for h in range(8):
    eq = (dq[h, :d, 0:32] - q[:, h] * scale).abs().max().item()
    ek = (dq[h, :d, 32:64] - k[:, h]).abs().max().item()
    ev = (dq[h, :d, 64:96] - v[:, h]).abs().max().item()
    s_ref = (q[:, h] * scale) @ k[:, h].T
    es = (ds[h, :d, :d] - s_ref).abs().max().item()
    w = torch.softmax(s_ref * float(np.log(2.0)), -1)
    o_ref = w @ v[:, h]
    eo = (do[h, :d] - o_ref).abs().max().item()
    print(f"head {h}: |dQ| {eq:.3e} |dK| {ek:.3e} |dV| {ev:.3e} |dS| {es:.3e} (max|S| {s_ref.abs().max().item():.2f}) |dO| {eo:.3e}"
          f"  nan: qkv {int(torch.isnan(dq[h, :d]).sum())} s {int(torch.isnan(ds[h, :d, :d]).sum())} o {int(torch.isnan(do[h, :d]).sum())}")
wq = torch.softmax(torch.einsum("the,The->htT", q, k) / 32 ** 0.5, -1)
attn = torch.einsum("htT,The->the", wq, v).reshape(d, 256)
ref = lin(attn, P(nm["mha"] + "/linear"))
if fuse_out:
    print("DELTA (seq 0) max err", (dd[:d] - ref).abs().max().item(), "nan", int(torch.isnan(dd[:d]).sum()))
full = lin(torch.einsum("bnhtT,bnThe->bnthe", torch.softmax(torch.einsum(
    "bnthe,bnThe->bnhtT", lin(ln(z, P(nm["ln_q"])), P(nm["mha"] + "/query")).reshape(B, n, d, 8, 32),
    lin(ln(z, P(nm["ln_k"])), P(nm["mha"] + "/key")).reshape(B, n, d, 8, 32)) / 32 ** 0.5, -1),
    lin(ln(z, P(nm["ln_v"])), P(nm["mha"] + "/value")).reshape(B, n, d, 8, 32)).reshape(B, n, d, 256), P(nm["mha"] + "/linear"))
err = (delta.float().reshape(B, n, d, 128) - full).abs()
print("delta all sequences: max err", err.max().item(), "nan rows", int(torch.isnan(delta.float()).any(-1).sum()), "of", B * n * d)
per_seq = err.reshape(n, -1).max(-1).values
bad = torch.nonzero(~(per_seq < 0.1)).flatten()[:20].tolist()
print("first bad sequences:", bad)
