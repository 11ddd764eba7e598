"""Timeline of dattn_tc_kernel's roles (CTA 0) from in-kernel clock64 stamps; needs a -DAVB_TRACE_ENABLE build
(AVB_LIB=<that build>).  Prints per-role event offsets for a window of units and the mean unit period."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import avici_b200

B, n, d = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 16000, int(sys.argv[2]) if len(sys.argv) > 2 else 100
fuse_out = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = avici_b200.random_init_model(dict(layers=1), seed=3, perturb=True, compute_dtype="bf16")
eng = model.engine
lib = eng.lib
g = torch.Generator(device="cuda").manual_seed(1)
z = torch.randn(B, n, d, 128, device="cuda", generator=g)
delta = torch.zeros(B * n * d, 128, device="cuda", dtype=torch.bfloat16)
guard = C.c_int32(-1)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
for rep in range(3):
    if rep == 2:
        lib.avb_debug_set_trace(C.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.avb_test_dattn_bf16(eng.handle, 0, z.data_ptr(), B, n, d, fuse_out, delta.data_ptr(), None, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"dattn S={B * n} T={d} fuse_out={fuse_out}: rc {rc} {ms:.3f} ms = {ms * 1e3 / (B * n) * 148:.2f} us per sequence per SM")
lib.avb_debug_set_trace(None)
t = trace.cpu().numpy().reshape(8, 512, 8)
if t.max() == 0:
    print("no stamps: not a -DAVB_TRACE_ENABLE build")
    sys.exit(0)
t0 = t[t > 0].min()
names = {0: ["top", "w_empty"], 1: ["top", "PV(u-1)", "S(u+1)"], 5: ["top", "G(u+2)", "OP(u-1)", "-", "-", "G:s_free"],
         2: ["top", "-", "conv:waits", "conv(u)"],
         3: ["top", "s_full", "ld+free", "exp", "pv_done", "st+p_full"], 4: ["top", "ln_loads", "oepi(u)", "oepi:pv_done", "ln_slice", "depi"]}
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 40
for role, rn in ((1, "attn issuer"), (5, "proj issuer"), (2, "conv"), (3, "softmax"), (4, "layernorm"), (0, "producer")):
    print(f"== {rn}: {names[role]}")
    rng = range(2 * lo, 2 * lo + 16) if role == 0 else range(lo, lo + 12)
    for step in rng:
        row = t[role, step, :len(names[role])].astype(np.int64)
        if row.max() == 0:
            continue
        rel = np.where(row > 0, row - t0, -1)
        print(f" step {step:3d} ", rel.tolist(), " next_in", int(t[role, step + 1, 0] - row[0]))
for role, rn in ((1, "attn issuer"), (3, "softmax")):
    tops = t[role, 16:400, 0].astype(np.int64)
    tops = tops[tops > 0]
    if len(tops) > 1:
        print(f"{rn}: mean unit period {(tops[-1] - tops[0]) / (len(tops) - 1):.0f} clocks over {len(tops)} units")
sm = t[3, 16:400]
ok = (sm[:, 0] > 0) & (sm[:, 5] > 0)
for a, b, nm in ((0, 1, "wait s_full"), (1, 2, "tmem ld"), (2, 3, "exps"), (3, 4, "wait pv_done"), (4, 5, "tmem st")):
    print(f"softmax {nm}: mean {(sm[ok, b] - sm[ok, a]).mean():.0f} clocks")
cv = t[2, 16:400].astype(np.int64)
ok = (cv[:, 0] > 0) & (cv[:, 3] > 0) & (cv[:, 2] > 0)
print(f"conv warps: waits {(cv[ok, 2] - cv[ok, 0]).mean():.0f}, conv {(cv[ok, 3] - cv[ok, 2]).mean():.0f} clocks per unit")
ln = t[4, 16:400].astype(np.int64)
ok = (ln[:, 0] > 0) & (ln[:, 5] > 0) & (ln[:, 3] > 0) & (ln[:, 2] > 0)
print(f"epilogue warps: wait pv_done {(ln[ok, 3] - ln[ok, 0]).mean():.0f}, oepi {(ln[ok, 2] - ln[ok, 3]).mean():.0f}, depi {(ln[ok, 5] - ln[ok, 2]).mean():.0f} clocks per unit")
for role, nm, cols in ((1, "attn issuer", ((0, 1, "PV"), (1, 2, "S"))), (5, "proj issuer", ((0, 1, "G"), (1, 2, "OP")))):
    m = t[role, 16:400].astype(np.int64)
    ok = (m[:, 0] > 0) & (m[:, 2] > 0)
    print(nm, ", ".join(f"{n} {(m[ok, b] - m[ok, a]).mean():.0f}" for a, b, n in cols))
oe = t[7, 16:400].astype(np.int64)
ok = (oe[:, 0] > 0) & (oe[:, 4] > 0)
print(f"oepi detail: tmem ld {(oe[ok, 1] - oe[ok, 0]).mean():.0f}, math {(oe[ok, 2] - oe[ok, 1]).mean():.0f}, wait op_done {(oe[ok, 3] - oe[ok, 2]).mean():.0f}, tmem st {(oe[ok, 4] - oe[ok, 3]).mean():.0f}")
de = t[6, 2:48].astype(np.int64)
ok = (de[:, 0] > 0) & (de[:, 5] > 0)
print(f"depi detail: wait d_full {(de[ok, 1] - de[ok, 0]).mean():.0f}, body {(de[ok, 5] - de[ok, 1]).mean():.0f}")
if (de[ok, 4] > 0).all():
    print(f"depi parts: chunks 0-1 tmem ld {(de[ok, 2] - de[ok, 1]).mean():.0f}, math + st.shared {(de[ok, 3] - de[ok, 2]).mean():.0f}, "
          f"chunks 2-3 + fence {(de[ok, 4] - de[ok, 3]).mean():.0f}, TMA issue {(de[ok, 5] - de[ok, 4]).mean():.0f}")
ok = (cv[:, 2] > 0) & (cv[:, 6] > 0)
print(f"conv detail: Q chunk {(cv[ok, 4] - cv[ok, 2]).mean():.0f}, K chunk {(cv[ok, 5] - cv[ok, 4]).mean():.0f}, V chunk {(cv[ok, 6] - cv[ok, 5]).mean():.0f}, fence+arrive {(cv[ok, 3] - cv[ok, 6]).mean():.0f}")
if len(sys.argv) > 5:
    # merged timeline of the events that concern units u0..u0+2, in time order
    u0 = int(sys.argv[5])
    ev = []
    for u in range(u0, u0 + (int(sys.argv[6]) if len(sys.argv) > 6 else 3)):
        add = lambda role, step, e, name: ev.append((int(t[role, step, e]), f"u{u} {name}")) if t[role, step, e] > 0 else None  # noqa: E731
        add(5, u - 2, 5, "G: s_free seen"); add(5, u - 2, 1, "G issued")
        add(2, u, 0, "conv top"); add(2, u, 2, "conv waits done (g_full, s_full(u-1), v_free)"); add(2, u, 3, "conv done -> qkv_ready")
        add(1, u - 1, 2, "S issued"); add(3, u, 0, "softmax top"); add(3, u, 1, "softmax: s_full"); add(3, u, 2, "softmax: S loaded -> s_free")
        add(3, u, 3, "softmax: exps done"); add(3, u, 4, "softmax: pv_done(u-1) seen"); add(3, u, 5, "softmax: P stored -> p_full")
        add(1, u + 1, 0, "attn issuer top (before PV)"); add(1, u + 1, 1, "PV issued")
        add(4, u, 0, "epi top"); add(4, u, 1, "epi: ln loads issued"); add(7, u, 0, "oepi: pv_done seen"); add(7, u, 1, "oepi: O loaded")
        add(7, u, 3, "oepi: op_done(u-1) seen"); add(7, u, 4, "oepi: O stored -> o_ready"); add(4, u, 4, "epi: ln slice done")
        add(5, u + 1, 1, "proj issuer: G(u+3) issued, now OP(u)"); add(5, u + 1, 2, "OP issued")
    for tt, nm in sorted(ev):
        print(f"{tt - t0:9d}  {nm}")
# per-head-index view: where in a sequence (h = u & 7) the unit period goes
for role, nm in ((1, "attn issuer"), (3, "softmax"), (4, "epilogue"), (2, "conv"), (5, "proj issuer")):
    tops = t[role, :, 0].astype(np.int64)
    per = {h: [] for h in range(8)}
    for u in range(16, 400):
        if tops[u] > 0 and tops[u + 1] > 0:
            per[u & 7].append(tops[u + 1] - tops[u])
    print(f"{nm}: period by h:", [int(np.mean(per[h])) if per[h] else -1 for h in range(8)])
ep, oe7 = t[4].astype(np.int64), t[7].astype(np.int64)
for h in range(8):
    us = [u for u in range(16, 400) if (u & 7) == h and ep[u, 0] > 0 and oe7[u, 0] > 0 and oe7[u, 4] > 0 and ep[u, 5] > 0]
    if not us:
        continue
    us = np.array(us)
    print(f"epilogue h={h}: top->pv_done seen {np.mean(oe7[us, 0] - ep[us, 0]):.0f}, ->O loaded {np.mean(oe7[us, 1] - oe7[us, 0]):.0f}, "
          f"->math {np.mean(oe7[us, 2] - oe7[us, 1]):.0f}, ->op_done {np.mean(oe7[us, 3] - oe7[us, 2]):.0f}, ->O stored {np.mean(oe7[us, 4] - oe7[us, 3]):.0f}, "
          f"->end of unit (ln slice / depi) {np.mean(ep[us, 5] - oe7[us, 4]):.0f}")
