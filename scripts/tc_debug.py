"""Structured diagnostics for the tcgen05 kernels (run on a B200): prints where errors sit (row / column
blocks) so that a wrong descriptor, swizzle or barrier phase can be told apart from one GPU round trip."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from avici_b200 import _lib  # noqa: E402

lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def blockmap(err, rb, cb):
    M, N = err.shape
    Mr, Nc = (M + rb - 1) // rb, (N + cb - 1) // cb
    out = np.zeros((Mr, Nc))
    for i in range(Mr):
        for j in range(Nc):
            out[i, j] = err[i * rb:(i + 1) * rb, j * cb:(j + 1) * cb].max()
    return out


def gemm_case(M, N, K, epi, pattern="rand"):
    g = torch.Generator(device="cuda").manual_seed(1)
    if pattern == "rand":
        a = torch.randn(M, K, device="cuda", generator=g)
        w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    else:  # structured: a[m,k] = (m%7)+k/64 ; w = identity-like -> out reveals the layout
        a = (torch.arange(M, device="cuda")[:, None] % 7 + torch.arange(K, device="cuda")[None, :] / 64.0).float()
        w = torch.zeros(N, K, device="cuda")
        w[torch.arange(min(N, K)), torch.arange(min(N, K))] = 1.0
    a, w = a.to(torch.bfloat16), w.to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    ref = a.float() @ w.float().T
    if epi == 2:
        out = torch.zeros(M, N, device="cuda")
    else:
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        if epi == 1:
            ref = ref.relu()
    rc = lib.avb_test_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epi, st())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().cpu().numpy()
    print(f"GEMM M={M} N={N} K={K} epi={epi} {pattern}: rc={rc} max_err={err.max():.4g} ref_max={ref.abs().max().item():.3g} "
          f"bad_frac={(err > 0.05 * max(1, ref.abs().max().item())).mean():.4f}")
    if err.max() > 0.05 * max(1, ref.abs().max().item()):
        bm = blockmap(err[:256], 32, 32)
        print(" block (32x32) max-err map of the first 256 rows:\n", np.array2string(bm, precision=2, max_line_width=200))
        print(" out[0,:16] =", out[0, :16].float().cpu().numpy())
        print(" ref[0,:16] =", ref[0, :16].cpu().numpy())
        print(" out[1,:16] =", out[1, :16].float().cpu().numpy())
        print(" ref[1,:16] =", ref[1, :16].cpu().numpy())
    return err.max()


def attn_case(B, n, d, H, axis, pattern="rand"):
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(B, n, d, 3 * H * 32, device="cuda", generator=g)
    if pattern == "uniform":  # q = 0 -> uniform attention -> out = mean of v over the sequence
        qkv[..., : H * 32] = 0
    qkv = qkv.to(torch.bfloat16)
    out = torch.zeros(B, n, d, H * 32, device="cuda", dtype=torch.bfloat16)
    rc = lib.avb_test_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, n, d, H, axis, st())
    torch.cuda.synchronize()
    x = qkv.float().reshape(B, n, d, 3, H, 32)
    q, k, v = x[..., 0, :, :], x[..., 1, :, :], x[..., 2, :, :]
    if axis == 1:
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    lg = torch.einsum("bsthe,bsThe->bshtT", q, k) * float(np.log(2.0))
    o = torch.einsum("bshtT,bsThe->bsthe", torch.softmax(lg, -1), v)
    if axis == 1:
        o = o.transpose(1, 2)
    ref = o.reshape(B, n, d, H * 32)
    err = (out.float() - ref).abs()
    print(f"ATTN B={B} n={n} d={d} H={H} axis={axis} {pattern}: rc={rc} max_err={err.max().item():.4g} "
          f"mean_err={err.mean().item():.4g} nan={torch.isnan(out.float()).sum().item()}")
    if err.max().item() > 0.05:
        T = d if axis == 0 else n
        e = err.amax(dim=(0, 3))  # [n,d]
        e_seq = (e.amax(0) if axis == 0 else e.amax(1)).cpu().numpy()
        print(" max err by position along the attended axis (first 160):", np.array2string(e_seq[:160], precision=2, max_line_width=200))
        eh = err.reshape(B, n, d, H, 32).amax(dim=(0, 1, 2)).cpu().numpy()
        print(" max err by (head, dim):\n", np.array2string(eh, precision=2, max_line_width=250))
        print(" out[0,0,0,:8] =", out[0, 0, 0, :8].float().cpu().numpy(), " ref =", ref[0, 0, 0, :8].cpu().numpy())
    return err.max().item()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("gemm", "all"):
        gemm_case(128, 128, 64, 2, "ident")
        gemm_case(128, 128, 64, 2)
        gemm_case(128, 128, 128, 2)
        gemm_case(128, 256, 128, 0)
        gemm_case(256, 128, 512, 2)
        gemm_case(1000, 768, 128, 0)
        gemm_case(100000, 768, 128, 0)
        gemm_case(100000, 128, 512, 2)
    if which in ("attn", "all"):
        attn_case(1, 1, 128, 1, 0, "uniform")
        attn_case(1, 1, 128, 1, 0)
        attn_case(1, 4, 100, 8, 0)
        attn_case(1, 128, 4, 2, 1)
        attn_case(1, 300, 8, 8, 1, "uniform")
        attn_case(1, 300, 8, 8, 1)
        attn_case(2, 1000, 100, 8, 0)
        attn_case(2, 1000, 100, 8, 1)
