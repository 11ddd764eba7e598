"""Multi-GPU execution of the forward: one process per GPU, `torch.distributed` (NCCL) plumbing.

Two levels:

* batches of independent datasets shard trivially: `batch_slice` gives each rank its datasets, no
  data-path collective;
* ONE large dataset is sharded axially: blocks that attend over d (even) run on an n-shard
  `[n/G, d, k]`, blocks that attend over n (odd) run on a d-shard `[n, d/G, k]`, and the residual
  stream is re-sharded between consecutive blocks with one all-to-all over NVLink
  (`reshard_n_to_d` / `reshard_d_to_n`).  The reference shards n only and leaves the collectives
  to XLA/GSPMD (avici/pretrain.py:129-139); per-token work (LayerNorm, projections, FFN) is
  layout-agnostic, so only the attention kernels care which axis is local.

The layout transforms below are backend-agnostic torch code (tested on CPU with gloo); the
per-shard compute goes through the stage-level C-ABI (`avb_standardize/embed/block/pool/head`).
"""
import ctypes as C

from . import _lib


def _dist():
    import torch.distributed as dist
    return dist


def world_size(group=None):
    dist = _dist()
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    dist = _dist()
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def batch_slice(B, r, G):
    """Contiguous, balanced slice of B independent datasets owned by rank r of G."""
    base, rem = divmod(B, G)
    lo = r * base + min(r, rem)
    return slice(lo, lo + base + (1 if r < rem else 0))


def reshard_n_to_d(z_loc, group=None):
    """[n/G, d, k] (rows of this rank) -> [n, d/G, k] (columns of this rank) by one all-to-all."""
    import torch
    dist = _dist()
    G = world_size(group)
    nl, d, k = z_loc.shape
    assert d % G == 0, "variables `d` must be divisible by the number of ranks for axial sharding"
    if G == 1:
        return z_loc
    # destination-major packing: chunk q = my rows x rank q's columns
    send = z_loc.reshape(nl, G, d // G, k).permute(1, 0, 2, 3).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # source-major = row-block order: [G, n/G, d/G, k] is already [n, d/G, k]
    return recv.reshape(G * nl, d // G, k)


def reshard_d_to_n(z_loc, group=None):
    """[n, d/G, k] (columns of this rank) -> [n/G, d, k] (rows of this rank) by one all-to-all."""
    import torch
    dist = _dist()
    G = world_size(group)
    n, dl, k = z_loc.shape
    assert n % G == 0, "observations `n` must be divisible by the number of ranks for axial sharding"
    if G == 1:
        return z_loc
    send = z_loc.contiguous()  # chunk q = rank q's rows x my columns, already contiguous
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # chunk q of recv = my rows x rank q's columns -> interleave the column blocks
    return recv.reshape(G, n // G, dl, k).permute(1, 0, 2, 3).reshape(n // G, G * dl, k).contiguous()


def forward_axial(model, x, interv=None, group=None):
    """One dataset x[n,d] (replicated on every rank) -> probs[d,d] on every rank.

    model: `AVICIModel`.  Requires G | n and G | d.
    """
    import torch
    dist = _dist()
    eng = model.engine
    lib, h = eng.lib, eng.handle
    G, r = world_size(group), rank(group)
    n, d = x.shape
    assert n % G == 0, "observations `n` must be divisible by `device_count` if sharding is enabled. "
    assert d % G == 0, "variables `d` must be divisible by `device_count` for axial sharding"
    k = eng.kw["dim"]
    nb = 2 * eng.kw["layers"]
    dev = x.device
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    x = x.contiguous()
    xs = torch.empty_like(x)
    scratch = torch.empty((n + d + 8) * 8, dtype=torch.uint8, device=dev)
    eng._check(lib.avb_standardize(h, x.data_ptr(), 1, n, d, int(model._standardizer), xs.data_ptr(),
                                   scratch.data_ptr(), scratch.numel(), st))
    nl, dl = n // G, d // G
    xs_loc = xs[r * nl:(r + 1) * nl].contiguous()
    iv_loc = None if interv is None else interv[r * nl:(r + 1) * nl].to(torch.float32).contiguous()
    z = torch.empty((nl, d, k), dtype=torch.float32, device=dev)
    eng._check(lib.avb_embed(h, xs_loc.data_ptr(), iv_loc.data_ptr() if iv_loc is not None else None,
                             nl * d, z.data_ptr(), st))
    ws_bytes = max(lib.avb_block_workspace_bytes(h, 1, nl, d), lib.avb_block_workspace_bytes(h, 1, n, dl))
    ws = eng.workspace(ws_bytes)
    for i in range(nb):
        if i % 2 == 0:   # attend over d on an n-shard
            if i > 0:
                z = reshard_d_to_n(z, group)
            eng._check(lib.avb_block(h, i, z.data_ptr(), 1, nl, d, _lib.AVB_AXIS_D, ws.data_ptr(), ws.numel(), st))
        else:            # attend over n on a d-shard
            z = reshard_n_to_d(z, group)
            eng._check(lib.avb_block(h, i, z.data_ptr(), 1, n, dl, _lib.AVB_AXIS_N, ws.data_ptr(), ws.numel(), st))
    # z is d-sharded [n, d/G, k]: final LN + max over n is local (model.py:86-89)
    pooled_loc = torch.empty((dl, k), dtype=torch.float32, device=dev)
    eng._check(lib.avb_pool(h, z.data_ptr(), 1, n, dl, pooled_loc.data_ptr(), 0, st))
    pooled = torch.empty((G, dl, k), dtype=torch.float32, device=dev)
    if G > 1:
        dist.all_gather_into_tensor(pooled, pooled_loc, group=group)
    else:
        pooled.copy_(pooled_loc.unsqueeze(0))
    probs = torch.empty((d, d), dtype=torch.float32, device=dev)
    hs = torch.empty(2 * d * eng.kw["out_dim"] * 4, dtype=torch.uint8, device=dev)
    eng._check(lib.avb_head(h, pooled.data_ptr(), 1, d, probs.data_ptr(), hs.data_ptr(), hs.numel(), st))
    return probs
