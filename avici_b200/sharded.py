"""Multi-GPU execution of the forward.

Two levels:

* batches of independent datasets shard trivially: `batch_slice` gives each rank (or each device of one process,
  `infer_batch_devices`) its datasets, no data-path collective;
* ONE large dataset is sharded axially (BASELINE config 4): blocks that attend over d (even) run on an n-shard
  `[n/G, d, k]`, blocks that attend over n (odd) run on a d-shard `[n, d/G, k]`, and the fp32 residual stream is
  re-sharded between consecutive blocks with one all-to-all over NVLink.  The reference shards n only and leaves the
  collectives to XLA/GSPMD (avici/pretrain.py:129-139); per-token work (LayerNorm, projections, FFN) is layout-agnostic,
  so only the attention kernels care which axis is local.

The whole sharded forward is ONE C-ABI call per rank, `avb_forward_sharded` (kernels + grouped ncclSend/ncclRecv on one
stream).  The n-shard keeps its tokens in G column blocks `[G][n/G][d/G][k]`, so that both shardings are G contiguous
chunks of `[n/G][d/G][k]` and the exchange needs no pack / unpack pass: chunk p goes to rank p, the chunk received from
rank p lands at position p.  Two ways to drive it:

* one process per GPU under `torchrun` (`forward_axial`): the NCCL unique id travels through `torch.distributed`;
* one process, several GPUs, like the reference's `devices=` / `shard_if_possible=True` (`forward_axial_devices`): one
  handle and one host thread per device (ctypes releases the GIL), the communicator is created in-process.

`to_column_blocks` / `exchange_*` below restate the layout algebra in backend-agnostic torch code for the CPU tests
(gloo, world_size 2); the product path does not use them.
"""
import ctypes as C
import threading

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


# ------------------------------------------------------------------------------------------ layout algebra (tests)
def to_column_blocks(z_rows, G):
    """[n/G, d, ...] rows of one rank -> the column-blocked n-shard [G, n/G, d/G, ...] (RowMap with dg > 0)."""
    nl, d = z_rows.shape[:2]
    assert d % G == 0, "variables `d` must be divisible by the number of ranks for axial sharding"
    return z_rows.reshape(nl, G, d // G, *z_rows.shape[2:]).transpose(0, 1).contiguous()


def from_column_blocks(z_blk):
    """Inverse of `to_column_blocks`: [G, n/G, d/G, ...] -> [n/G, d, ...]."""
    G, nl, dl = z_blk.shape[:3]
    return z_blk.transpose(0, 1).reshape(nl, G * dl, *z_blk.shape[3:]).contiguous()


def exchange(z_chunks, group=None):
    """The exchange of `avb_forward_sharded` on [G, n/G, d/G, ...] chunk tensors: chunk p -> rank p, the chunk from rank
    p lands at position p.  n-shard (column blocks) -> d-shard ([n, d/G] = row blocks) and back are the SAME call."""
    import torch
    dist = _dist()
    if world_size(group) == 1:
        return z_chunks
    send = z_chunks.contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def reshard_n_to_d(z_loc, group=None):
    """[n/G, d, k] (rows of this rank) -> [n, d/G, k] (columns of this rank): column blocks + one all-to-all."""
    G = world_size(group)
    if G == 1:
        return z_loc
    recv = exchange(to_column_blocks(z_loc, G), group)
    return recv.reshape(recv.shape[0] * recv.shape[1], *recv.shape[2:])  # [G, n/G, d/G, k] is [n, d/G, k]


def reshard_d_to_n(z_loc, group=None):
    """[n, d/G, k] (columns of this rank) -> [n/G, d, k] (rows of this rank)."""
    G = world_size(group)
    if G == 1:
        return z_loc
    n = z_loc.shape[0]
    assert n % G == 0, "observations `n` must be divisible by the number of ranks for axial sharding"
    recv = exchange(z_loc.reshape(G, n // G, *z_loc.shape[1:]), group)
    return from_column_blocks(recv)


# ------------------------------------------------------------------------------------------ product path
def _call_sharded(eng, x, interv, standardize, r, G, stream):
    """avb_forward_sharded on this engine's device with the engine's own communicator."""
    import torch
    lib, h = eng.lib, eng.handle
    n, d = x.shape
    need = lib.avb_sharded_workspace_bytes(h, n, d, G)
    assert need > 0, "axial sharding needs `n` and `d` divisible by the device count"
    ws = eng.workspace(need)
    probs = torch.empty((d, d), dtype=torch.float32, device=x.device)
    eng._check(lib.avb_forward_sharded(h, x.data_ptr(), interv.data_ptr() if interv is not None else None, n, d,
                                       int(standardize), probs.data_ptr(), ws.data_ptr(), ws.numel(),
                                       C.c_void_p(stream), None, r, G))
    return probs


def ensure_comm(eng, group=None):
    """One NCCL communicator per engine for the ranks of `group` (created on first use, collectively)."""
    dist = _dist()
    G, r = world_size(group), rank(group)
    key = ("dist", id(group), G, r)
    if getattr(eng, "_comm_key", None) == key:
        return
    if G > 1:
        buf = C.create_string_buffer(128)
        if r == 0:
            _lib.check(eng.lib, None, eng.lib.avb_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        eng._check(eng.lib.avb_comm_init(eng.handle, box[0], r, G))
    eng._comm_key = key


def forward_axial(model, x, interv=None, group=None):
    """One dataset x[n,d] (the full matrix on every rank) -> probs[d,d] on every rank; one process per GPU.

    model: `AVICIModel`.  Requires G | n and G | d.
    """
    import torch
    G, r = world_size(group), rank(group)
    n, d = x.shape
    assert n % G == 0, "observations `n` must be divisible by `device_count` if sharding is enabled. "
    assert d % G == 0, "variables `d` must be divisible by `device_count` for axial sharding"
    eng = model.engine_on(x.device)
    ensure_comm(eng, group)
    x = x.contiguous()
    iv = None if interv is None else interv.to(torch.float32).contiguous()
    with torch.cuda.device(x.device):
        return _call_sharded(eng, x, iv, model._standardizer, r, G, torch.cuda.current_stream(x.device).cuda_stream)


def _run_threads(fns):
    """Run one callable per device on its own host thread (ctypes calls release the GIL); re-raise the first error."""
    out, err = [None] * len(fns), [None] * len(fns)

    def work(i):
        try:
            out[i] = fns[i]()
        except BaseException as e:  # noqa: BLE001 - re-raised below
            err[i] = e

    ts = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def forward_axial_devices(model, x, interv=None, devices=None):
    """One dataset sharded axially over several GPUs of THIS process (the reference's `devices=` +
    `shard_if_possible=True`, pretrain.py:124-143): one handle and one host thread per device, an in-process NCCL
    communicator.  x, interv: host or device tensors [n, d]; returns probs[d, d] on devices[0]."""
    import torch
    devs = [torch.device(dv) for dv in devices]
    G = len(devs)
    n, d = x.shape
    assert n % G == 0, "observations `n` must be divisible by `device_count` if sharding is enabled. "
    assert d % G == 0, "variables `d` must be divisible by `device_count` for axial sharding"
    engs = [model.engine_on(dv) for dv in devs]
    key = ("local", tuple(dv.index for dv in devs))
    if any(getattr(e, "_comm_key", None) != key for e in engs):
        buf = C.create_string_buffer(128)
        _lib.check(engs[0].lib, None, engs[0].lib.avb_comm_unique_id(buf))
        idb = bytes(buf.raw)
        # ncclCommInitRank blocks until every rank has joined: one thread per rank
        _run_threads([lambda e=e, r=r: e._check(e.lib.avb_comm_init(e.handle, idb, r, G)) for r, e in enumerate(engs)])
        for e in engs:
            e._comm_key = key

    def rank_fn(r):
        dv = devs[r]
        with torch.cuda.device(dv):
            xr = x.to(device=dv, dtype=torch.float32, non_blocking=True).contiguous()
            ir = None if interv is None else interv.to(device=dv, dtype=torch.float32, non_blocking=True).contiguous()
            st = torch.cuda.current_stream(dv)
            p = _call_sharded(engs[r], xr, ir, model._standardizer, r, G, st.cuda_stream)
            st.synchronize()
            return p

    return _run_threads([lambda r=r: rank_fn(r) for r in range(G)])[0]


def infer_batch_devices(model, x, interv=None, devices=None, chunk_size=None):
    """Independent datasets x[B, n, d] spread over several GPUs of this process (no collective): each device runs its
    `batch_slice`; the results are gathered on devices[0]."""
    import torch
    devs = [torch.device(dv) for dv in devices]
    G = len(devs)
    B = x.shape[0]

    def dev_fn(r):
        sl = batch_slice(B, r, G)
        if sl.start == sl.stop:
            return None
        dv = devs[r]
        with torch.cuda.device(dv):
            xr = x[sl].to(device=dv, dtype=torch.float32, non_blocking=True)
            ir = None if interv is None else interv[sl].to(device=dv, dtype=torch.float32, non_blocking=True)
            p = model.engine_on(dv).forward(xr, ir, standardize=model._standardizer, chunk_size=chunk_size)
            torch.cuda.current_stream(dv).synchronize()
            return p

    parts = _run_threads([lambda r=r: dev_fn(r) for r in range(G)])
    return torch.cat([p.to(devs[0]) for p in parts if p is not None], 0)
