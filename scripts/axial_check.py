"""torchrun script: BASELINE config 4 (one dataset n=5000, d=1000, 16 blocks, bf16) through avb_forward_sharded on all
ranks; prints ms per forward, the NCCL share and (rank 0) the difference to the single-GPU forward.
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/axial_check.py [n d [single]]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import avici_b200  # noqa: E402
from avici_b200 import _lib, sharded  # noqa: E402

n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5000, 1000)
single = len(sys.argv) > 3
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G, r = dist.get_world_size(), dist.get_rank()
model = avici_b200.random_init_model(dict(layers=8), seed=0, perturb=True, compute_dtype="bf16")
eng = model.engine
x = torch.as_tensor(np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)).cuda()
f = lambda: sharded.forward_axial(model, x)  # noqa: E731
p = f()
torch.cuda.synchronize(); dist.barrier()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
eng.profile(True); eng.profile_read(reset=True); f(); prof = eng.profile_read(reset=True); eng.profile(False)
comm = torch.tensor([prof["comm"][0]], device="cuda")
dist.all_reduce(comm, op=dist.ReduceOp.MAX)
if r == 0:
    print(f"G={G} n={n} d={d}: {ms.item():.2f} ms per forward, comm {comm.item():.2f} ms ({100 * comm.item() / ms.item():.1f} %)",
          {k: round(v[0], 2) for k, v in prof.items()}, "NCCL env:", {k: v for k, v in os.environ.items() if k.startswith("NCCL_")})
    if single:
        p1 = eng.forward(x[None], None, standardize=_lib.AVB_STANDARDIZE_DEFAULT)[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.forward(x[None], None, standardize=_lib.AVB_STANDARDIZE_DEFAULT); e1.record(); torch.cuda.synchronize()
        print(f"single GPU {e0.elapsed_time(e1):.2f} ms, max|dp| sharded vs single {(p - p1).abs().max().item():.3e}")
dist.barrier()
dist.destroy_process_group()
