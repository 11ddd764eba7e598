"""torchrun -n G: axially-sharded forward of ONE dataset (all-to-all between blocks) vs the single-GPU forward."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import avici_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 256)
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g, x, iv = avici_b200.simulate_data(d=d, n=n - n // 5, n_interv=n // 5, seed=0, domain="lin-gauss", graph="erdos-renyi")
for dtype in ("fp32", "bf16"):
    model = avici_b200.random_init_model(dict(layers=layers), seed=0, perturb=True, compute_dtype=dtype)
    xt = torch.as_tensor(x, dtype=torch.float32).cuda()
    it = torch.as_tensor(iv, dtype=torch.float32).cuda()
    ref = model(x=xt, interv=it.clone(), shard_if_possible=False)   # single GPU, every rank
    for _ in range(2):
        out = model(x=xt, interv=it.clone())                        # sharded: world > 1 and shard_if_possible
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = model(x=xt, interv=it.clone())
    torch.cuda.synchronize(); dist.barrier()
    dt_sh = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = model(x=xt, interv=it.clone(), shard_if_possible=False)
    torch.cuda.synchronize()
    dt_1 = time.perf_counter() - t0
    err = (out - ref).abs().max().item()
    if rank == 0:
        print(f"axial[{dtype}] n={n} d={d} world={world}: max|dp| sharded vs single = {err:.3e}; "
              f"sharded {dt_sh * 1e3:.1f} ms, single-GPU {dt_1 * 1e3:.1f} ms", flush=True)
    assert err <= (1e-5 if dtype == "fp32" else 2e-2), err
dist.destroy_process_group()
