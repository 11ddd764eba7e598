"""Timing of the on-device graph metrics on the bench workload's output shape:
B = 64 predictions of [100, 100] at threshold 0.5 (what train.py:125-151 computes per evaluated batch), inputs resident
in HBM, against the numpy restatement of the reference's per-dataset loop on the host.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avici_b200 import _lib  # noqa: E402
from oracle import metrics_oracle as MO  # noqa: E402  (CPU baseline leg only)


def main():
    B, d = 64, 100
    rng = np.random.default_rng(0)
    probs = np.where(rng.random((B, d, d)) < 0.03, 0.75, 0.1).astype(np.float32)
    true = (rng.random((B, d, d)) < 0.02).astype(np.int32)
    lib = _lib.load()
    p, g = torch.from_numpy(probs).cuda(), torch.from_numpy(true).cuda()
    out = torch.empty((B, 8), dtype=torch.int64, device="cuda")
    nbytes = lib.avb_graph_metrics_workspace_bytes(B, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        import ctypes as C
        rc = lib.avb_graph_metrics(p.data_ptr(), g.data_ptr(), C.c_float(0.5), B, d, out.data_ptr(), ws.data_ptr(), nbytes, st)
        assert rc == 0
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    want = MO.graph_metrics(probs, true, 0.5)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = out.cpu().numpy()
    ok = bool(np.array_equal(got[:, 0], want["shd"]) and np.array_equal(got[:, 1], want["edges"]) and np.array_equal(got[:, 6], want["cyclic"]))
    # launches per call at d = 100: counts, init, 6 squarings + 2 products (+1 copy), trace
    print(json.dumps({"metric": "graph-metric batches/s", "workload": f"B={B}, d={d}, threshold 0.5", "gpu_ms_per_batch": gpu_ms,
                      "value": 1e3 / gpu_ms, "cpu_port_ms_per_batch": cpu_ms, "parity_vs_oracle": ok,
                      "algorithmic_bytes": int(B * d * d * 8), "fp64_flop": int(8 * 2 * B * d ** 3)}))


if __name__ == "__main__":
    main()
