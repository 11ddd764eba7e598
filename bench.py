#!/usr/bin/env python
"""bench.py — datasets/sec of the AVICI forward at n=1000, d=100 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W # CPU reference arm

A "step" is one forward pass (standardize -> 16 axial-attention blocks -> max-pool -> edge
probabilities) over one batch of `--batch` synthetic datasets (BASELINE config 2: rff-gauss d=100
n=1000, batch 64, bf16).  With N > 1 (torchrun, one rank per GPU) every rank runs its own batch of
independent datasets — weak scaling, no data-path collective (BASELINE config 3's partitioning).

One JSON line on rank 0:
  value      whole-job datasets/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through the host-buffer C-ABI call (avb_forward_host via
             AVICIModel.infer_batch): pinned host input -> H2D -> forward -> D2H inside the timed region
  roofline   dominant kernel (n-axis attention) algorithmic FLOP/s against the measured bf16 peak,
             from CUDA events on the launching stream over a second pass of the same K steps
  cpu_baseline  the torch-CPU restatement of the reference forward (oracle/) on a bounded sample
  axial      (WORLD_SIZE > 1 only, after the timed region) BASELINE config 4: one n=5000, d=1000 dataset sharded axially
             over all ranks vs the same forward on one GPU: ms, speed-up, NCCL share, max|dp| (bit-identical = 0)

The reference itself (JAX + Haiku) cannot be installed offline, so `--impl reference` times
the oracle restatement on the host cores — the tier's meaning of the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_OBS, N_VARS = 1000, 100
MODEL_KW = dict(layers=8, dim=128, key_size=32, num_heads=8, widening_factor=4, cosine_temp_init=2.0)


def flops_per_dataset(n, d, layers=8, k=128, H=8, dk=32, wf=4, out_dim=128):
    """Algorithmic FLOPs (MACs x 2; LayerNorm / softmax / elementwise excluded), SURVEY.md §8a."""
    nb = 2 * layers
    per_tok_lin = 2 * (3 * k * H * dk + H * dk * k + 2 * k * wf * k)
    return (n * d * (nb * per_tok_lin + (nb // 2) * 4 * H * dk * (n + d))
            + 2 * 2 * k * n * d + 2 * 2 * d * k * out_dim + 2 * d * d * out_dim)


def make_batch(B, seed0, domain="rff-gauss"):
    from avici_b200.synthetic import simulate_data
    xs = [simulate_data(N_VARS, N_OBS, seed=seed0 + s, domain=domain)[1] for s in range(B)]
    return np.stack(xs).astype(np.float32)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(p.get("bf16_tflops", 1590.0)), hbm=float(p.get("hbm_gbs", 6650.0)),
                    source="MEASURED_PEAKS.json")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_forward_timing(n_datasets, threads=None, x=None, return_probs=False):
    """Time the torch-CPU restated reference forward on `n_datasets` datasets of the bench workload (or on the given
    datasets x[B,n,d]); with return_probs also hand back the oracle's edge probabilities (the bench line's parity check)."""
    import torch
    from avici_b200.params import init_params
    from oracle import avici_oracle as O
    from oracle import avici_oracle_torch as OT
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must still use every host core it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(threads or avail)
    P = OT.to_torch(init_params(MODEL_KW, seed=0, perturb=True))
    x = torch.as_tensor(make_batch(n_datasets, 1000) if x is None else x[:n_datasets])
    probs = []
    with torch.no_grad():
        t0 = time.perf_counter()
        for b in range(n_datasets):
            probs.append(OT.forward_probs(P, x[b], None, O.DEFAULT_CFG))
        dt = time.perf_counter() - t0
    if return_probs:
        return n_datasets / dt, dt, torch.get_num_threads(), torch.stack(probs).numpy()
    return n_datasets / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 1  # bounded sample: one (n=1000, d=100) dataset per step (~10-30 s of CPU work each)
    for _ in range(min(args.warmup, 1)):
        cpu_forward_timing(1)
    t0 = time.perf_counter()
    cores = 0
    for _ in range(args.steps):
        _, _, cores = cpu_forward_timing(per_step)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": "datasets/sec fwd @ n=1000,d=100", "value": value, "unit": "datasets/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"rff-gauss d={N_VARS} n={N_OBS}, CPU restatement of the reference forward "
                               f"(JAX/Haiku not installable offline), {per_step} dataset per step"},
        "cpu_baseline": {"value": value, "unit": "datasets/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {per_step} dataset(s) of the bench workload, torch-CPU fp32"},
        "e2e": {"value": value, "unit": "datasets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if absent)."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            e = json.load(f)[key]
        return e["dram_read_bytes"] + e["dram_write_bytes"]
    except Exception:
        return None


AXIAL_N, AXIAL_D = 5000, 1000  # BASELINE config 4: one large dataset, axially sharded over the ranks


def axial_leg(model, world, rank, dist, timed, barrier):
    """BASELINE config 4 (runs behind the main timed region when WORLD_SIZE > 1): ONE rff-gauss dataset n=5000, d=1000
    through `avb_forward_sharded` on all ranks (n-shards <-> d-shards, an NCCL all-to-all of the fp32 residual stream
    between consecutive blocks) against the same forward on rank 0's GPU alone: ms, speed-up, comm share, max|dp|."""
    import torch
    from avici_b200 import _lib, sharded
    n, d = AXIAL_N, AXIAL_D
    if n % world or d % world:
        return {"skipped": f"n={n}, d={d} not divisible by {world} ranks"}
    x = torch.empty((n, d), dtype=torch.float32, device="cuda")
    if rank == 0:  # ~12 s of numpy on one rank; the others receive the matrix
        from avici_b200.synthetic import simulate_data
        x.copy_(torch.as_tensor(simulate_data(d, n, seed=4, domain="rff-gauss")[1].astype(np.float32)))
    dist.broadcast(x, src=0)
    eng = model.engine
    f = lambda: sharded.forward_axial(model, x)  # noqa: E731
    p_sh = f()  # creates the communicator, sizes the workspace
    steps = 3
    ms_sh = timed(f, steps) / steps
    eng.profile(True)
    eng.profile_read(reset=True)
    timed(f, 1)
    prof = eng.profile_read(reset=True)
    eng.profile(False)
    comm_ms = torch.tensor([prof["comm"][0]], device="cuda")
    dist.all_reduce(comm_ms, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:  # the same dataset on one GPU (the other ranks wait at the barrier)
        g = lambda: eng.forward(x[None], None, standardize=_lib.AVB_STANDARDIZE_DEFAULT)  # noqa: E731
        p_1 = g()[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g()
        e1.record()
        torch.cuda.synchronize()
        ms_1 = e0.elapsed_time(e1)
        F = flops_per_dataset(n, d)
        out = {"workload": f"rff-gauss n={n} d={d}, one dataset, 16 blocks, bf16, sharded axially over {world} GPUs "
                           "(n-shard <-> d-shard, fp32 all-to-all between blocks)",
               "ms_sharded": ms_sh, "ms_single_gpu": ms_1, "speedup": ms_1 / ms_sh, "n_gpus": world,
               "comm_ms_per_forward": float(comm_ms.item()), "comm_share": float(comm_ms.item()) / ms_sh,
               "exchange_bytes_per_gpu_per_forward": int(15 * (world - 1) / world * n * d * 128 * 4 / world),
               "tflops_sharded": F / (ms_sh * 1e-3) / 1e12, "tflops_single_gpu": F / (ms_1 * 1e-3) / 1e12,
               "max_abs_dp_vs_single_gpu": float((p_sh - p_1).abs().max().item()),
               "steps": steps, "timing": "CUDA events, max over ranks"}
    barrier()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import avici_b200
    from avici_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    x_host = make_batch(B, seed0=rank * B)  # every rank owns B independent datasets (weak scaling)
    model = avici_b200.random_init_model(MODEL_KW, seed=0, perturb=True, compute_dtype=args.dtype)
    eng = model.engine
    x_dev = torch.as_tensor(x_host).cuda()
    out = torch.empty((B, N_VARS, N_VARS), dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step = lambda: eng.forward(x_dev, None, standardize=_lib.AVB_STANDARDIZE_DEFAULT, out=out)  # noqa: E731
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    ms_total = timed(step, args.steps)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)

    # end to end through the host-buffer C-ABI entry point (pinned host memory in, host memory out)
    x_pin = torch.as_tensor(x_host).pin_memory()
    p_pin = torch.empty((B, N_VARS, N_VARS), dtype=torch.float32).pin_memory()
    e2e_step = lambda: eng.forward_host(x_pin.numpy(), None, out=p_pin.numpy())  # noqa: E731
    e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(dt.item())

    # per-kernel-category timing: same K steps again with CUDA events around every launch
    eng.profile(True)
    eng.profile_read(reset=True)
    ms_prof = timed(step, args.steps)
    prof = eng.profile_read(reset=True)
    eng.profile(False)

    # CUDA-graph variant of the same step: one graph launch instead of ~430 kernel launches; what it saves is
    # the launch gaps, so at batch 64 it should equal the plain number (reported, not the headline)
    graph_ms = None
    try:
        graph, _ = eng.capture(x_dev, None, standardize=_lib.AVB_STANDARDIZE_DEFAULT)
        graph.replay()
        graph_ms = timed(graph.replay, args.steps) / args.steps
        del graph
    except Exception as e:  # noqa: BLE001 - optional leg
        graph_ms = f"{type(e).__name__}: {e}"[:200]

    axial = None
    if world > 1 and not args.no_axial:
        del x_dev, out  # the axial leg needs the memory of rank 0 for the single-GPU run of the same dataset
        eng._ws = None
        torch.cuda.empty_cache()
        try:
            axial = axial_leg(model, world, rank, dist, timed, barrier)
        except Exception as e:  # the headline line must still be printed
            axial = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        peaks = measured_peaks()
        F = flops_per_dataset(N_OBS, N_VARS)
        # dominant kernel: attention over n.  Algorithmic FLOPs per launch = B*n*d tokens * 4*H*dk*n (QK^T + PV)
        attn_flops = B * N_OBS * N_VARS * 4 * 8 * 32 * N_OBS
        ms_attn, cnt_attn = prof["attn_n"]
        achieved = attn_flops / (ms_attn / max(cnt_attn, 1) * 1e-3) / 1e12 if ms_attn > 0 else 0.0
        breakdown = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
        # CPU leg: the oracle restatement on dataset 0 of this very batch, same weights -> timing AND parity of the benched
        # configuration (fp32 torch-CPU oracle; its own distance to the fp64 oracle is < 2e-5, tests/test_oracle.py)
        parity = None
        if not args.no_cpu_baseline:
            cpu_val, cpu_dt, cores, p_cpu = cpu_forward_timing(1, x=x_host, return_probs=True)
            tol = 1e-2 if args.dtype == "bf16" else 1e-4
            err = float(np.abs(p_pin.numpy()[0].astype(np.float64) - p_cpu[0]).max())
            margin = np.abs(p_cpu[0] - 0.5) <= tol
            parity = {"max_abs_dp": err, "tol": tol, "ok": bool(err <= tol),
                      "graph_equal_outside_margin": bool(np.array_equal((p_pin.numpy()[0] > 0.5)[~margin], (p_cpu[0] > 0.5)[~margin])),
                      "what": "GPU probs (host-buffer call) vs the fp32 torch-CPU oracle on dataset 0 of the benched batch, "
                              "same random-init weights (reference head init: temp 2, bias -3)"}
        else:
            cpu_val, cpu_dt, cores = None, None, None
        line = {
            "metric": "datasets/sec fwd @ n=1000,d=100", "value": value, "unit": "datasets/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"rff-gauss d={N_VARS} n={N_OBS} batch {B} datasets per GPU, scm-v0-shaped "
                                   f"random-init BaseModel (16 blocks, dim 128, 8 heads x 32, FFN 512), {args.dtype} forward",
                       "batch_per_gpu": B, "parallelism": f"dataset-parallel x{world}",
                       "l2": "activations per step (>20 GB at batch 64) far exceed the 126 MB L2; no flush needed"},
            "tensor_frac_of_measured_sustained": F * value / world / (peaks["tflops"] * 1e12),
            "flops_per_dataset": F,
            "e2e": {"value": e2e_value, "unit": "datasets/s", "h2d_bytes_per_step": int(x_host.nbytes),
                    "d2h_bytes_per_step": int(B * N_VARS * N_VARS * 4), "steps": e2e_steps},
            "gpu_launches": launches,
            "cuda_graph": ({"ms_per_step": graph_ms, "value": world * B / (graph_ms * 1e-3), "what": "the same step replayed as one CUDA graph"}
                           if isinstance(graph_ms, float) else {"error": graph_ms}),
            "parity": parity,
            "clocks": clocks,
            "roofline": {"kernel": "attention_tc_kernel (attend over n)", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         "traffic": ncu_traffic("attention_tc_kernel<1>:axis_n:B64") if B == 64 else None,
                         "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                         "algorithmic_bytes": B * N_OBS * N_VARS * 2048,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained",
                         "ms_per_launch": ms_attn / max(cnt_attn, 1), "launches_timed": cnt_attn,
                         # head dim 32: the softmax, not the MMA, is the floor of this kernel
                         "co_limit": {"what": "exp2 per launch on MUFU (16/clk/SM) + FMA-pipe polynomial (7 of every 16 exponentials under the score bound, 5 of 16 otherwise)",
                                      "exps_per_launch": B * N_OBS * N_VARS * 8 * N_OBS,
                                      "achieved_Texp_s": B * N_OBS * N_VARS * 8 * N_OBS / (ms_attn / max(cnt_attn, 1) * 1e-3) / 1e12
                                      if ms_attn > 0 else 0.0,
                                      "mufu_peak_Texp_s": 16 * 148 * 1.965e9 / 1e12}},
            "kernel_ms_per_step": breakdown, "profiled_ms_per_step": ms_prof / args.steps,
            "axial": axial,
            "cpu_baseline": None if cpu_val is None else {
                "value": cpu_val, "unit": "datasets/s", "cores": cores, "kind": "port",
                "sample": f"1 dataset of the bench workload (n={N_OBS}, d={N_VARS}), torch-CPU fp32 restatement, {cpu_dt:.1f} s"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-axial", action="store_true", help="skip the config-4 leg that runs when WORLD_SIZE > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
