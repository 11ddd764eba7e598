"""Generates tests/golden/full/*.npz: oracle outputs at the BASELINE configs' STATED sizes.

    python tests/golden/make_full_size_golden.py [case ...]

The fp64 torch-CPU oracle (oracle/avici_oracle_torch.py) takes minutes per case at these sizes, too long
to repeat inside every `-m gpu` run, so its result is committed: the seeded input (float32, exactly
what the CUDA path is fed), the oracle's cosine matrix u.v^T [d,d] in fp64 (the network output BEFORE
temperature / bias / sigmoid: p for any head follows exactly from it, model.py:136-141, 218-239), and
the seeds of the weights.  Like tests/golden/make_golden.py these pin the ORACLE, not the reference
("parity unpinned": JAX/Haiku cannot run here).

  case                what (BASELINE.json configs)
  c2_rff_n1000_d100   config 2: rff-gauss d=100 n=1000, two datasets of the bench batch (seeds 0, 1)
  c3_lin_iv_n1000_d100 config 3: lin-gauss d=100, 800 observational + 200 interventional rows, interv channel
  c4twin_n640_d256    config 4's twin (full size is hours of oracle time): one dataset n=640 d=256
  c5_gene_n2000_d200  config 5: count data d=200 n=2000, count standardizer
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from avici_b200.params import init_params  # noqa: E402
from avici_b200.synthetic import simulate_data  # noqa: E402
from oracle import avici_oracle as O  # noqa: E402
from oracle import avici_oracle_torch as OT  # noqa: E402

PARAM_SEED = 0
CASES = {
    "c2_rff_n1000_d100": dict(data=[dict(d=100, n=1000, seed=s, domain="rff-gauss") for s in (0, 1)], counts=False),
    "c3_lin_iv_n1000_d100": dict(data=[dict(d=100, n=800, n_interv=200, seed=3, domain="lin-gauss")], counts=False),
    "c4twin_n640_d256": dict(data=[dict(d=256, n=640, seed=4, domain="lin-gauss")], counts=False),
    "c5_gene_n2000_d200": dict(data=[dict(d=200, n=2000, seed=5, domain="gene-ecoli")], counts=True),
}


def oracle_cosines(params64, x32, iv32, counts):
    """fp64 oracle on the float32 inputs the CUDA path sees (promoted exactly)."""
    x = x32.astype(np.float64)
    iv = np.zeros_like(x) if iv32 is None else iv32.astype(np.float64)
    x2 = np.stack([x, iv], -1)
    x2 = O.standardize_count_simple(x2) if counts else O.standardize_default_simple(x2)
    with torch.no_grad():
        return OT.forward_cosines(params64, torch.as_tensor(x2), O.DEFAULT_CFG).numpy()


def main(names):
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    # the float32 weights the CUDA path is given (avici_b200.random_init_model(seed=PARAM_SEED, perturb=True)), promoted
    # exactly to fp64: oracle and CUDA path see identical weights and inputs
    params = OT.to_torch(init_params(dict(layers=8), seed=PARAM_SEED, perturb=True, dtype=np.float32), torch.float64)
    for name in names:
        case = CASES[name]
        xs, ivs, cos = [], [], []
        for dkw in case["data"]:
            g, x, iv = simulate_data(**dkw)
            x32 = x.astype(np.float32)
            iv32 = None if iv is None else iv.astype(np.float32)
            t0 = time.time()
            c = oracle_cosines(params, x32, iv32, case["counts"])
            print(f"{name} seed={dkw['seed']}: x{x32.shape} cos range [{c.min():.4f}, {c.max():.4f}] {time.time() - t0:.0f} s",
                  flush=True)
            xs.append(x32); ivs.append(iv32); cos.append(c)
        has_iv = ivs[0] is not None
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), x=np.stack(xs),
                            interv=np.stack(ivs).astype(np.uint8) if has_iv else np.zeros(0, np.uint8),
                            cos=np.stack(cos), param_seed=PARAM_SEED, layers=8, expects_counts=case["counts"])


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
