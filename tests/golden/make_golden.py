"""Generates tests/golden/*.npz: small seeded inputs + the fp64 oracle's edge probabilities.

The reference (JAX + Haiku) cannot be imported in this container, so these vectors
pin the ORACLE (oracle/avici_oracle.py, fp64), not the reference: "parity unpinned".  They guard the
oracle and the CUDA path against silent drift.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from avici_b200.params import init_params  # noqa: E402
from avici_b200.synthetic import simulate_data  # noqa: E402
from oracle import avici_oracle as O  # noqa: E402

CASES = {
    # name: (model kwargs, data kwargs, expects_counts, chunk)
    "lin_interv_l2": (dict(layers=2, cosine_temp_init=3.0, logit_bias_init=0.0), dict(d=12, n=30, n_interv=10, seed=3, domain="lin-gauss"), False, None),
    "rff_obs_l2": (dict(layers=2, cosine_temp_init=3.0, logit_bias_init=0.0), dict(d=16, n=64, seed=4, domain="rff-gauss"), False, None),
    "rff_obs_l2_chunk16": (dict(layers=2, cosine_temp_init=3.0, logit_bias_init=0.0), dict(d=16, n=64, seed=4, domain="rff-gauss"), False, 16),
    "gene_counts_l1": (dict(layers=1, cosine_temp_init=3.0, logit_bias_init=0.0), dict(d=10, n=48, seed=5, domain="gene-ecoli"), True, None),
    "wide_d130_l1": (dict(layers=1, cosine_temp_init=3.0, logit_bias_init=0.0), dict(d=130, n=20, seed=6, domain="lin-gauss"), False, None),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (mkw, dkw, counts, chunk) in CASES.items():
        g, x, interv = simulate_data(**dkw)
        params = init_params(mkw, seed=11, perturb=True, dtype=np.float64)
        cfg = dict(O.DEFAULT_CFG, layers=mkw["layers"])
        p = O.avici_call(params, x, cfg, interv=None if interv is None else interv.copy(), expects_counts=counts,
                         experimental_chunk_size=chunk, dtype=np.float64)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), x=x, interv=np.zeros(0) if interv is None else interv,
                            p=p, g=g, layers=mkw["layers"], param_seed=11, expects_counts=counts,
                            chunk=-1 if chunk is None else chunk)
        print(name, x.shape, "p range", float(p.min()), float(p.max()))


if __name__ == "__main__":
    main()
