"""Generates tests/golden/standardize/standardize_cases.npz with outputs of THE REFERENCE's own numpy standardizers.

`avici/utils/data.py` (the numpy twins of the JAX standardizers on the hot path: `_onp_standardize_default` :66-78 ==
`jax_standardize_default_simple` data_jax.py:72-79, `_onp_cpm_standardizer` / `_onp_cpm_shift_scale` /
`_onp_standardize_count` :10-64 == `jax_standardize_count_simple` data_jax.py:48-59) imports with numpy alone, so it is
loaded straight from the read-only reference tree by file path (importing the `avici` package would pull jax).  The
vectors pin oracle/avici_oracle.py's `standardize_default_simple` / `standardize_count_simple` (rows a3 / a4 of
SURVEY.md 8a) and, through the GPU test, `avb_standardize`.  The reference tree does not exist on the GPU box, hence
the committed file.  Run from the repo root:  python tests/golden/standardize/make_standardize_golden.py
"""
import importlib.util
import os
import warnings

import numpy as np

REF = "/root/reference/avici/utils/data.py"


def load_reference_data_utils():
    spec = importlib.util.spec_from_file_location("ref_avici_utils_data", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cases():
    rng = np.random.default_rng(20241020)
    out = []
    # ---- real-valued data (default standardizer)
    for n, d in ((1, 1), (1, 5), (2, 3), (7, 1), (50, 20), (200, 50), (333, 17)):
        x = rng.normal(size=(n, d)) * rng.uniform(0.1, 30.0, size=(1, d)) + rng.uniform(-50, 50, size=(1, d))
        out.append((f"real_n{n}_d{d}", x, False))
    x = rng.normal(size=(40, 6)); x[:, 2] = 3.25; x[:, 4] = 0.0          # constant and all-zero columns: std == 0 -> 1
    out.append(("real_const_cols", x, False))
    out.append(("real_all_zero", np.zeros((12, 4)), False))
    out.append(("real_all_equal", np.full((9, 3), -7.5), False))
    x = rng.normal(size=(64, 8)) * 1e-6 + 1e6                               # tiny spread on a large offset
    out.append(("real_large_offset", x, False))
    # ---- count data (count standardizer)
    for n, d in ((1, 4), (3, 3), (48, 10), (120, 40), (257, 33)):
        lam = rng.gamma(2.0, 3.0, size=(n, d))
        x = rng.poisson(lam).astype(np.float64)
        x[rng.random((n, d)) < 0.3] = 0.0                                   # dropout zeros
        out.append((f"count_n{n}_d{d}", x, True))
    x = rng.poisson(5.0, size=(30, 8)).astype(np.float64); x[4] = 0.0; x[17] = 0.0   # all-zero rows: library size 0
    out.append(("count_zero_rows", x, True))
    x = rng.poisson(5.0, size=(30, 8)).astype(np.float64); x[:, 3] = 0.0              # all-zero column
    out.append(("count_zero_col", x, True))
    out.append(("count_all_zero", np.zeros((10, 5)), True))                            # every entry NaN after log2-CPM
    x = np.zeros((6, 4)); x[2, 1] = 9.0
    out.append(("count_single_nonzero", x, True))                                      # global std == 0
    x = np.zeros((6, 4)); x[2, 1] = 9.0; x[3, 1] = 9.0
    out.append(("count_two_equal_nonzero", x, True))
    out.append(("count_all_ones", np.ones((5, 7)), True))                              # all cells equal: std == 0
    return out


def reference_standardize(ref, x, counts, split=None):
    """x[n,d] -> the reference's standardized channel 0, as one [n,d] matrix.  The reference standardizes an
    (observational, interventional) pair with statistics over their concatenation; `split` rows go to x_obs."""
    n = x.shape[0]
    split = n if split is None else split
    x2 = np.stack([x, np.zeros_like(x)], -1)
    x_obs, x_int = x2[:split].copy(), x2[split:].copy()
    fn = ref._onp_standardize_count if counts else ref._onp_standardize_default
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        with np.errstate(all="ignore"):
            x_obs, x_int, _, _ = fn(x_obs, x_int, None, None)
    assert np.all(x_obs[..., 1] == 0) and np.all(x_int[..., 1] == 0)  # channel 1 untouched
    return np.concatenate([x_obs[..., 0], x_int[..., 0]], 0)


def main():
    ref = load_reference_data_utils()
    blob = {}
    names = []
    for name, x, counts in cases():
        for dt in (np.float64, np.float32):
            xin = x.astype(dt)
            y = reference_standardize(ref, xin, counts)
            if x.shape[0] >= 2:  # the obs/int split does not change the result (statistics are over the concatenation)
                y2 = reference_standardize(ref, xin, counts, split=x.shape[0] // 2)
                assert np.array_equal(y, y2, equal_nan=True), name
            key = f"{name}__{np.dtype(dt).name}"
            blob[key + "__x"] = xin
            blob[key + "__y"] = y
            blob[key + "__counts"] = np.asarray(counts)
            names.append(key)
            print(key, xin.shape, "finite" if np.isfinite(y).all() else "NON-FINITE", float(np.nanmin(y)), float(np.nanmax(y)))
    blob["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "standardize_cases.npz"), **blob)


if __name__ == "__main__":
    main()
