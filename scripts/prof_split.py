"""One fp32-mode block pair on a config-5-shaped problem (for ncu captures of the split-precision kernels)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import avici_b200
from avici_b200 import _lib
d = np.load("tests/golden/full/c5_gene_n2000_d200.npz")
x = torch.as_tensor(d["x"]).cuda()
model = avici_b200.random_init_model(dict(layers=1), seed=0, perturb=True, expects_counts=True, compute_dtype="fp32")
eng = model.engine
for _ in range(2):
    eng.forward(x, None, standardize=_lib.AVB_STANDARDIZE_COUNT)
torch.cuda.synchronize()
