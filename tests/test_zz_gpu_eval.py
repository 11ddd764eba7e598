"""GPU test of the batched evaluation step (forward + eval-loop metrics on the device, train.py:51-69,125-151) against
the numpy oracle applied to the same predictions.  Runs last (file name) on purpose: it composes pieces the other GPU
tests check one by one."""
import numpy as np
import pytest
import torch

import avici_b200
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def test_eval_batch_on_device(built_lib):
    model = avici_b200.random_init_model(dict(layers=1, cosine_temp_init=2.0, logit_bias_init=0.0), seed=1, perturb=True,
                                         compute_dtype="fp32")
    B, n, d = 3, 40, 9
    xs, gs = [], []
    for b in range(B):
        g, x, _ = avici_b200.simulate_data(d=d, n=n, seed=b, domain="lin-gauss")
        xs.append(x[..., 0] if x.ndim == 3 else x)
        gs.append(g)
    x = torch.from_numpy(np.stack(xs).astype(np.float32)).cuda()
    g = torch.from_numpy(np.stack(gs).astype(np.int32)).cuda()
    scalars, probs = avici_b200.eval_batch(model, x, g, decision_thresholds=(0.5, 0.3))
    assert isinstance(probs, torch.Tensor) and probs.is_cuda and tuple(probs.shape) == (B, d, d)
    p, gt = probs.cpu().numpy(), g.cpu().numpy()
    for thres in (0.5, 0.3):
        key = f"{thres}".replace(".", "_")
        want = MO.graph_metrics(p, gt, np.float32(thres))
        for name, field in (("shd", "shd"), ("edges", "edges"), ("cyclic", "cyclic"), ("precision", "precision"),
                            ("recall", "recall"), ("f1", "f1")):
            assert scalars[f"{name}_{key}"] == pytest.approx(float(np.mean(want[field])), abs=1e-12), (name, thres)
    for k in ("auroc", "auprc", "ap"):
        want = np.mean([MO.threshold_metrics(gt[b], p[b])[k] for b in range(B)])
        assert scalars[k] == pytest.approx(float(want), abs=1e-12), k
