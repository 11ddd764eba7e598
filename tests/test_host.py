"""CPU tests of the host-side logic: Haiku naming / init, checkpoint IO, simulate_data, input validation
of the reference-facing wrapper, and the multi-rank layout transforms (gloo, world_size 2)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avici_b200
from avici_b200 import sharded
from avici_b200.params import flatten_params, init_params, param_shapes
from avici_b200.pretrain import ModelState, load_checkpoint, save_checkpoint


class _FakeOptState:
    """stands in for an optax state class inside a reference pickle"""
    def __init__(self):
        self.mu = np.zeros(3)


_FakeOptState.__qualname__ = _FakeOptState.__name__ = "ScaleByLambState"


def test_haiku_names_and_flatten():
    shapes = param_shapes()
    names = list(shapes)
    assert names[0] == "BaseModel/linear" and names[-1] == "BaseModel"
    assert "BaseModel/layer_norm_66" in shapes and "BaseModel/linear_34" in shapes
    assert "BaseModel/multi_head_attention_15/linear" in shapes and "BaseModel/layer_norm_67" not in shapes
    assert shapes["BaseModel/multi_head_attention/query"]["w"] == (128, 256)
    flat = flatten_params(init_params(dict(layers=1), seed=0))
    assert all(a.dtype == np.float32 and a.flags.c_contiguous for _, a in flat)
    assert ("BaseModel/learned_temp" in dict(flat)) and dict(flat)["BaseModel/linear/w"].shape == (2, 128)


def test_init_distributions():
    p = init_params(seed=0)
    w = p["BaseModel/linear_1"]["w"]  # kaiming-uniform, fan_in 128
    assert abs(np.abs(w).max() - np.sqrt(6 / 128)) < 1e-3 and abs(w.std() - np.sqrt(2 / 128)) < 2e-3
    q = p["BaseModel/multi_head_attention/query"]["w"]  # truncated normal, var 2/fan_in
    assert abs(q.std() - np.sqrt(2 / 128)) < 3e-3 and np.abs(q).max() <= 2 * np.sqrt(2 / 128) / 0.8796 + 1e-6
    assert np.all(p["BaseModel/layer_norm"]["scale"] == 1) and np.all(p["BaseModel/linear"]["b"] == 0)
    pp = init_params(seed=0, perturb=True)
    assert pp["BaseModel/layer_norm"]["scale"].std() > 0.05 and pp["BaseModel/linear_2"]["b"].std() > 0.005


def test_checkpoint_roundtrip_and_legacy_alias(tmp_path):
    kw = dict(layers=1, dim=128, key_size=32, num_heads=8, widening_factor=4, dropout=0.1, some_deprecated_key=3)
    params = init_params(kw, seed=3)
    folder = save_checkpoint(tmp_path / "ckpt", params, kw, {"mask_diag": True, "acyclicity": 0.1, "old_kwarg": 1})
    state, kwargs = load_checkpoint(folder)
    assert kwargs["inference_model_kwargs"]["acyclicity"] is None
    assert np.array_equal(state.params["BaseModel/linear"]["w"], params["BaseModel/linear"]["w"])
    # a pickle written by the reference names avici.backbone.ModelState (or the legacy amortibs alias) and
    # carries optax classes we do not have: emulate both
    import sys, types
    mod = types.ModuleType("amortibs"); sub = types.ModuleType("amortibs.backbone")
    fake_optax = types.ModuleType("optax_fake")
    ScaleByLambState = _FakeOptState
    from collections import namedtuple
    MS = namedtuple("ModelState", ModelState._fields)
    MS.__module__ = "amortibs.backbone"; sub.ModelState = MS
    ScaleByLambState.__module__ = "optax_fake"; fake_optax.ScaleByLambState = ScaleByLambState
    sys.modules.update({"amortibs": mod, "amortibs.backbone": sub, "optax_fake": fake_optax})
    try:
        blob = pickle.dumps(MS(7, None, ScaleByLambState(), params, None, None, None))
    finally:
        for k in ("amortibs", "amortibs.backbone", "optax_fake"):
            sys.modules.pop(k)
    (tmp_path / "legacy").mkdir()
    (tmp_path / "legacy" / "checkpoint_0000007.pkl").write_bytes(blob)
    (tmp_path / "legacy" / "kwargs.json").write_text(json.dumps(
        {"neural_net_kwargs": {"model_kwargs": kw}, "inference_model_kwargs": {}}))
    state2, _ = load_checkpoint(tmp_path / "legacy")
    assert state2.step == 7 and np.array_equal(state2.params["BaseModel/linear"]["w"], params["BaseModel/linear"]["w"])
    model = avici_b200.load_pretrained(checkpoint_dir=str(tmp_path / "legacy"), expects_counts=False, verbose=False)
    assert model._model.model_kwargs["layers"] == 1 and "some_deprecated_key" not in model._model.model_kwargs


class _FakeFlatMapping:
    """stands in for haiku._src.data_structures.FlatMapping (dm-haiku < 0.0.10): a Mapping that is NOT a dict and
    pickles as FlatMapping(<payload>)"""
    def __init__(self, payload):
        self._payload = payload

    def __reduce__(self):
        return (type(self), (self._payload,))


class _FakeFlatComponents:
    """stands in for haiku's FlatComponents(leaves, structure) NamedTuple; `structure` is an opaque pytree def"""
    def __init__(self, leaves, structure):
        self.leaves, self.structure = leaves, structure

    def __reduce__(self):
        return (type(self), (self.leaves, self.structure))


class _FakePyTreeDef:
    def __reduce__(self):
        return (type(self), ())


@pytest.mark.parametrize("form", ["dict", "components"])
def test_old_haiku_flatmapping_checkpoint(tmp_path, form):
    """A checkpoint whose params are haiku FlatMapping containers (published models, dm-haiku 0.0.8/0.0.9) loads: the
    restricted unpickler stubs the classes and `_plain_params` rebuilds the dict, for both pickled forms."""
    import sys, types
    kw = dict(layers=1, dim=128, key_size=32, num_heads=8, widening_factor=4)
    params = init_params(kw, seed=5, perturb=True)
    hk = types.ModuleType("haiku"); src = types.ModuleType("haiku._src"); ds = types.ModuleType("haiku._src.data_structures")
    jl = types.ModuleType("jaxlib_fake")
    for cls, name, mod in ((_FakeFlatMapping, "FlatMapping", ds), (_FakeFlatComponents, "FlatComponents", ds),
                           (_FakePyTreeDef, "PyTreeDef", jl)):
        cls.__qualname__ = cls.__name__ = name
        cls.__module__ = mod.__name__
        setattr(mod, name, cls)
    mods = {"haiku": hk, "haiku._src": src, "haiku._src.data_structures": ds, "jaxlib_fake": jl}
    sys.modules.update(mods)
    try:
        if form == "dict":  # FlatMapping(dict) at both levels
            wrapped = _FakeFlatMapping({m: _FakeFlatMapping(dict(l)) for m, l in params.items()})
        else:               # FlatMapping(FlatComponents(leaves in sorted key order, treedef))
            leaves = [params[m][l] for m in sorted(params) for l in sorted(params[m])]
            wrapped = _FakeFlatMapping(_FakeFlatComponents(tuple(leaves), _FakePyTreeDef()))
        blob = pickle.dumps(ModelState(3, None, None, wrapped, None, None, None))
    finally:
        for k in mods:
            sys.modules.pop(k)
    (tmp_path / "old").mkdir()
    (tmp_path / "old" / "checkpoint_0000003.pkl").write_bytes(blob)
    (tmp_path / "old" / "kwargs.json").write_text(json.dumps({"neural_net_kwargs": kw, "inference_model_kwargs": {}}))
    model = avici_b200.load_pretrained(checkpoint_dir=str(tmp_path / "old"), expects_counts=False, verbose=False)
    assert isinstance(model.params, dict) and set(model.params) == set(params)
    for m, leaves in params.items():
        assert isinstance(model.params[m], dict)
        for l, arr in leaves.items():
            assert np.array_equal(model.params[m][l], arr), (m, l)
    assert [n for n, _ in flatten_params(model.params)].count("BaseModel/linear/w") == 1


def test_unpickler_refuses_dangerous_builtins(tmp_path):
    """The allow-list is explicit: a pickle that names builtins.eval / os.system gets an inert stub, not the callable."""
    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('os').getpid()",))
    (tmp_path / "evil").mkdir()
    (tmp_path / "evil" / "checkpoint_0000001.pkl").write_bytes(pickle.dumps(ModelState(1, None, Evil(), {}, None, None, None)))
    (tmp_path / "evil" / "kwargs.json").write_text(json.dumps({"neural_net_kwargs": {}, "inference_model_kwargs": {}}))
    state, _ = load_checkpoint(tmp_path / "evil")
    assert type(state.opt_state).__name__ == "eval" and not isinstance(state.opt_state, int)  # stub instance, never called


def test_load_pretrained_argument_errors(tmp_path):
    with pytest.raises(AssertionError):
        avici_b200.load_pretrained()
    with pytest.raises(ValueError):
        avici_b200.load_pretrained(download="nope")
    with pytest.raises(AssertionError):
        avici_b200.load_pretrained(checkpoint_dir=str(tmp_path))  # expects_counts missing
    with pytest.raises(FileNotFoundError):
        avici_b200.load_pretrained(download="scm-v0", cache_path=str(tmp_path), verbose=False)


@pytest.mark.parametrize("domain", ["lin-gauss", "lin-gauss-heterosked", "lin-laplace-cauchy", "rff-gauss",
                                    "rff-gauss-heterosked", "rff-laplace-cauchy", "gene-ecoli"])
def test_simulate_data_contract(domain):
    g, x, interv = avici_b200.simulate_data(d=15, n=40, n_interv=10, seed=0, domain=domain)
    assert g.shape == (15, 15) and g.dtype.kind == "i" and x.shape == (50, 15) and interv.shape == (50, 15)
    assert set(np.unique(interv)) <= {0.0, 1.0} and interv.sum() == 10
    assert np.trace(np.linalg.matrix_power(np.eye(15) + g, 15)) == 15  # DAG (metrics.py:34-43 style check)
    g2, x2, none = avici_b200.simulate_data(d=15, n=40, seed=0, domain=domain)
    assert none is None and x2.shape == (40, 15)
    if domain == "gene-ecoli":
        assert np.all(x >= 0) and np.allclose(np.mod(x, 1), 0) and (x == 0).mean() > 0.02
    _, x3, _ = avici_b200.simulate_data(d=15, n=40, seed=0, domain=domain)
    assert np.array_equal(x2, x3)


def test_simulate_data_graph_families_and_yaml_path(tmp_path):
    """Every graph prior of the example domains gives a DAG with the requested size; a YAML domain file in the
    reference's DSL (`path=`, example.py:44-47) with option lists and a custom class from `module_paths` loads."""
    from avici_b200 import synthetic as S
    rng = np.random.default_rng(0)
    for gm in (S.ErdosRenyi(2), S.ScaleFree(2, 1.5), S.ScaleFreeTranspose(2, 0.5), S.WattsStrogatz(2, 1, 0.3),
               S.WattsStrogatz(3, 1, 0.3), S.SBM(2, 5, 0.1), S.GRG(0.1)):
        for d in (2, 7, 30, 100):
            g = gm(rng, d)
            assert g.shape == (d, d) and set(np.unique(g)) <= {0, 1}
            assert len(S._toporder(g)) == d
    specs = S.load_data_config(S.EXAMPLE_DOMAINS["lin-laplace-cauchy"], force_kwargs=dict(n_observations_obs=5, n_observations_int=0))
    assert len(specs) == 11 * 2 and all(sp["n_observations_obs"] == 5 for sp in specs)  # 11 graph options x {Cauchy, Laplace}
    (tmp_path / "mymod.py").write_text(
        "import numpy as np\n"
        "class Chain:\n"
        "    def __init__(self, flip=False):\n        self.flip = flip\n"
        "    def __call__(self, rng, n_vars):\n"
        "        g = np.eye(n_vars, k=1, dtype=int)\n        return g.T.copy() if self.flip else g\n")
    (tmp_path / "dom.yaml").write_text(
        "---\ndata:\n  - n_observations_obs: null\n    n_observations_int: null\n"
        "    graph:\n      - __class__: Chain\n        flip: [ false ]\n"
        "    mechanism:\n      - __class__: LinearAdditive\n"
        "        param:\n          - __class__: SignedUniform\n            low: 1.0\n            high: 2.0\n"
        "        bias:\n          - __class__: Uniform\n            low: -1.0\n            high: 1.0\n"
        "        noise:\n          - __class__: Gaussian\n          - __class__: Laplace\n"
        "        noise_scale:\n          - __class__: Uniform\n            low: 0.5\n            high: 1.0\n"
        "        n_interv_vars: -1\n        interv_dist:\n          - __class__: SignedUniform\n            low: 1.0\n            high: 5.0\n")
    g, x, iv = avici_b200.simulate_data(d=6, n=20, n_interv=12, seed=3, path=str(tmp_path / "dom.yaml"),
                                        module_paths=str(tmp_path / "mymod.py"))
    assert np.array_equal(g, np.eye(6, k=1, dtype=int)) and x.shape == (32, 6) and iv.sum() == 12
    with pytest.raises(SyntaxError):
        avici_b200.simulate_data(d=6, n=20, seed=3, path=str(tmp_path / "dom.yaml"))  # Chain is not registered
    with pytest.raises(KeyError):
        avici_b200.simulate_data(d=6, n=20, seed=3, domain="no-such-domain")
    with pytest.raises(KeyError):
        avici_b200.simulate_data(d=6, n=20, seed=3)


def test_wrapper_validation_matches_reference_asserts():
    model = avici_b200.random_init_model(dict(layers=1))
    x = np.zeros((10, 4), dtype=np.float32)
    with pytest.raises(AssertionError, match="2D array"):
        model(x=np.zeros((2, 10, 4), dtype=np.float32))
    with pytest.raises(AssertionError, match="same shape"):
        model(x=x, interv=np.zeros((10, 5)))
    with pytest.raises(AssertionError, match="binary"):
        model(x=x, interv=np.full((10, 4), 0.5))


def test_batch_slice_partitions():
    for B, G in ((64, 8), (10, 4), (3, 8), (512, 8)):
        idx = np.concatenate([np.arange(B)[sharded.batch_slice(B, r, G)] for r in range(G)])
        assert np.array_equal(idx, np.arange(B))
        sizes = [len(range(B)[sharded.batch_slice(B, r, G)]) for r in range(G)]
        assert max(sizes) - min(sizes) <= 1


def _reshard_worker(rank, world, port, n, d, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = torch.arange(n * d * k, dtype=torch.float32).reshape(n, d, k)
        nl, dl = n // world, d // world
        z_n = z[rank * nl:(rank + 1) * nl].clone()
        z_d = sharded.reshard_n_to_d(z_n)
        ok1 = torch.equal(z_d, z[:, rank * dl:(rank + 1) * dl])
        back = sharded.reshard_d_to_n(z_d)
        ok2 = torch.equal(back, z_n)
        # per-token work is layout agnostic: a token-wise function commutes with the re-shard
        f = lambda t: t * 2 + 1
        ok3 = torch.equal(sharded.reshard_d_to_n(f(z_d)), f(z_n))
        ret[rank] = bool(ok1 and ok2 and ok3 and sharded.world_size() == world and sharded.rank() == rank)
    finally:
        dist.destroy_process_group()


def test_axial_reshard_world2_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ret = mp.Manager().dict()
    mp.spawn(_reshard_worker, args=(world, port, 8, 6, 4, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_operand_layout_index_maps_are_bijections():
    """RowMap and the two operand layouts: every (row, head, channel) lands on its own element,
    inside the buffer the workspace reserves, for both attention axes and ragged shapes."""
    from avici_b200 import layouts as L
    for B, n, d in ((1, 5, 3), (2, 130, 7), (1, 50, 100), (3, 1, 4)):
        N = B * n * d
        for axis in (0, 1):
            m = np.arange(N)
            t = L.seq_major_to_token(m, n, d, axis)
            assert sorted(t.tolist()) == list(range(N))
            assert np.array_equal(L.token_to_seq_major(t, n, d, axis), m)
            # consecutive rows of one sequence are the tokens the block attends over
            T = d if axis == 0 else n
            step = 1 if axis == 0 else d
            first = t.reshape(-1, T)
            assert np.all(np.diff(first, axis=1) == step)
            mm, w, h, c = np.meshgrid(m, np.arange(3), np.arange(8), np.arange(32), indexing="ij")
            off = L.qkv_operand_offset(mm, w, h, c, n, d, axis).ravel()
            assert off.min() == 0 and off.max() == N * 768 - 1 and len(np.unique(off)) == off.size
            # a (sequence, q|k|v, head) block is contiguous: [T positions x 32 channels]
            blk = L.qkv_operand_offset(np.arange(T), 1, 3, 0, n, d, axis) // 32
            assert np.array_equal(blk, blk[0] + np.arange(T))
            mm, h, c = np.meshgrid(m, np.arange(8), np.arange(32), indexing="ij")
            off = L.attn_operand_offset(mm, h, c).ravel()
            tiles = (N + 127) // 128
            assert off.min() == 0 and off.max() < tiles * 128 * 256 and len(np.unique(off)) == off.size


def test_device_metrics_fail_loudly_without_a_gpu():
    """avici_b200.metrics has no CPU fallback: without a CUDA device every entry point raises (it never routes
    through the numpy oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine WITHOUT a GPU")
    from avici_b200 import metrics as GM
    g = np.zeros((4, 4), np.int32)
    for call in (lambda: GM.graph_metrics(g.astype(np.float32), g), lambda: GM.shd(g, g), lambda: GM.is_cyclic(g),
                 lambda: GM.threshold_metrics(g, g.astype(np.float32))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    import inspect
    assert "oracle" not in "".join(l for l in inspect.getsource(GM).splitlines() if "import" in l)


def test_eval_batch_composition_matches_the_reference_loop(monkeypatch):
    """avici_b200.evaluate.eval_batch (train.py:51-69,125-151): scalar names and batch means, checked on the CPU with
    the device entry points replaced by the numpy oracle (the real ones are exercised in tests/test_zz_gpu_eval.py)."""
    from avici_b200 import evaluate, metrics as GM
    from oracle import metrics_oracle as MO
    rng = np.random.default_rng(3)
    B, n, d = 5, 11, 7
    probs = rng.random((B, d, d)).astype(np.float32)
    g = (rng.random((B, d, d)) < 0.2).astype(np.int32)

    class FakeModel:
        def infer_batch(self, x, interv=None, return_probs=True, experimental_chunk_size=None):
            assert x.shape == (B, n, d) and return_probs
            return probs

    def fake_graph_metrics(pred, true, threshold=0.5):
        m = MO.graph_metrics(pred, true, threshold)
        return dict(m, tp=None)

    def fake_threshold_batch(true, pred):
        rows = [MO.threshold_metrics(true[b], pred[b]) for b in range(len(pred))]
        return {k: np.array([r[k] for r in rows]) for k in ("auroc", "auprc", "ap")}

    monkeypatch.setattr(GM, "graph_metrics", fake_graph_metrics)
    monkeypatch.setattr(GM, "threshold_metrics_batch", fake_threshold_batch)
    scalars, out = evaluate.eval_batch(FakeModel(), np.zeros((B, n, d), np.float32), g, decision_thresholds=(0.5, 0.25))
    assert out is probs
    # the reference's accumulation: scalars[name] += value / batch_size for every dataset j (train.py:125-151)
    want = {}
    for thres in (0.5, 0.25):
        key = f"{thres}".replace(".", "_")
        pred = (probs > thres).astype(np.int32)
        for j in range(B):
            want[f"shd_{key}"] = want.get(f"shd_{key}", 0.0) + MO.shd(g[j], pred[j]) / B
            want[f"edges_{key}"] = want.get(f"edges_{key}", 0.0) + MO.n_edges(pred[j]) / B
            want[f"cyclic_{key}"] = want.get(f"cyclic_{key}", 0.0) + float(MO.is_cyclic(pred[j])) / B
            for k, v in MO.classification_metrics(g[j], pred[j]).items():
                want[f"{k}_{key}"] = want.get(f"{k}_{key}", 0.0) + v / B
    for j in range(B):
        for k, v in MO.threshold_metrics(g[j], probs[j]).items():
            want[k] = want.get(k, 0.0) + v / B
    assert set(scalars) == set(want) and "f1_0_25" in scalars and "cyclic_0_5" in scalars
    for k in want:
        assert scalars[k] == pytest.approx(want[k], abs=1e-12), k


def test_shard_count_policy():
    """How many devices one dataset is sharded over (AVICIModel.__call__ with several devices): the largest G <= device_count
    with G | n and G | d, one device for small or chunked datasets.  The reference shards n only and asserts G | n
    (pretrain.py:129-131); the axial scheme also needs G | d."""
    from avici_b200.model import AVICIModel
    f = AVICIModel._shard_count
    assert f(5000, 1000, 8) == 8 and f(5000, 1000, 4) == 4 and f(5000, 1000, 3) == 2
    assert f(5000, 999, 8) == 1            # no common divisor with d
    assert f(4096, 96, 8) == 8 and f(4096, 100, 8) == 4
    assert f(1000, 100, 8) == 1            # below SHARD_MIN_CELLS: an exchange per block is not worth it
    assert f(5000, 1000, 8, chunk_size=500) == 1 and f(5000, 1000, 8, chunk_size=5000) == 8


def test_column_block_layout_roundtrip():
    """The column-blocked n-shard of avb_forward_sharded: token (i, j) at ((j // dg) * rows + i) * dg + j % dg, and back."""
    import torch
    from avici_b200 import sharded
    rows, d, G, k = 6, 12, 4, 3
    z = torch.arange(rows * d * k, dtype=torch.float32).reshape(rows, d, k)
    blk = sharded.to_column_blocks(z, G)
    assert blk.shape == (G, rows, d // G, k)
    dg = d // G
    flat = blk.reshape(-1, k)
    for i in (0, 3, 5):
        for j in (0, 2, 3, 7, 11):
            assert torch.equal(flat[((j // dg) * rows + i) * dg + j % dg], z[i, j])
    assert torch.equal(sharded.from_column_blocks(blk), z)
