"""Host-side mirror of the reference's inference interface, backed by the C-ABI CUDA library.

    reference                                        here
    -----------------------------------------------  -------------------------------------------
    avici.model.BaseModel          model.py:24-144   BaseModel        (shape holder; the network
                                                                       itself is csrc/*.cu)
    avici.model.InferenceModel     model.py:148-240  InferenceModel   (.infer_edge_probs, ...)
    avici.pretrain.AVICIModel      pretrain.py:23-156 AVICIModel      (__call__ drop-in + infer_batch)

PyTorch is only the carrier of device memory and streams; all arithmetic runs in the hand-written
sm_100a kernels.  No CPU fallback exists: calls fail if the extension or the GPU is missing.
"""
import ctypes as C
import warnings

import numpy as onp

from . import _lib
from .params import flatten_params, resolve_model_kwargs

_DTYPES = {"bf16": _lib.AVB_DTYPE_BF16, "bfloat16": _lib.AVB_DTYPE_BF16,
           "fp32": _lib.AVB_DTYPE_F32, "f32": _lib.AVB_DTYPE_F32, "float32": _lib.AVB_DTYPE_F32}


def _torch():
    import torch
    return torch


class BaseModel:
    """Carrier of the `BaseModel.__init__` kwargs (avici/model.py:26-50).  The forward pass lives in
    the CUDA library; this class exists so that `InferenceModel(model_class=BaseModel, ...)` reads
    like the reference."""

    def __init__(self, **kwargs):
        self.kwargs = resolve_model_kwargs(kwargs)


class Engine:
    """One C-ABI handle (`avb_handle`): a network shape + uploaded weights on the current GPU."""

    def __init__(self, model_kwargs, mask_diag=True, compute_dtype="bf16"):
        self.lib = _lib.load()
        self.kw = resolve_model_kwargs(model_kwargs)
        if compute_dtype not in _DTYPES:
            raise ValueError(f"compute_dtype must be one of {sorted(_DTYPES)}")
        self.compute_dtype = "bf16" if _DTYPES[compute_dtype] == _lib.AVB_DTYPE_BF16 else "fp32"
        cfg = _lib.AvbConfig(layers=self.kw["layers"], dim=self.kw["dim"], key_size=self.kw["key_size"],
                             num_heads=self.kw["num_heads"], widening_factor=self.kw["widening_factor"],
                             out_dim=self.kw["out_dim"], mask_diag=int(bool(mask_diag)),
                             compute_dtype=_DTYPES[compute_dtype])
        h = C.c_void_p()
        rc = self.lib.avb_create(C.byref(h), C.byref(cfg))
        _lib.check(self.lib, None, rc)
        self.handle = h
        self._ws = None
        self.device = _torch().device("cuda", _torch().cuda.current_device())  # avb_create binds the current device

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.avb_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _check(self, rc):
        _lib.check(self.lib, self.handle, rc)

    def load_params(self, params):
        flat = flatten_params(params)
        descs = (_lib.AvbTensorDesc * len(flat))()
        keep = []
        for i, (name, arr) in enumerate(flat):
            keep.append(arr)
            descs[i].name = name.encode()
            descs[i].data = arr.ctypes.data
            descs[i].numel = arr.size
        self._check(self.lib.avb_load_weights(self.handle, descs, len(flat)))

    @property
    def param_count(self):
        return int(self.lib.avb_param_count(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.avb_launch_count(self.handle))

    def profile(self, on=True):
        self._check(self.lib.avb_profile_enable(self.handle, int(bool(on))))

    def profile_read(self, reset=True):
        """{category: (total_ms, scopes)} accumulated since the last reset (synchronises the device)."""
        n = len(_lib.PROF_CATEGORIES)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        self._check(self.lib.avb_profile_read(self.handle, ms, cnt, int(bool(reset))))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROF_CATEGORIES)}

    def workspace(self, nbytes):
        torch = _torch()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, x, interv=None, standardize=_lib.AVB_STANDARDIZE_DEFAULT, chunk_size=None, out=None,
                log_probs=False):
        """x, interv: float32 CUDA tensors [B,n,d] on this engine's device -> probs (or log-probs) float32 CUDA
        tensor [B,d,d] (async on the current torch stream)."""
        torch = _torch()
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3, "x must be a float32 CUDA tensor [B,n,d]"
        assert x.device == self.device, f"this engine lives on {self.device}, got a tensor on {x.device}"
        x = x.contiguous()
        B, n, d = x.shape
        if interv is not None:
            assert interv.shape == x.shape and interv.device == x.device
            interv = interv.to(dtype=torch.float32).contiguous()
        cs = int(chunk_size) if chunk_size else 0
        need = self.lib.avb_workspace_bytes(self.handle, B, n, d, cs)
        if need == 0:
            raise AssertionError("observations `n` must be divisible by `experimental_chunk_size` ")
        ws = self.workspace(need)
        if out is None:
            out = torch.empty((B, d, d), dtype=torch.float32, device=x.device)
        assert out.device == self.device
        stream = torch.cuda.current_stream(x.device).cuda_stream
        fn = self.lib.avb_forward_logprobs if log_probs else self.lib.avb_forward
        self._check(fn(self.handle, x.data_ptr(), interv.data_ptr() if interv is not None else None,
                       B, n, d, int(standardize), cs, out.data_ptr(), ws.data_ptr(), ws.numel(), C.c_void_p(stream)))
        return out

    def capture(self, x, interv=None, standardize=_lib.AVB_STANDARDIZE_DEFAULT, chunk_size=None):
        """CUDA graph of one forward on the STATIC buffers `x` / `interv` (refill them in place, then `graph.replay()`):
        returns (graph, out).  avb_forward only enqueues work on the caller's stream - no allocation, no host
        synchronisation - so the whole forward (~110 launches at 16 blocks) captures as is; on small datasets the replay
        saves the launch gaps (n=200, d=50: 0.91 -> 0.79 ms).  The engine's workspace must not be resized by a larger call
        while the graph is alive."""
        torch = _torch()
        out = torch.empty((x.shape[0], x.shape[2], x.shape[2]), dtype=torch.float32, device=x.device)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):  # warm-up outside the capture: sizes the workspace, sets the kernel attributes
            self.forward(x, interv, standardize=standardize, chunk_size=chunk_size, out=out)
        torch.cuda.current_stream(x.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.forward(x, interv, standardize=standardize, chunk_size=chunk_size, out=out)
        return graph, out

    def forward_host(self, x, interv=None, standardize=_lib.AVB_STANDARDIZE_DEFAULT, chunk_size=None, out=None):
        """numpy float32 [B,n,d] in host memory -> numpy [B,d,d]; host<->device copies inside the call."""
        x = onp.ascontiguousarray(x, dtype=onp.float32)
        B, n, d = x.shape
        if interv is not None:
            interv = onp.ascontiguousarray(interv, dtype=onp.float32)
            assert interv.shape == x.shape
        if out is None:
            out = onp.empty((B, d, d), dtype=onp.float32)
        self._check(self.lib.avb_forward_host(self.handle, x.ctypes.data,
                                              interv.ctypes.data if interv is not None else None, B, n, d,
                                              int(standardize), int(chunk_size) if chunk_size else 0,
                                              out.ctypes.data))
        return out


class InferenceModel:
    """Mirror of `avici.model.InferenceModel` (model.py:148-240), inference methods only.

    `params` is passed per call like in the reference; the packed device copy is cached per params
    object.  `compute_dtype` ("bf16" tensor-core path, "fp32" parity path) is the one extra knob.
    """

    def __init__(self, *, model_class=BaseModel, model_kwargs=None, train_p_obs_only=0.0, acyclicity=None,
                 acyclicity_pow_iters=10, mask_diag=True, compute_dtype="bf16"):
        del train_p_obs_only, acyclicity, acyclicity_pow_iters  # training-only knobs (model.py:296-359)
        self.mask_diag = mask_diag
        self.model_class = model_class
        self.model_kwargs = resolve_model_kwargs(dict(model_kwargs or {}))
        self.compute_dtype = compute_dtype
        self._engines = {}  # device index -> (Engine, fingerprint of the params it holds, the params object)

    @staticmethod
    def _fingerprint(params):
        """Cheap identity of a parameter dict: the object, plus per leaf its object identity, shape and three values.
        The packed device copy is a SNAPSHOT taken at first use; replacing the dict, a leaf, or (for all practical
        purposes) editing a leaf in place changes the fingerprint and triggers a re-upload.  `reload()` forces one."""
        fp = [id(params)]
        add = fp.append
        for mod, leaves in params.items():
            for name, arr in leaves.items():
                a = arr if isinstance(arr, onp.ndarray) else onp.asarray(arr)
                n = a.size
                # (0.2 ms for the 334 leaves of the published architecture; it runs on every call)
                add((mod, name, id(arr), a.shape, a.item(0), a.item(n >> 1), a.item(n - 1)) if n else (mod, name))
        return hash(tuple(fp))

    def engine(self, params, device=None):
        """The C-ABI handle holding `params` on `device` (default: the current CUDA device); one per device."""
        torch = _torch()
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        idx = torch.cuda.current_device() if idx is None else idx
        fp = self._fingerprint(params)
        ent = self._engines.get(idx)
        if ent is None or ent[1] != fp:
            with torch.cuda.device(idx):
                eng = ent[0] if ent is not None else Engine(self.model_kwargs, mask_diag=self.mask_diag,
                                                            compute_dtype=self.compute_dtype)
                eng.load_params(params)
            self._engines[idx] = (eng, fp, params)  # the params reference keeps id(params) unique
        return self._engines[idx][0]

    def reload(self):
        """Drop every cached device copy of the weights (next call re-uploads)."""
        self._engines = {}

    def _infer(self, params, x, experimental_chunk_size, log_probs):
        torch = _torch()
        is_np = isinstance(x, onp.ndarray)
        xt = torch.as_tensor(x, dtype=torch.float32)
        xt = xt.cuda() if not xt.is_cuda else xt
        lead = xt.shape[:-3]
        n, d = xt.shape[-3], xt.shape[-2]
        xt = xt.reshape(-1, n, d, 2)
        eng = self.engine(params, xt.device)
        with torch.cuda.device(xt.device):
            p = eng.forward(xt[..., 0].contiguous(), xt[..., 1].contiguous(), standardize=_lib.AVB_STANDARDIZE_NONE,
                            chunk_size=experimental_chunk_size, log_probs=log_probs)
        p = p.reshape(*lead, d, d)
        return p.cpu().numpy() if is_np else p

    def infer_edge_probs(self, params, x, experimental_chunk_size=None):
        """x: [..., n, d, 2] (already standardized, channel 1 = intervention mask) -> [..., d, d]
        (model.py:225-240)."""
        return self._infer(params, x, experimental_chunk_size, False)

    def infer_edge_logprobs(self, params, rng, x, is_training=False, experimental_chunk_size=None):
        """`log_sigmoid(logits)` computed on the device from the logits themselves, diagonal = -inf when mask_diag
        (model.py:203-222); not log(p), which loses precision once p underflows."""
        assert not is_training, "the B200 path implements inference only"
        return self._infer(params, x, experimental_chunk_size, True)

    def sample_graphs(self, n_samples, params, rng, x, experimental_chunk_size=None):
        """Bernoulli(p) graph samples [..., n_samples, d, d] with zero diagonal (model.py:175-200).
        `rng`: int seed or torch.Generator (the reference's JAX threefry stream is not reproducible)."""
        torch = _torch()
        p = self.infer_edge_probs(params, torch.as_tensor(x), experimental_chunk_size=experimental_chunk_size)
        gen = rng if isinstance(rng, torch.Generator) else torch.Generator(device=p.device).manual_seed(int(rng))
        pe = p.unsqueeze(-3).expand(*p.shape[:-2], n_samples, *p.shape[-2:])
        return torch.bernoulli(pe, generator=gen).to(torch.int32)  # diag(p) = 0 -> never sampled


class AVICIModel:
    """Drop-in for `avici.pretrain.AVICIModel` (pretrain.py:23-156): same constructor keywords, same
    `__call__` signature and semantics (assertions, in-place rounding of `interv`, thresholding,
    numpy in -> numpy out)."""

    def __init__(self, *, params, model, expects_counts, standardizer=None):
        self.params = params
        self.expects_counts = expects_counts
        self._model = model
        if standardizer is None:
            standardizer = _lib.AVB_STANDARDIZE_COUNT if expects_counts else _lib.AVB_STANDARDIZE_DEFAULT
        self._standardizer = standardizer

    @property
    def engine(self):
        return self._model.engine(self.params)

    def engine_on(self, device):
        return self._model.engine(self.params, device)

    def __call_main__(self, x, interv=None, experimental_chunk_size=None):
        """pretrain.py:53-67 on device tensors [n,d] (or a batch [B,n,d])."""
        single = x.dim() == 2
        xb = x.unsqueeze(0) if single else x
        ib = None if interv is None else (interv.unsqueeze(0) if single else interv)
        with _torch().cuda.device(x.device):
            out = self.engine_on(x.device).forward(xb, ib, standardize=self._standardizer,
                                                   chunk_size=experimental_chunk_size)
        return out[0] if single else out

    def __call__(self, x, interv=None, return_probs=True, devices=None, shard_if_possible=True,
                 experimental_chunk_size=None):
        """
        Args:
            x: `[n, d]` real-valued data matrix (numpy array or torch tensor)
            interv (optional): binary matrix of the same shape, `interv[i,j] == 1` iff node `j` was
                intervened upon in observation `i`
            return_probs (optional): `False` thresholds the prediction at 0.5 (pretrain.py:149-150)
            devices (optional): "gpu" / "cuda", a CUDA device, or a list of them.  `None`: every visible GPU of this
                process, like the reference's `jax.local_devices()` (under an initialised `torch.distributed`
                group: this rank's GPU, with the ranks of the group playing the role of the devices).
            shard_if_possible (optional): with several devices, shard ONE dataset axially over them
                (`avici_b200.sharded`: n-shards for the blocks that attend over d, d-shards for those that attend
                over n, an all-to-all in between; results are bit-identical to one GPU).  Like the reference it
                requires `n` divisible by the device count; the number of devices actually used is the largest
                G <= device_count that also divides `d`, and 1 for small or chunked datasets.  `False`: first device.
            experimental_chunk_size (optional): as in the reference (model.py:99-116).
        Returns:
            `[d, d]` matrix of predicted edge probabilities (or 0/1 ints)
        """
        torch = _torch()
        x_type = type(x)
        assert x.ndim == 2, "`x` must be a 2D array of shape [n, d]."
        if interv is not None:
            assert x.shape == interv.shape, "`x` and `interv` must have the same shape when provided."
            assert interv.ndim == 2, "`interv` must be a 2D array of shape [n, d] when provided."
            if isinstance(interv, onp.ndarray):
                interv_is_binary = onp.all(onp.isclose(interv, 0) | onp.isclose(interv, 1))
                assert interv_is_binary, "Intervention mask `interv` has to be binary."
                interv = onp.around(interv, 0, out=interv)  # in place, like the reference (pretrain.py:115)
            else:
                interv_is_binary = bool((torch.isclose(interv, torch.zeros_like(interv), atol=1e-8) |
                                         torch.isclose(interv, torch.ones_like(interv), atol=1e-8)).all())
                assert interv_is_binary, "Intervention mask `interv` has to be binary."
                interv.round_()
        if self.expects_counts:
            xm = x if isinstance(x, onp.ndarray) else x.detach().cpu().numpy()
            if not onp.allclose(onp.mod(xm, 1), 0):
                warnings.warn("Model expects count data but `x` contains non-integer values. ")

        from . import sharded
        devs = self._resolve_devices(devices)
        dev = devs[0]
        world = sharded.world_size()
        n_obs, n_vars = x.shape
        with torch.cuda.device(dev):
            if shard_if_possible and world > 1 and devices is None:
                # one process per GPU (torchrun): shard over the ranks of the default group
                assert not n_obs % world, ("observations `n` must be divisible by `device_count` if sharding "
                                           "is enabled. ")
                assert not n_vars % world, ("variables `d` must be divisible by `device_count` for the axial "
                                            "(n-shard <-> d-shard) sharding of this implementation. ")
                assert not experimental_chunk_size or experimental_chunk_size >= n_obs, (
                    "`experimental_chunk_size` is not supported together with axial sharding; pass "
                    "`shard_if_possible=False` to run the chunked forward on one device. ")
                xt = torch.as_tensor(x).to(device=dev, dtype=torch.float32)
                it = None if interv is None else torch.as_tensor(interv).to(device=dev, dtype=torch.float32)
                out = sharded.forward_axial(self, xt, it)
            elif shard_if_possible and len(devs) > 1:
                # one process, several GPUs - the reference's own mode (pretrain.py:124-143)
                assert not n_obs % len(devs), ("observations `n` must be divisible by `device_count` if sharding "
                                               "is enabled. ")
                G = self._shard_count(n_obs, n_vars, len(devs), experimental_chunk_size)
                xt = torch.as_tensor(x).to(dtype=torch.float32)
                it = None if interv is None else torch.as_tensor(interv).to(dtype=torch.float32)
                if G > 1:
                    out = sharded.forward_axial_devices(self, xt, it, devs[:G])
                else:
                    out = self.__call_main__(xt.to(dev), None if it is None else it.to(dev),
                                             experimental_chunk_size=experimental_chunk_size)
            else:
                xt = torch.as_tensor(x).to(device=dev, dtype=torch.float32)
                it = None if interv is None else torch.as_tensor(interv).to(device=dev, dtype=torch.float32)
                out = self.__call_main__(xt, it, experimental_chunk_size=experimental_chunk_size)
            if not return_probs:
                out = (out > 0.5).to(torch.int64)
            if x_type == onp.ndarray:
                out = out.cpu().numpy()
                if not return_probs:
                    out = out.astype(int)
        return out

    # datasets smaller than this many cells are not worth an exchange per block: they run on the first device
    SHARD_MIN_CELLS = 1 << 17

    @staticmethod
    def _resolve_devices(devices):
        """`devices` of the reference's `__call__` -> list of CUDA devices.  None / "gpu" / "cuda": every visible GPU in a
        plain process, the current device under an initialised `torch.distributed` group (one process per GPU)."""
        torch = _torch()
        from . import sharded
        if devices is None or (isinstance(devices, str) and devices in ("gpu", "cuda")):
            if sharded.world_size() > 1:
                return [torch.device("cuda", torch.cuda.current_device())]
            cur = torch.cuda.current_device()
            idx = [cur] + [i for i in range(torch.cuda.device_count()) if i != cur]
            return [torch.device("cuda", i) for i in idx]
        if not isinstance(devices, (list, tuple)):
            devices = [devices]
        out = []
        for dv in devices:
            dv = torch.device("cuda", dv) if isinstance(dv, int) else torch.device(dv)
            assert dv.type == "cuda", "avici_b200 runs on CUDA (sm_100a) devices only"
            out.append(torch.device("cuda", torch.cuda.current_device() if dv.index is None else dv.index))
        assert out, "`devices` must not be empty"
        return out

    @classmethod
    def _shard_count(cls, n, d, device_count, chunk_size=None):
        """Number of devices one dataset is sharded over: the largest G <= device_count with G | n and G | d (the
        axial scheme re-shards between the n and the d axis, so both must divide; the reference shards n only).
        1 for small datasets and for the chunked forward."""
        if (chunk_size and 0 < chunk_size < n) or n * d < cls.SHARD_MIN_CELLS:
            return 1
        for G in range(device_count, 0, -1):
            if n % G == 0 and d % G == 0:
                return G
        return 1

    def infer_batch(self, x, interv=None, return_probs=True, experimental_chunk_size=None, devices=None):
        """Batched forward `[B,n,d] -> [B,d,d]` (the reference's batched call site is train.py:61).
        `devices`: a list of CUDA devices spreads the independent datasets over them (one host thread and one
        handle per device, no collective); default: the current device."""
        torch = _torch()
        is_np = isinstance(x, onp.ndarray)
        assert x.ndim == 3, "`x` must be a 3D array of shape [B, n, d]."
        if interv is not None:
            assert x.shape == interv.shape, "`x` and `interv` must have the same shape when provided."
        if devices is not None and len(self._resolve_devices(devices)) > 1:
            from . import sharded
            xt = torch.as_tensor(x)
            out = sharded.infer_batch_devices(self, xt, None if interv is None else torch.as_tensor(interv),
                                              self._resolve_devices(devices), chunk_size=experimental_chunk_size)
            out = out if return_probs else (out > 0.5).to(torch.int64)
            return out.cpu().numpy() if is_np else out
        if is_np:
            out = self.engine.forward_host(x, interv, standardize=self._standardizer,
                                           chunk_size=experimental_chunk_size)
            return out if return_probs else (out > 0.5).astype(int)
        xt = x.to(dtype=torch.float32)
        out = self.__call_main__(xt, None if interv is None else interv.to(dtype=torch.float32),
                                 experimental_chunk_size=experimental_chunk_size)
        return out if return_probs else (out > 0.5).to(torch.int64)
