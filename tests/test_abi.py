"""CPU tests of the drop-in boundary: the shared library builds, loads and exports every symbol that
include/avici_b200.h declares; without a GPU the library fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

from avici_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "avici_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 17
    raw = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/avici_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in avici_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    assert b"sm_100a" in built_lib.avb_version()


def test_invalid_config_is_rejected_not_emulated(built_lib):
    h = C.c_void_p()
    for bad in (dict(dim=64), dict(key_size=64), dict(layers=0), dict(compute_dtype=7), dict(out_dim=256)):
        kw = dict(layers=8, dim=128, key_size=32, num_heads=8, widening_factor=4, out_dim=128, mask_diag=1,
                  compute_dtype=1)
        kw.update(bad)
        cfg = _lib.AvbConfig(**kw)
        rc = built_lib.avb_create(C.byref(h), C.byref(cfg))
        assert rc == _lib.AVB_ERR_INVALID and not h.value
        assert built_lib.avb_last_error(None)
        with pytest.raises(AssertionError):
            _lib.check(built_lib, None, rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_no_cpu_fallback(built_lib):
    h = C.c_void_p()
    cfg = _lib.AvbConfig(layers=1, dim=128, key_size=32, num_heads=8, widening_factor=4, out_dim=128, mask_diag=1,
                         compute_dtype=1)
    rc = built_lib.avb_create(C.byref(h), C.byref(cfg))
    assert rc == _lib.AVB_ERR_CUDA and not h.value
    assert b"no CPU fallback" in built_lib.avb_last_error(None)
    import avici_b200
    with pytest.raises(Exception):
        avici_b200.random_init_model(dict(layers=1)).engine


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
