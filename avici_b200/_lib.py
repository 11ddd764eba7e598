"""ctypes binding of the C-ABI in include/avici_b200.h (`avici_b200/csrc/libavici_b200.so`).

There is no CPU fallback and no alternative backend: if the shared library is missing the import of
any compute entry point raises, and without a CUDA device `avb_create` returns AVB_ERR_CUDA.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVB_LIB") or os.path.join(_HERE, "csrc", "libavici_b200.so")  # AVB_LIB: A/B builds
SRC_DIR = os.path.join(_HERE, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

AVB_OK, AVB_ERR_INVALID, AVB_ERR_CUDA, AVB_ERR_MISSING, AVB_ERR_WORKSPACE = 0, 1, 2, 3, 4
AVB_DTYPE_F32, AVB_DTYPE_BF16 = 0, 1
AVB_STANDARDIZE_NONE, AVB_STANDARDIZE_DEFAULT, AVB_STANDARDIZE_COUNT = 0, 1, 2
AVB_AXIS_D, AVB_AXIS_N = 0, 1
PROF_CATEGORIES = ("prologue", "qkv", "attn_d", "attn_n", "out", "ffn", "pool_head", "dattn", "comm")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class AvbConfig(C.Structure):
    _fields_ = [(name, C.c_int32) for name in
                ("layers", "dim", "key_size", "num_heads", "widening_factor", "out_dim", "mask_diag",
                 "compute_dtype")]


class AvbTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


# symbol -> (restype, argtypes); must list every function include/avici_b200.h declares
SIGNATURES = {
    "avb_version": (C.c_char_p, []),
    "avb_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(AvbConfig)]),
    "avb_destroy": (None, [C.c_void_p]),
    "avb_last_error": (C.c_char_p, [C.c_void_p]),
    "avb_load_weights": (C.c_int, [C.c_void_p, C.POINTER(AvbTensorDesc), C.c_int32]),
    "avb_param_count": (C.c_int64, [C.c_void_p]),
    "avb_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "avb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "avb_forward_logprobs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "avb_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_void_p]),
    "avb_standardize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "avb_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "avb_block_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "avb_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "avb_block_sharded": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "avb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "avb_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "avb_comm_destroy": (C.c_int, [C.c_void_p]),
    "avb_sharded_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "avb_forward_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "avb_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                           C.c_void_p]),
    "avb_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                           C.c_void_p]),
    "avb_head_logprobs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "avb_graph_metrics_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "avb_graph_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "avb_threshold_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "avb_launch_count": (C.c_int64, [C.c_void_p]),
    "avb_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "avb_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "avb_debug_set_trace": (C.c_int, [C.c_void_p]),
    "avb_debug_set_dattn_dump": (C.c_int, [C.c_void_p]),
    "avb_test_set_dattn_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "avb_test_set_f32_tc": (C.c_int, [C.c_void_p, C.c_int32]),
    "avb_test_set_head_tc": (C.c_int, [C.c_void_p, C.c_int32]),
    "avb_test_set_dattn_pack": (C.c_int, [C.c_void_p, C.c_int32]),
    "avb_test_dattn_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "avb_test_ffn_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "avb_test_qkv_layout_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_void_p]),
    "avb_test_outproj_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p]),
    "avb_test_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p]),
}

_lib = None


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(SRC_DIR, f) for f in sorted(os.listdir(SRC_DIR)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(INCLUDE_DIR, "avici_b200.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    extra = os.environ.get("AVB_NVCC_EXTRA", "").split()  # e.g. -DAVB_ATTN_TURNS=0 for an A/B build
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-o", LIB_PATH, os.path.join(SRC_DIR, "avici_b200.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=SRC_DIR)
    return LIB_PATH


def load():
    """Load the shared library and bind every declared symbol; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"avici_b200: CUDA extension {LIB_PATH} is missing. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            f"There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AvbError(RuntimeError):
    pass


def check(lib, handle, rc):
    """Map C status codes to the reference's Python exception types."""
    if rc == AVB_OK:
        return
    msg = lib.avb_last_error(handle)
    msg = msg.decode() if msg else f"avici_b200 error {rc}"
    if rc == AVB_ERR_INVALID:
        raise AssertionError(msg)
    if rc == AVB_ERR_MISSING:
        raise KeyError(msg)
    raise AvbError(f"[{rc}] {msg}")
