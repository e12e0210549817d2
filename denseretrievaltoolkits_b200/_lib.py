"""ctypes binding of the C ABI declared in include/drt_b200.h.

The shared library is built in-tree (`denseretrievaltoolkits_b200/libdrt_b200.so`, see
`build_library`) so it travels with the repository snapshot.  There is no CPU fallback: if the
library is missing `load()` raises, and every compute entry point returns an error status on a
machine without an sm_100 device, which `check()` turns into `RuntimeError` (the same exception
type faiss' SWIG layer raises, DRT/evaluator/index.py callers see no new exception classes).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
LIB_PATH = os.path.join(_PKG_DIR, "libdrt_b200.so")
_SRC = os.path.join(_PKG_DIR, "csrc", "drt_b200.cu")

# every symbol include/drt_b200.h declares (tests/test_abi.py checks the header against this)
SYMBOLS = [
    "drt_abi_version", "drt_last_error", "drt_device_count",
    "drt_store_create", "drt_store_destroy", "drt_store_add", "drt_store_ntotal", "drt_store_dim",
    "drt_store_device", "drt_store_reset", "drt_store_reconstruct", "drt_store_set_exact_tail",
    "drt_search", "drt_search_async", "drt_search_stats", "drt_plan_params", "drt_plan_chunks", "drt_merge_topk",
    "drt_merge_topk_peers", "drt_merge_topk_peers2",
    "drt_inbatch_ce_fwd", "drt_inbatch_ce_bwd", "drt_inbatch_ce_bwd_needs_work", "drt_filter_negatives",
]

SEARCH_DEFAULT = 0
SEARCH_NO_RESCORE = 1
SEARCH_FORCE_1CTA = 2
SEARCH_FORCE_2CTA = 4
SEARCH_TIME_KERNELS = 8
MERGE_DEFAULT = 0
MERGE_SORTED_UNIQUE = 1
MAX_K = 2048

_lib = None


def nvcc_command(out: str = LIB_PATH) -> list[str]:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "--expt-extended-lambda", "-shared", "-Xcompiler", "-fPIC", "-o", out, _SRC,
    ]


def _sources_mtime() -> float:
    srcs = [os.path.join(_PKG_DIR, "csrc", f) for f in os.listdir(os.path.join(_PKG_DIR, "csrc"))]
    srcs.append(os.path.join(_REPO_DIR, "include", "drt_b200.h"))
    return max(os.path.getmtime(s) for s in srcs)


def build_library(force: bool = False) -> str:
    """Compile csrc/ for sm_100a into the in-tree shared library (cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _sources_mtime():
        return LIB_PATH
    tmp = LIB_PATH + ".tmp"
    subprocess.check_call(nvcc_command(tmp), cwd=_PKG_DIR)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). denseretrievaltoolkits_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    i64p = POINTER(c_int64)
    lib.drt_abi_version.restype = c_int
    lib.drt_last_error.restype = c_char_p
    lib.drt_device_count.restype = c_int
    lib.drt_store_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int64]
    lib.drt_store_destroy.argtypes = [c_void_p]
    lib.drt_store_add.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p]
    lib.drt_store_ntotal.argtypes = [c_void_p]
    lib.drt_store_ntotal.restype = c_int64
    lib.drt_store_dim.argtypes = [c_void_p]
    lib.drt_store_device.argtypes = [c_void_p]
    lib.drt_store_reset.argtypes = [c_void_p]
    lib.drt_store_set_exact_tail.argtypes = [c_void_p, c_int]
    lib.drt_store_reconstruct.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]
    lib.drt_search.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int,
                               c_int64, c_uint32, c_void_p]
    lib.drt_search_async.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_uint32, c_void_p, c_void_p]
    lib.drt_search_stats.argtypes = [c_void_p, i64p]
    lib.drt_plan_params.argtypes = [c_int, c_int, POINTER(c_int), POINTER(c_int)]
    lib.drt_plan_chunks.argtypes = [c_int64, c_int64, c_int, c_int, i64p, c_int]
    lib.drt_merge_topk.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                   c_void_p, c_uint32, c_int, c_void_p]
    lib.drt_merge_topk_peers.argtypes = [c_int, POINTER(c_void_p), POINTER(c_void_p), c_int64, c_int64, c_int, c_int,
                                         POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int, c_void_p]
    lib.drt_merge_topk_peers2.argtypes = [c_int, POINTER(c_void_p), POINTER(c_void_p), c_int64, c_int64, c_int, c_int,
                                          POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_void_p,
                                          c_int, c_void_p]
    lib.drt_inbatch_ce_fwd.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                       c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p]
    lib.drt_inbatch_ce_bwd.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                       c_void_p, c_void_p, c_int, c_void_p]
    lib.drt_inbatch_ce_bwd_needs_work.argtypes = [c_int64, c_int64, c_int, c_int]
    lib.drt_filter_negatives.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p, c_int, c_void_p]
    for name in SYMBOLS:
        getattr(lib, name)   # AttributeError here = the library does not match the header
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().drt_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


_raw_stream = None


def current_stream_ptr(device_index: int | None = None) -> int:
    """The caller's current torch CUDA stream as a raw cudaStream_t (0 when torch has no CUDA)."""
    global _raw_stream
    if _raw_stream is None:
        import torch

        if not torch.cuda.is_available():
            return 0
        fast = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if fast is not None:
            _raw_stream = lambda idx: int(fast(torch.cuda.current_device() if idx is None else idx))
        else:
            _raw_stream = lambda idx: int(torch.cuda.current_stream(idx).cuda_stream)
    return _raw_stream(device_index)
