"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when there is no B200."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from denseretrievaltoolkits_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "drt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drt_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    for name in _header_symbols():
        assert hasattr(built_lib, name), name
    assert built_lib.drt_abi_version() == 2


def test_build_command_targets_sm100a():
    cmd = " ".join(_lib.nvcc_command())
    assert "arch=compute_100a,code=sm_100a" in cmd and "-lineinfo" in cmd


def test_no_cpu_fallback(built_lib):
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    assert built_lib.drt_device_count() == 0
    h = ctypes.c_void_p()
    rc = built_lib.drt_store_create(ctypes.byref(h), 768, 0, 0)
    assert rc == -3 and "no CPU fallback" in _lib.last_error()
    from denseretrievaltoolkits_b200 import faiss_compat

    with pytest.raises(RuntimeError):
        faiss_compat.IndexFlatIP(768)
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

    with pytest.raises(RuntimeError):
        SimpleContrastiveLoss()(torch.zeros(2, 64), torch.zeros(4, 64))


def test_argument_validation_without_device(built_lib):
    h = ctypes.c_void_p()
    assert built_lib.drt_store_create(ctypes.byref(h), 100000, 0, 0) == -5   # dim > 8192
    assert built_lib.drt_store_create(ctypes.byref(h), -1, 0, 0) == -1
    assert built_lib.drt_store_create(ctypes.byref(h), 768, 0, 100) == -1    # seg_rows % 256
    assert built_lib.drt_merge_topk(0, None, None, 1, 1, 1, None, None, 0, 0, None) == -1
    assert built_lib.drt_search(None, None, 1, 1, None, None, 0, 0, 0, None) == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "denseretrievaltoolkits_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f
