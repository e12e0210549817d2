"""ctypes loader for the plain-C oracle (oracle/flat_ip_c.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_flatip.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flat_ip_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def search(corpus: np.ndarray, q: np.ndarray, k: int):
    lib = ctypes.CDLL(build())
    corpus = np.ascontiguousarray(corpus, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    nq, d = q.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib.oracle_flat_ip_search(
        corpus.ctypes.data_as(fp), ctypes.c_int64(corpus.shape[0]), ctypes.c_int(d),
        q.ctypes.data_as(fp), ctypes.c_int64(nq), ctypes.c_int(k),
        D.ctypes.data_as(fp), I.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise RuntimeError(f"oracle_flat_ip_search failed: {rc}")
    return D, I
