"""denseretrievaltoolkits_b200 — B200-native (sm_100a) exact maximum-inner-product top-k search
and in-batch-negative loss for DenseRetrievalToolkits, behind the reference's own call
signatures.  CUDA kernels + C ABI live in csrc/ (see include/drt_b200.h); this package is the
thin Python host side.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_faiss() -> None:
    """Register `faiss_compat` as `sys.modules['faiss']` so the reference's
    `DRT/evaluator/index.py`, `retrieval.py` and `DRT/trainer/trainer.py` (all `import faiss`)
    run unmodified on the B200 store."""
    from . import faiss_compat

    sys.modules["faiss"] = faiss_compat


def build() -> str:
    from . import _lib

    return _lib.build_library()
