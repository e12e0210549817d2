"""NVTX ranges around the hot-path phases (ingest / search / exchange+merge / loss), for Nsight
timelines.  Off unless DRT_B200_NVTX=1: a range push/pop costs ~1 us of host time per call, which
matters at the loss's call rate."""
from __future__ import annotations

import contextlib
import os

_ON = os.environ.get("DRT_B200_NVTX", "0") == "1"


@contextlib.contextmanager
def _range(name: str):
    import torch

    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


_NULL = contextlib.nullcontext()


def rng(name: str):
    """`with rng("drt.search"):` — an NVTX range when enabled, a no-op context otherwise."""
    return _range(name) if _ON else _NULL
