"""A `faiss`-shaped module covering exactly the faiss API subset DenseRetrievalToolkits touches,
backed by the device-resident B200 corpus store (C ABI in include/drt_b200.h).

Reference call sites this module serves, unmodified:
  faiss.IndexFlatIP(d)                 DRT/evaluator/index.py:19,23
  index.add(x)                         index.py:28, DRT/trainer/trainer.py:235
  index.search(x, k) -> (D, I)         index.py:32
  faiss.index_factory(d, str)          index.py:50   (only exact "Flat" + inner product)
  index.is_trained / train / verbose   index.py:52-54
  faiss.write_index / read_index       trainer.py:245,257

Drop-in use:  `denseretrievaltoolkits_b200.install_as_faiss()` registers this module as
`sys.modules["faiss"]`, after which `import DRT.evaluator.index` runs as written.

Differences from faiss, all additive: `add` and `search` also accept CUDA `torch.Tensor`s
(zero-copy: rows/queries stay on the device, results come back as CUDA tensors), and the index
lives in GPU memory.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

from . import _lib

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class IndexFlatIP:
    """Exact inner-product index; rows are kept on one B200 as an fp32 plane (exact rescoring)
    plus a bf16 plane (tensor-core first pass)."""

    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int, device: int | None = None, seg_rows: int = 0):
        self.d = int(d)
        self.is_trained = True
        self.verbose = False
        if device is None:
            device = _default_device()
        self._device = int(device)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.drt_store_create(ctypes.byref(h), self.d, self._device, int(seg_rows)),
                   "IndexFlatIP: creating the device store")
        self._h = h

    # ---- faiss attributes -------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(self._lib.drt_store_ntotal(self._h))

    @property
    def device(self) -> int:
        return self._device

    def train(self, x) -> None:  # flat index: nothing to train (index.py:54)
        return None

    def reset(self) -> None:
        _lib.check(self._lib.drt_store_reset(self._h), "reset")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.drt_store_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- add --------------------------------------------------------------------------------
    def add(self, x) -> None:
        if _is_torch_tensor(x):
            import torch

            if x.dim() != 2 or x.shape[1] != self.d:
                raise RuntimeError(f"add: expected [n,{self.d}], got {tuple(x.shape)}")
            if x.is_cuda:
                if x.device.index != self._device:
                    raise RuntimeError(f"add: rows are on cuda:{x.device.index}, store is on cuda:{self._device}")
                x = x.detach().to(torch.float32).contiguous()
                _lib.check(self._lib.drt_store_add(self._h, x.data_ptr(), x.shape[0], 1,
                                                   _lib.current_stream_ptr(self._device)), "add")
                return
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"add: expected [n,{self.d}] float32, got {x.shape}")
        _lib.check(self._lib.drt_store_add(self._h, x.ctypes.data, x.shape[0], 0,
                                           _lib.current_stream_ptr(self._device)), "add")

    # ---- search -----------------------------------------------------------------------------
    def search(self, x, k: int, *, id_offset: int = 0, flags: int = 0):
        k = int(k)
        if k <= 0:
            raise RuntimeError(f"search: k must be positive, got {k}")
        if _is_torch_tensor(x) and x.is_cuda:
            import torch

            if x.dim() != 2 or x.shape[1] != self.d:
                raise RuntimeError(f"search: expected [nq,{self.d}], got {tuple(x.shape)}")
            if x.device.index != self._device:
                raise RuntimeError("search: queries and store are on different devices")
            x = x.detach().to(torch.float32).contiguous()
            nq = x.shape[0]
            D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
            _lib.check(self._lib.drt_search(self._h, x.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), 1,
                                            int(id_offset), int(flags),
                                            _lib.current_stream_ptr(self._device)), "search")
            return D, I
        if _is_torch_tensor(x):
            x = x.detach().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"search: expected [nq,{self.d}] float32, got {x.shape}")
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        _lib.check(self._lib.drt_search(self._h, x.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data, 0,
                                        int(id_offset), int(flags),
                                        _lib.current_stream_ptr(self._device)), "search")
        return D, I

    def search_stats(self) -> dict:
        buf = (ctypes.c_int64 * 12)()
        _lib.check(self._lib.drt_search_stats(self._h, buf), "search_stats")
        names = ["launches", "filter_launches", "overflow_retries", "kprime", "flagged_queries",
                 "ctas_per_tile", "chunks", "filter_ns", "exact_queries"]
        return dict(zip(names, [int(v) for v in buf]))

    # ---- reconstruct ------------------------------------------------------------------------
    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.ntotal - i0
        out = np.empty((n, self.d), dtype=np.float32)
        _lib.check(self._lib.drt_store_reconstruct(self._h, int(i0), int(n), out.ctypes.data, 0,
                                                   _lib.current_stream_ptr(self._device)), "reconstruct_n")
        return out

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]


def _default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


def index_factory(d: int, description: str, metric: int = METRIC_INNER_PRODUCT):
    """faiss.index_factory (index.py:50).  Only the exact configuration is served: "Flat" with
    the inner-product metric; anything else (IVF/PQ/HNSW..., L2) is approximate or a different
    metric and is refused rather than silently approximated."""
    if description.strip() != "Flat":
        raise RuntimeError(f"index_factory: only 'Flat' is supported by the exact B200 path, got {description!r}")
    if metric != METRIC_INNER_PRODUCT:
        raise RuntimeError("index_factory: only METRIC_INNER_PRODUCT is supported")
    return IndexFlatIP(d)


# faiss IndexFlat on-disk layout (faiss/impl/index_write.cpp, from the published format):
# fourcc "IxFI", d:int32, ntotal:int64, dummy:int64 x2 (1<<20), is_trained:uint8,
# metric_type:int32, then the raw vector as count:uint64 (number of floats) + float32 data.
_FOURCC = b"IxFI"
_CHUNK_ROWS = 1 << 18


def write_index(index: IndexFlatIP, path: str) -> None:
    """faiss.write_index (trainer.py:245): streams the fp32 plane back to disk."""
    n, d = index.ntotal, index.d
    with open(path, "wb") as f:
        f.write(_FOURCC)
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, METRIC_INNER_PRODUCT))
        f.write(struct.pack("<Q", n * d))
        for r0 in range(0, n, _CHUNK_ROWS):
            index.reconstruct_n(r0, min(_CHUNK_ROWS, n - r0)).tofile(f)


def read_index(path: str, device: int | None = None) -> IndexFlatIP:
    """faiss.read_index (trainer.py:257)."""
    with open(path, "rb") as f:
        if f.read(4) != _FOURCC:
            raise RuntimeError(f"read_index: {path} is not an IndexFlatIP file")
        d, n, _, _, _, metric = struct.unpack("<iqqqBi", f.read(struct.calcsize("<iqqqBi")))
        (count,) = struct.unpack("<Q", f.read(8))
        if metric != METRIC_INNER_PRODUCT or count != n * d:
            raise RuntimeError(f"read_index: unsupported header in {path}")
        index = IndexFlatIP(d, device=device)
        for r0 in range(0, n, _CHUNK_ROWS):
            rows = min(_CHUNK_ROWS, n - r0)
            buf = np.fromfile(f, dtype=np.float32, count=rows * d)
            if buf.size != rows * d:
                raise RuntimeError(f"read_index: {path} is truncated")
            index.add(buf.reshape(rows, d))
    return index
