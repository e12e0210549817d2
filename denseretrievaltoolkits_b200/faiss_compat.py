"""A `faiss`-shaped module covering exactly the faiss API subset DenseRetrievalToolkits touches,
backed by the device-resident B200 corpus store (C ABI in include/drt_b200.h).

Reference call sites this module serves, unmodified:
  faiss.IndexFlatIP(d)                 DRT/evaluator/index.py:19,23
  index.add(x)                         index.py:28, DRT/trainer/trainer.py:235
  index.search(x, k) -> (D, I)         index.py:32
  faiss.index_factory(d, str)          index.py:50   (only exact "Flat"; faiss' default metric L2 -> IndexFlatL2)
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
from ._nvtx import rng as _nvtx

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
                with _nvtx("drt.store_add"):
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
    def search(self, x, k: int, *, id_offset: int = 0, flags: int = 0, out=None, status=None):
        """`out=(D, I)`: preallocated contiguous CUDA tensors [nq,k] float32 / int64 the results
        are written into (device queries only; used by the peer-memory exchange of store.py).
        `status`: a 1-element CUDA uint8 tensor -> the search runs without any host round trip
        (`drt_search_async`) and sets status[0] = 1 when its result is not final and the caller
        must search again without `status` (candidate overflow / a query the certificate flagged)."""
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
            if out is not None:
                D, I = out
                if (tuple(D.shape) != (nq, k) or tuple(I.shape) != (nq, k) or D.dtype is not torch.float32
                        or I.dtype is not torch.int64 or not D.is_contiguous() or not I.is_contiguous()
                        or D.device != x.device or I.device != x.device):
                    raise RuntimeError("search: out=(D, I) must be contiguous CUDA [nq,k] float32 / int64 tensors")
            else:
                D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
                I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
            if status is not None:
                with _nvtx("drt.search_async"):
                    _lib.check(self._lib.drt_search_async(self._h, x.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(),
                                                          int(id_offset), int(flags), status.data_ptr(),
                                                          _lib.current_stream_ptr(self._device)), "search_async")
                return D, I
            with _nvtx("drt.search"):
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
        with _nvtx("drt.search_host"):
            _lib.check(self._lib.drt_search(self._h, x.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data, 0,
                                            int(id_offset), int(flags),
                                            _lib.current_stream_ptr(self._device)), "search")
        return D, I

    def search_stats(self) -> dict:
        buf = (ctypes.c_int64 * 12)()
        _lib.check(self._lib.drt_search_stats(self._h, buf), "search_stats")
        names = ["launches", "filter_launches", "overflow_retries", "kprime", "flagged_queries",
                 "ctas_per_tile", "chunks", "filter_ns", "exact_queries", "refined_queries", "rescored_rows"]
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

    def reconstruct_n_device(self, i0: int = 0, n: int | None = None):
        """`reconstruct_n` into a CUDA tensor on the store's device (no host round trip)."""
        import torch

        if n is None:
            n = self.ntotal - i0
        out = torch.empty((int(n), self.d), dtype=torch.float32, device=torch.device("cuda", self._device))
        if n:
            _lib.check(self._lib.drt_store_reconstruct(self._h, int(i0), int(n), out.data_ptr(), 1,
                                                       _lib.current_stream_ptr(self._device)), "reconstruct_n_device")
        return out


def _default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


_FLT_MAX = np.float32(3.4028234663852886e38)
_L2_CHUNK = 1 << 18


def _bf16_round(a: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32 (finite inputs)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return u.view(np.float32)


class IndexFlatL2:
    """Exact squared-L2 index — what `faiss.index_factory(d, "Flat")` returns with faiss' default
    metric, i.e. what the reference's `FaissRetriever(init_reps, "Flat")` (index.py:47-54) holds.

    Runs on the inner-product store: argmin |q - x|^2 = argmax (q.x - |x|^2 / 2), so every row is
    stored as [x, n1, n2, n3] with n1 + n2 + n3 = -|x|^2 / 2 split exactly into three
    bf16-representable pieces (the tensor-core first pass sees the norm term to fp32 accuracy),
    queries are searched as [q, 1, 1, 1], and D = |q|^2 - 2 * score, ascending."""

    metric_type = METRIC_L2

    def __init__(self, d: int, device: int | None = None, seg_rows: int = 0):
        self.d = int(d)
        self.is_trained = True
        self.verbose = False
        self._ip = IndexFlatIP(self.d + 3, device=device, seg_rows=seg_rows)
        # the three norm columns (and the query's ones) are exact in bf16: let the exactness
        # certificate bound the rounding error of the head dims only
        _lib.check(self._ip._lib.drt_store_set_exact_tail(self._ip._h, 3), "set_exact_tail")

    @property
    def ntotal(self) -> int:
        return self._ip.ntotal

    @property
    def device(self) -> int:
        return self._ip.device

    def train(self, x) -> None:
        return None

    def reset(self) -> None:
        self._ip.reset()

    def add(self, x) -> None:
        if _is_torch_tensor(x) and x.is_cuda:
            import torch

            if x.dim() != 2 or x.shape[1] != self.d:
                raise RuntimeError(f"add: expected [n,{self.d}], got {tuple(x.shape)}")
            for r0 in range(0, x.shape[0], _L2_CHUNK):
                xb = x[r0:r0 + _L2_CHUNK].detach().to(torch.float32)
                nrm = -0.5 * (xb.double() * xb.double()).sum(1)
                n1 = nrm.float().bfloat16().float()
                n2 = (nrm - n1.double()).float().bfloat16().float()
                n3 = (nrm - n1.double() - n2.double()).float().bfloat16().float()
                self._ip.add(torch.cat([xb, n1[:, None], n2[:, None], n3[:, None]], dim=1))
            return
        if _is_torch_tensor(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"add: expected [n,{self.d}] float32, got {x.shape}")
        for r0 in range(0, x.shape[0], _L2_CHUNK):
            xb = x[r0:r0 + _L2_CHUNK]
            nrm = -0.5 * np.einsum("ij,ij->i", xb, xb, dtype=np.float64)
            n1 = _bf16_round(nrm.astype(np.float32))
            n2 = _bf16_round((nrm - n1).astype(np.float32))
            n3 = _bf16_round((nrm - n1 - n2).astype(np.float32))
            self._ip.add(np.concatenate([xb, n1[:, None], n2[:, None], n3[:, None]], axis=1))

    def search(self, x, k: int, *, id_offset: int = 0, flags: int = 0):
        if _is_torch_tensor(x) and x.is_cuda:
            import torch

            if x.dim() != 2 or x.shape[1] != self.d:
                raise RuntimeError(f"search: expected [nq,{self.d}], got {tuple(x.shape)}")
            xq = x.detach().to(torch.float32)
            S, I = self._ip.search(torch.cat([xq, torch.ones((xq.shape[0], 3), dtype=torch.float32, device=xq.device)], dim=1),
                                   k, id_offset=id_offset, flags=flags)
            qn = (xq.double() * xq.double()).sum(1, keepdim=True)
            D = torch.where(I < 0, torch.full_like(S, float(_FLT_MAX), dtype=torch.float64),
                            (qn - 2.0 * S.double()).clamp_min_(0.0)).float()
            return D, I
        if _is_torch_tensor(x):
            x = x.detach().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"search: expected [nq,{self.d}] float32, got {x.shape}")
        S, I = self._ip.search(np.concatenate([x, np.ones((x.shape[0], 3), np.float32)], axis=1), k,
                               id_offset=id_offset, flags=flags)
        qn = np.einsum("ij,ij->i", x, x, dtype=np.float64)[:, None]
        D = np.where(I < 0, np.float64(_FLT_MAX), np.maximum(qn - 2.0 * S.astype(np.float64), 0.0)).astype(np.float32)
        return D, I

    def search_stats(self) -> dict:
        return self._ip.search_stats()

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        return np.ascontiguousarray(self._ip.reconstruct_n(i0, n)[:, :self.d])

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]


def index_factory(d: int, description: str, metric: int = METRIC_L2):
    """faiss.index_factory (index.py:50), same signature and default metric as faiss (L2).  Only
    the exact configuration is served: "Flat" -> IndexFlatL2 / IndexFlatIP; anything else
    (IVF/PQ/HNSW...) is approximate and is refused rather than silently made exact or approximated."""
    if description.strip() != "Flat":
        raise RuntimeError(f"index_factory: only 'Flat' is supported by the exact B200 path, got {description!r}")
    if metric == METRIC_L2:
        return IndexFlatL2(d)
    if metric == METRIC_INNER_PRODUCT:
        return IndexFlatIP(d)
    raise RuntimeError(f"index_factory: unsupported metric {metric}")


# faiss IndexFlat on-disk layout (faiss/impl/index_write.cpp, from the published format):
# fourcc "IxFI", d:int32, ntotal:int64, dummy:int64 x2 (1<<20), is_trained:uint8,
# metric_type:int32, then the raw vector as count:uint64 (number of floats) + float32 data.
_FOURCC = {METRIC_INNER_PRODUCT: b"IxFI", METRIC_L2: b"IxF2"}
_CHUNK_ROWS = 1 << 18


def write_index(index, path: str) -> None:
    """faiss.write_index (trainer.py:245): streams the fp32 plane back to disk."""
    n, d = index.ntotal, index.d
    metric = getattr(index, "metric_type", METRIC_INNER_PRODUCT)
    with open(path, "wb") as f:
        f.write(_FOURCC[metric])
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric))
        f.write(struct.pack("<Q", n * d))
        for r0 in range(0, n, _CHUNK_ROWS):
            index.reconstruct_n(r0, min(_CHUNK_ROWS, n - r0)).tofile(f)


def read_index(path: str, device: int | None = None):
    """faiss.read_index (trainer.py:257)."""
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in _FOURCC.values():
            raise RuntimeError(f"read_index: {path} is not an IndexFlatIP / IndexFlatL2 file")
        d, n, _, _, _, metric = struct.unpack("<iqqqBi", f.read(struct.calcsize("<iqqqBi")))
        (count,) = struct.unpack("<Q", f.read(8))
        if _FOURCC.get(metric) != fourcc or count != n * d:
            raise RuntimeError(f"read_index: unsupported header in {path}")
        index = IndexFlatIP(d, device=device) if metric == METRIC_INNER_PRODUCT else IndexFlatL2(d, device=device)
        for r0 in range(0, n, _CHUNK_ROWS):
            rows = min(_CHUNK_ROWS, n - r0)
            buf = np.fromfile(f, dtype=np.float32, count=rows * d)
            if buf.size != rows * d:
                raise RuntimeError(f"read_index: {path} is truncated")
            index.add(buf.reshape(rows, d))
    return index
