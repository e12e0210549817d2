"""Row-sharded, device-resident corpus-embedding store with multi-GPU exact search.

Replaces the reference's corpus hand-off — per-rank `.npy` + id JSONL, rank-0 `index.add` of
every shard, `faiss.write_index`, `faiss.read_index` on every other rank, then each rank
searching a full CPU replica (`DRT/trainer/trainer.py:191-262,287-297`) — with: every rank keeps
the rows it encoded on its own GPU, a search runs the local shard on each GPU, the per-shard
top-k candidate lists ([Q,k] fp32 scores + int64 global ids) are exchanged with one NCCL
all-gather over NVLink, and a merge kernel (`drt_merge_topk`, the device form of
`merge_retrieval_results_by_score`, `DRT/model/utils.py:215-229`) produces the global top-k.

Global ids follow the reference's concatenation order: rank-major, insertion order inside a
rank (rank 0's rows first, trainer.py:225-241 iterates the per-rank files and appends).

`ShardedCorpusStore(d, num_virtual_shards=G)` without an initialised process group keeps G
shards on ONE GPU and runs the same search + merge path, so the merge is testable on one GPU.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def _cuda_merge(scores: torch.Tensor, ids: torch.Tensor, k_out: int, sorted_unique: bool = False):
    """scores/ids: CUDA tensors [G, Q, k_in] -> (D [Q,k_out], I [Q,k_out]) via drt_merge_topk.
    `sorted_unique`: the lists are per-shard search results (ordered, disjoint ids)."""
    if not scores.is_cuda:
        raise RuntimeError("merge needs CUDA tensors: there is no CPU fallback")
    lib = _lib.load()
    G, Q, k_in = scores.shape
    dev = scores.device
    scores = scores.contiguous().to(torch.float32)
    ids = ids.contiguous().to(torch.int64)
    # the kernel merges <= 8192 entries per query: fold lists hierarchically above that
    while G * k_in > 8192 and G > 1:
        half = (G + 1) // 2
        d0, i0 = _cuda_merge(scores[:half], ids[:half], min(k_out, half * k_in), sorted_unique)
        d1, i1 = _cuda_merge(scores[half:], ids[half:], min(k_out, (G - half) * k_in), sorted_unique)
        kk = max(d0.shape[1], d1.shape[1])
        pad = lambda d, i: (torch.nn.functional.pad(d, (0, kk - d.shape[1]), value=-3.4028234663852886e38),
                            torch.nn.functional.pad(i, (0, kk - i.shape[1]), value=-1))
        d0, i0 = pad(d0, i0)
        d1, i1 = pad(d1, i1)
        scores, ids = torch.stack([d0, d1]), torch.stack([i0, i1])
        G, k_in = 2, kk
    D = torch.empty((Q, k_out), dtype=torch.float32, device=dev)
    I = torch.empty((Q, k_out), dtype=torch.int64, device=dev)
    _lib.check(lib.drt_merge_topk(G, scores.data_ptr(), ids.data_ptr(), Q, k_in, k_out, D.data_ptr(),
                                  I.data_ptr(), _lib.MERGE_SORTED_UNIQUE if sorted_unique else _lib.MERGE_DEFAULT,
                                  dev.index, _lib.current_stream_ptr(dev.index)), "merge_topk")
    return D, I


def _to_host(dev, *tensors):
    """Device tensors -> numpy arrays backed by (cached) pinned host blocks, one sync for all."""
    outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
    for o, t in zip(outs, tensors):
        o.copy_(t, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return tuple(o.numpy() for o in outs)


def shard_offsets(counts) -> list[int]:
    """Global id of each shard's first row, rank-major (exclusive prefix sum) + total."""
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    return off


class ShardedCorpusStore:
    def __init__(self, d: int, group=None, device: Optional[int] = None, num_virtual_shards: int = 0,
                 seg_rows: int = 0, index_factory: Optional[Callable] = None,
                 merge_fn: Optional[Callable] = None):
        self.d = int(d)
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized() and num_virtual_shards == 0
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else max(1, int(num_virtual_shards))
        if index_factory is None:
            from .faiss_compat import IndexFlatIP

            index_factory = lambda: IndexFlatIP(self.d, device=device, seg_rows=seg_rows)
        # shard results are ordered and id-disjoint: use the rank-based merge
        self._merge = merge_fn or (lambda s, i, k: _cuda_merge(s, i, k, sorted_unique=True))
        n_local = 1 if self.distributed else self.world
        self.shards = [index_factory() for _ in range(n_local)]
        self._offsets: Optional[list[int]] = None
        self._next_virtual = 0

    # ---- ingest -----------------------------------------------------------------------------
    def add(self, rows, shard: Optional[int] = None) -> None:
        """Append rows to this rank's shard (distributed) or to virtual shard `shard`."""
        if self.distributed:
            self.shards[0].add(rows)
        else:
            self.shards[0 if shard is None else shard].add(rows)
        self._offsets = None

    def add_split(self, rows) -> None:
        """Single-process helper: split `rows` contiguously over the virtual shards."""
        n = rows.shape[0]
        per = -(-n // self.world)
        for g in range(self.world):
            part = rows[g * per:(g + 1) * per]
            if part.shape[0]:
                self.shards[g].add(part)
        self._offsets = None

    def finalize(self) -> list[int]:
        """Agree on the global id ranges (one tiny all-gather of the local row counts)."""
        if self.distributed:
            local = torch.tensor([self.shards[0].ntotal], dtype=torch.int64)
            dev = getattr(self.shards[0], "device", None)
            backend = dist.get_backend(self.group)
            if backend == "nccl":
                local = local.cuda(dev)
            counts = [torch.zeros_like(local) for _ in range(self.world)]
            dist.all_gather(counts, local, group=self.group)
            self._offsets = shard_offsets([int(c.item()) for c in counts])
        else:
            self._offsets = shard_offsets([s.ntotal for s in self.shards])
        return self._offsets

    @property
    def ntotal(self) -> int:
        if self._offsets is None:
            self.finalize()
        return self._offsets[-1]

    # ---- search -----------------------------------------------------------------------------
    def search(self, q, k: int):
        """Every rank passes the SAME queries (CLI / benchmark use) and receives the global
        (D [Q,k], I [Q,k]).  Collective: all ranks must call with equal (Q, k)."""
        if self._offsets is None:
            self.finalize()
        host_in = isinstance(q, np.ndarray)
        dev = getattr(self.shards[0], "device", None)
        if host_in and dev is not None and torch.cuda.is_available():
            # host queries: the device path end to end, one D2H copy of the merged result (instead
            # of bouncing every shard's candidates through host memory)
            qd = self._upload_queries(np.ascontiguousarray(q, dtype=np.float32), dev)
            Dm, Im = self.search(qd, k)
            return _to_host(dev, Dm, Im)
        if self.distributed:
            D, I = self.shards[0].search(q, k, id_offset=self._offsets[self.rank])
            as_numpy = isinstance(D, np.ndarray)
            if as_numpy:
                D, I = torch.from_numpy(D), torch.from_numpy(I)
                if dist.get_backend(self.group) == "nccl":
                    D, I = D.cuda(self.shards[0].device), I.cuda(self.shards[0].device)
            Dm, Im = self._exchange_and_merge(D, I, k)
            if as_numpy:
                return Dm.cpu().numpy(), Im.cpu().numpy()
            return Dm, Im
        parts = [s.search(q, k, id_offset=self._offsets[g]) for g, s in enumerate(self.shards)]
        as_numpy = isinstance(parts[0][0], np.ndarray)
        if as_numpy:
            dev = getattr(self.shards[0], "device", None)
            to_t = (lambda a: torch.from_numpy(a).cuda(dev)) if dev is not None and torch.cuda.is_available() else torch.from_numpy
            parts = [(to_t(d), to_t(i)) for d, i in parts]
        Dm, Im = self._merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
        if as_numpy:
            return Dm.cpu().numpy(), Im.cpu().numpy()
        return Dm, Im

    # host query bytes above which a rank uploads only its 1/W slice over PCIe and the ranks
    # all-gather the slices over NVLink (every rank passes the same queries to `search`)
    SLICE_UPLOAD_MIN_BYTES = 1 << 20

    def _upload_queries(self, qn: np.ndarray, dev) -> torch.Tensor:
        Q, d = qn.shape
        W = self.world
        if not (self.distributed and W > 1 and Q >= W and qn.nbytes >= self.SLICE_UPLOAD_MIN_BYTES
                and dist.get_backend(self.group) == "nccl"):
            return torch.from_numpy(qn).cuda(dev, non_blocking=True)
        per = -(-Q // W)
        lo = min(self.rank * per, Q)
        hi = min(lo + per, Q)
        local = torch.zeros((per, d), dtype=torch.float32, device=torch.device("cuda", dev))
        if hi > lo:
            local[:hi - lo].copy_(torch.from_numpy(qn[lo:hi]), non_blocking=True)
        full = torch.empty((W * per, d), dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(full, local, group=self.group)
        return full[:Q]

    # candidate entries per rank above which the exchange switches from all-gather (every rank
    # merges all Q queries) to all-to-all (every rank merges Q/W queries, then the merged
    # [Q/W, k] slices are all-gathered): W x fewer bytes received and W x less merge work
    A2A_MIN_ENTRIES = 1 << 20
    A2A_MIN_WORLD = 4

    def _exchange_and_merge(self, D: torch.Tensor, I: torch.Tensor, k: int):
        W = self.world
        Q = D.shape[0]
        if W >= self.A2A_MIN_WORLD and Q * k >= self.A2A_MIN_ENTRIES and Q >= W:
            per = -(-Q // W)
            if per * W != Q:     # pad the query axis so it splits evenly; padding rows are dropped below
                padD = torch.full((per * W - Q, k), -3.4028234663852886e38, dtype=D.dtype, device=D.device)
                padI = torch.full((per * W - Q, k), -1, dtype=I.dtype, device=I.device)
                D, I = torch.cat([D, padD]), torch.cat([I, padI])
            rD, rI = torch.empty_like(D), torch.empty_like(I)
            dist.all_to_all_single(rD, D.contiguous(), group=self.group)    # rD[g] = rank g's list for MY query slice
            dist.all_to_all_single(rI, I.contiguous(), group=self.group)
            mD, mI = self._merge(rD.view(W, per, k), rI.view(W, per, k), k)
            oD, oI = torch.empty_like(D), torch.empty_like(I)
            dist.all_gather_into_tensor(oD, mD.contiguous(), group=self.group)
            dist.all_gather_into_tensor(oI, mI.contiguous(), group=self.group)
            return oD[:Q], oI[:Q]
        if D.is_cuda:
            # gather straight into the [W, Q, k] layout the merge kernel reads (no list copies)
            gD = torch.empty((W * Q, k), dtype=D.dtype, device=D.device)
            gI = torch.empty((W * Q, k), dtype=I.dtype, device=I.device)
            dist.all_gather_into_tensor(gD, D.contiguous(), group=self.group)
            dist.all_gather_into_tensor(gI, I.contiguous(), group=self.group)
            return self._merge(gD.view(W, Q, k), gI.view(W, Q, k), k)
        Ds = [torch.empty_like(D) for _ in range(W)]
        Is = [torch.empty_like(I) for _ in range(W)]
        dist.all_gather(Ds, D, group=self.group)
        dist.all_gather(Is, I, group=self.group)
        return self._merge(torch.stack(Ds), torch.stack(Is), k)

    def search_local_queries(self, q_local, k: int):
        """Trainer.evaluate use (trainer.py:287-297): every rank holds its OWN query batch
        (equal sizes).  Queries are all-gathered, searched against every shard, and each rank
        gets back the global top-k of its own queries."""
        if not self.distributed:
            return self.search(q_local, k)
        qs = [torch.empty_like(q_local) for _ in range(self.world)]
        dist.all_gather(qs, q_local.contiguous(), group=self.group)
        D, I = self.search(torch.cat(qs, dim=0), k)
        n = q_local.shape[0]
        return D[self.rank * n:(self.rank + 1) * n], I[self.rank * n:(self.rank + 1) * n]
