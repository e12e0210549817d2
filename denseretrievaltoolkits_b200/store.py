"""Row-sharded, device-resident corpus-embedding store with multi-GPU exact search.

Replaces the reference's corpus hand-off — per-rank `.npy` + id JSONL, rank-0 `index.add` of
every shard, `faiss.write_index`, `faiss.read_index` on every other rank, then each rank
searching a full CPU replica (`DRT/trainer/trainer.py:191-262,287-297`) — with: every rank keeps
the rows it encoded on its own GPU, a search runs the local shard on each GPU (without a host
round trip: `drt_search_async`), and the per-shard top-k candidate lists ([Q,k_l] fp32 scores +
int64 global ids) are exchanged and merged by ONE kernel over peer-mapped memory
(`drt_merge_topk_peers2`; NCCL all-gather / all-to-all + `drt_merge_topk` as the alternative) —
the device form of `merge_retrieval_results_by_score`, `DRT/model/utils.py:215-229`.

Global ids follow the reference's concatenation order: rank-major, insertion order inside a
rank (rank 0's rows first, trainer.py:225-241 iterates the per-rank files and appends).

`ShardedCorpusStore(d, num_virtual_shards=G)` without an initialised process group keeps G
shards on ONE GPU and runs the same search + merge path, so the merge is testable on one GPU.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._nvtx import rng as _nvtx

_NEG = -3.4028234663852886e38     # faiss' "no result" score


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


class _PeerExchange:
    """Candidate exchange + merge as one kernel over peer-mapped memory (`drt_merge_topk_peers`).

    One symmetric buffer per rank (torch symmetric memory: the same allocation mapped into every
    peer over NVLink / NVSwitch) holds this rank's [Q,kl] candidate lists — the local search
    writes them there directly — and the [Q,k] result.  After a cross-rank barrier rank r merges
    the queries of its slice reading all W lists from the peers and stores the merged rows into
    every peer's result region; a second barrier completes the step.  No all-gather, no
    all-to-all, and each rank merges Q/W queries."""

    def __init__(self, group, device: int, rank: int, world: int):
        import torch.distributed._symmetric_memory as symm_mem

        self._symm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device("cuda", device)
        self.rank, self.world = rank, world
        self.buf = None
        self.hdl = None
        self.cap = 0
        self._ptr_cache = {}       # (offsets, local_only) -> the five ctypes pointer tables of one launch
        self._view_cache = {}      # (Q, kl, k) -> tensor views into the symmetric buffer

    @staticmethod
    def _layout(Q: int, kl: int, k: int):
        al = lambda n: (n + 255) // 256 * 256
        sizes = [Q * kl * 4, Q * kl * 8, Q * k * 4, Q * k * 8, Q, 512]      # last: status byte (peers read it) | redo byte
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += al(n)
        return offs, o

    def views(self, Q: int, kl: int, k: int):
        cached = self._view_cache.get((Q, kl, k))
        if cached is not None:
            self.status, self.redo = cached[2]
            return cached[0], cached[1]
        offs, total = self._layout(Q, kl, k)
        if total > self.cap:                                   # collective: every rank sees the same shapes
            # everything queued on this rank — including the trailing barrier of the previous
            # search, i.e. every peer's reads of the old buffer — is done before it is dropped
            torch.cuda.current_stream(self.device).synchronize()
            cap = max(total + total // 4, 1 << 20)
            self.buf = self._symm.empty(cap, dtype=torch.uint8, device=self.device)
            self.hdl = self._symm.rendezvous(self.buf, self.group)
            self.cap = cap
            self._ptr_cache = {}
            self._view_cache = {}
        b = self.buf
        cD = b[offs[0]:offs[0] + Q * kl * 4].view(torch.float32).view(Q, kl)
        cI = b[offs[1]:offs[1] + Q * kl * 8].view(torch.int64).view(Q, kl)
        oD = b[offs[2]:offs[2] + Q * k * 4].view(torch.float32).view(Q, k)
        oI = b[offs[3]:offs[3] + Q * k * 8].view(torch.int64).view(Q, k)
        bad = b[offs[4]:offs[4] + Q]
        self.status, self.redo = b[offs[5]:offs[5] + 1], b[offs[5] + 256:offs[5] + 257]
        self._view_cache[(Q, kl, k)] = ((cD, cI, oD, oI, bad), offs, (self.status, self.redo))
        return (cD, cI, oD, oI, bad), offs

    def my_slice(self, Q: int):
        per = -(-Q // self.world)
        q0 = min(self.rank * per, Q)
        return q0, max(0, min(per, Q - q0))

    def merge(self, offs, Q: int, kl: int, k: int, local_only: bool = False, with_status: bool = False):
        """Barrier, merge my query slice from the peers' lists into everyone's result (or, with
        `local_only`, into mine alone — the flags still go to every rank), barrier.
        `with_status`: the kernel also ORs every rank's "local result not final" byte into `redo`."""
        import ctypes

        lib = _lib.load()
        W = self.world
        key = (tuple(offs), bool(local_only))
        tabs = self._ptr_cache.get(key)
        if tabs is None:
            bases = [int(p) for p in self.hdl.buffer_ptrs]
            arr = lambda off: (ctypes.c_void_p * W)(*[b + off for b in bases])
            out = (lambda off: (ctypes.c_void_p * W)(*[(b + off) if (not local_only or g == self.rank) else None
                                                        for g, b in enumerate(bases)]))
            tabs = self._ptr_cache[key] = (arr(offs[0]), arr(offs[1]), out(offs[2]), out(offs[3]), arr(offs[4]), arr(offs[5]),
                                           bases[self.rank] + offs[5] + 256)
        q0, qn = self.my_slice(Q)
        self.hdl.barrier(channel=0)
        _lib.check(lib.drt_merge_topk_peers2(W, tabs[0], tabs[1], q0, qn, kl, k, tabs[2], tabs[3], tabs[4],
                                             tabs[5] if with_status else None, tabs[6] if with_status else None,
                                             self.device.index, _lib.current_stream_ptr(self.device.index)),
                   "merge_topk_peers")
        self.hdl.barrier(channel=0)


def plan_rebalance(counts, times, max_shift: float = 0.10, granularity: int = 256):
    """Row counts proportional to the measured speed counts[g] / times[g] of every shard, each
    moved by at most `max_shift` of its current size, in multiples of `granularity` (the last
    shard takes the rounding remainder).  Returns the old counts when the plan is degenerate."""
    counts = [int(c) for c in counts]
    total = sum(counts)
    if not counts or min(counts) <= 0 or min(times) <= 0:
        return counts
    rate = [c / t for c, t in zip(counts, times)]
    want = [total * r / sum(rate) for r in rate]
    new = []
    for c, w in zip(counts, want):
        lo, hi = c * (1.0 - max_shift), c * (1.0 + max_shift)
        new.append(int(min(max(w, lo), hi)) // granularity * granularity)
    new[-1] += total - sum(new)
    return new if min(new) > 0 else counts


def shard_offsets(counts) -> list[int]:
    """Global id of each shard's first row, rank-major (exclusive prefix sum) + total."""
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    return off


class ShardedCorpusStore:
    def __init__(self, d: int, group=None, device: Optional[int] = None, num_virtual_shards: int = 0,
                 seg_rows: int = 0, _test_index_factory: Optional[Callable] = None,
                 _test_merge_fn: Optional[Callable] = None):
        # `_test_*`: hooks for the CPU (gloo) tests of the host logic only — they let a test stand
        # in for the device index / merge kernel.  The product path never sets them.
        index_factory, merge_fn = _test_index_factory, _test_merge_fn
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
        self._reduce_depth = True
        self._peer = None            # _PeerExchange | False (disabled) | None (not decided yet)
        self._local_events = None    # list of (start, end) CUDA events while rebalance() calibrates
        self.last_search = {}

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

    def rebalance(self, q_probe, k: int = 100, reps: int = 5, max_shift: float = 0.10, granularity: int = 256):
        """Speed-proportional row split (collective, NCCL groups).

        The ranks' GPUs do not run at the same clock under the power cap (measured: 3-5 % between
        the fastest and the slowest shard search of one step), and a step ends when the slowest
        rank does.  This measures every rank's shard-search time for `q_probe`, moves the shard
        boundaries so that rows per rank are proportional to the measured speed (by at most
        `max_shift` of a shard, in multiples of `granularity` rows), and ships the boundary rows to
        the neighbouring ranks with one all-to-all.  Global row ids do not change: shards stay
        contiguous ranges of the same global order, only `finalize()`'s offsets move.
        Returns the new per-rank row counts."""
        if not (self.distributed and self.world > 1 and dist.get_backend(self.group) == "nccl"):
            return [s.ntotal for s in self.shards]
        if self._offsets is None:
            self.finalize()
        shard = self.shards[0]
        dev = torch.device("cuda", shard.device)
        W, r, d = self.world, self.rank, self.d
        old = list(self._offsets)
        total = old[-1]
        # Calibration under the conditions of real use: whole search steps back to back (shard
        # search, exchange, merge, the per-step host word), timing only this rank's shard search.
        # Isolated, barrier-separated probes let the GPUs boost between repetitions and did not
        # predict the sustained per-GPU speed under the power cap.
        times = []
        if self._peer_ok(q_probe, self.local_depth(k), k):
            for i in range(reps + 3):
                self._local_events = []
                self.search(q_probe, k, local_results=True)
                torch.cuda.synchronize(dev)
                if i >= 3:
                    times.append(sum(a.elapsed_time(b) for a, b in self._local_events))
            self._local_events = None
        else:
            kl = self.local_depth(k)
            for i in range(reps + 2):
                dist.barrier(group=self.group)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                shard.search(q_probe, kl, id_offset=old[r])
                e1.record()
                torch.cuda.synchronize(dev)
                if i >= 2:
                    times.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(times)[len(times) // 2]], dtype=torch.float64, device=dev)
        ts = [torch.zeros_like(t) for _ in range(W)]
        dist.all_gather(ts, t, group=self.group)
        ts = [float(x.item()) for x in ts]
        counts = [old[g + 1] - old[g] for g in range(W)]
        new_counts = plan_rebalance(counts, ts, max_shift, granularity)
        if new_counts == counts:
            return counts
        new = shard_offsets(new_counts)
        # rows of mine that go to rank j: [old_r, old_r+1) ∩ [new_j, new_j+1)
        def overlap(a0, a1, b0, b1):
            lo, hi = max(a0, b0), min(a1, b1)
            return (lo, hi) if hi > lo else (lo, lo)
        send = [overlap(old[r], old[r + 1], new[j], new[j + 1]) for j in range(W)]
        recv = [overlap(old[j], old[j + 1], new[r], new[r + 1]) for j in range(W)]
        send_rows = [hi - lo for lo, hi in send]
        recv_rows = [hi - lo for lo, hi in recv]
        keep_lo, keep_hi = send[r]
        kept = shard.reconstruct_n_device(keep_lo - old[r], keep_hi - keep_lo)
        send_rows[r] = recv_rows[r] = 0                          # my own rows do not travel
        outbuf = torch.cat([shard.reconstruct_n_device(send[j][0] - old[r], send_rows[j]) for j in range(W)], dim=0) \
            if sum(send_rows) else torch.empty((0, d), dtype=torch.float32, device=dev)
        inbuf = torch.empty((sum(recv_rows), d), dtype=torch.float32, device=dev)
        dist.all_to_all_single(inbuf, outbuf, output_split_sizes=recv_rows, input_split_sizes=send_rows, group=self.group)
        pieces, pos = [], 0
        for j in range(W):                                       # ascending source rank = ascending global row id
            if j == r:
                pieces.append(kept)
            elif recv_rows[j]:
                pieces.append(inbuf[pos:pos + recv_rows[j]])
                pos += recv_rows[j]
        torch.cuda.synchronize(dev)
        shard.reset()
        for p in pieces:
            for r0 in range(0, p.shape[0], 1 << 20):
                shard.add(p[r0:r0 + (1 << 20)])
        torch.cuda.synchronize(dev)
        del kept, outbuf, inbuf, pieces
        self._offsets = None
        offs = self.finalize()
        if offs != new:
            raise RuntimeError(f"rebalance: shard sizes {offs} do not match the plan {new}")
        self.last_rebalance = {"search_ms": ts, "rows_before": counts, "rows_after": new_counts}
        return new_counts

    # ---- persistence ------------------------------------------------------------------------
    def save(self, directory: str) -> None:
        """Every rank writes its own shard as a faiss IndexFlat file (`shard{g}.faiss`, readable
        by faiss.read_index) — in parallel, instead of rank 0 writing the whole index and the
        others reading it back (trainer.py:245,257) — plus `store.json` with the id offsets."""
        import json
        import os

        from .faiss_compat import write_index

        offs = self.finalize() if self._offsets is None else self._offsets
        os.makedirs(directory, exist_ok=True)
        for j, shard in enumerate(self.shards):
            g = self.rank if self.distributed else j
            write_index(shard, os.path.join(directory, f"shard{g}.faiss"))
        if self.rank == 0:
            with open(os.path.join(directory, "store.json"), "w", encoding="utf-8") as f:
                json.dump({"d": self.d, "world": self.world, "offsets": offs}, f)
        if self.distributed:
            dist.barrier(group=self.group)

    @classmethod
    def load(cls, directory: str, group=None, device: Optional[int] = None, num_virtual_shards: int = 0,
             seg_rows: int = 0) -> "ShardedCorpusStore":
        """Inverse of `save` for the same number of shards (ranks, or virtual shards on one GPU)."""
        import json
        import os

        from .faiss_compat import read_index

        with open(os.path.join(directory, "store.json"), "r", encoding="utf-8") as f:
            meta = json.load(f)
        st = cls(int(meta["d"]), group=group, device=device, num_virtual_shards=num_virtual_shards, seg_rows=seg_rows)
        if st.world != int(meta["world"]):
            raise RuntimeError(f"store in {directory} has {meta['world']} shards, this process group has {st.world}")
        for j in range(len(st.shards)):
            g = st.rank if st.distributed else j
            st.shards[j] = read_index(os.path.join(directory, f"shard{g}.faiss"), device=device)
        if st.finalize() != [int(v) for v in meta["offsets"]]:
            raise RuntimeError(f"store in {directory}: shard sizes do not match store.json")
        return st

    # ---- search -----------------------------------------------------------------------------
    def local_depth(self, k: int) -> int:
        """Depth every shard is searched to for a global top-k.

        With rows dealt to W shards independently of the queries, a shard that owns a fraction p of
        the rows holds Binomial(k, p) of a query's global top-k (p = 1/W for even shards), so mean + 6 sigma entries per shard suffice almost surely
        (k=100, W=8: 40 instead of 100; k=1000, W=8: 192) — and rescoring, the candidate
        exchange and the merge all shrink with it.  The result stays EXACT: after the merge a
        query whose global k-th score does not lie strictly above the last entry of every
        truncated shard list is searched again at full depth (`_truncated`), and a store whose
        row order turns out to be correlated with the queries stops reducing the depth."""
        W = self.world
        if W <= 1 or not self._reduce_depth:
            return k
        if self._offsets is None:
            self.finalize()
        total = self._offsets[-1]
        if total <= 0:
            return k
        # the largest shard's share of the rows (1/W for even shards); every shard uses the same depth
        p = max(self._offsets[g + 1] - self._offsets[g] for g in range(W)) / float(total)
        kl = int(math.ceil((k * p + 6.0 * math.sqrt(k * p * (1.0 - p)) + 4.0) / 8.0) * 8)
        return min(k, max(kl, -(-k // W)))

    @staticmethod
    def _truncated(S: torch.Tensor, Dm: torch.Tensor, k: int) -> torch.Tensor:
        """S [W,Q,kl] shard lists, Dm [Q,k] merged: True where some shard's list is full and its
        last score is not strictly below the global k-th (it may hold further top-k rows)."""
        last = S[:, :, -1]
        return ((last >= Dm[:, k - 1].unsqueeze(0)) & (last > _NEG)).any(0)

    def search(self, q, k: int, flags: int = 0, local_results: bool = False):
        """Every rank passes the SAME queries (CLI / benchmark use) and receives the global
        (D [Q,k], I [Q,k]).  Collective: all ranks must call with equal (Q, k).

        `local_results=True`: every rank receives only the rows of ITS slice of the queries
        (`result_slice(Q)`; the reference evaluates per rank, trainer.py:287-297), so the merged
        rows are not broadcast and only Q/W rows are downloaded per rank."""
        if self._offsets is None:
            self.finalize()
        host_in = isinstance(q, np.ndarray)
        dev = getattr(self.shards[0], "device", None)
        on_gpu = dev is not None and torch.cuda.is_available()
        qd = self._upload_queries(np.ascontiguousarray(q, dtype=np.float32), dev) if (host_in and on_gpu) else q
        Q = qd.shape[0]
        kl = self.local_depth(k)
        local = bool(local_results) and self.distributed and self.world > 1
        q0, qn = self.result_slice(Q) if local else (0, Q)
        Dm, Im, bad, owned, redo = self._search_merged(qd, k, kl, flags, local)
        self.last_search = {"local_depth": kl, "requeried": 0, "redone": 0}
        to_host = host_in and on_gpu
        host = None
        # the ONE host round trip of a search: re-query count (+ the redo word of the asynchronous
        # shard searches), and with the host API the result rows in the same transfer
        nb = bad.sum(dtype=torch.int64).reshape(1) if bad is not None else None
        if redo is not None:
            nb = torch.cat([nb if nb is not None else torch.zeros(1, dtype=torch.int64, device=redo.device), redo.to(torch.int64)])
        if to_host:
            if nb is None:
                nb = torch.zeros(1, dtype=torch.int64, device=Dm.device)
            host = _to_host(dev, Dm[q0:q0 + qn], Im[q0:q0 + qn], nb)
            words = host[2].tolist()
        else:
            words = nb.tolist() if nb is not None else [0]   # same values on every rank (computed from exchanged data)
        if redo is not None and words[-1]:
            # some rank's first pass was not final (candidate overflow / flagged query): all ranks
            # repeat the step on the synchronous path, which retries and refines
            Dm, Im, bad, owned, _ = self._search_merged(qd, k, kl, flags, local, allow_async=False)
            words = [int(bad.sum().item()) if bad is not None else 0]
            host = None
            self.last_search["redone"] = 1
        nbad = int(words[0])
        if nbad:
            idx = bad.nonzero().squeeze(1)
            qt = torch.from_numpy(qd) if isinstance(qd, np.ndarray) else qd
            qb = qt.index_select(0, idx.to(qt.device))
            Db, Ib, _, _, _ = self._search_merged(qb.numpy() if isinstance(qd, np.ndarray) else qb, k, k, flags, False, allow_async=False)
            if not owned:
                Dm, Im = Dm.clone(), Im.clone()
            Dm[idx] = Db
            Im[idx] = Ib
            owned, host = True, None
            self.last_search["requeried"] = nbad
            if 4 * nbad > Dm.shape[0]:
                self._reduce_depth = False     # row order correlates with the queries
        if to_host:
            return (host[0], host[1]) if host is not None else _to_host(dev, Dm[q0:q0 + qn], Im[q0:q0 + qn])
        if host_in:
            return Dm[q0:q0 + qn].cpu().numpy(), Im[q0:q0 + qn].cpu().numpy()
        if not owned:                          # the peer result region is reused by the next search
            return Dm[q0:q0 + qn].clone(), Im[q0:q0 + qn].clone()
        return Dm[q0:q0 + qn], Im[q0:q0 + qn]

    def result_slice(self, Q: int):
        """(first query, count) of the rows this rank keeps with `local_results=True`."""
        per = -(-Q // self.world)
        q0 = min(self.rank * per, Q)
        return q0, max(0, min(per, Q - q0))

    # Exchange + merge as one kernel over peer-mapped memory (torch symmetric memory over NVLink /
    # NVSwitch) instead of NCCL all-gather / all-to-all + merge.  Default on for NCCL groups of up
    # to 16 ranks; DRT_B200_PEER_EXCHANGE=0 selects the NCCL path, and so does a box on which the
    # symmetric allocation cannot be set up (the ranks agree on that with one all-reduce).
    def _peer_ok(self, q, kl: int, k: int) -> bool:
        if self._peer is False or not torch.is_tensor(q) or not q.is_cuda or q.shape[0] == 0:
            return False
        if self._peer is None:
            import os

            self._peer = False
            if os.environ.get("DRT_B200_PEER_EXCHANGE", "1") != "0" and 1 < self.world <= 16 and dist.get_backend(self.group) == "nccl":
                peer, ok = None, 1
                try:
                    peer = _PeerExchange(self.group, self.shards[0].device, self.rank, self.world)
                    peer.views(1, 1, 1)                                    # first rendezvous
                except Exception:
                    ok = 0
                flag = torch.tensor([ok], dtype=torch.int32, device=q.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
                if int(flag.item()) == 1:
                    self._peer = peer
        return self._peer is not False and self.world * kl <= 8192 and k <= 4096 and k <= self.world * kl

    def _search_merged(self, q, k: int, kl: int, flags: int, local: bool = False, allow_async: bool = True):
        """Depth-kl search of every shard, candidate exchange, merge to depth k.
        Returns torch (D [Q,k], I [Q,k], truncated-mask [Q] or None when kl == k, owned, redo) —
        `owned` False: D / I are views of the peer-exchange result region, valid until the next
        search; with `local` only this rank's `result_slice` rows of D / I are filled; `redo`
        (1-element uint8 tensor or None): set when a rank's asynchronous shard search was not
        final and the step has to be repeated with `allow_async=False`."""
        if self.distributed and self._peer_ok(q, kl, k):
            (cD, cI, oD, oI, bad), offs = self._peer.views(q.shape[0], kl, k)
            # shard search without a host round trip when the shape allows it: the search, the
            # barriers and the merge kernel are then enqueued back to back
            use_async = (allow_async and q.shape[0] <= 16384 and self.d % 64 == 0 and q.is_contiguous()
                         and q.dtype is torch.float32 and q.data_ptr() % 16 == 0 and self.shards[0].ntotal > 0)
            if self._local_events is not None:           # rebalance(): time the shard search under real step conditions
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self.shards[0].search(q, kl, id_offset=self._offsets[self.rank], flags=flags, out=(cD, cI),
                                  status=self._peer.status if use_async else None)
            if self._local_events is not None:
                ev[1].record()
                self._local_events.append(ev)
            with _nvtx("drt.exchange_merge_peers"):
                self._peer.merge(offs, q.shape[0], kl, k, local_only=local, with_status=use_async)
            return oD, oI, (bad.bool() if kl < k else None), False, (self._peer.redo if use_async else None)
        if self.distributed:
            D, I = self.shards[0].search(q, kl, id_offset=self._offsets[self.rank], flags=flags)
            if isinstance(D, np.ndarray):
                D, I = torch.from_numpy(D), torch.from_numpy(I)
                if dist.get_backend(self.group) == "nccl":
                    D, I = D.cuda(self.shards[0].device), I.cuda(self.shards[0].device)
            with _nvtx("drt.exchange_merge_nccl"):
                return self._exchange_and_merge(D, I, k, check=kl < k) + (True, None)
        parts = [s.search(q, kl, id_offset=self._offsets[g], flags=flags) for g, s in enumerate(self.shards)]
        if isinstance(parts[0][0], np.ndarray):
            dev = getattr(self.shards[0], "device", None)
            to_t = (lambda a: torch.from_numpy(a).cuda(dev)) if dev is not None and torch.cuda.is_available() else torch.from_numpy
            parts = [(to_t(d), to_t(i)) for d, i in parts]
        S = torch.stack([p[0] for p in parts])
        Dm, Im = self._merge(S, torch.stack([p[1] for p in parts]), k)
        return Dm, Im, (self._truncated(S, Dm, k) if kl < k else None), True, None

    def profile_phases(self, q, k: int, reps: int = 5):
        """Measurement helper (bench.py): CUDA-event times of the two phases of a peer-exchange
        search on THIS rank — the local shard search, and barrier + exchange/merge kernel +
        barrier — in ms, averaged over `reps`.  The second includes the wait for the slowest
        rank (rank skew shows up there).  Returns None when the peer path is not in use."""
        kl = self.local_depth(k)
        if not (self.distributed and self._peer_ok(q, kl, k)):
            return None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_local = t_merge = 0.0
        for _ in range(reps):
            (cD, cI, oD, oI, bad), offs = self._peer.views(q.shape[0], kl, k)
            dist.barrier(group=self.group)
            ev[0].record()
            self.shards[0].search(q, kl, id_offset=self._offsets[self.rank], out=(cD, cI))
            ev[1].record()
            self._peer.merge(offs, q.shape[0], kl, k)
            ev[2].record()
            torch.cuda.synchronize()
            t_local += ev[0].elapsed_time(ev[1])
            t_merge += ev[1].elapsed_time(ev[2])
        W, Q = self.world, q.shape[0]
        return {"local_search_ms": t_local / reps, "exchange_merge_ms": t_merge / reps,
                # bytes this rank's merge kernel moves over NVLink: it reads the lists of its Q/W
                # queries from the W-1 peers and writes the merged rows + flags to the W-1 peers
                "nvlink_bytes_read": (Q // W) * (W - 1) * kl * 12, "nvlink_bytes_written": (Q // W) * (W - 1) * (k * 12 + 1)}

    # candidate entries per rank above which the exchange switches from all-gather (every rank
    # merges all Q queries) to all-to-all (every rank merges Q/W queries, then the merged
    # [Q/W, k] slices are all-gathered): W x fewer bytes received and W x less merge work
    A2A_MIN_ENTRIES = 1 << 20
    A2A_MIN_WORLD = 4

    def _exchange_and_merge(self, D: torch.Tensor, I: torch.Tensor, k: int, check: bool = False):
        """D, I: this rank's [Q, kl] lists.  Returns (Dm [Q,k], Im [Q,k], truncated-mask | None)."""
        W = self.world
        Q, kl = D.shape
        if W >= self.A2A_MIN_WORLD and Q * kl >= self.A2A_MIN_ENTRIES and Q >= W:
            per = -(-Q // W)
            if per * W != Q:     # pad the query axis so it splits evenly; padding rows are dropped below
                padD = torch.full((per * W - Q, kl), _NEG, dtype=D.dtype, device=D.device)
                padI = torch.full((per * W - Q, kl), -1, dtype=I.dtype, device=I.device)
                D, I = torch.cat([D, padD]), torch.cat([I, padI])
            rD, rI = torch.empty_like(D), torch.empty_like(I)
            dist.all_to_all_single(rD, D.contiguous(), group=self.group)    # rD[g] = rank g's list for MY query slice
            dist.all_to_all_single(rI, I.contiguous(), group=self.group)
            mD, mI = self._merge(rD.view(W, per, kl), rI.view(W, per, kl), k)
            oD = torch.empty((W * per, k), dtype=D.dtype, device=D.device)
            oI = torch.empty((W * per, k), dtype=I.dtype, device=I.device)
            dist.all_gather_into_tensor(oD, mD.contiguous(), group=self.group)
            dist.all_gather_into_tensor(oI, mI.contiguous(), group=self.group)
            bad = None
            if check:
                mine = self._truncated(rD.view(W, per, kl), mD, k).to(torch.uint8)
                allbad = torch.empty((W * per,), dtype=torch.uint8, device=D.device)
                dist.all_gather_into_tensor(allbad, mine, group=self.group)
                bad = allbad[:Q].bool()
            return oD[:Q], oI[:Q], bad
        if D.is_cuda:
            # gather straight into the [W, Q, kl] layout the merge kernel reads (no list copies)
            gD = torch.empty((W * Q, kl), dtype=D.dtype, device=D.device)
            gI = torch.empty((W * Q, kl), dtype=I.dtype, device=I.device)
            dist.all_gather_into_tensor(gD, D.contiguous(), group=self.group)
            dist.all_gather_into_tensor(gI, I.contiguous(), group=self.group)
            S, Si = gD.view(W, Q, kl), gI.view(W, Q, kl)
        else:
            Ds = [torch.empty_like(D) for _ in range(W)]
            Is = [torch.empty_like(I) for _ in range(W)]
            dist.all_gather(Ds, D, group=self.group)
            dist.all_gather(Is, I, group=self.group)
            S, Si = torch.stack(Ds), torch.stack(Is)
        Dm, Im = self._merge(S, Si, k)
        return Dm, Im, (self._truncated(S, Dm, k) if check else None)

    def search_local_queries(self, q_local, k: int):
        """Trainer.evaluate use (trainer.py:287-297): every rank holds its OWN query batch
        (equal sizes).  Queries are all-gathered, searched against every shard, and each rank
        gets back the global top-k of its own queries."""
        if not self.distributed:
            return self.search(q_local, k)
        n = q_local.shape[0]
        qs = torch.empty((self.world * n,) + tuple(q_local.shape[1:]), dtype=q_local.dtype, device=q_local.device)
        dist.all_gather_into_tensor(qs, q_local.contiguous(), group=self.group)
        # rank r's queries are rows [r n, (r+1) n) of the gathered batch = its `result_slice`
        return self.search(qs, k, local_results=True)

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
