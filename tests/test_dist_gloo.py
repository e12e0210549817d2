"""CPU, world_size 2 over gloo: the multi-rank host logic of the sharded store (id offsets,
candidate all-gather, merge ordering) and of the rank-major gather used by the distributed loss.
Compute is injected from the oracle; the product's defaults are the CUDA paths."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip
from oracle import merge as omerge


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _OracleIndex:
    def __init__(self, d):
        self.inner = flat_ip.IndexFlatIP(d)
        self.device = None

    @property
    def ntotal(self):
        return self.inner.ntotal

    def add(self, x):
        self.inner.add(np.asarray(x))

    def search(self, q, k, id_offset=0, flags=0):
        D, I = self.inner.search(np.asarray(q), k)
        return D, np.where(I >= 0, I + id_offset, I)


def _oracle_merge(scores, ids, k):
    D, I = omerge.merge_topk(scores.numpy(), ids.numpy(), k)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from denseretrievaltoolkits_b200.losses import gather_rank_major
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 32)).astype(np.float32)
    x[900:950] = x[100:150]                       # cross-shard exact ties
    q = rng.standard_normal((6, 32)).astype(np.float32)
    split = [0, 430, 1000]                        # uneven shards
    st = ShardedCorpusStore(32, _test_index_factory=lambda: _OracleIndex(32), _test_merge_fn=_oracle_merge)
    st.add(x[split[rank]:split[rank + 1]])
    offs = st.finalize()
    assert offs == split, offs
    D, I = st.search(q, 20)
    D2, I2 = st.search(q, 200)                     # reduced per-shard depth (168 of 200: the larger shard owns 57 % of the rows)
    depth2 = dict(st.last_search)
    # rows correlated with the queries: shard 0 owns the whole top-200 -> truncation check -> re-query
    xc = x.copy()
    xc[:430] *= 5.0
    stc = ShardedCorpusStore(32, _test_index_factory=lambda: _OracleIndex(32), _test_merge_fn=_oracle_merge)
    stc.add(xc[split[rank]:split[rank + 1]])
    Dc, Ic = stc.search(q, 200)
    depthc = dict(stc.last_search)
    Dl, Il = st.search_local_queries(torch.from_numpy(q[rank * 3:(rank + 1) * 3]), 20)
    Ds, Is = st.search(q, 20, local_results=True)                 # every rank keeps its slice of the queries
    assert st.result_slice(6) == (rank * 3, 3) and np.array_equal(np.asarray(Is), np.asarray(I)[rank * 3:(rank + 1) * 3])
    t = torch.full((2, 3), float(rank), requires_grad=True)
    g = gather_rank_major(t, rank, world)
    g.sum().backward()
    from denseretrievaltoolkits_b200.evaluation import reduce_metrics

    red = reduce_metrics({"Recall@5": 2.0 + rank, "MRR@5": 1.0, "query_num": 0}, 3 + rank)   # per-rank sums
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), D=D, I=I, I2=I2, Ic=Ic,
             depth=np.array([depth2["local_depth"], depth2["requeried"], depthc["local_depth"], depthc["requeried"]]), Dl=np.asarray(Dl), Il=np.asarray(Il),
             g=g.detach().numpy(), grad=t.grad.numpy(), red=np.array([red["Recall@5"], red["MRR@5"], red["query_num"]]))
    dist.destroy_process_group()


def test_sharded_store_two_ranks_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 32)).astype(np.float32)
    x[900:950] = x[100:150]
    q = rng.standard_normal((6, 32)).astype(np.float32)
    Dr, Ir = flat_ip.flat_ip_search(x, q, 20)
    for rank in range(2):
        r = np.load(tmp_path / f"r{rank}.npz")
        np.testing.assert_array_equal(r["I"], Ir)          # G-shard result == 1-shard result, ids bit-for-bit
        np.testing.assert_allclose(r["D"], Dr, rtol=1e-6)
        np.testing.assert_array_equal(r["Il"], Ir[rank * 3:(rank + 1) * 3])
        np.testing.assert_array_equal(r["I2"], flat_ip.flat_ip_search(x, q, 200)[1])
        xc = x.copy()
        xc[:430] *= 5.0
        np.testing.assert_array_equal(r["Ic"], flat_ip.flat_ip_search(xc, q, 200)[1])
        assert r["depth"][0] == 168 and r["depth"][2] == 168 and r["depth"][3] >= 1, r["depth"]   # largest shard owns 57 % of the rows
        # rank-major gather; gradient flows only into the local slot (biencoder.py:251)
        np.testing.assert_array_equal(r["g"], np.repeat([[0.0], [1.0]], 2, axis=0).repeat(3, axis=1).reshape(4, 3))
        np.testing.assert_array_equal(r["grad"], np.ones((2, 3)))
        # metrics reduced across ranks: (2+3)/7 and (1+1)/7 over 3+4 queries (the reference never reduces)
        np.testing.assert_allclose(r["red"], [5.0 / 7.0, 2.0 / 7.0, 7.0])


def test_shard_offsets_rank_major():
    from denseretrievaltoolkits_b200.store import shard_offsets

    assert shard_offsets([3, 0, 5]) == [0, 3, 3, 8]


def test_rebalance_plan_is_speed_proportional_bounded_and_conserves_rows():
    from denseretrievaltoolkits_b200.store import plan_rebalance

    counts = [1_100_000] * 8
    times = [8.0, 8.1, 8.4, 8.0, 8.2, 8.05, 8.3, 8.0]           # ms per shard search
    new = plan_rebalance(counts, times, max_shift=0.10, granularity=256)
    assert sum(new) == sum(counts) and all(c % 256 == 0 for c in new[:-1])
    assert new[2] < new[0] and new[6] < new[3]                     # slower GPUs get fewer rows
    assert all(abs(n - c) <= 0.10 * c + 256 * 8 for n, c in zip(new, counts))
    est = [t * n / c for t, n, c in zip(times, new, counts)]
    assert max(est) - min(est) < 0.02 * max(est)                   # predicted times equalised
    assert plan_rebalance([10, 0, 5], [1.0, 1.0, 1.0]) == [10, 0, 5]      # degenerate: unchanged
    assert plan_rebalance([1000, 1000], [1.0, 100.0], max_shift=0.25, granularity=1) == [1250, 750]
