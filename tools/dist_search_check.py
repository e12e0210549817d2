"""torchrun -N ranks: ShardedCorpusStore (NCCL candidate exchange + merge kernel) against a single
full index built on every rank: ids bit for bit, scores bit for bit, both exchange modes.
Run: torchrun --nproc-per-node N tools/dist_search_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from denseretrievaltoolkits_b200 import faiss_compat
from denseretrievaltoolkits_b200.store import ShardedCorpusStore

n, d, nq = 300_000, 768, 1000
g = torch.Generator(device=dev).manual_seed(5)          # same corpus on every rank
x = torch.randn(n, d, device=dev, generator=g)
x[-500:] = x[:500]                                        # ties across shards
q = torch.randn(nq, d, device=dev, generator=g)
full = faiss_compat.IndexFlatIP(d, device=lr, seg_rows=1 << 16)
full.add(x)
bounds = np.linspace(0, n, world + 1).astype(int)
bounds[1:-1] += np.arange(1, world) * 37                  # uneven shards
os.environ["DRT_B200_PEER_EXCHANGE"] = "0"                 # this store: the NCCL exchange modes
store = ShardedCorpusStore(d, device=lr, seg_rows=1 << 15)
store.add(x[bounds[rank]:bounds[rank + 1]])
offs = store.finalize()
ok = offs == [int(b) for b in bounds]
res = {"rank": rank, "world": world, "offsets_ok": ok}
for k in (100, 1000):
    Df, If = full.search(q, k)
    for mode, thr in (("all_gather", 1 << 62), ("all_to_all", 0)):
        store.A2A_MIN_ENTRIES, store.A2A_MIN_WORLD = thr, 2
        D, I = store.search(q, k)
        same = bool(torch.equal(I, If) and torch.equal(D, Df))
        res[f"k{k}_{mode}"] = same
        ok = ok and same
    per = nq // world
    Dl, Il = store.search_local_queries(q[rank * per:(rank + 1) * per].contiguous(), k)
    same = bool(torch.equal(Il, If[rank * per:(rank + 1) * per]))
    res[f"k{k}_local_queries"] = same
    ok = ok and same
Dn, In = store.search(q.cpu().numpy(), 100)               # host-buffer API
Df, If = full.search(q, 100)
res["numpy_api"] = bool(np.array_equal(In, If.cpu().numpy()) and np.array_equal(Dn, Df.cpu().numpy()))
ok = ok and res["numpy_api"]
Dn, In = store.search(q.cpu().numpy()[:997], 100)         # ragged slice upload (997 % world != 0)
res["numpy_api_ragged"] = bool(np.array_equal(In, If.cpu().numpy()[:997]) and np.array_equal(Dn, Df.cpu().numpy()[:997]))
ok = ok and res["numpy_api_ragged"]
store.SLICE_UPLOAD_MIN_BYTES = 1 << 62                    # whole-batch upload on every rank
Dn, In = store.search(q.cpu().numpy(), 100)
res["numpy_api_full_upload"] = bool(np.array_equal(In, If.cpu().numpy()))
ok = ok and res["numpy_api_full_upload"]
# exchange + merge as one kernel over peer-mapped memory (torch symmetric memory)
try:
    os.environ["DRT_B200_PEER_EXCHANGE"] = "1"
    store.A2A_MIN_ENTRIES, store.A2A_MIN_WORLD, store.SLICE_UPLOAD_MIN_BYTES = 1 << 20, 4, 1 << 20
    peer = ShardedCorpusStore(d, device=lr, seg_rows=1 << 15)
    peer.add(x[bounds[rank]:bounds[rank + 1]])
    peer.finalize()
    for k in (100, 1000):
        Df, If = full.search(q, k)
        D, I = peer.search(q, k)
        same = bool(torch.equal(I, If) and torch.equal(D, Df))
        res[f"k{k}_peer_exchange"] = same and peer._peer not in (None, False)
        ok = ok and res[f"k{k}_peer_exchange"]
    Dn, In = peer.search(q.cpu().numpy()[:997], 100)
    res["peer_numpy_ragged"] = bool(np.array_equal(In, full.search(q, 100)[1].cpu().numpy()[:997]))
    ok = ok and res["peer_numpy_ragged"]
    # rank-local results: every rank keeps only its slice of the queries (device and host API)
    Df, If = full.search(q, 100)
    for st_, name in ((peer, "peer"), (store, "nccl")):
        q0, qn = st_.result_slice(997)
        Dl, Il = st_.search(q[:997], 100, local_results=True)
        Dh, Ih = st_.search(q.cpu().numpy()[:997], 100, local_results=True)
        same = bool(torch.equal(Il, If[q0:q0 + qn]) and torch.equal(Dl, Df[q0:q0 + qn]) and
                    np.array_equal(Ih, If[q0:q0 + qn].cpu().numpy()) and np.array_equal(Dh, Df[q0:q0 + qn].cpu().numpy()))
        res[f"{name}_local_results"] = same
        ok = ok and same
    per = nq // world
    Dl, Il = peer.search_local_queries(q[rank * per:(rank + 1) * per].contiguous(), 100)
    res["peer_local_queries"] = bool(torch.equal(Il, If[rank * per:(rank + 1) * per]))
    ok = ok and res["peer_local_queries"]
    res["peer_async_clean"] = peer.last_search.get("redone") == 0       # the asynchronous shard search was final
    # Trainer.evaluate with queries batched across loader steps (collective DeferredSearch): every
    # rank queues its OWN 16-query batches, one corpus pass per 128 queued queries per rank
    from denseretrievaltoolkits_b200.deferred import DeferredSearch
    mine = q[rank * per:(rank + 1) * per]
    steps = [mine[i:i + 16].contiguous() for i in range(0, min(320, per // 16 * 16), 16)]
    ds = DeferredSearch(peer, 100, max_queries=128)
    got = list(ds.results(enumerate(steps)))
    same = len(got) == len(steps) and ds.searches == -(-len(steps) * 16 // 128)
    for (tag, (Dd, Id)), st_ in zip(got, steps):
        lo = rank * per + tag * 16
        same = same and bool(torch.equal(Id, If[lo:lo + 16]) and torch.equal(Dd, Df[lo:lo + 16]))
    res["peer_deferred_evaluate"] = same
    ok = ok and same
    ok = ok and res["peer_async_clean"]
    # speed-proportional re-split of the shards: rows move between neighbouring ranks, global ids do not
    before = list(peer._offsets)
    new_counts = peer.rebalance(q, 100, reps=3, max_shift=0.25)
    Df, If = full.search(q, 100)
    Dp, Ip = peer.search(q, 100)
    res["peer_rebalance"] = bool(torch.equal(Ip, If) and torch.equal(Dp, Df)) and sum(new_counts) == n and peer._offsets[-1] == n
    res["rebalance_rows"] = [before, list(peer._offsets)]
    ok = ok and res["peer_rebalance"]
    # a corpus of near-duplicates (integer rows differing by sparse +1s): the bf16 first pass cannot
    # be certified, the asynchronous shard searches publish "not final", every rank sees the OR of
    # the status bytes and the step is repeated on the synchronous path (retry / refine / fp32 pass)
    gi = torch.Generator(device=dev).manual_seed(9)
    base = torch.randint(256, 768, (d,), generator=gi, device=dev).float()
    xi = base[None, :] + (torch.rand(40_000, d, generator=gi, device=dev) < 0.05).float()
    qi = torch.randint(0, 3, (64, d), generator=gi, device=dev).float()
    fulli = faiss_compat.IndexFlatIP(d, device=lr, seg_rows=1 << 14)
    fulli.add(xi)
    peeri = ShardedCorpusStore(d, device=lr, seg_rows=1 << 14)
    bi = np.linspace(0, 40_000, world + 1).astype(int)
    peeri.add(xi[bi[rank]:bi[rank + 1]])
    peeri.finalize()
    Dfi, Ifi = fulli.search(qi, 50)
    Dpi, Ipi = peeri.search(qi, 50)
    res["peer_redo_path"] = bool(torch.equal(Ipi, Ifi) and torch.equal(Dpi, Dfi)) and peeri.last_search.get("redone") == 1
    ok = ok and res["peer_redo_path"]

    def timeit(st, reps=30):
        for _ in range(5):
            st.search(q, 100)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            st.search(q, 100)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    res["us_nccl"], res["us_peer"] = round(timeit(store), 1), round(timeit(peer), 1)
except Exception as e:  # symmetric memory unavailable on this box
    res["peer_exchange_error"] = repr(e)[:300]
    ok = False
res["ok"] = ok
sys.stdout.write(json.dumps(res) + "\n")   # one write: lines of different ranks must not interleave
sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
