"""Runs the five BASELINE.json configs on the GPUs it is launched on and prints one JSON line per
config (rank 0).  cfg1/cfg3 run on rank 0 only; cfg2/cfg4/cfg5 shard the corpus over all ranks.

  python tools/run_configs.py [cfg ...]                                   # 1 GPU
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_configs.py [cfg ...]

Parity at scale: a fixed query subset is checked against an independent GPU fp32
reference (torch.matmul with TF32 off + topk over the same seeded chunks).
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench
from denseretrievaltoolkits_b200 import faiss_compat
from denseretrievaltoolkits_b200.store import ShardedCorpusStore

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False


def out(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def build_store(n):
    per = -(-n // world)
    r0, r1 = min(n, rank * per), min(n, (rank + 1) * per)
    store = ShardedCorpusStore(bench.DIM, device=local_rank) if world > 1 else None
    index = store.shards[0] if store else faiss_compat.IndexFlatIP(bench.DIM, device=local_rank)
    bench.fill_rows(torch, (store.add if store else index.add), r0, r1, dev)
    if store:
        store.finalize()
    barrier()
    return store, index


def search(store, index, q, k):
    return store.search(q, k) if store else index.search(q, k)


def torch_reference(q, n, k):
    best_d = torch.full((q.shape[0], 0), 0.0, device=dev)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=dev)
    for c in range(0, n, bench.CHUNK):
        rows = bench.make_corpus_chunk(torch, c // bench.CHUNK, dev)[: min(bench.CHUNK, n - c)]
        d, i = torch.topk(q @ rows.t(), min(k, rows.shape[0]), dim=1)
        best_d, best_i = torch.cat([best_d, d], 1), torch.cat([best_i, i + c], 1)
        d2, sel = torch.topk(best_d, min(k, best_d.shape[1]), dim=1)
        best_d, best_i = d2, torch.gather(best_i, 1, sel)
        del rows
    return best_d, best_i


def timed_search(store, index, q, k, reps):
    for _ in range(2):
        search(store, index, q, k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        D, I = search(store, index, q, k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, D, I


def parity(D, I, q, n, k, nsub=32):
    Dr, Ir = torch_reference(q[:nsub], n, k)
    got = I[:nsub].cpu()
    recall = float(np.mean([len(set(Ir[r].tolist()) & set(got[r].tolist())) / k for r in range(nsub)]))
    same = float((I[:nsub] == Ir).float().mean())
    rel = float(((D[:nsub] - Dr).abs() / Dr.abs().clamp_min(1e-6)).max())
    return dict(recall_at_k=recall, identical_ids=same, max_rel_score_err=rel, subset=nsub)


def cfg1():
    """1k queries x 100k x 768, k=100 — the reference's own CPU-runnable case; full parity vs
    the numpy oracle."""
    if rank != 0:
        return
    from oracle import flat_ip

    rng = np.random.default_rng(0)
    x = rng.standard_normal((100_000, 768), dtype=np.float32)
    q = rng.standard_normal((1000, 768), dtype=np.float32)
    idx = faiss_compat.IndexFlatIP(768, device=local_rank)
    idx.add(x)
    idx.search(q, 100)
    t0 = time.perf_counter()
    D, I = idx.search(q, 100)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    Dr, Ir = flat_ip.torch_flat_ip_search(torch.from_numpy(x), torch.from_numpy(q), 100)
    cpu_dt = time.perf_counter() - t1
    Do, Io = flat_ip.flat_ip_search(x, q, 100)
    recall = float(np.mean([len(set(a) & set(b)) / 100 for a, b in zip(I, Io)]))
    out(config="cfg1", nq=1000, n=100_000, k=100, host_api_ms=dt * 1e3, qps=1000 / dt, cpu_torch_ms=cpu_dt * 1e3,
        cpu_qps=1000 / cpu_dt, cpu_threads=torch.get_num_threads(), recall_at_k=recall, identical_ids=float((I == Io).mean()),
        max_rel_score_err=float(np.max(np.abs(D - Do) / np.abs(Do))))


def cfg2(store, index):
    q = bench.make_queries(torch, 6980, dev)
    ms, D, I = timed_search(store, index, q, 1000, 3)
    out(config="cfg2", nq=6980, n=8_800_000, k=1000, n_gpus=world, ms=ms, qps=6980 / ms * 1e3, stats=index.search_stats(),
        **parity(D, I, q, 8_800_000, 1000, 16))


def cfg3():
    if rank != 0:
        return
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

    B, n = 128, 8
    x = torch.randn(B, 768, device=dev, requires_grad=True)
    y = torch.randn(B * n, 768, device=dev, requires_grad=True)
    tgt = torch.arange(0, B * n, n, device=dev)
    ours = SimpleContrastiveLoss()

    def run(fn, reps=300):
        for _ in range(30):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    def fb_ours():
        x.grad = None; y.grad = None
        ours(x, y).backward()

    def fb_ref():
        x.grad = None; y.grad = None
        torch.nn.functional.cross_entropy(x @ y.t(), tgt).backward()

    lo, lr = ours(x, y).item(), torch.nn.functional.cross_entropy(x @ y.t(), tgt).item()
    out(config="cfg3", B=B, P=B * n, d=768, fwdbwd_us_ours=run(fb_ours), fwdbwd_us_torch_eager=run(fb_ref),
        loss_ours=lo, loss_torch=lr, rel_err=abs(lo - lr) / abs(lr))


def cfg4():
    n = 21_000_000
    store, index = build_store(n)
    q = bench.make_queries(torch, 3600, dev)
    ms, D, I = timed_search(store, index, q, 100, 3)
    out(config="cfg4", nq=3600, n=n, k=100, n_gpus=world, ms=ms, qps=3600 / ms * 1e3, stats=index.search_stats(),
        **parity(D, I, q, n, 100, 16))
    del store, index
    torch.cuda.empty_cache()


def cfg5(store, index):
    """Hard-negative mining: 500k queries x 8.8M, top-200 + positive exclusion, streamed in
    batches of 16,384 queries."""
    from denseretrievaltoolkits_b200.mining import filter_negatives

    nq_total, k, batch = int(os.environ.get("CFG5_NQ", 500_000)), 200, 16384
    g = torch.Generator(device=dev).manual_seed(99)
    barrier()
    t0 = time.perf_counter()
    done = 0
    while done < nq_total:
        nb = min(batch, nq_total - done)
        q = torch.randn((nb, bench.DIM), generator=g, device=dev)
        D, I = search(store, index, q, k)
        pb = torch.randint(0, 8_800_000 - 8, (nb,), device=dev, generator=g)
        neg = filter_negatives(I, pb, pb + 4, 196)
        done += nb
    barrier()
    dt = time.perf_counter() - t0
    out(config="cfg5", nq=nq_total, n=8_800_000, k=k, n_gpus=world, seconds=dt, qps=nq_total / dt,
        last_batch_unfilled=int((neg < 0).sum().item()), stats=index.search_stats())


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    if "cfg1" in which:
        cfg1()
    if "cfg3" in which:
        cfg3()
    if "cfg2" in which or "cfg5" in which:
        store, index = build_store(8_800_000)
        if "cfg2" in which:
            cfg2(store, index)
        if "cfg5" in which:
            cfg5(store, index)
        del store, index
        torch.cuda.empty_cache()
    if "cfg4" in which:
        cfg4()
    if world > 1:
        dist.destroy_process_group()
