#!/usr/bin/env python
"""bench.py — exact top-k MIPS throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

A "step" is one pass of the hot path over one query batch: every query of the batch against the
whole corpus, top-k out.  Workload at N = 1: the headline shape of BASELINE.json / SURVEY.md
§8d — 6,980 queries x 8.8M x 768 corpus, k = 100 (cfg2's shape at the metric's depth).  At
N > 1 the same corpus is row-sharded over the ranks ("scaling": "strong"): every rank scans its
shard for all queries (to the reduced per-shard depth, `config.shard_depth`), then the candidate
lists are exchanged and merged by one kernel over peer-mapped memory (`config.exchange`; NCCL
all-gather + merge kernel when DRT_B200_PEER_EXCHANGE=0).
Synthetic data: corpus rows iid N(0,1) generated on the device per 2^20-row chunk with seed
1234 + chunk index (identical corpus for every N), queries N(0,1) seed 4321.  The corpus
(13.5 GB bf16 streamed per pass) is far larger than L2, so no explicit L2 flush is needed.

Prints ONE JSON line (rank 0).  `value` = queries/s with inputs resident in HBM; `e2e` = the
same through the host-buffer API (numpy in / numpy out, H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU path (real faiss.IndexFlatIP when importable, else
its stated equivalent: blocked fp32 torch.mm + topk on all host threads) over the FULL 8.8M-row
corpus held in host memory, a bounded number of queries per step, q/s scaled linearly in Q.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 768
HEADLINE = dict(nq=6980, n=8_800_000, k=100, name="headline: 6980 queries x 8.8Mx768 corpus, exact top-100 (cfg2 shape at the metric's k)")
CHUNK = 1 << 20
METRIC = "queries/sec, exact top-100 MIPS, 8.8Mx768 corpus"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"), hbm=p["hbm_gbs"],
                    source="MEASURED_PEAKS.json (measured on this pool: burst figure; the sustained one is peak_sustained)")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def mark(self):
        """Start of the timed region: earlier samples (warm-up, sampler start-up) are dropped."""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        t1 = time.time()
        for ts, r in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def make_corpus_chunk(torch, chunk_index: int, device):
    g = torch.Generator(device=device)
    g.manual_seed(1234 + chunk_index)
    return torch.randn((CHUNK, DIM), generator=g, device=device, dtype=torch.float32)


def make_queries(torch, nq: int, device):
    g = torch.Generator(device=device)
    g.manual_seed(4321)
    return torch.randn((nq, DIM), generator=g, device=device, dtype=torch.float32)


def fill_rows(torch, add_fn, row0: int, row1: int, device):
    """Add global corpus rows [row0,row1) (chunk-seeded, so any sharding sees the same corpus)."""
    c = row0 // CHUNK
    while c * CHUNK < row1:
        lo, hi = max(row0, c * CHUNK), min(row1, (c + 1) * CHUNK)
        chunk = make_corpus_chunk(torch, c, device)
        add_fn(chunk[lo - c * CHUNK: hi - c * CHUNK])
        del chunk
        c += 1


def parity_check(torch, dist, world, rank, device, row0, row1, q_dev, k, D, I, nsub=128):
    """Result check OUTSIDE the timed region, at every N: a fixed subset of the queries against an
    independent exact search — fp32 torch.matmul (TF32 off) + topk over this rank's rows,
    regenerated from the chunk seeds (not read back from the store), the per-rank top-k lists
    all-gathered and merged with torch.topk.  The global top-k must equal the single-index top-k
    (reference semantic: DRT/model/utils.py:215-229)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    sel = torch.linspace(0, q_dev.shape[0] - 1, min(nsub, q_dev.shape[0]), device=device).long()
    qs = q_dev.index_select(0, sel)
    bd = torch.full((qs.shape[0], k), float("-inf"), device=device)
    bi = torch.full((qs.shape[0], k), -1, dtype=torch.int64, device=device)
    c = row0 // CHUNK
    while c * CHUNK < row1:
        lo, hi = max(row0, c * CHUNK), min(row1, (c + 1) * CHUNK)
        rows = make_corpus_chunk(torch, c, device)[lo - c * CHUNK: hi - c * CHUNK]
        kk = min(k, rows.shape[0])
        d, i = torch.topk(qs @ rows.t(), kk, dim=1)
        md, mi = torch.cat([bd, d], 1), torch.cat([bi, i + lo], 1)
        bd, o = torch.topk(md, k, dim=1)
        bi = torch.gather(mi, 1, o)
        del rows
        c += 1
    if world > 1:
        gd = torch.empty((world,) + tuple(bd.shape), device=device)
        gi = torch.empty((world,) + tuple(bi.shape), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gd, bd.contiguous())
        dist.all_gather_into_tensor(gi, bi.contiguous())
        md, mi = gd.permute(1, 0, 2).reshape(qs.shape[0], -1), gi.permute(1, 0, 2).reshape(qs.shape[0], -1)
        bd, o = torch.topk(md, k, dim=1)
        bi = torch.gather(mi, 1, o)
    Ds, Is = D.index_select(0, sel), I.index_select(0, sel)
    got, ref = Is.cpu().tolist(), bi.cpu().tolist()
    recall = sum(len(set(a) & set(b)) for a, b in zip(got, ref)) / float(len(got) * k)
    return {"queries": len(got), "recall": recall, "identical_ids": float((Is == bi).float().mean().item()),
            "max_rel_err": float(((Ds - bd).abs() / bd.abs().clamp_min(1e-6)).max().item()),
            "reference": "fp32 torch.matmul (TF32 off) + topk per shard over re-generated rows, all-gathered and merged"}


def cpu_search_fn(torch):
    """The reference's CPU search path: real faiss.IndexFlatIP when `import faiss` works on this
    box, else its stated equivalent — blocked fp32 torch.mm (MKL, all host
    threads) + torch.topk (oracle/flat_ip.py::torch_flat_ip_search).  Returns (kind, label, build, search)."""
    try:
        import faiss  # noqa: F401

        if not hasattr(faiss, "omp_get_max_threads"):
            raise ImportError("not the real faiss")

        def build(corpus_t):
            idx = faiss.IndexFlatIP(corpus_t.shape[1])
            idx.add(corpus_t.numpy())
            return idx

        return "faiss", f"faiss.IndexFlatIP {getattr(faiss, '__version__', '?')}", build, (lambda idx, q_t, k: idx.search(q_t.numpy(), k))
    except Exception:
        from oracle import flat_ip

        return ("port", "faiss absent -> blocked fp32 torch.mm (MKL) + torch.topk",
                (lambda corpus_t: corpus_t), (lambda corpus_t, q_t, k: flat_ip.torch_flat_ip_search(corpus_t, q_t, k)))


def host_threads(torch):
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


def host_corpus(torch, n, threads):
    """The full synthetic corpus in host memory (n x 768 fp32 = 27 GB at n = 8.8M), chunk-seeded
    N(0,1) rows generated by a pool of host threads."""
    from concurrent.futures import ThreadPoolExecutor

    corpus = torch.empty((n, DIM), dtype=torch.float32)

    def fill(c):
        lo, hi = c * CHUNK, min(n, (c + 1) * CHUNK)
        corpus[lo:hi].normal_(generator=torch.Generator().manual_seed(1234 + c))

    torch.set_num_threads(1)
    with ThreadPoolExecutor(max_workers=max(1, min(threads, 16))) as ex:
        list(ex.map(fill, range(-(-n // CHUNK))))
    torch.set_num_threads(threads)
    return corpus


def host_mem_available():
    try:
        import psutil

        return psutil.virtual_memory().available
    except Exception:
        return 0


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = HEADLINE
    cores = host_threads(torch)
    kind, label, build, search = cpu_search_fn(torch)
    n_full = cfg["n"]
    full = host_mem_available() > n_full * DIM * 4 + (12 << 30)
    n_host = n_full if full else 1 << 20           # not enough host memory for 27 GB: a 2^20-row sample, scaled in N as well
    t0 = time.perf_counter()
    corpus = host_corpus(torch, n_host, cores)
    index = build(corpus)
    gen_s = time.perf_counter() - t0
    q_all = torch.randn((256, DIM), generator=torch.Generator().manual_seed(4321), dtype=torch.float32)
    # queries per step: the largest power of two <= 256 that keeps a step under ~6 s on this host
    search(index, q_all[:16], cfg["k"])              # first call warms the BLAS threads
    t0 = time.perf_counter()
    search(index, q_all[:16], cfg["k"])
    t16 = time.perf_counter() - t0
    qs = 256
    while qs > 16 and t16 * qs / 16.0 > 6.0:
        qs //= 2
    q = q_all[:qs].contiguous()
    for _ in range(args.warmup):
        search(index, q, cfg["k"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        search(index, q, cfg["k"])
    total = time.perf_counter() - t0
    value = (qs * args.steps / total) * (n_host / float(n_full))
    sample = (f"{qs} of the {cfg['nq']} queries x {'ALL' if full else 'the first'} {n_host} corpus rows per step, "
              f"q/s scaled linearly in Q{'' if full else ' and N'}; corpus generated on the host in {gen_s:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "nq": cfg["nq"], "n": cfg["n"], "dim": DIM, "k": cfg["k"]},
        "reference_impl": f"reference CPU path: {label}",
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nq", type=int, default=HEADLINE["nq"])
    ap.add_argument("--n", type=int, default=HEADLINE["n"])
    ap.add_argument("--k", type=int, default=HEADLINE["k"])
    ap.add_argument("--ctas", type=int, default=0, help="force the 1- or 2-CTA tile variant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from denseretrievaltoolkits_b200 import _lib, faiss_compat
    from denseretrievaltoolkits_b200.store import ShardedCorpusStore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    nq, n, k = args.nq, args.n, args.k
    first_pass = "f16" if os.environ.get("DRT_B200_FIRST_PASS") == "f16" else "bf16"
    flags = _lib.SEARCH_TIME_KERNELS
    if args.ctas == 1:
        flags |= _lib.SEARCH_FORCE_1CTA
    elif args.ctas == 2:
        flags |= _lib.SEARCH_FORCE_2CTA

    # ---- build the (sharded) device-resident store ----
    per = -(-n // world)
    row0, row1 = min(n, rank * per), min(n, (rank + 1) * per)
    t_build0 = time.perf_counter()
    if world > 1:
        store = ShardedCorpusStore(DIM, device=local_rank)
        index = store.shards[0]
        fill_rows(torch, store.add, row0, row1, device)
        store.finalize()
    else:
        store = None
        index = faiss_compat.IndexFlatIP(DIM, device=local_rank)
        fill_rows(torch, index.add, row0, row1, device)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build0
    q_dev = make_queries(torch, nq, device)
    shard_rows = None
    if store is not None and os.environ.get("DRT_B200_REBALANCE", "0") == "1":
        # optional (off by default: measured at N=2 the per-GPU speed differences under the power cap
        # drift between the calibration and the run, 35.50 -> 35.36 ms per step, within noise):
        # rows per rank proportional to the measured speed of the rank's GPU; global row ids unchanged
        shard_rows = store.rebalance(q_dev, k)
        row0, row1 = store._offsets[rank], store._offsets[rank + 1]
    q_host = torch.empty((nq, DIM), dtype=torch.float32).pin_memory()
    q_host.copy_(q_dev)
    q_np = q_host.numpy()

    def step_device():
        if store is not None:
            # shard search + exchange + merge; every rank keeps the merged rows of its own slice of
            # the queries on its device (rank-local results, as in the e2e leg)
            return store.search(q_dev, k, flags=flags, local_results=True)
        return index.search(q_dev, k, flags=flags)

    def step_host():
        if store is not None:
            # every rank uploads its slice of the queries and downloads the result rows of that
            # slice (the reference evaluates per rank, DRT/trainer/trainer.py:287-297)
            return store.search(q_np, k, local_results=True)
        return index.search(q_np, k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        filt_ns, launches = 0, 0
        for _ in range(steps):
            fn()
            st = index.search_stats()
            filt_ns += st["filter_ns"]
            launches += st["launches"] + (1 if store is not None else 0)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, filt_ns, launches, index.search_stats()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # started before the warm-up so it is sampling by the timed region
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler.mark()
    ms, filt_ns, launches, stats = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = nq * args.steps / (ms / 1e3)

    step_host()
    ms_e2e, _, _, _ = timed(step_host, args.steps)
    e2e_value = nq * args.steps / (ms_e2e / 1e3)

    # ---- parity, outside the timed region: the (merged) result against an independent search ----
    Dm, Im = store.search(q_dev, k, flags=flags) if store is not None else step_device()     # full result on every rank
    last = dict(store.last_search) if store is not None else {}
    st_now = index.search_stats()
    cert = torch.tensor([st_now["flagged_queries"], st_now["exact_queries"], st_now["refined_queries"]],
                        dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(cert)                       # summed over the shards
    parity = parity_check(torch, dist, world, rank, device, row0, row1, q_dev, k, Dm, Im)
    parity.update(requeried=int(last.get("requeried", 0)), flagged=int(cert[0]), exact_queries=int(cert[1]),
                  refined_queries=int(cert[2]))
    if parity["recall"] < 0.999:
        raise SystemExit(f"parity check failed: {parity}")
    phases = None
    if store is not None:
        ph = store.profile_phases(q_dev, k)
        if ph is not None:       # max / min over ranks of the two phases (untimed extra steps)
            t = torch.tensor([ph["local_search_ms"], ph["exchange_merge_ms"], -ph["local_search_ms"]], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            phases = {"local_search_ms_max": float(t[0]), "local_search_ms_min": float(-t[2]), "exchange_merge_ms_max": float(t[1]),
                      "nvlink_bytes_read_per_rank": ph["nvlink_bytes_read"], "nvlink_bytes_written_per_rank": ph["nvlink_bytes_written"],
                      "note": "CUDA events per rank over 5 untimed steps; exchange_merge includes waiting for the slowest rank"}

    # ---- roofline of the dominant kernel (the tcgen05 filter), timed live with CUDA events ----
    pk = peaks()
    n_local = row1 - row0
    flops_per_step = 2.0 * nq * n_local * DIM                       # this rank's share, counted once
    filt_s = filt_ns / 1e9 / max(1, args.steps)
    achieved = flops_per_step / filt_s / 1e12 if filt_s > 0 else 0.0
    ridge_q = pk["bf16"] * 1e12 * 2 / (2 * pk["hbm"] * 1e9)          # bf16 plane: bytes = 2/elem
    roof = {"bound": "tensor" if nq >= ridge_q else "hbm", "achieved": achieved, "peak": pk["bf16"],
            "unit": "TFLOP/s", "frac": achieved / pk["bf16"], "traffic": None,
            "kernel": "mips_filter_kernel", "kernel_ms_per_step": filt_s * 1e3,
            "kernel_share_of_step": filt_s * 1e3 / (ms / args.steps),
            "launches_per_step": stats["filter_launches"], "peak_source": pk["source"],
            "peak_sustained": pk["bf16_sustained"],
            "frac_sustained": achieved / pk["bf16_sustained"] if pk["bf16_sustained"] else None}
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):   # dram bytes per corpus row from the committed single-pass ncu metric run
        roof["traffic"] = json.load(open(tpath))["dram_bytes_per_corpus_row"] * n_local
        roof["traffic_unit"] = "bytes per step (sum over the step's K1 launches; ncu dram read+write per row x rows)"
        roof["traffic_source"] = ("constant from a committed single-pass ncu dram-bytes run over one headline search (profiles/k1_traffic.json) "
                                  "scaled by this rank's rows; NOT measured in this run")
        roof["algorithmic_bytes"] = n_local * DIM * 2 + nq * DIM * 2
    if roof["bound"] == "hbm":
        gbs = n_local * DIM * 2 / filt_s / 1e9 if filt_s > 0 else 0.0
        roof.update(achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path on this box's host cores, bounded sample: the store's whole fp32
        # plane copied back to the host when memory allows (same N; q/s scales in Q only), else
        # its first 2^20 rows (scaled in N as well)
        cores = host_threads(torch)
        kind, label, build, search = cpu_search_fn(torch)
        full = host_mem_available() > n * DIM * 4 + (12 << 30)
        ns = n if full else min(n, 1 << 20)
        corpus_s = torch.empty((ns, DIM), dtype=torch.float32)
        for r0 in range(0, ns, 1 << 19):
            m = min(1 << 19, ns - r0)
            corpus_s[r0:r0 + m] = torch.from_numpy(index.reconstruct_n(r0, m))
        ref_index = build(corpus_s)
        qs = min(nq, 64 if full else 512)
        q_s = q_host[:qs].clone()
        search(ref_index, q_s[:8], k)                                    # warm the BLAS threads
        dts, reps = 0.0, 0
        while dts < 10.0 and reps < 8:
            t0 = time.perf_counter()
            search(ref_index, q_s, k)
            dts += time.perf_counter() - t0; reps += 1
        cpu = {"value": (qs * reps / dts) * (ns / float(n)), "unit": "queries/s", "cores": cores, "kind": kind,
               "impl": label,
               "sample": f"{qs} queries x {'all' if full else 'first'} {ns} corpus rows x {reps} reps ({dts:.1f} s), "
                         f"q/s scaled linearly in Q{'' if full else ' and N'}"}
        del corpus_s, ref_index

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": first_pass,
            "dtype_detail": f"{first_pass} tcgen05 first pass (kind::f16, fp32 accumulate; same MMA rate as the bf16 peak it is "
                            "measured against) + f32 exact rescoring of the k' candidates, every query certified exact",
            "data": "synthetic",
            "config": {"workload": HEADLINE["name"] if (nq, n, k) == (HEADLINE["nq"], HEADLINE["n"], HEADLINE["k"])
                       else f"custom: {nq} queries x {n}x{DIM}, k={k}",
                       "nq": nq, "n": n, "dim": DIM, "k": k,
                       "results": "every rank keeps the merged rows of its slice of the queries" if store is not None else "one device",
                       "sharding": f"rows/{world}" + ("" if not shard_rows else f", contiguous shards sized by measured GPU speed: {shard_rows}"),
                       "shard_depth": store.last_search.get("local_depth") if store is not None else k,
                       "exchange": (None if store is None else
                                    "one kernel over peer-mapped memory (drt_merge_topk_peers)" if store._peer not in (None, False)
                                    else "NCCL all-gather + merge kernel"),
                       "l2": "inputs larger than L2 (13.5 GB bf16 corpus streamed per step); no flush",
                       "ctas_per_tile": stats["ctas_per_tile"], "kprime": stats["kprime"],
                       "corpus_chunks": stats["chunks"], "store_build_s": build_s},
            "roofline": roof,
            "parity": parity,
            "multi_gpu_phases": phases,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "queries/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * k * 12,
                    "note": "whole-job bytes: each query row crosses PCIe once (ranks upload 1/N slices and "
                            "all-gather them over NVLink when N>1) and each merged result row is downloaded once, "
                            "by the rank that owns that query slice (rank-local results)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "search_stats": stats,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
