"""Builds one 8.8M x 768 store and times search variants (dev tool; numbers go to gpurun_out/)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from denseretrievaltoolkits_b200 import _lib, faiss_compat

n = int(os.environ.get("SWEEP_N", 8_800_000))
ks = [int(v) for v in os.environ.get("SWEEP_K", "100").split(",")]
nqs = [int(v) for v in os.environ.get("SWEEP_NQ", "128,1024,6980").split(",")]
ctas_list = [int(v) for v in os.environ.get("SWEEP_CTAS", "1,2").split(",")]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
t0 = time.time()
bench.fill_rows(torch, index.add, 0, n, dev)
torch.cuda.synchronize()
print(f"store build {time.time() - t0:.1f}s ntotal={index.ntotal}", flush=True)
for k in ks:
    for nq in nqs:
        q = bench.make_queries(torch, nq, dev)
        for ctas in ctas_list:
            flags = _lib.SEARCH_TIME_KERNELS | (_lib.SEARCH_FORCE_2CTA if ctas == 2 else _lib.SEARCH_FORCE_1CTA)
            for _ in range(2):
                index.search(q, k, flags=flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = int(os.environ.get("SWEEP_REPS", 3))
            e0.record()
            fns = 0
            for _ in range(reps):
                index.search(q, k, flags=flags)
                fns += index.search_stats()["filter_ns"]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fms = fns / 1e6 / reps
            tf = 2.0 * nq * n * bench.DIM / (fms / 1e3) / 1e12
            pk = bench.peaks()
            roof_ms = max(n * bench.DIM * 2 / (pk["hbm"] * 1e9), 2.0 * nq * n * bench.DIM / (pk["bf16"] * 1e12)) * 1e3
            print(json.dumps(dict(k=k, nq=nq, ctas=ctas, ms=round(ms, 3), qps=round(nq / ms * 1e3, 1), filter_ms=round(fms, 3),
                                  filter_tflops=round(tf, 1), roofline_ms=round(roof_ms, 3),
                                  frac_of_roofline_step=round(roof_ms / ms, 3), frac_of_roofline_k1=round(roof_ms / fms, 3), stats=index.search_stats())), flush=True)
