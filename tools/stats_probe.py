"""Dev probe: search stats (flags / retries) across k on the full corpus."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from denseretrievaltoolkits_b200 import _lib, faiss_compat
n = int(os.environ.get("SWEEP_N", 8_800_000))
dev = torch.device("cuda", 0)
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
bench.fill_rows(torch, index.add, 0, n, dev)
for k, nq in [(100, 6980), (200, 700), (200, 4096), (1000, 512), (1000, 6980), (10, 1024), (2048, 256)]:
    q = bench.make_queries(torch, nq, dev)
    for _ in range(2):
        index.search(q, k, flags=_lib.SEARCH_TIME_KERNELS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    index.search(q, k, flags=_lib.SEARCH_TIME_KERNELS)
    e1.record(); torch.cuda.synchronize()
    st = index.search_stats()
    ms = e0.elapsed_time(e1)
    print(json.dumps(dict(k=k, nq=nq, ms=round(ms, 2), qps=round(nq / ms * 1e3), filter_ms=round(st["filter_ns"] / 1e6, 2), stats=st)), flush=True)
