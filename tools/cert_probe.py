"""Dev probe: certificate statistics (k', refined / flagged queries, rescored rows) and step time
for one search shape.  python tools/cert_probe.py [n] [nq] [k]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from denseretrievaltoolkits_b200 import faiss_compat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
bench.fill_rows(torch, index.add, 0, n, dev)
q = bench.make_queries(torch, nq, dev)
for _ in range(3):
    D, I = index.search(q, k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    D, I = index.search(q, k)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
st = index.search_stats()
print(json.dumps(dict(n=n, nq=nq, k=k, ms=round(ms, 3), rho=os.environ.get("DRT_B200_KPRIME_RHO"), kprime=st["kprime"],
                      refined=st["refined_queries"], flagged=st["flagged_queries"], exact=st["exact_queries"],
                      rescored_frac=round(st["rescored_rows"] / float(nq * st["kprime"]), 3), launches=st["launches"])))
