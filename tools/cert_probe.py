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
fam = sys.argv[4] if len(sys.argv) > 4 else "gauss"          # gauss | lognormal (row norms spread 100x) | huge (1 in 5000 rows x1000)
dev = torch.device("cuda", 0)
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)


def add_rows(rows):
    if fam == "lognormal":
        g = torch.Generator(device=dev).manual_seed(int(rows[0, 0].item() * 1e6) % (1 << 30))
        rows = rows * torch.exp(torch.randn((rows.shape[0], 1), generator=g, device=dev) * 1.15)
    elif fam == "huge":
        rows = rows.clone()
        rows[::5000] *= 1000.0
    index.add(rows)


bench.fill_rows(torch, add_rows, 0, n, dev)
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
print(json.dumps(dict(fam=fam, n=n, nq=nq, k=k, ms=round(ms, 3), rho=os.environ.get("DRT_B200_KPRIME_RHO"), kprime=st["kprime"],
                      refined=st["refined_queries"], flagged=st["flagged_queries"], exact=st["exact_queries"],
                      rescored_frac=round(st["rescored_rows"] / float(nq * st["kprime"]), 3), launches=st["launches"])))
