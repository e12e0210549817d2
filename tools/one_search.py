"""Dev: build the 8.8M store and run a few searches (target for ncu launch lists / captures).
  python tools/one_search.py K NQ [REPS] [CTAS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from denseretrievaltoolkits_b200 import _lib, faiss_compat
k, nq = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = int(os.environ.get("SWEEP_N", 8_800_000))
dev = torch.device("cuda", 0)
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
bench.fill_rows(torch, index.add, 0, n, dev)
q = bench.make_queries(torch, nq, dev)
flags = {0: 0, 1: _lib.SEARCH_FORCE_1CTA, 2: _lib.SEARCH_FORCE_2CTA}[ctas]
for _ in range(reps):
    index.search(q, k, flags=flags)
torch.cuda.synchronize()
print(index.search_stats())
