"""torchrun -N ranks: stress of the peer-memory exchange (DRT_B200_PEER_EXCHANGE=1) — many searches
with changing (Q, k) so the symmetric buffer is re-used and re-grown, every result compared bit for
bit with a single full index.  Run: torchrun --nproc-per-node N tools/peer_stress.py [iters]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DRT_B200_PEER_EXCHANGE"] = "1"
import numpy as np
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from denseretrievaltoolkits_b200 import faiss_compat
from denseretrievaltoolkits_b200.store import ShardedCorpusStore

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, d = 200_000, 256
g = torch.Generator(device=dev).manual_seed(11)
x = torch.randn(n, d, device=dev, generator=g)
full = faiss_compat.IndexFlatIP(d, device=lr, seg_rows=1 << 15)
full.add(x)
per = -(-n // world)
store = ShardedCorpusStore(d, device=lr, seg_rows=1 << 14)
store.add(x[rank * per:(rank + 1) * per])
store.finalize()
rng = np.random.default_rng(3)            # same sequence on every rank
bad, used_peer, requeried = 0, 0, 0
for it in range(iters):
    Q = int(rng.integers(1, 3000))
    k = int(rng.choice([1, 10, 100, 200, 1000]))
    q = torch.randn(Q, d, device=dev, generator=g)
    dist.broadcast(q, 0)                   # identical queries on every rank
    D, I = store.search(q, k)
    Df, If = full.search(q, k)
    used_peer += int(store._peer not in (None, False))
    requeried += store.last_search["requeried"]
    if not (torch.equal(I, If) and torch.equal(D, Df)):
        bad += 1
res = dict(rank=rank, world=world, iters=iters, mismatches=bad, used_peer=used_peer, requeried=requeried, ok=(bad == 0 and used_peer == iters))
sys.stdout.write(json.dumps(res) + "\n")
sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
