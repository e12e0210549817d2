"""Trainer-shaped evaluation calls on the full corpus (8.8M x 768): `eval_batch_size` queries per
loader step (16 in the reference's run.sh:30, 128 by default, DRT/arguments.py:189), encoder
outputs on the device, ranked ids needed on the host (trainer.py:298-311).  Per-step searches vs
`DeferredSearch`.  python tools/deferred_bench.py [eval_batch_size ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from denseretrievaltoolkits_b200 import faiss_compat
from denseretrievaltoolkits_b200.deferred import DeferredSearch

dev = torch.device("cuda", 0)
n, k = bench.HEADLINE["n"], 100
index = faiss_compat.IndexFlatIP(bench.DIM, device=0)
bench.fill_rows(torch, index.add, 0, n, dev)
for ebs in [int(a) for a in sys.argv[1:]] or [16, 128]:
    nsteps = 4096 // ebs
    q = bench.make_queries(torch, nsteps * ebs, dev)
    steps = [q[i * ebs:(i + 1) * ebs] for i in range(nsteps)]

    def per_step(m):
        out = []
        for s in steps[:m]:
            _, I = index.search(s, k)
            out.append(I.cpu().numpy())          # the trainer maps ids to documents on the host
        return out

    def deferred(max_queries):
        ds = DeferredSearch(index, k, max_queries=max_queries)
        return [I.cpu().numpy() for _, (D, I) in ds.results(enumerate(steps))]

    m = min(nsteps, 64)
    per_step(4); deferred(4096)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a = per_step(m); torch.cuda.synchronize(); t_ps = time.perf_counter() - t0
    res = dict(eval_batch_size=ebs, n=n, k=k, per_step_qps=round(m * ebs / t_ps), per_step_ms_per_call=round(t_ps / m * 1e3, 3))
    for mq in (1024, 4096):
        torch.cuda.synchronize(); t0 = time.perf_counter(); b = deferred(mq); torch.cuda.synchronize(); t_d = time.perf_counter() - t0
        res[f"deferred_{mq}_qps"] = round(nsteps * ebs / t_d)
        res[f"identical_{mq}"] = all((x == y).all() for x, y in zip(a, b))
    print(json.dumps(res), flush=True)
