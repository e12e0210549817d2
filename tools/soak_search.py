"""Randomised soak of the search path against a float64 brute force (torch, on the GPU):
random sizes / dims / k / segment sizes / add splits and data families that stress the
certificate and the select paths.  python tools/soak_search.py [iterations] [seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from denseretrievaltoolkits_b200 import faiss_compat

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
dev = torch.device("cuda", 0)
bad = []
fams = ["gauss", "lognormal", "dups", "huge_rows", "neardup_int", "aniso", "tiny_scale", "mixed_sign_big"]
agg = dict(refined=0, flagged=0, exact=0, overflow=0)
t0 = time.time()
for it in range(iters):
    fam = fams[it % len(fams)]
    n = int(rng.integers(1, 200_000))
    d = int(rng.choice([64, 96, 128, 200, 768, 1000]))
    k = int(rng.integers(1, 260))
    nq = int(rng.integers(1, 300))
    seg_rows = int(rng.choice([256, 4096, 65536, 1 << 20]))
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn((n, d), generator=g, device=dev)
    q = torch.randn((nq, d), generator=g, device=dev)
    if fam == "lognormal":
        x *= torch.exp(torch.randn((n, 1), generator=g, device=dev) * 1.15)
    elif fam == "dups" and n > 10:
        m = max(1, n // 20)
        x[-m:] = x[:m]
    elif fam == "huge_rows":
        idx = torch.randint(0, n, (max(1, n // 3000),), generator=g, device=dev)
        x[idx] *= 1000.0
    elif fam == "neardup_int":
        base = torch.randint(200, 700, (d,), generator=g, device=dev).float()
        x = base[None, :] + (torch.rand((n, d), generator=g, device=dev) < 0.05).float()
        q = torch.randint(0, 3, (nq, d), generator=g, device=dev).float()
        n = min(n, 40_000); x = x[:n]; nq = min(nq, 24); q = q[:nq]
    elif fam == "aniso":
        sc = torch.exp(torch.rand((d,), generator=g, device=dev) * 2.77 - 1.386)
        x = x * sc + 0.3
        q = q * sc
    elif fam == "tiny_scale":
        x *= 1e-6; q *= 1e-5
    elif fam == "mixed_sign_big":
        x *= 300.0; q *= 0.01
    index = faiss_compat.IndexFlatIP(d, device=0, seg_rows=seg_rows)
    cuts = sorted(set(int(c) for c in rng.integers(0, n + 1, size=int(rng.integers(0, 4)))) | {0, n})
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            index.add(x[a:b])
    D, I = index.search(q, k)
    st = index.search_stats()
    agg["refined"] += st["refined_queries"]; agg["flagged"] += st["flagged_queries"]; agg["exact"] += st["exact_queries"]
    agg["overflow"] += st["overflow_retries"]
    if st["overflow_retries"]:
        agg.setdefault("overflow_cases", []).append((fam, n, d, k, nq, seg_rows))
    s = q.double() @ x.double().t()
    kk = min(k, n)
    Dr, Ir = torch.topk(s, kk, dim=1)
    ok = True
    why = ""
    if kk < k and not ((I[:, kk:] == -1).all() and (D[:, kk:] < -3e38).all()):
        ok, why = False, "padding"
    Dk, Ik = D[:, :kk].double(), I[:, :kk]
    scale = float(q.double().norm(dim=1).max() * x.double().norm(dim=1).max())
    tol = 3e-7 * scale
    if ok and not torch.all((Dk - Dr).abs() <= 3e-6 * Dr.abs() + tol):
        ok, why = False, f"scores max diff {float((Dk - Dr).abs().max()):.3e} tol {tol:.3e}"
    diff = Ik != Ir
    if ok and diff.any():
        # a differing id must sit in an fp32 tie: its exact score equals the reference's at that rank
        sd = torch.gather(s, 1, Ik.clamp_min(0))
        if not torch.all(((sd - Dr).abs() <= 3e-6 * Dr.abs() + tol)[diff]):
            ok, why = False, f"ids differ outside ties ({int(diff.sum())})"
    if ok and st["flagged_queries"] != 0:
        ok, why = False, "flagged queries left"
    if not ok:
        bad.append(dict(it=it, fam=fam, n=n, d=d, k=k, nq=nq, seg_rows=seg_rows, why=why, stats=st))
    del index, x, q, s
print(json.dumps(dict(iterations=iters, seed=seed, failures=len(bad), agg=agg, seconds=round(time.time() - t0, 1), bad=bad[:5])))
sys.exit(1 if bad else 0)
