"""Staged on-GPU self test used while bringing kernels up: each stage runs in its own process
(a device-side trap poisons the CUDA context) under a timeout.
  python tools/gpu_selftest.py            # all stages
  python tools/gpu_selftest.py <stage>    # one stage in this process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["tiny1", "tiny2", "mid1", "mid2", "ties1", "big1", "big2", "k1000", "ce", "merge", "mining"]


def _search_case(nq, n, k, ctas, seg_rows=8192, dup=False, seed=0):
    import numpy as np

    from denseretrievaltoolkits_b200 import _lib, faiss_compat
    from oracle import flat_ip

    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    if dup:
        x[-n // 100:] = x[: n // 100]
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    index = faiss_compat.IndexFlatIP(768, device=0, seg_rows=seg_rows)
    index.add(x)
    flags = _lib.SEARCH_FORCE_2CTA if ctas == 2 else _lib.SEARCH_FORCE_1CTA
    t0 = time.time()
    D, I = index.search(q, k, flags=flags)
    t1 = time.time()
    Dr, Ir = flat_ip.flat_ip_search(x, q, k)
    kk = min(k, n)
    recall = np.mean([len(set(a[:kk]) & set(b[:kk])) / kk for a, b in zip(I, Ir)])
    same = (I == Ir)
    err = np.abs(D[same] - Dr[same]) / np.maximum(1e-6, np.abs(Dr[same]))
    print(f"  nq={nq} n={n} k={k} ctas={ctas}: recall={recall:.5f} same_ids={same.mean():.5f} "
          f"max_rel_err={err.max():.2e} time={t1 - t0:.3f}s stats={index.search_stats()}")
    if recall < 0.999 or err.max() > 1e-4:
        # diagnostics
        print("   first row gpu ids", I[0][:10], "scores", D[0][:5])
        print("   first row ref ids", Ir[0][:10], "scores", Dr[0][:5])
        raise SystemExit(1)


def run_stage(name):
    import numpy as np
    import torch

    torch.cuda.set_device(0)
    if name == "tiny1":
        _search_case(7, 1000, 10, 1)
        _search_case(3, 5, 8, 1)          # k > ntotal
    elif name == "tiny2":
        _search_case(7, 1000, 10, 2)
        _search_case(3, 5, 8, 2)
    elif name == "mid1":
        _search_case(200, 20000, 100, 1)
        _search_case(33, 5000, 1000, 1)
    elif name == "mid2":
        _search_case(200, 20000, 100, 2)
        _search_case(300, 20000, 100, 2)
        _search_case(33, 5000, 1000, 2)
    elif name == "ties1":
        _search_case(64, 30000, 100, 1, dup=True)
        _search_case(64, 30000, 100, 2, dup=True)
    elif name == "big1":
        _search_case(1000, 100000, 100, 1, seg_rows=1 << 16)
    elif name == "big2":
        _search_case(1000, 100000, 100, 2, seg_rows=1 << 16)
    elif name == "k1000":
        _search_case(128, 100000, 1000, 1, seg_rows=1 << 16)
        _search_case(128, 100000, 1000, 2, seg_rows=1 << 16)
    elif name == "ce":
        from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss
        from oracle import inbatch_loss

        rng = np.random.default_rng(1)
        for B, n in [(128, 8), (16, 2), (5, 3)]:
            x = rng.standard_normal((B, 768), dtype=np.float32)
            y = rng.standard_normal((B * n, 768), dtype=np.float32)
            xt = torch.from_numpy(x).cuda().requires_grad_(True)
            yt = torch.from_numpy(y).cuda().requires_grad_(True)
            loss = SimpleContrastiveLoss()(xt, yt)
            loss.backward()
            ref, _, _ = inbatch_loss.contrastive_loss(x, y)
            dx, dy = inbatch_loss.contrastive_loss_grads(x, y)
            e1 = abs(loss.item() - ref) / abs(ref)
            e2 = np.abs(xt.grad.cpu().numpy() - dx).max() / np.abs(dx).max()
            e3 = np.abs(yt.grad.cpu().numpy() - dy).max() / np.abs(dy).max()
            print(f"  ce B={B} n={n}: loss={loss.item():.6f} ref={ref:.6f} rel={e1:.2e} dx={e2:.2e} dy={e3:.2e}")
            assert e1 < 1e-4 and e2 < 1e-4 and e3 < 1e-4
    elif name == "merge":
        from denseretrievaltoolkits_b200.store import ShardedCorpusStore
        from oracle import flat_ip

        rng = np.random.default_rng(2)
        x = rng.standard_normal((30000, 768), dtype=np.float32)
        q = rng.standard_normal((100, 768), dtype=np.float32)
        st = ShardedCorpusStore(768, num_virtual_shards=4, device=0, seg_rows=4096)
        st.add_split(x)
        D, I = st.search(q, 100)
        Dr, Ir = flat_ip.flat_ip_search(x, q, 100)
        recall = np.mean([len(set(a) & set(b)) / 100 for a, b in zip(I, Ir)])
        print(f"  merge 4 virtual shards: recall={recall:.5f} same={np.mean(I == Ir):.5f}")
        assert recall >= 0.999
    elif name == "mining":
        from denseretrievaltoolkits_b200.mining import filter_negatives
        from oracle import merge as omerge

        rng = np.random.default_rng(3)
        ids = rng.integers(0, 1000, size=(50, 200)).astype(np.int64)
        pb = rng.integers(0, 900, size=50).astype(np.int64)
        pe = pb + rng.integers(1, 100, size=50)
        out = filter_negatives(torch.from_numpy(ids).cuda(), torch.from_numpy(pb), torch.from_numpy(pe), 64)
        ref = omerge.filter_negatives(ids, pb, pe, 64)
        assert (out.cpu().numpy() == ref).all()
        print("  mining filter ok")
    else:
        raise SystemExit(f"unknown stage {name}")
    print(f"stage {name}: PASS")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_stage(sys.argv[1])
    else:
        failed = []
        for s in STAGES:
            print(f"=== {s}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, s], timeout=240)
                if r.returncode != 0:
                    failed.append(s)
                    print(f"stage {s}: FAIL rc={r.returncode}", flush=True)
            except subprocess.TimeoutExpired:
                failed.append(s)
                print(f"stage {s}: TIMEOUT", flush=True)
        print("FAILED:", failed)
        sys.exit(1 if failed else 0)
