"""Steady-state DEVICE time of the in-batch loss (no host in the loop): N iterations captured in
one CUDA graph, replayed, timed with CUDA events.  Ours vs torch eager (fp32 matmul + F.cross_entropy).
python tools/ce_device_time.py [BxN ...]   e.g. 128x8 256x8"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
ITERS = 20


def graph_time(step, reps=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(5):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):      # same stream as the warm-up: the library's scratch is per (device, stream)
        for _ in range(ITERS):
            step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * ITERS) * 1e3


shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(128, 8), (256, 8), (16, 8), (1024, 8)]
for B, n in shapes:
    x = torch.randn(B, 768, device=dev, requires_grad=True)
    y = torch.randn(B * n, 768, device=dev, requires_grad=True)
    tgt = torch.arange(0, B * n, n, device=dev)
    ours = SimpleContrastiveLoss()

    def f_ours():
        with torch.no_grad():
            return ours(x, y)

    def fb_ours():
        loss = ours(x, y)
        return torch.autograd.grad(loss, (x, y))

    def f_ref():
        with torch.no_grad():
            return torch.nn.functional.cross_entropy(x @ y.t(), tgt)

    def fb_ref():
        loss = torch.nn.functional.cross_entropy(x @ y.t(), tgt)
        return torch.autograd.grad(loss, (x, y))

    print(json.dumps(dict(B=B, P=B * n, d=768, fwd_dev_us_ours=round(graph_time(f_ours), 2), fwd_dev_us_torch=round(graph_time(f_ref), 2),
                          fwdbwd_dev_us_ours=round(graph_time(fb_ours), 2), fwdbwd_dev_us_torch=round(graph_time(fb_ref), 2))), flush=True)
