"""Dev timing of the fused in-batch CE (cfg3 shapes) against torch eager on the same GPU."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)


def timeit(fn, reps=int(os.environ.get("CE_REPS", "200")), warm=int(os.environ.get("CE_WARM", "20"))):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


shapes = [(128, 8), (1024, 8), (16, 2)]
if os.environ.get("CE_SHAPES"):
    shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["CE_SHAPES"].split(",")]
for B, n in shapes:
    x = torch.randn(B, 768, device=dev, requires_grad=True)
    y = torch.randn(B * n, 768, device=dev, requires_grad=True)
    tgt = torch.arange(0, B * n, n, device=dev)
    ours = SimpleContrastiveLoss()

    def f_ours():
        return ours(x, y)

    def fb_ours():
        x.grad = None; y.grad = None
        ours(x, y).backward()

    def f_ref():
        return torch.nn.functional.cross_entropy(x @ y.t(), tgt)

    def fb_ref():
        x.grad = None; y.grad = None
        torch.nn.functional.cross_entropy(x @ y.t(), tgt).backward()

    with torch.no_grad():
        t_f_ours, t_f_ref = timeit(f_ours), timeit(f_ref)
    t_fb_ours, t_fb_ref = timeit(fb_ours), timeit(fb_ref)
    print(json.dumps(dict(B=B, P=B * n, fwd_us_ours=round(t_f_ours, 1), fwd_us_torch=round(t_f_ref, 1),
                          fwdbwd_us_ours=round(t_fb_ours, 1), fwdbwd_us_torch=round(t_fb_ref, 1))), flush=True)
