"""Dev: where the host time of the loss call goes (cfg3 shape)."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

dev = torch.device("cuda", 0)
B, n = 128, 8
x = torch.randn(B, 768, device=dev, requires_grad=True)
y = torch.randn(B * n, 768, device=dev, requires_grad=True)
fn = SimpleContrastiveLoss()


def wall(f, reps=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6    # host issue time, wall incl. drain


def fwd():
    return fn(x, y)

def fwdbwd():
    x.grad = None; y.grad = None
    fn(x, y).backward()

class Null(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a.new_zeros(())
    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return a, b

def null_fwdbwd():
    x.grad = None; y.grad = None
    Null.apply(x, y).backward()

with torch.no_grad():
    print("fwd no_grad   host/wall us", wall(fwd))
print("fwd with grad host/wall us", wall(fwd))
print("fwd+bwd       host/wall us", wall(fwdbwd))
print("null Function fwd+bwd host/wall us", wall(null_fwdbwd))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): fwdbwd()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
