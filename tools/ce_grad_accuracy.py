"""Gradient accuracy of the in-batch loss against a float64 reference, for the shapes that run on
the tensor-core (bf16x3) path; torch's fp32 eager result is measured the same way beside it."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
for B, P in [(128, 8192), (256, 2048), (1024, 8192)]:
    g = torch.Generator(device=dev).manual_seed(B + P)
    x = torch.randn(B, 768, device=dev, generator=g)
    y = torch.randn(P, 768, device=dev, generator=g)
    tgt = torch.arange(0, P, P // B, device=dev)[:B]
    def run(fn, dt):
        a, b = x.to(dt).clone().requires_grad_(True), y.to(dt).clone().requires_grad_(True)
        l = fn(a, b); l.backward()
        return l.detach().double(), a.grad.double(), b.grad.double()
    l64, dx64, dy64 = run(lambda a, b: torch.nn.functional.cross_entropy(a @ b.t(), tgt), torch.float64)
    l32, dx32, dy32 = run(lambda a, b: torch.nn.functional.cross_entropy(a @ b.t(), tgt), torch.float32)
    lo, dxo, dyo = run(lambda a, b: SimpleContrastiveLoss()(a, b, target=tgt), torch.float32)
    rel = lambda a, r: ((a - r).abs().max() / r.abs().max()).item()
    print(json.dumps(dict(B=B, P=P, loss_rel_ours=abs((lo - l64) / l64).item(), loss_rel_torch32=abs((l32 - l64) / l64).item(),
                          dx_rel_ours=rel(dxo, dx64), dx_rel_torch32=rel(dx32, dx64),
                          dy_rel_ours=rel(dyo, dy64), dy_rel_torch32=rel(dy32, dy64))), flush=True)
