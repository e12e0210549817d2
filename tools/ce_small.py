"""Dev: a few CE fwd+bwd calls (ours, then torch eager) as an ncu target for device-time comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
B, n = 128, 8
x = torch.randn(B, 768, device=dev, requires_grad=True)
y = torch.randn(B * n, 768, device=dev, requires_grad=True)
tgt = torch.arange(0, B * n, n, device=dev)
ours = SimpleContrastiveLoss()
for _ in range(3):
    ours(x, y).backward()
torch.cuda.synchronize()
for _ in range(3):
    torch.nn.functional.cross_entropy(x @ y.t(), tgt).backward()
torch.cuda.synchronize()
print("done")
