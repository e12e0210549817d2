"""Dev: phase timeline of gemm_tc_small_kernel<kCe> (globaltimer stamps per CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
B, n = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "128x8").split("x"))
dev = torch.device("cuda", 0)
dbg = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
os.environ["DRT_B200_CE_TRACE"] = hex(dbg.data_ptr())
from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss
x = torch.randn(B, 768, device=dev); y = torch.randn(B * n, 768, device=dev)
fn = SimpleContrastiveLoss()
with torch.no_grad():
    for _ in range(5): fn(x, y)
    torch.cuda.synchronize(); dbg.zero_(); torch.cuda.synchronize()
    fn(x, y); torch.cuda.synchronize()
t = dbg.cpu()
t = t[t[:, 0] > 0]
t0 = int(t[:, 0].min())
names = ["start", "setup", "mma_done", "ticket", "summed", "tile_done", "fold_done", "exit"]
print("ctas", t.shape[0])
for j, nm in enumerate(names):
    col = t[:, j][t[:, j] > 0]
    if len(col): print(f"{nm:10s} n={len(col):4d}  min {int(col.min()) - t0:7d} ns  median {int(col.median()) - t0:7d} ns  max {int(col.max()) - t0:7d} ns")
