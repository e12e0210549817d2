"""torchrun -N ranks: DistributedContrastiveLoss (sharded evaluation, fused kernels) against the
reference formulation (all-gather both, local slot keeps autograd, torch CE on the full matrix,
x world_size) — values, gradients and timing.  Run: torchrun --nproc-per-node N tools/dist_loss_check.py"""
import os
import sys
import json
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False
from denseretrievaltoolkits_b200.losses import DistributedContrastiveLoss, gather_rank_major

B, n, d = 128, 8, 768
g = torch.Generator(device=dev).manual_seed(100 + rank)
x = torch.randn(B, d, device=dev, generator=g)
y = torch.randn(B * n, d, device=dev, generator=g)


def ref_loss(xr, yr):
    X = gather_rank_major(xr, rank, world)
    Y = gather_rank_major(yr, rank, world)
    tgt = torch.arange(0, X.shape[0] * n, n, device=dev)
    return torch.nn.functional.cross_entropy(X @ Y.t(), tgt) * world


ours = DistributedContrastiveLoss()
x1, y1 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
x2, y2 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
l1 = ours(x1, y1)
l1.backward()
l2 = ref_loss(x2, y2)
l2.backward()
res = dict(rank=rank, loss_ours=l1.item(), loss_ref=l2.item(),
           dx_rel=((x1.grad - x2.grad).abs().max() / x2.grad.abs().max()).item(),
           dy_rel=((y1.grad - y2.grad).abs().max() / y2.grad.abs().max()).item())
ok = abs(res["loss_ours"] - res["loss_ref"]) <= 1e-4 * abs(res["loss_ref"]) and res["dx_rel"] < 1e-3 and res["dy_rel"] < 1e-3


def timeit(fn, reps=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def fb_ours():
    x1.grad = None; y1.grad = None
    ours(x1, y1).backward()


def fb_ref():
    x2.grad = None; y2.grad = None
    ref_loss(x2, y2).backward()


res.update(fwdbwd_us_ours=timeit(fb_ours), fwdbwd_us_reference_formulation=timeit(fb_ref), ok=ok, world=world)
sys.stdout.write(json.dumps(res) + "\n")   # one write: lines of different ranks must not interleave
sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
