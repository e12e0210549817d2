"""GPU: the in-batch loss on its three code paths — the small-shape tcgen05 kernel with the cross
entropy fused into the accumulator read-out (cfg3's path), the large-shape persistent tcgen05
kernel with the online (max, sum-exp) epilogue, and the fp32 SIMT kernels for shapes whose
contraction length is not a multiple of 4 — against a float64 evaluation of
`F.cross_entropy(x @ y.T, target)` (DRT/trainer/losses.py:11-17) and its autograd."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref64(x, y, target, reduction):
    xd = x.detach().double().requires_grad_(True)
    yd = y.detach().double().requires_grad_(True)
    if target is None:
        tpq = y.shape[0] // x.shape[0]
        target = torch.arange(0, x.shape[0] * tpq, tpq, device=x.device)
    loss = torch.nn.functional.cross_entropy(xd @ yd.t(), target, reduction=reduction)
    (loss.sum() if reduction == "none" else loss).backward()
    return loss.detach(), xd.grad, yd.grad


def _check(x, y, target=None, reduction="mean", rtol=5e-6):
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

    fn = SimpleContrastiveLoss()
    xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    loss = fn(xs, ys) if (target is None and reduction == "mean") else fn(xs, ys, target=target, reduction=reduction)
    (loss.sum() if reduction == "none" else loss).backward()
    lr, gx, gy = _ref64(x, y, target, reduction)
    torch.testing.assert_close(loss.double(), lr, rtol=rtol, atol=rtol * float(lr.abs().max()))
    for got, want in ((xs.grad, gx), (ys.grad, gy)):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) <= 3e-5 * scale, (float((got.double() - want).abs().max()), scale)
    return loss


@pytest.mark.parametrize("B,P,d", [(128, 1024, 768), (16, 32, 64), (100, 300, 96), (256, 2048, 768), (8, 1024, 768),
                                   (130, 520, 132), (4, 4, 16)])
def test_small_tensor_core_path_against_float64(B, P, d):
    g = torch.Generator(device="cuda").manual_seed(B * 31 + P)
    x = torch.randn((B, d), generator=g, device="cuda")
    y = torch.randn((P, d), generator=g, device="cuda") * 0.7 + 0.05
    _check(x, y)
    tgt = torch.randint(0, P, (B,), generator=g, device="cuda")
    _check(x, y, target=tgt, reduction="sum")
    _check(x, y, target=tgt, reduction="none")


@pytest.mark.parametrize("B,P,d", [(5, 15, 64), (16, 32, 66), (33, 70, 50)])
def test_simt_path_for_odd_contraction_lengths(B, P, d):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((B, d), generator=g, device="cuda")
    y = torch.randn((P, d), generator=g, device="cuda")
    _check(x, y)


def test_backward_twice_over_one_graph_matches_torch():
    """the backward used to overwrite the saved logits in place, so a second backward
    over the same graph (retain_graph=True) returned wrong gradients without any error."""
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss, inbatch_scores_and_loss

    g = torch.Generator(device="cuda").manual_seed(3)
    for B, P, d in [(128, 1024, 768), (16, 48, 66), (1024, 8192, 256)]:
        x = torch.randn((B, d), generator=g, device="cuda", requires_grad=True)
        y = torch.randn((P, d), generator=g, device="cuda", requires_grad=True)
        loss = SimpleContrastiveLoss()(x, y)
        loss.backward(retain_graph=True)
        g1x, g1y = x.grad.clone(), y.grad.clone()
        x.grad = None; y.grad = None
        loss.backward()
        assert torch.equal(x.grad, g1x) and torch.equal(y.grad, g1y)
        # two losses sharing one node, and the returned scores stay intact after the backward
        x.grad = None; y.grad = None
        l2, scores = inbatch_scores_and_loss(x, y, P // B)
        keep = scores.clone()
        (l2 * 2.0 + l2).backward()
        assert torch.equal(scores, keep)
        torch.testing.assert_close(x.grad, 3.0 * g1x, rtol=1e-5, atol=3e-6 * float(g1x.abs().max()))
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = (x.detach() @ y.detach().t())
        assert float((scores - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def test_forward_without_grad_never_needs_the_score_matrix():
    """No backward and no DROutput.scores requested: the logits pointer handed to the C ABI is
    NULL on every path, and the loss is the same as with it."""
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss, inbatch_scores_and_loss

    g = torch.Generator(device="cuda").manual_seed(4)
    for B, P, d in [(128, 1024, 768), (1024, 8192, 768), (16, 32, 66)]:
        x = torch.randn((B, d), generator=g, device="cuda")
        y = torch.randn((P, d), generator=g, device="cuda")
        with torch.no_grad():
            l0 = SimpleContrastiveLoss()(x, y)
        l1, sc = inbatch_scores_and_loss(x, y, P // B)
        assert sc is not None and sc.shape == (B, P)
        torch.testing.assert_close(l0, l1, rtol=1e-6, atol=0)
        lr, _, _ = _ref64(x, y, None, "mean")
        torch.testing.assert_close(l0.double(), lr, rtol=5e-6, atol=0)


def test_losses_on_two_streams_do_not_share_scratch():
    """one global workspace per device let two losses in flight on different streams
    corrupt each other's tickets / partials.  Workspaces are keyed by (device, stream)."""
    from denseretrievaltoolkits_b200.losses import SimpleContrastiveLoss

    fn = SimpleContrastiveLoss()
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = [torch.randn((128, 768), generator=g, device="cuda") for _ in range(2)]
    ys = [torch.randn((1024, 768), generator=g, device="cuda") for _ in range(2)]
    with torch.no_grad():
        want = [fn(x, y).item() for x, y in zip(xs, ys)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [[], []]
    for it in range(50):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s), torch.no_grad():
                got[i].append(fn(xs[i], ys[i]))
    torch.cuda.synchronize()
    for i in range(2):
        vals = torch.stack(got[i]).cpu().numpy()
        assert np.all(vals == np.float32(want[i])), (vals[:5], want[i])
