"""Fused Adam (csrc/optim.cu) against torch.optim.Adam on the same tensors after 5 steps with a decaying lr (fp32 tolerance: 1e-5
relative on the parameters, 5e-5 on the moments -- the two evaluate the same formula with different contraction / scalar rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    from pointnerf2studio_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(1, 5001, 32), (1, 5001, 1), (256, 284), (256,), (3,), (1, 7, 3)]
    a = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    b = [t.detach().clone().requires_grad_(True) for t in a]
    oa = torch.optim.Adam(a, lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    ob = FusedAdam(b, lr=2e-3)
    sa = torch.optim.lr_scheduler.LambdaLR(oa, lr_lambda=lambda s: 0.1 ** (s / 10))
    sb = torch.optim.lr_scheduler.LambdaLR(ob, lr_lambda=lambda s: 0.1 ** (s / 10))
    for it in range(5):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, device="cuda", generator=g) * (10.0 ** (it - 2))
            if it == 3 and x.numel() == 3:
                gr = None                                   # a parameter without a gradient is skipped by both
            x.grad = gr
            y.grad = None if gr is None else gr.clone()
        oa.step(); ob.step(); sa.step(); sb.step()
    for x, y in zip(a, b):
        torch.testing.assert_close(y, x, rtol=1e-5, atol=1e-7)
    for x, y in zip(a, b):
        torch.testing.assert_close(ob.state[y]["exp_avg_sq"], oa.state[x]["exp_avg_sq"], rtol=5e-5, atol=1e-12)
        torch.testing.assert_close(ob.state[y]["exp_avg"], oa.state[x]["exp_avg"], rtol=1e-5, atol=2e-4)   # |m| ~ 10: cancellation near 0
