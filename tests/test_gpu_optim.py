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


def test_masked_mse_matches_torch_expression():
    """get_loss_dict's masked-ray MSE (studio_model.py:415-426) as one kernel each way == MSELoss over masked_select rows + 1e-6."""
    from pointnerf2studio_b200 import native
    g = torch.Generator().manual_seed(3)
    for R in (1, 777, 4096, 100_000):
        pred = torch.rand((R, 3), generator=g).cuda().requires_grad_(True)
        image = torch.rand((R, 3), generator=g).cuda()
        mask = (torch.rand((R,), generator=g) > 0.4).to(torch.int8).cuda()
        if R == 1:
            mask[:] = 1
        loss = native.masked_mse(pred, image, mask)
        (loss * 1.7).backward()
        p2 = pred.detach().clone().requires_grad_(True)
        m = mask > 0
        ref = torch.nn.functional.mse_loss(p2[m], image[m]) + 1e-6
        (ref * 1.7).backward()
        torch.testing.assert_close(loss, ref, rtol=2e-5, atol=1e-7)
        torch.testing.assert_close(pred.grad, p2.grad, rtol=2e-5, atol=1e-9)
    # no masked ray: NaN, as MSELoss over an empty selection
    pred = torch.rand((8, 3)).cuda()
    assert torch.isnan(native.masked_mse(pred, pred.clone(), torch.zeros(8, dtype=torch.int8).cuda()))
