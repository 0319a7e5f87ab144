"""Host-side logic that needs no GPU: the rows-per-sample classes, the step constants of a graph replay, the op argument packing,
the lazily materialised confidence tensor."""
import types

import numpy as np
import torch


def test_class_rows():
    from pointnerf2studio_b200.native_tc import class_rows
    assert class_rows(8) == [8, 4, 2] and class_rows(16) == [16, 8, 4, 2] and class_rows(3) == [4, 2] and class_rows(1) == [2]
    assert class_rows(32) == [32, 16, 8, 4, 2]


def test_adam_step_scalars_follow_torch_and_the_schedule():
    """TrainEngine._hyper: what the graph replay writes for step k = (lr_k / bc1, lr_k / bc1, 1 / sqrt(bc2)) with the reference's
    schedule lr_k = lr0 * 0.1 ** ((k - 1) / 1e6) (LambdaLR: optimiser step k runs with lambda(k - 1))."""
    from pointnerf2studio_b200.parallel import TrainEngine, _as_i32
    eng = types.SimpleNamespace(lr0=(2e-3, 5e-4), decay=(0.1, 1e6), betas=(0.9, 0.999))
    for k in (1, 2, 10, 250000):
        lp, lf, ib = TrainEngine._hyper(eng, k)
        f = 0.1 ** ((k - 1) / 1e6)
        assert abs(lp - 2e-3 * f / (1 - 0.9 ** k)) < 1e-15 and abs(lf - 5e-4 * f / (1 - 0.9 ** k)) < 1e-15
        assert abs(ib - 1 / np.sqrt(1 - 0.999 ** k)) < 1e-12
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=2e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda s: 0.1 ** (s / 1e6))
    for k in range(1, 4):
        assert abs(opt.param_groups[0]["lr"] - 2e-3 * 0.1 ** ((k - 1) / 1e6)) < 1e-18
        opt.step(); sched.step()
    assert _as_i32(0xffffffff) == -1 and _as_i32(5) == 5 and _as_i32(0x80000000) == -(1 << 31)


def test_op_argument_packing_layout():
    from pointnerf2studio_b200 import native, ops
    frame = native.GridFrame(lo=np.array([1, 2, 3], np.float32), hi=np.zeros(3, np.float32), sv=np.array([.5, .25, .125], np.float32),
                             dim=np.array([7, 8, 9], np.int32))
    mode = native.make_mode("original", training=False, bg=(0.1, 0.2, 0.3), vsize_z=0.004)
    fl, it = ops.fl_it(frame, [4, 5, 6], np.arange(9, dtype=np.float32).reshape(3, 3), list(range(10, 19)), 2.0, 6.0, 0.3, 0.016, mode, 400, 80, 8,
                       3, seed=(7 << 32) | 9, t_stride=400, event=123, ws_limit_mib=5)
    assert len(fl) == ops.N_FL and len(it) == ops.N_IT
    assert fl[ops.FL_LO:ops.FL_LO + 3] == [1, 2, 3] and fl[ops.FL_SV:ops.FL_SV + 3] == [.5, .25, .125] and fl[ops.FL_ORIGIN:ops.FL_ORIGIN + 3] == [4, 5, 6]
    assert fl[ops.FL_RC2W:ops.FL_RC2W + 9] == list(range(9)) and fl[ops.FL_RW2C:ops.FL_RW2C + 9] == list(range(10, 19))
    assert fl[ops.FL_NEAR] == 2.0 and fl[ops.FL_FAR] == 6.0 and fl[ops.FL_JITTER] == 0.3 and abs(fl[ops.FL_SLOPE] - 0.01) < 1e-9
    assert it[ops.IT_DIM:ops.IT_DIM + 3] == [7, 8, 9] and it[ops.IT_SEED_LO] == 9 and it[ops.IT_SEED_HI] == 7
    assert (it[ops.IT_D], it[ops.IT_SR], it[ops.IT_K], it[ops.IT_KS0]) == (400, 80, 8, 3)
    assert (it[ops.IT_SOFTPLUS], it[ops.IT_WCONF], it[ops.IT_BGMODE], it[ops.IT_CLAMP]) == (1, 1, 1, 0)
    assert it[ops.IT_TSTRIDE] == 400 and it[ops.IT_EVENT] == 123 and it[ops.IT_WS_LIMIT_MIB] == 5
    assert sum(ops.MLP_NUMEL) == 341764 and len(ops.MLP_SHAPES) == 18


def test_conf_coefficient_is_a_lazy_tensor():
    from pointnerf2studio_b200.model import ConfCoefficient
    conf = torch.tensor([[[0.5], [2.0], [0.00001], [0.9]]])                    # (1, N, 1)
    pidx = torch.tensor([[[0, 1], [-1, 3]], [[2, 2], [1, -1]], [[3, 3], [3, 3]]], dtype=torch.int32)     # (R=3, SR=2, K=2)
    ray_mask = torch.tensor([1, 0, 1], dtype=torch.int8)
    cc = ConfCoefficient(conf, pidx, ray_mask, torch.tensor([2], dtype=torch.int32))
    assert cc._t is None
    assert tuple(cc.shape) == (1, 2, 2, 2)                                     # R'' = 2 surviving rays
    want = torch.tensor([[[0.5, 1.0], [0.5, 0.9]], [[0.9, 0.9], [0.9, 0.9]]])[None]     # invalid slots read point 0, clamp to [1e-4, 1]
    assert torch.allclose(torch.clamp(cc, 1e-3, 1 - 1e-3), want.clamp(1e-3, 1 - 1e-3))
    assert torch.allclose(cc.materialize(), want) and cc._t is not None
