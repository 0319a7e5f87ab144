"""The oracle against fixtures produced by executing the reference's own code
(tests/golden/make_golden.py): positional encoding, coarse positions, aggregator + ray_march
forward, and the autograd gradients of points and MLP weights (original mode, shipped weights)."""
import os

import numpy as np
import pytest
import torch

from oracle import field as of
from oracle import grid_query as gq
from scenes import GOLDEN_CFG as CFG, golden_scene

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    return dict(np.load(os.path.join(HERE, "golden", "pointnerf_golden.npz")))


@pytest.fixture(scope="module")
def W():
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(HERE, "golden", "aggregator_weights.npz")).items()}
    return of.FieldWeights.from_aggregator(sd, prefix="")


def test_positional_encoding_matches_reference(G):
    for x, y, F, ori in (("pe_x6", "pe_y6", 5, False), ("pe_x32", "pe_y32", 3, False), ("pe_x3", "pe_y3", 4, True)):
        got = of.positional_encoding(torch.from_numpy(G[x]), F, ori=ori).numpy()
        assert got.shape == G[y].shape
        np.testing.assert_array_equal(got, G[y])


def test_coarse_positions_match_reference(G):
    cloud, cam, pix = golden_scene()
    np.testing.assert_array_equal(pix, G["pix"])
    raypos, t_mid = of.coarse_positions(torch.from_numpy(cam.origin), torch.from_numpy(cam.rays(pix)), CFG["D"],
                                        CFG["near"], CFG["far"], jitter=0.0)
    np.testing.assert_array_equal(t_mid[0].numpy(), G["t_mid"])
    np.testing.assert_array_equal(raypos[:4].numpy(), G["raypos_first4"])


@pytest.fixture(scope="module")
def oracle_run(G, W):
    cloud, cam, pix = golden_scene()
    frame = gq.hyperparameters(cloud.xyz, CFG["vsize"], CFG["vscale"], CFG["kernel_size"], CFG["ranges"])
    origin, dirs = torch.from_numpy(cam.origin), torch.from_numpy(cam.rays(pix))
    raypos, _ = of.coarse_positions(origin, dirs, CFG["D"], CFG["near"], CFG["far"])
    radius = np.float32(4 * max(CFG["vsize"][:2]))
    pidx, loc_w, ray_mask = gq.woord_query_grid_point_index(raypos.numpy(), cloud.xyz, CFG["kernel_size"],
                                                            CFG["query_size"], CFG["SR"], CFG["K"], frame, CFG["P"], radius)
    pts = {"xyz": torch.from_numpy(cloud.xyz), "Rw2c": torch.from_numpy(cloud.Rw2c)}
    for k in ("embed", "color", "dir", "conf"):
        pts[k] = torch.from_numpy(getattr(cloud, k)).clone().requires_grad_(True)
    W.requires_grad_(True)
    for v in W.p.values():
        v.grad = None
    out = of.render(pts, W, origin, dirs, torch.from_numpy(cam.R_c2w), pidx, loc_w, ray_mask, CFG["vsize"][2], CFG["SR"],
                    mode="original")
    return dict(frame=frame, pidx=pidx, loc_w=loc_w, ray_mask=ray_mask, out=out, pts=pts)


def test_query_fixture_stable(G, oracle_run):
    np.testing.assert_array_equal(oracle_run["frame"].dim, G["frame_dim"])
    np.testing.assert_array_equal(oracle_run["frame"].lo, G["frame_lo"])
    np.testing.assert_array_equal(oracle_run["pidx"], G["pidx"].astype(np.int32))
    np.testing.assert_array_equal(oracle_run["loc_w"], G["loc_w"])
    np.testing.assert_array_equal(oracle_run["ray_mask"], G["ray_mask"])
    assert 0 < G["ray_mask"].sum() < len(G["ray_mask"])


def test_field_and_march_forward_match_reference(G, oracle_run):
    o = oracle_run["out"]
    np.testing.assert_array_equal(o["valid"].numpy(), G["ray_valid"])
    np.testing.assert_allclose(o["extras"]["weight"].detach().numpy(), G["weight"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(o["conf_coefficient"].detach().numpy(), G["conf_coefficient"], rtol=0, atol=0)
    np.testing.assert_allclose(o["delta"].numpy(), G["ray_dist"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(o["decoded"].detach().numpy(), G["decoded"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(o["blend_weight"].detach().numpy(), G["blend_weight"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(o["C_valid"].detach().numpy(), G["ray_color"], rtol=2e-5, atol=2e-6)


def test_backward_matches_reference_autograd(G, W, oracle_run):
    o = oracle_run["out"]
    gt = torch.from_numpy(G["gt"])
    val = o["conf_coefficient"].clamp(1e-3, 1 - 1e-3)
    loss = torch.nn.functional.mse_loss(o["C_valid"], gt) + 1e-4 * torch.mean(torch.log(val) + torch.log(1 - val))
    assert abs(loss.item() - float(G["loss"])) < 1e-6
    loss.backward()
    for k in ("embed", "color", "dir", "conf"):
        g = oracle_run["pts"][k].grad.numpy()
        ref = G["grad_" + k]
        scale = np.abs(ref).max()
        assert scale > 0
        np.testing.assert_allclose(g, ref, rtol=1e-3, atol=2e-5 * scale, err_msg=k)
    for new, old, _, _ in of.FieldWeights.NAMES:
        for s in ("weight", "bias"):
            g = W.p[f"{new}.{s}"].grad.numpy()
            ref = G[f"gradw_{old}.{s}"]
            scale = np.abs(ref).max()
            np.testing.assert_allclose(g, ref, rtol=1e-3, atol=5e-5 * scale, err_msg=new)
