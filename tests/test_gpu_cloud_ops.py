"""GPU parity of the point-cloud maintenance kernels (csrc/cloud_ops.cu) against the golden vectors produced by executing the
reference's source and against the oracle on larger seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import cloud_ops as oc
from test_oracle_cloud_ops import G, assert_same_argmin

pytestmark = pytest.mark.gpu


def test_probe_filter_kernel_matches_golden_and_oracle():
    from pointnerf2studio_b200 import cloud_ops
    for tag in ("nofar", "far"):
        k = lambda n: G[f"probe_{tag}_{n}"]
        H, W = k("ray_mask").shape
        c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        got = cloud_ops.probe_filter(c(k("ray_mask")), c(k("gt")), c(k("color")), c(k("far_dist")), c(k("opacity")), c(k("edge")),
                                     [1.0, 1.0, 1.0], H, W, float(k("far_thresh")), 0.7)
        np.testing.assert_array_equal(got.cpu().numpy().reshape(H, W), k("keep"))
    # a full 800 x 800 frame against the oracle
    rng = np.random.default_rng(3)
    H = W = 800
    ray_mask = (rng.random((H, W)) > 0.3).astype(np.int8)
    gt = rng.random((H, W, 3)).astype(np.float32)
    gt[rng.random((H, W)) > 0.5] = 1.0
    color = np.clip(gt + 0.1 * rng.standard_normal((H, W, 3)), 0, 1).astype(np.float32)
    far, op = (rng.random((H, W)) * 0.03).astype(np.float32), rng.random((H, W)).astype(np.float32)
    edge = rng.random((H, W)) > 0.05
    want = oc.probe_filter(ray_mask, gt, color, far, op, edge, [1.0, 1.0, 1.0], 0.012, 0.6)
    c = lambda a: torch.from_numpy(a).cuda()
    got = cloud_ops.probe_filter(c(ray_mask), c(gt), c(color), c(far), c(op), c(edge), torch.ones(3), H, W, 0.012, 0.6)
    np.testing.assert_array_equal(got.cpu().numpy().reshape(H, W), want)
    got = cloud_ops.probe_filter(c(ray_mask), c(gt), None, None, c(op), None, [1.0, 1.0, 1.0], H, W, -1.0, 0.6)
    np.testing.assert_array_equal(got.cpu().numpy().reshape(H, W), oc.probe_filter(ray_mask, gt, color, far, op, np.ones((H, W), bool),
                                                                                  [1.0, 1.0, 1.0], -1.0, 0.6))


def test_vox_closest_kernel_matches_golden_and_oracle():
    from pointnerf2studio_b200 import cloud_ops
    for tag in ("a", "b"):
        xyz, res = G[f"vox_{tag}_xyz"], int(G[f"vox_{tag}_res"])
        cen, grid, amin = cloud_ops.construct_vox_points_closest(torch.from_numpy(xyz).cuda(), res)
        np.testing.assert_array_equal(grid.cpu().numpy(), G[f"vox_{tag}_grid"])
        np.testing.assert_allclose(cen.cpu().numpy(), G[f"vox_{tag}_centroid"], rtol=0, atol=2e-7)
        assert amin.dtype == torch.int64
        assert_same_argmin(xyz, G[f"vox_{tag}_centroid"], amin.cpu().numpy(), G[f"vox_{tag}_min_idx"])
    # an MVS-sized cloud (1.5 M points, vox_res 320 as in the NeRF-synthetic scripts) against the oracle's frame / voxels / centroids
    from pointnerf2studio_b200.synth import make_cloud
    xyz = make_cloud(1_500_000, seed=9, P=1000).xyz
    cen, grid, amin = cloud_ops.construct_vox_points_closest(torch.from_numpy(xyz).cuda(), 320)
    f32 = np.float32
    mn, mx = xyz.min(0), xyz.max(0)
    edge = f32((mx - mn).max() * f32(1.05))
    smin = ((mx + mn) / f32(2) - edge / f32(2)).astype(f32)
    vi = np.floor(((xyz - smin[None]).astype(f32) / f32(edge / f32(320))).astype(f32)).astype(np.int64)
    key = (vi[:, 0] * 400 + vi[:, 1]) * 400 + vi[:, 2]
    uk, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    g = grid.cpu().numpy().astype(np.int64)
    np.testing.assert_array_equal((g[:, 0] * 400 + g[:, 1]) * 400 + g[:, 2], uk)
    sums = np.zeros((len(uk), 3))
    np.add.at(sums, inv, xyz.astype(np.float64))
    want_c = (sums / cnt[:, None]).astype(f32)
    np.testing.assert_allclose(cen.cpu().numpy(), want_c, rtol=0, atol=2e-7)
    a = amin.cpu().numpy()
    assert np.array_equal(inv[a], np.arange(len(uk)))                              # the chosen point lies in its voxel
    res_all = np.linalg.norm(xyz - want_c[inv], axis=1)
    best = np.full(len(uk), np.inf)
    np.minimum.at(best, inv, res_all)
    assert np.all(res_all[a] <= best * (1 + 1e-5) + 1e-9)                          # ... and is (one of) the closest to the centroid


def test_probe_hole_feeds_grow_points():
    """probe -> filter -> grow on a cloud with a hole punched into it: the new points land in the hole's neighbourhood."""
    from pointnerf2studio_b200 import RayBundle, cloud_ops
    from test_gpu_parity import _make_model, _scene
    from pointnerf2studio_b200.synth import make_camera
    s, cloud, cam, _ = _scene("config1")
    model = _make_model(cloud, "bf16", "original", SR=s["SR"], K=s["K"], P=s["P"]).eval()
    with torch.no_grad():       # make the surface opaque so that the densest sample passes the opacity threshold
        model.field_output_density.net.bias.fill_(400.0)
    H = W = 96
    small = make_camera(H=H, W=W, focal=cam.focal * H / cam.H * 6.0)       # zoom in: the 0.2-radius object fills the frame
    rb = RayBundle.for_camera(torch.from_numpy(small.rays(None)).cuda(), small.origin, small.R_c2w, small.near, small.far)
    out = model.probe(rb)
    hit = out["ray_mask"].reshape(H, W).bool()
    assert 0.2 < float(hit.float().mean()) < 0.98, float(hit.float().mean())
    gt = torch.ones((H, W, 3), device="cuda")
    gt[~hit] = 0.3                       # the ground truth says there is an object where no neural point was found
    n0 = model.neural_points.points_xyz.shape[0]
    add = cloud_ops.probe_hole(model, rb, gt, H, W, opacity_thresh=0.3)
    keep = cloud_ops.probe_filter(out["ray_mask"], gt, out["coarse_raycolor"], out["ray_max_far_dist"], out["ray_max_shading_opacity"], None,
                                  [1.0, 1.0, 1.0], H, W, -1.0, 0.3)
    assert add[0].shape[0] == int(keep.sum()) > 0 and add[1].shape == (add[0].shape[0], 32) and add[4].shape == (add[0].shape[0], 1)
    assert bool((hit.reshape(-1)[keep]).all())                                     # only rays that found points are grown from
    model.neural_points.grow_points(*add)
    assert model.neural_points.points_xyz.shape[0] == n0 + add[0].shape[0]
    model.get_outputs(rb)                                                          # the grid cache is rebuilt, rendering still works
