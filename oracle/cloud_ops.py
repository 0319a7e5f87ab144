"""Oracle (numpy, CPU) for the point-cloud maintenance rows of SURVEY.md 8f -- TEST INFRASTRUCTURE ONLY.

  probe_filter                  restates the tensor code of probe_hole, /root/reference/pointnerf/run/train_studio.py:414-423 with
                                bloat_inds :447-455 (TS below)
  construct_vox_points_closest  restates /root/reference/pointnerf/models/mvs/mvs_utils.py:537-561 (MU below); torch_scatter
                                (scatter_mean / scatter_min, not installed here) is restated from its documented semantics:
                                mean per index; minimum per index with the FIRST minimal element's position as the argument.

Pinned by tests/golden/cloud_ops_golden.npz, which tests/golden/make_golden_cloud_ops.py produced by EXECUTING the reference's own
source lines on CPU (tests/test_oracle_cloud_ops.py).
"""
from __future__ import annotations

import numpy as np


def bloat_mask(miss: np.ndarray) -> np.ndarray:
    """TS:447-455 + TS:418-419: pixels within one step (8-neighbourhood, indices clamped at the border) of a True pixel."""
    H, W = miss.shape
    out = np.zeros((H, W), dtype=bool)
    ys, xs = np.nonzero(miss)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            out[np.clip(ys + dy, 0, H - 1), np.clip(xs + dx, 0, W - 1)] = True
    return out


def probe_filter(ray_mask, gt, color, far_dist, opacity, edge_mask, bg, far_thresh, opacity_thresh):
    """ray_mask (H,W) int, gt / color (H,W,3) f32, far_dist / opacity (H,W) f32, edge_mask (H,W) bool, bg (3,) -> keep (H,W) bool."""
    f32 = np.float32
    gt, color, bg = gt.astype(f32), color.astype(f32), np.asarray(bg, f32)
    n_bg = np.sqrt(((gt - bg) ** 2).sum(-1, dtype=f32))
    miss = (ray_mask < 1) & (n_bg > f32(0.002)) & edge_mask                     # TS:414-415
    near_miss = bloat_mask(miss).astype(np.int32)                                # TS:417-419
    if far_thresh > 0:                                                           # TS:420-422
        n_c = np.sqrt(((gt - color) ** 2).sum(-1, dtype=f32))
        near_miss = near_miss + ((ray_mask > 0) & (far_dist > f32(far_thresh)) & (n_c < f32(0.1)))
    return (ray_mask > 0) & (near_miss > 0) & (opacity > f32(opacity_thresh))    # TS:423


def scatter_mean(src, index, n):
    out = np.zeros((n, src.shape[1]), np.float64)
    np.add.at(out, index, src.astype(np.float64))
    cnt = np.bincount(index, minlength=n).astype(np.float64)
    return (out / cnt[:, None]).astype(np.float32)


def scatter_min_arg(src, index, n):
    best = np.full(n, np.inf, np.float32)
    arg = np.full(n, -1, np.int64)
    for i in range(len(src)):                 # first minimum wins (strict <), as torch_scatter's CPU kernel
        if src[i] < best[index[i]]:
            best[index[i]], arg[index[i]] = src[i], i
    return arg


def construct_vox_points_closest(xyz_val, vox_res):
    """MU:537-561 with space_min = None.  -> (xyz_centroid (V,3) f32, sparse_grid_idx (V,3) i32, min_idx (V) i64, space_min, vox_sz)."""
    f32 = np.float32
    xyz = xyz_val.astype(f32)
    xyz_min, xyz_max = xyz.min(0), xyz.max(0)
    space_edge = f32((xyz_max - xyz_min).max() * f32(1.05))
    xyz_mid = (xyz_max + xyz_min) / f32(2)
    space_min = (xyz_mid - space_edge / f32(2)).astype(f32)
    vox_sz = f32(space_edge / f32(vox_res))
    vi = np.floor(((xyz - space_min[None]).astype(f32) / vox_sz).astype(f32)).astype(np.int32)
    grid_idx, inv = np.unique(vi, axis=0, return_inverse=True)                   # sorted lexicographically, like torch.unique(dim=0)
    inv = inv.reshape(-1)
    centroid = scatter_mean(xyz, inv, len(grid_idx))
    d = xyz - centroid[inv]
    residual = np.sqrt((d * d).sum(-1, dtype=f32)).astype(f32)
    return centroid, grid_idx.astype(np.int32), scatter_min_arg(residual, inv, len(grid_idx)), space_min, vox_sz
