"""Seeded synthetic neural-point clouds and camera rays.

No neural-point cloud ships with the reference (they are produced per scene by
run/gen_pnts.py from datasets that are not in the repo), so every test and
benchmark uses these: points on a union of three spheres with a small Gaussian
offset along the normal, thinned to at most P points per scaled voxel so the
reference's random bucket overflow (query_worldcoords.cu:152-158) never fires.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SynthCloud:
    xyz: np.ndarray      # (N,3) f32
    embed: np.ndarray    # (N,C) f32   points_embeding
    color: np.ndarray    # (N,3) f32   points_color
    dir: np.ndarray      # (N,3) f32   points_dir (unit normal)
    conf: np.ndarray     # (N,1) f32   points_conf
    Rw2c: np.ndarray     # (3,3) f32   identity (normview=0 in the NeRF-synthetic scripts)
    stats: dict

    def state_dict(self):
        """Checkpoint layout the plugin reads (studio_utils.py:84-90)."""
        import torch
        return {
            "neural_points.xyz": torch.from_numpy(self.xyz),
            "neural_points.points_embeding": torch.from_numpy(self.embed)[None],
            "neural_points.points_conf": torch.from_numpy(self.conf)[None],
            "neural_points.points_dir": torch.from_numpy(self.dir)[None],
            "neural_points.points_color": torch.from_numpy(self.color)[None],
            "neural_points.Rw2c": torch.from_numpy(self.Rw2c),
        }


def _frame(xyz, sv, ks, ranges):
    mn = np.maximum(xyz.min(0), np.float32(ranges[:3]))
    mx = np.minimum(xyz.max(0), np.float32(ranges[3:]))
    half = (sv.astype(np.float64) * np.asarray(ks) / 2).astype(np.float32)
    return (mn - half).astype(np.float32), (mx + half).astype(np.float32)


def make_cloud(n_points: int, seed: int = 1234, feat_dim: int = 32, scaled_vsize: float = 0.008, P: int = 12,
               radii=(0.35, 0.5, 0.65), kernel_size=(3, 3, 3),
               ranges=(-1.2, -1.2, -1.2, 1.2, 1.2, 1.2)) -> SynthCloud:
    rng = np.random.default_rng(seed)
    radii = np.asarray(radii, dtype=np.float64)
    centres = rng.uniform(-0.15, 0.15, size=(len(radii), 3)) * (radii.max() / 0.65)
    area = radii ** 2
    which = rng.choice(len(radii), size=n_points, p=area / area.sum())
    nrm = rng.normal(size=(n_points, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    off = rng.normal(scale=0.5 * scaled_vsize, size=(n_points, 1))
    xyz = (centres[which] + nrm * (radii[which][:, None] + off)).astype(np.float32)
    keep = np.ones(n_points, dtype=bool)
    sv = np.full(3, scaled_vsize, dtype=np.float32)
    for _ in range(8):   # thin to <= P per voxel of the frame the querier will derive from the result
        lo, _hi = _frame(xyz[keep], sv, kernel_size, ranges)
        v = np.floor((xyz - lo) / sv).astype(np.int64)
        key = (v[:, 0] * 4096 + v[:, 1]) * 4096 + v[:, 2]
        key[~keep] = -1
        order = np.argsort(key, kind="stable")
        sk = key[order]
        start = np.r_[0, np.nonzero(sk[1:] != sk[:-1])[0] + 1]
        rank = np.arange(n_points) - np.repeat(start, np.diff(np.r_[start, n_points]))
        drop = order[(rank >= P) & (sk >= 0)]
        if len(drop) == 0:
            break
        keep[drop] = False
    xyz = np.ascontiguousarray(xyz[keep])
    nrm = nrm[keep]
    n = len(xyz)
    lo, hi = _frame(xyz, sv, kernel_size, ranges)
    v = np.floor((xyz - lo) / sv).astype(np.int64)
    key = (v[:, 0] * 4096 + v[:, 1]) * 4096 + v[:, 2]
    _, cnt = np.unique(key, return_counts=True)
    stats = {"n_points": int(n), "occupied_voxels": int(len(cnt)), "mean_pts_per_voxel": float(cnt.mean()),
             "max_pts_per_voxel": int(cnt.max()), "seed": seed, "scaled_vsize": scaled_vsize}
    return SynthCloud(
        xyz=xyz,
        embed=rng.normal(scale=0.3, size=(n, feat_dim)).astype(np.float32),
        color=rng.uniform(0, 1, size=(n, 3)).astype(np.float32),
        dir=nrm.astype(np.float32),
        conf=rng.uniform(0.1, 1.0, size=(n, 1)).astype(np.float32),
        Rw2c=np.eye(3, dtype=np.float32),
        stats=stats,
    )


def make_cloud_state_dict_torch(n_points: int, seed: int = 1234, device="cuda", feat_dim: int = 32, scaled_vsize: float = 0.008, P: int = 12,
                                radii=(0.35, 0.5, 0.65), kernel_size=(3, 3, 3), ranges=(-1.2, -1.2, -1.2, 1.2, 1.2, 1.2)):
    """The same kind of cloud as make_cloud (points on a union of spheres + Gaussian normal offset, thinned to <= P points per scaled
    voxel), generated with torch on `device` -- seconds instead of minutes for 10 M points.  Different random stream than the numpy
    version: same statistics, not the same points.  -> (checkpoint-layout state dict on `device`, stats)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    dev = torch.device(device)
    rad = torch.tensor(radii, dtype=torch.float64, device=dev)
    centres = (torch.rand((len(radii), 3), generator=g, device=dev, dtype=torch.float64) * 0.3 - 0.15) * float(max(radii) / 0.65)
    area = rad ** 2
    which = torch.multinomial(area / area.sum(), n_points, replacement=True, generator=g)
    nrm = torch.randn((n_points, 3), generator=g, device=dev, dtype=torch.float32)
    nrm = nrm / nrm.norm(dim=1, keepdim=True)
    off = torch.randn((n_points, 1), generator=g, device=dev, dtype=torch.float32) * (0.5 * scaled_vsize)
    xyz = (centres[which].float() + nrm * (rad[which].float()[:, None] + off)).contiguous()
    keep = torch.ones(n_points, dtype=torch.bool, device=dev)
    sv = torch.full((3,), scaled_vsize, dtype=torch.float32, device=dev)
    half = (sv.double() * torch.tensor(kernel_size, device=dev) / 2).float()
    r_lo = torch.tensor(ranges[:3], dtype=torch.float32, device=dev)
    ar = torch.arange(n_points, device=dev)
    for _ in range(4):          # thin to <= P per voxel of the frame the querier will derive from the result
        lo = torch.maximum(xyz[keep].min(0)[0], r_lo) - half
        v = torch.floor((xyz - lo) / sv).long()
        key = (v[:, 0] * 8192 + v[:, 1]) * 8192 + v[:, 2]
        key = torch.where(keep, key, torch.full_like(key, -1))
        sk, order = torch.sort(key, stable=True)
        new_run = torch.ones(n_points, dtype=torch.bool, device=dev)
        new_run[1:] = sk[1:] != sk[:-1]
        run_start = torch.cummax(torch.where(new_run, ar, torch.zeros_like(ar)), 0)[0]
        drop = order[((ar - run_start) >= P) & (sk >= 0)]
        if drop.numel() == 0:
            break
        keep[drop] = False
    xyz, nrm = xyz[keep].contiguous(), nrm[keep].contiguous()
    n = xyz.shape[0]
    _, cnt = torch.unique(key[keep], return_counts=True)
    stats = {"n_points": int(n), "occupied_voxels": int(cnt.numel()), "mean_pts_per_voxel": float(cnt.float().mean()),
             "max_pts_per_voxel": int(cnt.max()), "seed": seed, "scaled_vsize": scaled_vsize, "generator": "torch/" + str(dev.type)}
    sd = {"neural_points.xyz": xyz,
          "neural_points.points_embeding": (torch.randn((n, feat_dim), generator=g, device=dev) * 0.3)[None],
          "neural_points.points_conf": (torch.rand((n, 1), generator=g, device=dev) * 0.9 + 0.1)[None],
          "neural_points.points_dir": nrm[None],
          "neural_points.points_color": torch.rand((n, 3), generator=g, device=dev)[None],
          "neural_points.Rw2c": torch.eye(3, device=dev)}
    return sd, stats


@dataclass
class SynthCamera:
    origin: np.ndarray    # (3,) f32
    R_c2w: np.ndarray     # (3,3) f32, columns = camera x (right), y (down), z (forward) in world
    H: int
    W: int
    focal: float
    near: float = 2.0
    far: float = 6.0

    def rays(self, pix=None):
        """Unit world-space directions for pixel ids `pix` (row-major), or for the whole image."""
        if pix is None:
            pix = np.arange(self.H * self.W)
        pix = np.asarray(pix)
        i, j = pix // self.W, pix % self.W
        d = np.stack([(j + 0.5 - self.W / 2) / self.focal, (i + 0.5 - self.H / 2) / self.focal,
                      np.ones(len(pix))], axis=-1)
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        return (d @ self.R_c2w.astype(np.float64).T).astype(np.float32)


def make_camera(H=800, W=800, focal=1111.1, radius=4.0, azim_deg=30.0, elev_deg=20.0, near=2.0, far=6.0):
    """Pin-hole camera on a sphere of `radius` looking at the origin (Blender-like framing)."""
    a, e = np.deg2rad(azim_deg), np.deg2rad(elev_deg)
    o = radius * np.array([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)])
    z = -o / np.linalg.norm(o)
    x = np.cross(z, np.array([0.0, 0.0, 1.0]))
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return SynthCamera(origin=o.astype(np.float32), R_c2w=np.stack([x, y, z], axis=1).astype(np.float32),
                       H=H, W=W, focal=focal, near=near, far=far)
