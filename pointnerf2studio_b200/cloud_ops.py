"""Point-cloud maintenance either side of the per-ray path on libpnerf_b200.so.

  construct_vox_points_closest   models/mvs/mvs_utils.py:537-561 -- the voxel down-sample of neural-point initialisation
                                 (run/gen_pnts.py): same arguments, same return triple
  probe_filter / probe_hole      run/train_studio.py:335-444 -- which probed rays become new points, and their attributes
No CPU fallback: a missing library or a failed launch raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, native
from ._lib import check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def construct_vox_points_closest(xyz_val: torch.Tensor, vox_res, partition_xyz=None, space_min=None, space_max=None):
    """-> (xyz_centroid (V,3) f32, sparse_grid_idx (V,3) i32, min_idx (V) i64): per occupied voxel, in lexicographic voxel order,
    the mean of its points, its integer coordinates, and the index (into xyz_val) of the point closest to the mean.
    With space_min / space_max the reference first drops the points outside the box (MU:545-549) and indexes into the filtered
    array; the same is done here.  `partition_xyz` (a second cloud that only defines the frame, MU:539) is honoured for the
    frame; the reference then mixes the two clouds' lengths (MU:552-554) and only works when both are the same points."""
    lib = _lib.load()
    assert xyz_val.is_cuda and xyz_val.dim() == 2 and xyz_val.shape[1] == 3
    xyz_val = xyz_val.detach().float().contiguous()
    frame_xyz = xyz_val if partition_xyz is None else partition_xyz.detach().float().contiguous()
    f32 = np.float32
    if space_min is None:
        mm = torch.empty(6, dtype=torch.float32, device=xyz_val.device)
        check(lib.pnerf_bbox(_p(frame_xyz), frame_xyz.shape[0], _p(mm), _stream()), "pnerf_bbox")
        mm = mm.cpu().numpy()
        xyz_min, xyz_max = mm[:3], mm[3:]
        edge = f32((xyz_max - xyz_min).max() * f32(1.05))                       # MU:541-542, fp32 like torch
        smin = ((xyz_max + xyz_min) / f32(2) - edge / f32(2)).astype(f32)       # MU:543-544
        edge3 = np.full(3, edge, dtype=f32)
    else:
        smin = np.asarray(space_min.detach().cpu() if torch.is_tensor(space_min) else space_min, dtype=f32).reshape(3)
        smax = np.asarray(space_max.detach().cpu() if torch.is_tensor(space_max) else space_max, dtype=f32).reshape(3)
        edge3 = (smax - smin).astype(f32)
        lo, hi = torch.as_tensor(smin, device=xyz_val.device), torch.as_tensor(smax, device=xyz_val.device)
        keep = torch.prod((xyz_val - lo[None]) * (hi[None] - xyz_val), dim=-1) > 0                             # MU:546-549
        xyz_val = xyz_val[keep].contiguous()
    res = np.broadcast_to(np.asarray(vox_res, dtype=f32), (3,))
    vsz = (edge3 / res).astype(f32)                                              # MU:551
    dim = (C.c_int * 3)(*[int(np.ceil(r)) + 1 for r in res])
    n = xyz_val.shape[0]
    dev = xyz_val.device
    max_out = max(n, 1)
    centroid = torch.empty((max_out, 3), dtype=torch.float32, device=dev)
    grid_idx = torch.empty((max_out, 3), dtype=torch.int32, device=dev)
    min_idx = torch.empty((max_out,), dtype=torch.int32, device=dev)
    counts = torch.zeros((2,), dtype=torch.int32, device=dev)
    ws_bytes = lib.pnerf_vox_closest_workspace_bytes(dim)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    check(lib.pnerf_vox_closest(_p(xyz_val), n, (C.c_float * 3)(*smin), (C.c_float * 3)(*vsz), dim, max_out, _p(centroid), _p(grid_idx),
                                _p(min_idx), C.c_void_p(counts.data_ptr()), C.c_void_p(counts.data_ptr() + 4), _p(ws), ws_bytes, _stream()),
          "pnerf_vox_closest")
    native.LAUNCHES["n"] += 8
    v, outside = [int(x) for x in counts.tolist()]
    if outside:
        raise RuntimeError(f"{outside} points fall outside the voxel frame")
    return centroid[:v], grid_idx[:v], min_idx[:v].long()


def probe_filter(ray_mask, gt_image, coarse_raycolor, ray_max_far_dist, ray_max_shading_opacity, edge_mask, bg, H, W,
                 far_thresh: float = -1.0, opacity_thresh: float = 0.7) -> torch.Tensor:
    """The candidate mask of probe_hole (TS:414-423) over the H x W probe maps -> (H*W,) bool."""
    lib = _lib.load()
    dev = gt_image.device
    rm = ray_mask.reshape(-1).to(torch.int8).contiguous()
    gt = gt_image.reshape(-1, 3).float().contiguous()
    col = None if coarse_raycolor is None else coarse_raycolor.reshape(-1, 3).float().contiguous()
    far = None if ray_max_far_dist is None else ray_max_far_dist.reshape(-1).float().contiguous()
    op = ray_max_shading_opacity.reshape(-1).float().contiguous()
    edge = None if edge_mask is None else edge_mask.reshape(-1).to(torch.uint8).contiguous()
    assert rm.numel() == H * W == op.numel() and gt.shape[0] == H * W
    keep = torch.empty((H * W,), dtype=torch.uint8, device=dev)
    bgv = (C.c_float * 3)(*[float(v) for v in (bg.reshape(-1).tolist() if torch.is_tensor(bg) else bg)])
    check(lib.pnerf_probe_filter(_p(rm), _p(gt), _p(col), _p(far), _p(op), _p(edge), bgv, int(H), int(W), C.c_float(far_thresh),
                                 C.c_float(opacity_thresh), _p(keep), _stream()), "pnerf_probe_filter")
    native.LAUNCHES["n"] += 1
    return keep.bool()


@torch.no_grad()
def probe_hole(model, ray_bundle, gt_image, H, W, edge_mask=None, bg=None, far_thresh: float = -1.0, opacity_thresh: float = 0.7,
               prob_mul: float = 1.0):
    """One frame of the reference's probe_hole (run/train_studio.py:335-444): probe the H x W rays of `ray_bundle` (row-major,
    one camera), select the rays next to a hole (or far from their neighbours) whose densest sample is opaque enough, and return the
    new points (add_xyz, add_embedding, add_color, add_dir, add_conf) ready for NeuralPoints.grow_points (TS:428-432, 706)."""
    out = model.probe(ray_bundle)
    bg = model.get_background_color() if bg is None else bg
    keep = probe_filter(out["ray_mask"], gt_image, out["coarse_raycolor"], out["ray_max_far_dist"], out["ray_max_shading_opacity"],
                        edge_mask, bg, H, W, far_thresh, opacity_thresh)
    idx = torch.nonzero(keep).reshape(-1)
    return (out["ray_max_sample_loc_w"][idx], out["shading_avg_embedding"][idx], out["shading_avg_color"][idx],
            out["shading_avg_dir"][idx], out["shading_avg_conf"][idx] * prob_mul)
