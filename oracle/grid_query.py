"""Oracle (numpy) for the voxel-grid neighbour querier.  TEST INFRASTRUCTURE ONLY.

Restates rows H, G1, G2, Q of SURVEY.md section 8a.  Reference files (under
/root/reference/pointnerf/):
  SU  = nerfstudio/studio_utils.py
  CU  = models/neural_points/cuda/query_worldcoords.cu

Deterministic rule where the reference is racy (DESIGN.md "Tie-break rule"):
  * a voxel keeps its first P points in ascending point index (CU:149-158 keeps a
    random subset when a voxel holds more than P);
  * every occupied voxel is kept (CU:64-73 drops random voxels beyond max_o) and the
    voxel that wins id 0 keeps its points (CU:147 loses them);
  * candidates are visited in (layer, ux, uy, uz, point index) order, the K kept are
    the K smallest by (d2, visit order), and they are emitted sorted by (d2, index).

All float arithmetic is IEEE fp32 exactly as the kernels do it; the squared distance
uses the FMA chain nvcc emits for ``x*x + y*y + z*z`` (fmul, ffma, ffma), emulated in
float64 here (exact in ``query_oracle.c`` through fmaf).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


@dataclass
class GridFrame:
    """Grid origin / voxel size / dims (SU:115-127)."""
    lo: np.ndarray      # (3,) f32  ranges_tensor[:3]
    hi: np.ndarray      # (3,) f32  ranges_tensor[3:]
    sv: np.ndarray      # (3,) f32  scaled voxel size
    dim: np.ndarray     # (3,) i32  scaled_vdim

    @property
    def cells(self) -> int:
        return int(self.dim[0]) * int(self.dim[1]) * int(self.dim[2])


def hyperparameters(xyz, vsize, vscale, kernel_size, ranges) -> GridFrame:
    """Grid frame from the point bounding box (SU:115-127, twin PQ:47-71).

    dtype walk of the reference: min/max and the +-margin are fp32 torch ops; the
    margin itself is float64 numpy (f32 array * python list / 2) cast to fp32; the
    dims are float64 numpy: ceil(((hi-lo) as f32 -> f64) / vsize / vscale).
    """
    xyz = _f32(xyz).reshape(-1, 3)
    mn, mx = xyz.min(0), xyz.max(0)
    if ranges is not None:
        r = _f32(ranges)
        mn, mx = np.maximum(mn, r[:3]), np.minimum(mx, r[3:])
    vscale_i = np.asarray(vscale, dtype=np.int32)
    sv = (np.asarray(vsize, dtype=np.float64) * vscale_i).astype(np.float32)          # SU:112
    half = (sv.astype(np.float64) * np.asarray(kernel_size, dtype=np.int64) / 2).astype(np.float32)  # SU:121
    lo = (mn - half).astype(np.float32)
    hi = (mx + half).astype(np.float32)
    vdim = (hi - lo).astype(np.float32).astype(np.float64) / np.asarray(vsize, dtype=np.float64)     # SU:125
    dim = np.ceil(vdim / vscale_i).astype(np.int32)                                     # SU:126
    return GridFrame(lo=lo, hi=hi, sv=sv, dim=dim)


def voxel_of(pos, frame: GridFrame):
    """(int) floor((p - lo) / sv) in fp32 (CU:38-40) and the inside-grid test (CU:44)."""
    pos = _f32(pos)
    with np.errstate(invalid="ignore", over="ignore"):
        v = np.floor((pos - frame.lo) / frame.sv)
        inside = np.all((v >= 0) & (v < frame.dim.astype(np.float32)), axis=-1)
        vi = np.where(inside[..., None], v, 0).astype(np.int32)
    return vi, inside


def cell_id(vi, frame: GridFrame):
    d = frame.dim.astype(np.int64)
    vi = vi.astype(np.int64)
    return vi[..., 0] * (d[1] * d[2]) + vi[..., 1] * d[2] + vi[..., 2]                 # CU:45


@dataclass
class Grid:
    frame: GridFrame
    P: int
    cell_bucket: np.ndarray   # (G,) i32: bucket id of the cell or -1   (coor_2_occ)
    bucket_pts: np.ndarray    # (V,P) i32 point ids ascending, -1 padded (occ_2_pnts)
    bucket_cnt: np.ndarray    # (V,) i32 min(count, P)                   (occ_numpnts)
    bucket_cell: np.ndarray   # (V,) i64 cell id of each bucket, ascending
    occ: np.ndarray           # (dimx,dimy,dimz) bool dilated occupancy  (coor_occ)
    n_dropped: int            # points outside the clipped grid


def build_grid(xyz, frame: GridFrame, P: int, query_size) -> Grid:
    """Buckets + dilated occupancy (claim_occ CU:18-78, map_coor2occ CU:80-115,
    fill_occ2pnts CU:117-162) under the deterministic rule in the module docstring."""
    xyz = _f32(xyz).reshape(-1, 3)
    vi, inside = voxel_of(xyz, frame)
    cid = cell_id(vi, frame)
    idx = np.nonzero(inside)[0]
    order = idx[np.argsort(cid[idx], kind="stable")]          # ascending (cell, point index)
    scid = cid[order]
    ucell, start, cnt = np.unique(scid, return_index=True, return_counts=True)
    V = len(ucell)
    bucket_pts = np.full((V, P), -1, dtype=np.int32)
    rank = np.arange(len(order)) - np.repeat(start, cnt)
    keep = rank < P
    bidx = np.repeat(np.arange(V), cnt)
    bucket_pts[bidx[keep], rank[keep]] = order[keep].astype(np.int32)
    cell_bucket = np.full(frame.cells, -1, dtype=np.int32)
    cell_bucket[ucell] = np.arange(V, dtype=np.int32)
    # dilation (CU:105-112): v in [u - q/2, u + (q+1)/2)
    d = frame.dim
    occ0 = np.zeros(tuple(int(x) for x in d), dtype=bool)
    occ0.reshape(-1)[ucell] = True
    occ = np.zeros_like(occ0)
    q = [int(x) for x in query_size]
    for ox in range(-(q[0] // 2), (q[0] + 1) // 2):
        for oy in range(-(q[1] // 2), (q[1] + 1) // 2):
            for oz in range(-(q[2] // 2), (q[2] + 1) // 2):
                src = occ0[max(0, -ox):d[0] - max(0, ox), max(0, -oy):d[1] - max(0, oy), max(0, -oz):d[2] - max(0, oz)]
                occ[max(0, ox):d[0] - max(0, -ox), max(0, oy):d[1] - max(0, -oy), max(0, oz):d[2] - max(0, -oz)] |= src
    return Grid(frame=frame, P=P, cell_bucket=cell_bucket, bucket_pts=bucket_pts,
                bucket_cnt=np.minimum(cnt, P).astype(np.int32), bucket_cell=ucell.astype(np.int64),
                occ=occ, n_dropped=int((~inside).sum()))


def select_samples(raypos, grid: Grid, SR: int):
    """First SR coarse positions per ray that fall in dilated occupancy
    (mask_raypos CU:165-189, cumsum CU:390-391, get_shadingloc CU:192-214).

    raypos (R,D,3) -> sample_loc (R,SR,3) zero filled, sample_mask (R,SR) bool,
    ray_hit (R,) bool (membership of R'), sample_t_index (R,SR) i32 (-1 if empty).
    """
    raypos = _f32(raypos)
    R, D, _ = raypos.shape
    vi, inside = voxel_of(raypos, grid.frame)
    hit = inside & grid.occ[vi[..., 0], vi[..., 1], vi[..., 2]]
    rank = np.cumsum(hit, axis=1)
    take = hit & (rank <= SR)
    sample_loc = np.zeros((R, SR, 3), dtype=np.float32)
    sample_mask = np.zeros((R, SR), dtype=bool)
    sample_j = np.full((R, SR), -1, dtype=np.int32)
    r, j = np.nonzero(take)
    s = rank[r, j] - 1
    sample_loc[r, s] = raypos[r, j]
    sample_mask[r, s] = True
    sample_j[r, s] = j
    return sample_loc, sample_mask, hit.any(axis=1), sample_j


def _d2_fma(p, q):
    """fl(fma(dz,dz, fl(fma(dy,dy, fl(dx*dx))))) with fp32 dx,dy,dz (CU:268-271 as nvcc contracts it)."""
    d = (p.astype(np.float32) - q.astype(np.float32)).astype(np.float64)
    t = (d[..., 0] * d[..., 0]).astype(np.float32).astype(np.float64)
    t = (d[..., 1] * d[..., 1] + t).astype(np.float32).astype(np.float64)
    return (d[..., 2] * d[..., 2] + t).astype(np.float32)


def layer_offsets(layer: int):
    """Voxel offsets of Chebyshev shell `layer`, in the kernel's x,y,z ascending order (CU:257-263)."""
    rng = range(-layer, layer + 1)
    return [(x, y, z) for x in rng for y in rng for z in rng if max(abs(x), abs(y), abs(z)) == layer]


def query_neighbours(sample_loc, sample_mask, xyz, grid: Grid, K: int, kernel_size, radius: float,
                     chunk: int = 8192):
    """Layer-truncated, bucket-capped, radius-limited K nearest (CU:217-302).

    Returns sample_pidx (R,SR,K) i32 (-1 padded, sorted by (d2, index)) and, for
    statistics, per-sample (#voxel entries visited, #candidate points examined).
    """
    xyz = _f32(xyz).reshape(-1, 3)
    R, SR, _ = sample_loc.shape
    r2 = np.float32(np.float32(radius) * np.float32(radius))                 # CU:410
    L = (int(kernel_size[0]) + 1) // 2                                          # CU:256
    out = np.full((R * SR, K), -1, dtype=np.int32)
    n_vis = np.zeros(R * SR, dtype=np.int32)
    n_cand = np.zeros(R * SR, dtype=np.int32)
    flat_loc = _f32(sample_loc).reshape(-1, 3)
    valid_ids = np.nonzero(np.asarray(sample_mask).reshape(-1))[0]
    dim = grid.frame.dim.astype(np.int64)
    P = grid.P
    for c0 in range(0, len(valid_ids), chunk):
        ids = valid_ids[c0:c0 + chunk]
        q = flat_loc[ids]
        v, _ = voxel_of(q, grid.frame)
        v = v.astype(np.int64)
        S = len(ids)
        seen = np.zeros(S, dtype=np.int64)
        active = np.ones(S, dtype=bool)
        d2_cols, id_cols = [], []
        for layer in range(L):
            for off in layer_offsets(layer):
                u = v + np.asarray(off, dtype=np.int64)
                ins = np.all((u >= 0) & (u < dim), axis=1) & active
                cid = np.where(ins, u[:, 0] * dim[1] * dim[2] + u[:, 1] * dim[2] + u[:, 2], 0)
                b = np.where(ins, grid.cell_bucket[cid], -1)
                n_vis[ids] += ins
                pts = np.where((b >= 0)[:, None], grid.bucket_pts[np.maximum(b, 0)], -1)    # (S,P)
                has = pts >= 0
                n_cand[ids] += has.sum(1)
                d2 = _d2_fma(xyz[np.maximum(pts, 0)], q[:, None, :])
                inr = has & ((r2 == 0) | (d2 <= r2))                                        # CU:273
                seen += inr.sum(1)
                d2_cols.append(np.where(inr, d2, np.float32(np.inf)))
                id_cols.append(np.where(inr, pts, -1))
            active &= seen < K                                                              # CU:300
        d2_all = np.concatenate(d2_cols, axis=1)
        id_all = np.concatenate(id_cols, axis=1)
        pick = np.argsort(d2_all, axis=1, kind="stable")[:, :K]       # K smallest by (d2, visit order)
        pd2 = np.take_along_axis(d2_all, pick, 1)
        pid = np.take_along_axis(id_all, pick, 1)
        if pd2.shape[1] < K:
            pad = K - pd2.shape[1]
            pd2 = np.pad(pd2, ((0, 0), (0, pad)), constant_values=np.inf)
            pid = np.pad(pid, ((0, 0), (0, pad)), constant_values=-1)
        key_id = np.where(pid >= 0, pid, np.iinfo(np.int32).max)
        emit = np.lexsort((key_id, pd2), axis=1)                      # emit sorted by (d2, index)
        out[ids] = np.take_along_axis(pid, emit, 1)
    return out.reshape(R, SR, K), n_vis.reshape(R, SR), n_cand.reshape(R, SR)


def compact_rays(sample_pidx, sample_loc, ray_hit):
    """Drop rays without any neighbour (CU:425-432): returns compact pidx/loc of the R''
    surviving rays and ray_mask (R,) int8."""
    R = len(ray_hit)
    has = (sample_pidx >= 0).any(axis=(1, 2)) & ray_hit
    return sample_pidx[has], sample_loc[has], has.astype(np.int8)


def woord_query_grid_point_index(raypos, xyz, kernel_size, query_size, SR, K, frame: GridFrame, P, radius):
    """Whole native op (CPP:33-50 -> CU:305-433) on the CPU: (pidx (R'',SR,K), loc (R'',SR,3), ray_mask (R,))."""
    grid = build_grid(xyz, frame, P, query_size)
    loc, mask, ray_hit, _ = select_samples(raypos, grid, SR)
    pidx, _, _ = query_neighbours(loc, mask, xyz, grid, K, kernel_size, radius)
    return compact_rays(pidx, loc, ray_hit)
