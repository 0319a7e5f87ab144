"""Cross-check of rows G1/G2/Q against the UNMODIFIED reference CUDA extension
(oracle/_ref/query_worldcoords_cuda.so, built by oracle/build_ref.py from the sources under /root/reference).

The reference is racy by construction; it is deterministic as a SET per sample when every voxel holds <= P
points and occupied voxels <= max_o, except that the voxel which wins id 0 loses all its points
(query_worldcoords.cu:147 tests `voxel_idx > 0`).  Which voxel that is changes from run to run, so the test
infers it from the first comparison, removes that voxel's points from OUR buckets (occupancy untouched, as in
the reference) and then demands exact equality of every per-sample index set, sample position and ray mask.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref, field as of, grid_query as gq

pytestmark = pytest.mark.gpu


def _sorted_sets(p):
    return np.sort(p, axis=-1)


@pytest.mark.parametrize("jitter", [0.0, 0.3])
def test_against_reference_extension(jitter):
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref = build_ref.load_module()
    from pointnerf2studio_b200 import native
    from pointnerf2studio_b200.synth import make_camera, make_cloud
    SR, K, P, D = 40, 8, 12, 400
    cloud = make_cloud(50000, seed=77, radii=(0.11, 0.16, 0.2), P=P)
    cam = make_camera()
    rng = np.random.default_rng(9)
    c = cam.H // 2
    pix = rng.integers(c - 70, c + 70, size=1024) * cam.W + rng.integers(c - 70, c + 70, size=1024)
    raypos, _ = of.coarse_positions(torch.from_numpy(cam.origin), torch.from_numpy(cam.rays(pix)), D, 2.0, 6.0, jitter=jitter,
                                    generator=torch.Generator().manual_seed(2))
    frame = gq.hyperparameters(cloud.xyz, [0.004] * 3, [2, 2, 2], [3, 3, 3], [-1.2] * 3 + [1.2] * 3)
    xyz = torch.from_numpy(cloud.xyz).cuda()
    i32 = lambda v: torch.as_tensor(np.asarray(v, dtype=np.int32)).cuda()
    r_pidx, r_loc, r_mask = ref.woord_query_grid_point_index(
        raypos[None].cuda().contiguous(), xyz[None].contiguous(), i32([len(cloud.xyz)]), i32([3, 3, 3]), i32([3, 3, 3]), SR, K,
        len(pix), D, i32(frame.dim), 1000000, P, 0.016, torch.as_tensor(np.concatenate([frame.lo, frame.hi])).cuda(),
        torch.as_tensor(frame.sv).cuda(), 1024, 2)
    torch.cuda.synchronize()
    r_pidx, r_loc, r_mask = r_pidx[0].cpu().numpy(), r_loc[0].cpu().numpy(), r_mask[0].cpu().numpy()

    nf = native.GridFrame(lo=frame.lo, hi=frame.hi, sv=frame.sv, dim=frame.dim)
    full = native.VoxelGrid(xyz, nf, P, [3, 3, 3])
    q = native.sample_and_query(full, len(pix), D, SR, K, 3, 0.016, raypos=raypos.cuda().contiguous())
    m_pidx, m_loc, m_mask, _, _ = native.compact_rays(q)
    m_pidx, m_loc, m_mask = m_pidx.cpu().numpy(), m_loc.cpu().numpy(), m_mask.cpu().numpy()

    same_rays = np.array_equal(m_mask, r_mask)
    diff_pts = np.array([], dtype=np.int64)
    if same_rays:
        a, b = _sorted_sets(m_pidx), _sorted_sets(r_pidx)
        bad = (a != b).any(-1)
        if not bad.any():
            return   # the id-0 voxel was not near any sample: already identical
        diff_pts = np.setdiff1d(m_pidx[bad].reshape(-1), r_pidx[bad].reshape(-1))
    else:
        diff_pts = np.setdiff1d(m_pidx.reshape(-1), r_pidx.reshape(-1))
    diff_pts = diff_pts[diff_pts >= 0]
    assert len(diff_pts) > 0
    vi, _ = gq.voxel_of(cloud.xyz[diff_pts], frame)
    cells = np.unique(gq.cell_id(vi, frame))
    assert len(cells) == 1, f"differences are not confined to one voxel (the reference's id-0 voxel): {cells}"
    # drop that voxel's points from our buckets, keep the occupancy of the full cloud (as the reference does)
    all_vi, all_in = gq.voxel_of(cloud.xyz, frame)
    in_v0 = all_in & (gq.cell_id(all_vi, frame) == cells[0])
    moved = cloud.xyz.copy()
    moved[in_v0] = 50.0
    holed = native.VoxelGrid(torch.from_numpy(moved).cuda(), nf, P, [3, 3, 3])
    holed.occ_bits = full.occ_bits
    holed.view.occ_bits = full.occ_bits.data_ptr()
    q = native.sample_and_query(holed, len(pix), D, SR, K, 3, 0.016, raypos=raypos.cuda().contiguous())
    m_pidx, m_loc, m_mask, _, _ = native.compact_rays(q)
    np.testing.assert_array_equal(m_mask.cpu().numpy(), r_mask)
    np.testing.assert_array_equal(m_loc.cpu().numpy(), r_loc)
    np.testing.assert_array_equal(_sorted_sets(m_pidx.cpu().numpy()), _sorted_sets(r_pidx))
