#!/usr/bin/env python
"""Histogram of valid neighbours per valid sample for the bench view (how many MMA rows are padding at KP = 8?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    cloud, _ = bench.make_scene()
    model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict()).eval()
    for rank in range(2):
        cam = bench.view(rank)
        rb = RayBundle.for_camera(torch.from_numpy(cam.rays(None)).cuda(), cam.origin, cam.R_c2w, cam.near, cam.far)
        with torch.no_grad():
            q, _, _, _ = model.neural_points.query(rb, compact=True)
        cnt = (q.sample_pidx >= 0).sum(-1).reshape(-1)
        cnt = cnt[cnt > 0]
        h = torch.bincount(cnt, minlength=9).cpu().numpy()
        S, M = int(cnt.numel()), int(cnt.sum())
        print(f"view {rank}: S = {S}, M = {M}, rows at KP=8: {8 * S} ({100 * (8 * S - M) / (8 * S):.1f} % padding)")
        print("  neighbours per sample 1..8:", h[1:].tolist(), " fractions:", np.round(h[1:] / S, 3).tolist())
        le4 = int(h[1:5].sum())
        rows_split = 4 * le4 + 8 * (S - le4)
        print(f"  KP=4 tiles for the {le4} samples with <= 4 neighbours: {rows_split} rows ({100 * (rows_split - M) / rows_split:.1f} % padding, "
              f"{100 * (1 - rows_split / (8 * S)):.1f} % fewer rows)")
        # per-ray structure: samples per hit ray
        per_ray = (q.sample_valid > 0).sum(1)
        print("  valid samples per hit ray: mean %.1f, max %d" % (float(per_ray.float().mean()), int(per_ray.max())))


if __name__ == "__main__":
    main()
