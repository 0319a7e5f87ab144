#!/usr/bin/env python
"""BASELINE.json configs[4] -- Tanks-and-Temples-scale stress: ~10 M-point synthetic cloud, K = 16, SR = 80, vsize 0.002 x vscale 2,
kernel 5^3 (3 layers, 125 voxels), P = 10 (dev_scripts/w_tt_ft/truck_points.sh:53-63); sweep rays in {4 k, 64 k, 1 M} and time the
neighbour query and the aggregation (fused field kernels) separately with CUDA events.  Prints one JSON line per ray count."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--rays", type=int, nargs="+", default=[4096, 65536, 1048576])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle, native
    from pointnerf2studio_b200.synth import make_camera, make_cloud
    t0 = time.time()
    # surface area of the three spheres ~ 10 units^2 -> ~2.5 M voxels of 0.004 on the shells: ~4 points per voxel at 10 M
    cloud = make_cloud(args.points, seed=1239, scaled_vsize=0.004, P=10, radii=(0.45, 0.65, 0.85), kernel_size=(5, 5, 5))
    print(f"cloud: {cloud.stats} ({time.time() - t0:.0f} s)", file=sys.stderr)
    cfg = PointNerfConfig(K=16, SR=80, P=10, vsize=[0.002] * 3, kernel_size=[5, 5, 5], max_o=1600000, precision="bf16")
    model = PointNerf(cfg, state_dict=cloud.state_dict()).eval()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = make_camera(H=1024, W=1024, focal=1422.0)
    peaks = bench.load_peaks()
    for R in args.rays:
        pix = np.random.default_rng(R).choice(cam.H * cam.W, size=R, replace=False) if R < cam.H * cam.W else np.arange(cam.H * cam.W)
        rb = bench.to_device(bench.host_bundle(cam, np.sort(pix), pinned=False), RayBundle)
        with torch.no_grad():
            model.get_outputs(rb)                         # builds the grid, warms up
            q, _, _, _ = model.neural_points.query(rb, want_stats=True)
            torch.cuda.synchronize()
            S, M = int((q.sample_valid > 0).sum()), int((q.sample_pidx >= 0).sum())
            filled = int(q.sample_cnt.sum())
            vis, cand = [int(x) for x in q.stats.tolist()]
            del q
            native.Timers.enabled, native.Timers.spans = True, []
            for _ in range(args.reps):
                model.get_outputs(rb)
            torch.cuda.synchronize()
            sp = {k: sum(v) / len(v) for k, v in native.Timers.collect().items()}
            native.Timers.enabled = False
        q_bytes = 12.0 * filled + 4.0 * vis + 16.0 * cand + 4.0 * cfg.K * filled
        flops = (542208.0 + 512.0) * M + 137984.0 * S
        print(json.dumps({"rays": R, "n_points": int(cloud.xyz.shape[0]), "K": 16, "SR": 80, "kernel": "5x5x5", "filled_slots": filled,
                          "valid_samples": S, "neighbour_rows": M, "mean_voxels_visited": vis / max(filled, 1),
                          "mean_candidates": cand / max(filled, 1), "ms": sp,
                          "query_samples_per_s": filled / (sp["query"] * 1e-3), "query_algorithmic_GBps": q_bytes / (sp["query"] * 1e-3) / 1e9,
                          "query_frac_of_hbm": q_bytes / (sp["query"] * 1e-3) / 1e9 / peaks["hbm"],
                          "aggregate_rows_per_s": M / (sp["field"] * 1e-3), "aggregate_TFLOPs": flops / (sp["field"] * 1e-3) / 1e12,
                          "aggregate_frac_of_bf16": flops / (sp["field"] * 1e-3) / 1e12 / peaks["tensor"]}), flush=True)


if __name__ == "__main__":
    main()
