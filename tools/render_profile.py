#!/usr/bin/env python
"""torch.profiler over a few full-image render steps: every kernel (ours and torch's glue) with its device time."""
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
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = bench.view(0)
    d = torch.from_numpy(cam.rays(None)).cuda()
    rb = RayBundle.for_camera(d, cam.origin, cam.R_c2w, cam.near, cam.far)
    for _ in range(3):
        model.get_outputs_for_camera_ray_bundle(rb)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    n = 5
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            model.get_outputs_for_camera_ray_bundle(rb)
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.self_device_time_total / n) for e in prof.key_averages() if e.self_device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print(f"device time per render: {tot / 1e3:.3f} ms over {sum(r[1] for r in rows) // n} launches")
    for k, c, t in rows[:40]:
        print(f"{t:10.1f} us  x{c // n:<3d} {k[:110]}")


if __name__ == "__main__":
    main()
