#!/usr/bin/env python
"""torch.profiler over a few TrainEngine steps: which kernels own the device time of a training step."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    from pointnerf2studio_b200.parallel import TrainEngine
    cloud, _ = bench.make_scene()
    model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict()).train()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = bench.view(0)
    pix = np.random.default_rng(100).choice(cam.H * cam.W, size=4096, replace=False)
    rb = bench.to_device(bench.host_bundle(cam, pix), RayBundle)
    gt = torch.rand((4096, 3)).cuda()
    eng = TrainEngine(model, None)
    for _ in range(5):
        eng.step(rb, gt)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            eng.step(rb, gt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))


if __name__ == "__main__":
    main()
