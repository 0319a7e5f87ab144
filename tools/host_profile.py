#!/usr/bin/env python
"""cProfile of the host side of bench.py's training step (where does the CPU time between kernel launches go?)."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
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
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        out = model.get_outputs(rb)
        ld = model.get_loss_dict(out, {"image": gt})
        sum(ld.values()).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
    print(s.getvalue()[:6000])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
    print(s.getvalue()[:5000])


if __name__ == "__main__":
    main()
