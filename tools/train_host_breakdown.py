#!/usr/bin/env python
"""Where a TrainEngine step's time goes: host enqueue time per phase (no syncs) next to the device time per phase (sync after each)."""
import os
import sys
import time

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

    def phases(sync):
        t = [time.perf_counter()]
        out = model.get_outputs(rb)
        if sync: torch.cuda.synchronize()
        t.append(time.perf_counter())
        ld = model.get_loss_dict(out, {"image": gt})
        loss = sum(ld.values())
        if sync: torch.cuda.synchronize()
        t.append(time.perf_counter())
        loss.backward()
        if sync: torch.cuda.synchronize()
        t.append(time.perf_counter())
        eng.update()
        if sync: torch.cuda.synchronize()
        t.append(time.perf_counter())
        return np.diff(t) * 1e3

    for _ in range(5):
        phases(True)
    for sync in (True, False):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc = np.zeros(4)
        n = 30
        for _ in range(n):
            acc += phases(sync)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / n
        print(f"sync={sync}: get_outputs {acc[0]/n:.3f}  losses {acc[1]/n:.3f}  backward {acc[2]/n:.3f}  update {acc[3]/n:.3f}  | wall per step {wall:.3f} ms", flush=True)
    # memory allocator traffic per step
    s0 = torch.cuda.memory_stats()
    for _ in range(10):
        phases(False)
    torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated"):
        print(k, s1.get(k, 0) - s0.get(k, 0), flush=True)
    print("reserved GB", torch.cuda.memory_reserved() / 2**30, "allocated GB", torch.cuda.memory_allocated() / 2**30)


if __name__ == "__main__":
    main()
