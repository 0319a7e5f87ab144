#!/usr/bin/env python
"""Where a training step's time goes at N ranks: device time of forward+backward / gradient all-reduce / Adam (CUDA events)
next to the host wall clock of the same phases.  Same workload as `bench.py --workload train`.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
           tools/train_breakdown.py [--steps 20] [--sparse]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    from pointnerf2studio_b200.optim import make_optimizers
    from pointnerf2studio_b200.parallel import allreduce_gradients
    cloud, _ = bench.make_scene()
    model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict())
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    model.train()
    cam = bench.view(rank)
    pix = np.random.default_rng(100 + rank).choice(cam.H * cam.W, size=bench.TRAIN_RAYS, replace=False)
    rb = bench.to_device(bench.host_bundle(cam, pix), RayBundle)
    gt = torch.rand((bench.TRAIN_RAYS, 3), generator=torch.Generator().manual_seed(9)).cuda()
    params = [p for p in model.parameters() if p.requires_grad]
    opts, scheds = make_optimizers(model)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    dev = np.zeros(3)
    host = np.zeros(3)
    for it in range(args.steps + 5):
        e = [ev() for _ in range(4)]
        t = [0.0] * 4
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t[0] = time.perf_counter(); e[0].record()
        for p in params:
            p.grad = None
        out = model.get_outputs(rb)
        loss = sum(model.get_loss_dict(out, {"image": gt}).values())
        loss.backward()
        t[1] = time.perf_counter(); e[1].record()
        if dist is not None:
            allreduce_gradients(params, dist)
        t[2] = time.perf_counter(); e[2].record()
        for k in opts:
            opts[k].step()
            scheds[k].step()
        t[3] = time.perf_counter(); e[3].record()
        torch.cuda.synchronize()
        if it >= 5:
            dev += [e[i].elapsed_time(e[i + 1]) for i in range(3)]
            host += [1e3 * (t[i + 1] - t[i]) for i in range(3)]
    dev /= args.steps
    host /= args.steps
    print(f"rank {rank}/{world} cores {os.cpu_count()}: device ms fwd+bwd {dev[0]:.2f} allreduce {dev[1]:.2f} adam {dev[2]:.2f} | "
          f"host ms fwd+bwd {host[0]:.2f} allreduce {host[1]:.2f} adam {host[2]:.2f}", flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
