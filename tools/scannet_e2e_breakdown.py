#!/usr/bin/env python
"""Host wall-clock breakdown of the end-to-end step of `bench.py --workload scannet` at N ranks (upload / render / all-gather /
download), to see where an end-to-end number that is far above the device-timed one spends its time.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 tools/scannet_e2e_breakdown.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    from pointnerf2studio_b200.parallel import gather_interleaved_image, interleaved_rows
    from pointnerf2studio_b200.synth import make_camera, make_cloud
    cloud = make_cloud(3_000_000, seed=1237, scaled_vsize=0.016, P=30, radii=(0.5, 0.72, 0.93))
    model = PointNerf(PointNerfConfig(precision="bf16", vsize=[0.008] * 3, P=30, SR=24), state_dict=cloud.state_dict()).eval()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = make_camera(H=968, W=1296, focal=1170.0, azim_deg=30.0, elev_deg=20.0)
    rows = np.asarray(interleaved_rows(cam.H, rank, world))
    pix = (rows[:, None] * cam.W + np.arange(cam.W)[None]).reshape(-1)
    host = bench.host_bundle(cam, pix)
    out_host = torch.empty((cam.H * cam.W, 3), dtype=torch.float32).pin_memory()
    acc = np.zeros(4)
    n = 0
    for it in range(13):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rb = bench.to_device(host, RayBundle)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        o = model.get_outputs_for_camera_ray_bundle(rb)["coarse_raycolor"]
        torch.cuda.synchronize(); t2 = time.perf_counter()
        img = gather_interleaved_image(o, cam.H, cam.W, dist)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        out_host.copy_(img, non_blocking=True)
        torch.cuda.synchronize(); t4 = time.perf_counter()
        if it >= 3:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
            n += 1
    acc *= 1e3 / n
    print(f"rank {rank}/{world}: upload {acc[0]:.2f} render {acc[1]:.2f} all-gather {acc[2]:.2f} download {acc[3]:.2f} ms", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
