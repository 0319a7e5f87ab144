#!/usr/bin/env python
"""Multi-GPU check of parallel.TrainEngine: the three gradient-exchange routes (NCCL all-reduce, peer-memory kernel with plain
loads / stores, peer-memory kernel with NVLS multimem) must leave the same parameters on every rank after the same steps.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/dp_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    from pointnerf2studio_b200.parallel import TrainEngine
    from pointnerf2studio_b200.synth import make_cloud
    cloud = make_cloud(200000, seed=1241, radii=(0.2, 0.3, 0.38), P=12)
    weights = bench.make_weights()
    cam = bench.view(rank)
    pix = np.random.default_rng(100 + rank).choice(cam.H * cam.W, size=4096, replace=False)
    pix = pix[np.abs(pix // cam.W - 400) < 120]             # rays that hit the small cloud
    gt = torch.rand((len(pix), 3), generator=torch.Generator().manual_seed(rank)).cuda()
    res = {}
    for mode, graph in (("nccl", False), ("p2p_plain", False), ("p2p_multicast", False), ("p2p_multicast", True)):
        os.environ["PNERF_DP_MULTICAST"] = "0" if mode == "p2p_plain" else "1"
        model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict()).train()
        own = dict(model.named_parameters())
        with torch.no_grad():
            for k, v in weights.items():
                own[k].copy_(v)
        eng = TrainEngine(model, dist, exchange="nccl" if mode == "nccl" else "p2p", lr_fields=5e-3, lr_points=2e-2, use_graph=graph)
        losses = []
        for it in range(4):
            rb = RayBundle.for_camera(torch.from_numpy(cam.rays(pix)).cuda(), cam.origin, cam.R_c2w, cam.near, cam.far)
            losses.append(float(eng.step(rb, gt).detach()))
        torch.cuda.synchronize()
        eng.timing = []
        for _ in range(10):
            dist.barrier()
            eng.update()
        torch.cuda.synchronize()
        upd = float(np.median([a.elapsed_time(b) for a, b in eng.timing]))
        eng.timing = None
        flat = eng.P[:eng.total].clone()
        # identical on every rank?
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, flat))
        res[(mode, graph)] = (flat, losses)
        print(f"rank {rank} {mode} graph={graph}: losses {[round(l, 5) for l in losses]} update {upd:.3f} ms, params equal to rank 0: {same}", flush=True)
        assert same, mode
        del eng, model
        torch.cuda.empty_cache()
    base = res[("nccl", False)][0]
    for k, (flat, losses) in res.items():
        d = (flat - base).abs()
        print(f"rank {rank} {k} vs nccl: mean |dp| {float(d.mean()):.3e}, frac > 1e-3: {float((d > 1e-3).float().mean()):.2e}", flush=True)
        assert float(d.mean()) <= 3e-4 and float((d > 2e-3).float().mean()) <= 0.02, k
    dist.barrier()
    if rank == 0:
        print("dp_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
