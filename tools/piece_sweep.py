#!/usr/bin/env python
"""does running the colour kernel right behind its field-kernel piece (F kept in L2) pay?
native_tc.MAX_SAMPLES_PER_LAUNCH bounds the samples per (field launches + colour launch) piece and the F workspace is reused by
every piece, so a small piece size IS that experiment: render the bench view with several piece sizes on one box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle, native, native_tc
    cloud, _ = bench.make_scene()
    model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict()).eval()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = bench.view(0)
    rb = bench.to_device(bench.host_bundle(cam, np.arange(cam.H * cam.W)), RayBundle)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for piece in (1 << 23, 1 << 20, 400_000, 200_000, 100_000, 1 << 23):
        native_tc.MAX_SAMPLES_PER_LAUNCH = piece
        for _ in range(2):
            out = model.get_outputs_for_camera_ray_bundle(rb)["coarse_raycolor"]
        torch.cuda.synchronize()
        native.Timers.enabled, native.Timers.spans = True, []
        evs = []
        for _ in range(8):
            flush.add_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = model.get_outputs_for_camera_ray_bundle(rb)["coarse_raycolor"]
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        sp = native.Timers.collect()
        native.Timers.enabled = False
        ms = sorted(x.elapsed_time(y) for x, y in evs)[len(evs) // 2]
        field = sum(sp["field"]) / 8
        # (piece size does not change the bits -- tests/test_gpu_tc.py::test_tc_sample_pieces_are_bit_identical; here every call
        # draws a new jitter, so the pixels of two calls are not comparable)
        print(f"piece {piece:8d} samples ({piece * 512 / 1e6:7.1f} MB of F): step {ms:6.3f} ms, field stage {field:6.3f} ms", flush=True)


if __name__ == "__main__":
    main()
