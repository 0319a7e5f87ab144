#!/usr/bin/env python
"""Decode the pipeline trace of field_tc_kernel (pnerf_tc_set_trace): run the bench render workload once with the
trace on and print, for CTA 0, where each role warp spends its cycles in steady state."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle, _lib
    lib = _lib.load()
    cloud, _ = bench.make_scene()
    model = PointNerf(PointNerfConfig(precision="bf16"), state_dict=cloud.state_dict()).eval()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in bench.make_weights().items():
            own[k].copy_(v)
    cam = bench.view(0)
    rb = bench.to_device(bench.host_bundle(cam, np.arange(cam.H * cam.W)), RayBundle)
    model.get_outputs_for_camera_ray_bundle(rb)
    torch.cuda.synchronize()
    n = lib.pnerf_tc_trace_bytes() // 8
    buf = torch.zeros(n, dtype=torch.int64, device="cuda")
    lib.pnerf_tc_set_trace(buf.data_ptr())
    model.get_outputs_for_camera_ray_bundle(rb)
    torch.cuda.synchronize()
    lib.pnerf_tc_set_trace(None)
    t = buf.cpu().numpy().reshape(32, -1)
    names = {0: "encoder", 4: "epilogue slot0 lo", 12: "epilogue slot1 lo", 21: "mma issuer"}
    for w, name in names.items():
        cnt = int(t[w, 0])
        ev = t[w, 1:cnt]
        ids, clk = ev & 0xff, ev >> 8
        if cnt < 40:
            continue
        lo, hi = cnt // 4, 3 * cnt // 4          # steady-state window
        ids, clk = ids[lo:hi], clk[lo:hi]
        d = np.diff(clk)
        print(f"--- warp {w} ({name}): {cnt - 1} events, window of {len(ids)}; cycles between consecutive events")
        span = {}
        for a, b, dt in zip(ids[:-1], ids[1:], d):
            span.setdefault((int(a), int(b)), []).append(int(dt))
        tot = float(d.sum())
        for k in sorted(span):
            v = np.array(span[k])
            print(f"   {k[0]:3d} -> {k[1]:3d}: n={len(v):4d} mean={v.mean():8.0f} p10={np.percentile(v, 10):7.0f} p90={np.percentile(v, 90):7.0f} share={v.sum() / tot:5.1%}")
    print("events: 1/2 encoder start/end; 10+L epilogue of layer L starts (acc_full seen), 20+L ends; issuer: 30+4s+L waits for "
          "slot s layer L operand, 40+.. got it, 50+.. layer issued+committed")


if __name__ == "__main__":
    main()
