#!/usr/bin/env python
"""bench.py -- the Point-NeRF per-ray hot path on B200, measured on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload render|train] [--precision bf16|fp32]
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference's algorithm)

Workloads (BASELINE.json `configs`):
  render (default, configs[1]) : NeRF-Synthetic-shaped render -- one 800x800 view (640 000 rays) of a ~1 M-point
                                 synthetic neural cloud, K=8, SR=80, scaled voxel 0.008.  A step = one full image
                                 through the hot path (coarse positions, sample selection, neighbour query, field
                                 networks, compositing).  N GPUs: every rank renders its own view (ray-sharded by
                                 view, cloud replicated, no data-path collective) -> weak scaling.
  train  (configs[2])          : one training step fwd+bwd on 4096 rays per rank (point feature / colour / dir /
                                 confidence grads + MLP grads), NCCL all-reduce of the gradients when N > 1.

One JSON line on stdout (rank 0).  `value` = rays/s with the rays resident in HBM; `e2e` = the same through the
public API with pinned-host rays and a host read-back of the pixels (loss for train) inside the timed region.
The oracle is imported only by the `cpu_baseline` leg and by `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

N_POINTS = 1_000_000
CLOUD_SEED = 1236          # 1234 + config id
IMG = 800
TRAIN_RAYS = 4096


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workload
def make_weights(seed=7):
    """Xavier-uniform weights of the shipped network shape (helpers/networks.py:121-173), seeded."""
    g = torch.Generator().manual_seed(seed)
    shapes = [("mlp_base.layers.0", 256, 284), ("mlp_base.layers.1", 256, 256), ("mlp_head.layers.0", 256, 263),
              ("mlp_head.layers.1", 256, 256), ("field_output_density.net", 1, 256), ("mlp_color.layers.0", 128, 280),
              ("mlp_color.layers.1", 128, 128), ("mlp_color.layers.2", 128, 128), ("field_output_color.net", 3, 128)]
    out = {}
    for name, o, i in shapes:
        bound = (6.0 / (i + o)) ** 0.5
        out[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        out[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * 0.05
    out["field_output_density.net.bias"] += 0.3      # keep a good share of the densities above the ReLU
    return out


def make_scene(n_points=N_POINTS):
    from pointnerf2studio_b200.synth import make_cloud
    t0 = time.time()
    cloud = make_cloud(n_points, seed=CLOUD_SEED, P=12)
    return cloud, time.time() - t0


def view(rank, H=IMG, W=IMG):
    from pointnerf2studio_b200.synth import make_camera
    return make_camera(H=H, W=W, azim_deg=30.0 + 45.0 * rank, elev_deg=20.0)


def host_bundle(cam, pix, pinned=True):
    d = torch.from_numpy(cam.rays(pix))
    R = d.shape[0]
    o = torch.from_numpy(cam.origin)[None].expand(R, 3).contiguous()
    nears, fars = torch.full((R, 1), cam.near), torch.full((R, 1), cam.far)
    rot = torch.from_numpy(cam.R_c2w).contiguous()
    ts = [o, d, nears, fars, rot]
    if pinned:
        ts = [t.pin_memory() for t in ts]
    return ts


def to_device(ts, RayBundle):
    o, d, n, f, rot = [t.cuda(non_blocking=True) for t in ts]
    # the host still holds the camera: tell the model, so that it does not read ray 0 back from the device (a host sync per call)
    host_cam = {"origin": ts[0][0].numpy(), "camrotc2w": ts[4].numpy(), "near": float(ts[2][0, 0]), "far": float(ts[3][0, 0])}
    return RayBundle(origins=o, directions=d, nears=n, fars=f, metadata={"camrotc2w": rot, "camera_host": host_cam})


def timed_steps(step_fn, K, flush, dist):
    """K steps, each bracketed by CUDA events on the launching stream; an L2 flush (not timed) runs between steps."""
    evs = []
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(K):
        flush.add_(1.0)                      # > L2 (126 MB) write: the next step starts cold
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    return sum(a.elapsed_time(b) for a, b in evs)    # ms over the K steps


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(cloud, cam, weights, n_rays, SR, K, seed=5, mode="plugin"):
    """The reference's algorithm on the host cores (oracle port): C grid querier (rebuilds the grid on every call,
    as the reference does) + torch fp32 field networks + compositing, on `n_rays` pixels drawn uniformly from the
    same view.  -> (seconds, rays)"""
    from oracle import field as of, grid_query as gq, query_c
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    pix = rng.choice(cam.H * cam.W, size=n_rays, replace=False)
    W = of.FieldWeights({k: v.clone() for k, v in weights.items()})
    pts = {"xyz": torch.from_numpy(cloud.xyz), "Rw2c": torch.from_numpy(cloud.Rw2c)}
    for k in ("embed", "color", "dir", "conf"):
        pts[k] = torch.from_numpy(getattr(cloud, k))
    rays = torch.from_numpy(cam.rays(pix))
    t0 = time.perf_counter()
    with torch.no_grad():
        frame = gq.hyperparameters(cloud.xyz, [0.004] * 3, [2, 2, 2], [3, 3, 3], [-1.2] * 3 + [1.2] * 3)
        raypos, _ = of.coarse_positions(torch.from_numpy(cam.origin), rays, 400, cam.near, cam.far, jitter=0.3,
                                        generator=torch.Generator().manual_seed(seed))
        pidx, loc, mask, hit = query_c.woord_query_grid_point_index(raypos.numpy(), cloud.xyz, [3, 3, 3], [3, 3, 3], SR, K, frame, 12,
                                                                    np.float32(0.016))
        cp, cl, cm = gq.compact_rays(pidx, loc, hit)
        if cp.shape[0] > 0:
            of.render(pts, W, torch.from_numpy(cam.origin), rays, torch.from_numpy(cam.R_c2w), cp, cl, cm, 0.004, SR, mode=mode,
                      training=False)
    return time.perf_counter() - t0, n_rays


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port; the reference's own querier is CUDA-only and its
    plugin needs nerfstudio, neither runs on a CPU) on rank 0's host cores, same workload, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cloud, _ = make_scene()
    cam = view(0)
    weights = make_weights()
    n = args.cpu_rays
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_arm(cloud, cam, weights, min(n, 256), 80, 8)
    tot = 0.0
    for i in range(args.steps):
        dt, _ = cpu_arm(cloud, cam, weights, n, 80, 8, seed=5 + i)
        tot += dt
    v = n * args.steps / tot
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "render rays/s", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "render 800x800 view, 1M-point synthetic cloud, K=8, SR=80, voxel 0.008 (configs[1])",
                       "n_points": int(cloud.xyz.shape[0]), "rays_per_step": n},
            "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{n} pixels drawn uniformly from the 800x800 view per step; C grid querier (grid rebuilt per "
                                       f"call like the reference) + torch fp32 field/compositing on {cores} threads"},
            "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="render", choices=["render", "train", "scannet"])
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--cpu-rays", type=int, default=131072, help="pixels per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))     # N ranks share the host cores
    import importlib.util
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle, native
    precision = args.precision
    if precision is None:
        precision = "bf16" if importlib.util.find_spec("pointnerf2studio_b200.native_tc") is not None else "fp32"
    peaks = load_peaks()

    weights = make_weights()
    if args.workload == "scannet":
        # BASELINE configs[3]: ~3 M points, 1296 x 968 image, scaled voxel 0.016, radius 0.032, P = 30, SR = 24
        # (dev_scripts/w_scannet_etf/scene101_points.sh:24-37); ONE image split in row blocks over the ranks + all-gather
        from pointnerf2studio_b200.synth import make_cloud
        cloud = make_cloud(3_000_000 if args.points == N_POINTS else args.points, seed=1237, scaled_vsize=0.016, P=30,
                           radii=(0.5, 0.72, 0.93))
        cfg = PointNerfConfig(precision=precision, vsize=[0.008] * 3, P=30, SR=24)
    else:
        cloud, t_cloud = make_scene(args.points)
        cfg = PointNerfConfig(precision=precision)    # plugin defaults: SR=80, K=8, P=12, vsize .004 x vscale 2, jitter 0.3
    model = PointNerf(cfg, state_dict=cloud.state_dict())
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in weights.items():
            own[k].copy_(v)
    cam = view(0)      # weak scaling = fixed work per GPU: every rank renders the same view (training: its own pixels of it)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")     # 256 MB > 126 MB L2

    if args.workload == "scannet":
        from pointnerf2studio_b200.parallel import gather_interleaved_image, interleaved_rows
        from pointnerf2studio_b200.synth import make_camera
        model.eval()
        cam = make_camera(H=968, W=1296, focal=1170.0, azim_deg=30.0, elev_deg=20.0)      # the same view on every rank
        n_img = cam.H * cam.W
        rows = np.asarray(interleaved_rows(cam.H, rank, world))                          # rows rank, rank + N, ...: balanced work
        pix = (rows[:, None] * cam.W + np.arange(cam.W)[None]).reshape(-1)
        host = host_bundle(cam, pix)
        rb_dev = to_device(host, RayBundle)
        R = len(pix)
        out_host = torch.empty((n_img, 3), dtype=torch.float32).pin_memory()

        def render_rows(rb):
            o = model.get_outputs_for_camera_ray_bundle(rb)
            if dist is None:
                return o["coarse_raycolor"]
            return gather_interleaved_image(o["coarse_raycolor"], cam.H, cam.W, dist)

        def step_dev():
            render_rows(rb_dev)

        def step_e2e():
            img = render_rows(to_device(host, RayBundle))
            out_host[:img.shape[0]].copy_(img, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = out_host.numel() * 4
        metric = "render rays/s"
        workload = ("ScanNet-scale render: one 1296x968 image, rows interleaved over the ranks + pixel all-gather, 3M-point "
                    "synthetic cloud, K=8, SR=24, voxel 0.016, P=30 (configs[3])")
    elif args.workload == "render":
        model.eval()
        pix = np.arange(cam.H * cam.W)
        host = host_bundle(cam, pix)
        rb_dev = to_device(host, RayBundle)
        R = len(pix)
        out_host = torch.empty((R, 3), dtype=torch.float32).pin_memory()

        def step_dev():
            model.get_outputs_for_camera_ray_bundle(rb_dev)

        def step_e2e():
            # one camera per call: the directions are the per-ray input; origin / rotation / near / far are 14 floats
            rb = RayBundle.for_camera(host[1].cuda(non_blocking=True), cam.origin, cam.R_c2w, cam.near, cam.far)
            o = model.get_outputs_for_camera_ray_bundle(rb)
            out_host.copy_(o["coarse_raycolor"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        h2d = host[1].numel() * host[1].element_size() + 14 * 4
        d2h = out_host.numel() * 4
        metric = "render rays/s"
        workload = "render 800x800 view, 1M-point synthetic cloud, K=8, SR=80, voxel 0.008 (configs[1])"
    else:
        model.train()
        rng = np.random.default_rng(100 + rank)
        pix = rng.choice(cam.H * cam.W, size=TRAIN_RAYS, replace=False)
        host = host_bundle(cam, pix)
        rb_dev = to_device(host, RayBundle)
        R = TRAIN_RAYS
        gt_host = torch.rand((R, 3), generator=torch.Generator().manual_seed(9)).pin_memory()
        gt_dev = gt_host.cuda()
        from pointnerf2studio_b200.optim import make_optimizers
        from pointnerf2studio_b200.parallel import allreduce_gradients
        params = [p for p in model.parameters() if p.requires_grad]
        opts, scheds = make_optimizers(model)       # the plugin's two Adam groups + exponential decay (studio_config.py:33-48)

        def train_step(rb, gt):
            for p in params:
                p.grad = None
            out = model.get_outputs(rb)
            ld = model.get_loss_dict(out, {"image": gt})
            loss = sum(ld.values())
            loss.backward()
            if dist is not None:
                allreduce_gradients(params, dist)
            for k in opts:
                opts[k].step()
                scheds[k].step()
            return loss

        def step_dev():
            train_step(rb_dev, gt_dev)

        def step_e2e():
            rb = to_device(host, RayBundle)
            loss = train_step(rb, gt_host.cuda(non_blocking=True))
            loss.item()

        h2d = sum(t.numel() * t.element_size() for t in host) + gt_host.numel() * 4
        d2h = 4
        metric = "train rays/s"
        workload = ("training step fwd+bwd+Adam (fields 5e-4, neural points 2e-3), 4096 rays per rank, 1M-point synthetic cloud, "
                    "K=8, SR=80 (configs[2])")

    # ---- warm-up (also builds the cached voxel grid) + occupancy statistics of this view (not timed)
    # multi-rank training: NCCL finishes its channel / buffer set-up over the first ~10 all-reduces of these tensors (a 5-step run
    # at 4 GPUs read 6.7 ms/step against 3.6 ms in steady state), so those steps are added to the untimed warm-up
    n_warm = max(args.warmup, 1) + (10 if (dist is not None and args.workload == "train") else 0)
    for _ in range(n_warm):
        step_dev()
    torch.cuda.synchronize()
    with torch.no_grad():
        q, _, _, _ = model.neural_points.query(rb_dev, want_stats=True)
        torch.cuda.synchronize()
        S = int(q.sample_valid.sum().item())
        M = int((q.sample_pidx >= 0).sum().item())
        filled = int(q.sample_cnt.sum().item())
        rays_hit = int((q.sample_valid.sum(1) > 0).sum().item())
        vis, cand = [int(x) for x in q.stats.tolist()]
        del q

    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    native.Timers.enabled = True
    native.Timers.spans = []
    l0 = native.LAUNCHES["n"]
    ms_total = timed_steps(step_dev, args.steps, flush, dist)
    launches = native.LAUNCHES["n"] - l0
    spans = native.Timers.collect()
    native.Timers.enabled = False
    # ---- end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_steps(step_e2e, args.steps, flush, dist)
    clocks = sampler.stop()

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    rays_all = (n_img if args.workload == "scannet" else R * world) * args.steps
    value = rays_all / (ms_total * 1e-3)
    e2e_value = rays_all / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel(s): the field networks (tensor bound); the query stage against HBM
    stage_ms = {k: sum(v) / args.steps for k, v in spans.items()}
    mult = 3.0 if args.workload == "train" else 1.0
    flops = (542208.0 + 512.0) * M + 137984.0 * S
    field_ms = stage_ms.get("field", 0.0) + (stage_ms.get("field_bwd", 0.0) if args.workload == "train" else 0.0)
    ach_tf = flops * mult / (field_ms * 1e-3) / 1e12 if field_ms > 0 else 0.0
    peak_tf = peaks["tensor"]
    q_bytes = 12.0 * filled + 4.0 * vis + 16.0 * cand + 4.0 * cfg.K * filled
    q_ms = stage_ms.get("query", 0.0)
    q_gbs = q_bytes / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0
    # dram__bytes_read + dram__bytes_write of field_tc_kernel<8,0> for one launch of this very workload (ncu --set full,
    # profiles/r01_final_ncu_full.md): 0.58 GB read + 1.62 GB written, against 5.3 GB of algorithmic gather + output bytes
    # (neighbouring samples share points, the gathers hit L2).  Only quoted for the workload it was captured on.
    traffic = 2.194e9 if (args.workload == "render" and args.points == N_POINTS and precision == "bf16") else None
    roofline = {"kernel": "field networks (gather + per-neighbour MLP + aggregation + colour MLP)", "bound": "tensor",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
                "traffic_unit": "bytes of DRAM traffic per launch of field_tc_kernel (ncu, profiles/r01_final_ncu_full.md)",
                "peak_source": peaks["src"] + " (sustained cuBLAS bf16)", "flops_per_launch": flops * mult, "ms_per_launch": field_ms}
    stages = {"ms": stage_ms,
              "query_hbm": {"achieved": q_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": q_gbs / peaks["hbm"],
                            "algorithmic_bytes": q_bytes, "mean_voxels_visited": vis / max(filled, 1),
                            "mean_candidates": cand / max(filled, 1)}}

    scaling = "strong" if args.workload == "scannet" else "weak"      # scannet: one image of fixed size shared by all ranks
    line = {"metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload, "n_points": int(cloud.xyz.shape[0]), "rays_per_step_per_gpu": R,
                       "rays_hit": rays_hit, "filled_slots": filled, "valid_samples_S": S, "neighbour_rows_M": M,
                       "cloud": cloud.stats, "l2": "256 MB flush write between timed steps (outside the event pairs)",
                       "precision": precision, "jitter": cfg.jitter, "parallelism": f"ray-sharded x{world}, cloud replicated; every rank works on the same view (fixed work per GPU)"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "stages": stages}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload != "scannet":
        n = args.cpu_rays
        cpu_arm(cloud, cam, weights, 256, cfg.SR, cfg.K)                        # warm-up (builds nothing that persists)
        dt, _ = cpu_arm(cloud, cam, weights, n, cfg.SR, cfg.K)
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {"value": n / dt, "unit": "rays/s", "cores": cores, "kind": "port",
                                "sample": f"{n} pixels drawn uniformly from the same 800x800 view ({dt:.1f} s); C grid querier (grid "
                                          f"rebuilt per call like the reference) + torch fp32 field/compositing on {cores} threads"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
