#!/usr/bin/env python
"""bench.py -- the Point-NeRF per-ray hot path on B200, measured on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload render|train|scannet] [--precision bf16|fp32]
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference's algorithm)

The default line (`--workload render`, BASELINE configs[1]) carries, besides the render measurement:
  value / e2e / roofline / stages : one 800x800 view (640 000 rays) of a ~1 M-point synthetic neural cloud, K=8, SR=80,
                                    scaled voxel 0.008 through the whole hot path (coarse positions, sample selection,
                                    neighbour query, field networks, compositing).  N GPUs: every rank renders a DIFFERENT
                                    view (azimuth 30 + 45 * rank degrees; ray-sharded by view, cloud replicated) and the
                                    pixels of all views are all-gathered inside the device-timed region -> weak scaling; end to
                                    end every rank uploads its rays and downloads its own view (no collective on that route).
  train   (configs[2])            : fwd + bwd + gradient all-reduce (N > 1) + Adam on 4096 rays per rank, the reference's own
                                    per-process batch (studio_config.py:20-21); `update_ms` (exchange + Adam) is reported separately.
  parity  (N = 1)                 : the CPU oracle on a pixel sample of the SAME view with the very t table the GPU's in-kernel
                                    jitter generated (pnerf_coarse_t): neighbour-index mismatches, pixel error, PSNR.  The same
                                    CPU run is the `cpu_baseline`.
  stress  (configs[4], N = 1)     : ~10 M-point cloud, K = 16, SR = 80, 5^3 kernel: query GB/s and aggregation TFLOP/s on 262 144 rays.
  scannet (configs[3], N > 1 or --with-scannet) : ONE 1296x968 image of a 3 M-point cloud split over the ranks by interleaved
                                    rows (strong scaling); device-timed with the pixel all-gather, end to end with every rank
                                    copying its rows straight into one shared pinned host image.

One JSON line on stdout (rank 0).  `value` = rays/s with the rays resident in HBM; `e2e` = the same through the public API with
pinned-host rays and a host read-back of the pixels (loss for train) inside the timed region.  The oracle is imported only by
the `cpu_baseline` / `parity` leg and by `--impl reference`.
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
FIELD_FLOP_ROW = 542208.0 + 512.0     # per valid neighbour row
COLOR_FLOP_SAMPLE = 137984.0          # per valid sample


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


def load_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this very workload
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when there is no capture for it."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    e = d.get(f"{workload}:{kernel}")
    return (e["dram_bytes"], e["source"]) if e else (None, None)


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
        busy = [s for s, p in zip(sm, pw) if p > 400.0] or sm          # samples taken while the GPU was working
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "samples_under_load": len(busy) if pw else 0,
                "reasons": sorted(reasons)}


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
    # Xavier-sized density logits (|raw| ~ 0.3) times steps of 0.004-0.008 would leave every ray transparent and the image white --
    # parity figures measured on it would say nothing.  Scale the density head so that sigma * delta ~ 1 at surfaces, as in a
    # trained scene (sigma of a few hundred): the pixels then carry the colour network's output and the bf16 error shows.
    out["field_output_density.net.weight"] *= 300.0
    out["field_output_density.net.bias"] = (out["field_output_density.net.bias"] + 0.3) * 300.0
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


def timed_steps(step_fn, K, flush, dist, per_step=None):
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
    ms = [a.elapsed_time(b) for a, b in evs]
    if per_step is not None:
        per_step.extend(ms)
    return sum(ms)    # ms over the K steps


def max_over_ranks(vals, dist):
    t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(cloud, cam, weights, pix, SR, K, seed=5, mode="plugin", t_mid=None, want_outputs=False):
    """The reference's algorithm on the host cores (oracle port): C grid querier (rebuilds the grid on every call,
    as the reference does) + torch fp32 field networks + compositing on the pixels `pix` of the view.
    `t_mid` (R,D): coarse t mid-points to use instead of drawing them with torch.rand (the parity leg passes the table
    the GPU kernel generated).  -> (seconds, outputs or None)"""
    from oracle import field as of, grid_query as gq, query_c
    torch.set_num_threads(os.cpu_count() or 1)
    W = of.FieldWeights({k: v.clone() for k, v in weights.items()})
    pts = {"xyz": torch.from_numpy(cloud.xyz), "Rw2c": torch.from_numpy(cloud.Rw2c)}
    for k in ("embed", "color", "dir", "conf"):
        pts[k] = torch.from_numpy(getattr(cloud, k))
    rays = torch.from_numpy(cam.rays(pix))
    origin = torch.from_numpy(cam.origin)
    t0 = time.perf_counter()
    out = None
    with torch.no_grad():
        frame = gq.hyperparameters(cloud.xyz, [0.004] * 3, [2, 2, 2], [3, 3, 3], [-1.2] * 3 + [1.2] * 3)
        if t_mid is None:
            raypos, _ = of.coarse_positions(origin, rays, 400, cam.near, cam.far, jitter=0.3, generator=torch.Generator().manual_seed(seed))
        else:
            raypos = origin.float().view(1, 1, 3) + rays[:, None, :] * torch.as_tensor(t_mid)[:, :, None]     # RM:330
        pidx, loc, mask, hit = query_c.woord_query_grid_point_index(raypos.numpy(), cloud.xyz, [3, 3, 3], [3, 3, 3], SR, K, frame, 12,
                                                                    np.float32(0.016))
        cp, cl, cm = gq.compact_rays(pidx, loc, hit)
        if cp.shape[0] > 0:
            r = of.render(pts, W, origin, rays, torch.from_numpy(cam.R_c2w), cp, cl, cm, 0.004, SR, mode=mode, training=False)
            if want_outputs:
                out = {"pixels": r["coarse_raycolor"].numpy(), "pidx": pidx, "ray_mask": np.asarray(cm)}
    return time.perf_counter() - t0, out


def parity_leg(model, cloud, cam, weights, n_rays, SR, K, seed=5):
    """The GPU path and the CPU oracle on the same `n_rays` pixels of `cam` with the same coarse t table.
    -> (parity dict, cpu seconds)"""
    from pointnerf2studio_b200 import RayBundle, native
    rng = np.random.default_rng(seed)
    pix = np.sort(rng.choice(cam.H * cam.W, size=n_rays, replace=False))
    was_training = model.training
    model.eval()
    with torch.no_grad():
        rb = RayBundle.for_camera(torch.from_numpy(cam.rays(pix)).cuda(), cam.origin, cam.R_c2w, cam.near, cam.far)
        out = model.get_outputs_for_camera_ray_bundle(rb)
        near, far, jitter, jseed = model.neural_points._last_jitter
        t = native.coarse_t(near, far, jitter, jseed, n_rays, model.config.z_depth_dim, "cuda")
        got_pix = out["coarse_raycolor"].cpu().numpy()
        got_mask = out["ray_mask"].cpu().numpy()
        got_pidx = model.last_query_dense().sample_pidx.cpu().numpy()
        t = t.cpu().numpy()
    model.train(was_training)
    dt, ref = cpu_arm(cloud, cam, weights, pix, SR, K, t_mid=t, want_outputs=True)
    mism = int((got_pidx != ref["pidx"]).sum())
    mask_mism = int((got_mask.astype(bool) != ref["ray_mask"].astype(bool)).sum())
    err = np.abs(got_pix - ref["pixels"])
    hit = ref["ray_mask"].astype(bool)
    mse = float((err ** 2).mean())
    mse_hit = float((err[hit] ** 2).mean()) if hit.any() else 0.0
    par = {"rays": int(n_rays), "rays_hit": int(hit.sum()), "idx_entries": int(got_pidx.size), "idx_mismatch": mism,
           "ray_mask_mismatch": mask_mism, "max_abs_err": float(err.max()), "psnr_db": 10 * np.log10(1.0 / max(mse, 1e-20)),
           "psnr_hit_rays_db": 10 * np.log10(1.0 / max(mse_hit, 1e-20)),
           "against": "CPU oracle (fp32, C querier + torch field/compositing) fed the GPU kernel's own jittered t table"}
    return par, dt


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
    pick = lambda m, s: np.random.default_rng(s).choice(cam.H * cam.W, size=m, replace=False)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_arm(cloud, cam, weights, pick(min(n, 256), 4), 80, 8)
    tot = 0.0
    for i in range(args.steps):
        dt, _ = cpu_arm(cloud, cam, weights, pick(n, 5 + i), 80, 8, seed=5 + i)
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


# ------------------------------------------------------------------------------------------------ GPU arm: pieces
class Ctx:
    pass


def query_stats(model, rb):
    with torch.no_grad():
        q, _, _, _ = model.neural_points.query(rb, want_stats=True)
        torch.cuda.synchronize()
        S = int((q.sample_valid > 0).sum().item())
        M = int((q.sample_pidx >= 0).sum().item())
        filled = int(q.sample_cnt.sum().item())
        rays_hit = int(((q.sample_valid > 0).sum(1) > 0).sum().item())
        vis, cand = [int(x) for x in q.stats.tolist()]
    return {"S": S, "M": M, "filled": filled, "rays_hit": rays_hit, "vis": vis, "cand": cand}


def bench_render(c, cam, steps, warmup):
    """configs[1]: one full 800x800 view per step on every rank (a different view per rank), pixels of all ranks all-gathered."""
    from pointnerf2studio_b200 import RayBundle, native
    model, dist, world = c.model, c.dist, c.world
    model.eval()
    pix = np.arange(cam.H * cam.W)
    host = host_bundle(cam, pix)
    rb_dev = to_device(host, RayBundle)
    R = len(pix)
    out_host = torch.empty((R, 3), dtype=torch.float32).pin_memory()
    all_pix = torch.empty((world * R, 3), dtype=torch.float32, device="cuda") if dist is not None else None

    def gather(rgb):
        if dist is not None:       # every rank ends up with every view (what a device-side consumer of the frames needs)
            dist.all_gather_into_tensor(all_pix, rgb.contiguous())

    def step_dev():
        gather(model.get_outputs_for_camera_ray_bundle(rb_dev)["coarse_raycolor"])

    def step_e2e():
        # one camera per call: the directions are the per-ray input; origin / rotation / near / far are 14 floats
        # (every rank serves its own view from and to ITS host buffers: no collective on this route -- the all-gather above is for a
        # device-side consumer of all the frames and belongs to the device-timed figure)
        rb = RayBundle.for_camera(host[1].cuda(non_blocking=True), cam.origin, cam.R_c2w, cam.near, cam.far)
        o = model.get_outputs_for_camera_ray_bundle(rb)
        out_host.copy_(o["coarse_raycolor"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    t0 = time.perf_counter()
    model.neural_points.grid()
    torch.cuda.synchronize()
    grid_ms = 1e3 * (time.perf_counter() - t0)      # first build of the cached voxel grid (outside every timed region)
    for _ in range(max(warmup, 1)):
        step_dev()
    torch.cuda.synchronize()
    st = query_stats(model, rb_dev)
    native.Timers.enabled = True
    native.Timers.spans = []
    l0 = native.LAUNCHES["n"]
    per_step = []
    ms_total = timed_steps(step_dev, steps, c.flush, dist, per_step)
    launches = native.LAUNCHES["n"] - l0
    spans = native.Timers.collect()
    native.Timers.enabled = False
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_steps(step_e2e, steps, c.flush, dist)
    mine = ms_total
    ms_total, ms_e2e = max_over_ranks([ms_total, ms_e2e], dist)
    stage_ms = {k: sum(v) / steps for k, v in spans.items()}
    h2d = host[1].numel() * host[1].element_size() + 14 * 4
    d2h = out_host.numel() * 4
    return {"R": R, "ms_total": ms_total, "ms_e2e": ms_e2e, "ms_this_rank": mine / steps, "stage_ms": stage_ms, "stats": st,
            "launches": launches, "h2d": h2d, "d2h": d2h, "grid_build_ms": grid_ms, "rb_dev": rb_dev}


def bench_train(c, cam, steps, warmup):
    """configs[2]: fwd + bwd + gradient all-reduce + Adam (both groups) on 4096 rays per rank."""
    from pointnerf2studio_b200 import RayBundle, native
    from pointnerf2studio_b200.parallel import TrainEngine
    model, dist, world, rank = c.model, c.dist, c.world, c.rank
    model.train()
    rng = np.random.default_rng(100 + rank)
    pix = rng.choice(cam.H * cam.W, size=TRAIN_RAYS, replace=False)
    host = host_bundle(cam, pix)
    rb_dev = to_device(host, RayBundle)
    gt_host = torch.rand((TRAIN_RAYS, 3), generator=torch.Generator().manual_seed(9)).pin_memory()
    gt_dev = gt_host.cuda()
    # the plugin's two Adam groups + exponential decay (studio_config.py:33-48); the step is captured as one CUDA graph
    engine = TrainEngine(model, dist, exchange=c.exchange, use_graph=not c.no_graph)

    def step_dev():
        engine.step(rb_dev, gt_dev)

    def step_e2e():
        rb = RayBundle.for_camera(host[1].cuda(non_blocking=True), cam.origin, cam.R_c2w, cam.near, cam.far)
        loss = engine.step(rb, gt_host.cuda(non_blocking=True))
        loss.item()

    # NCCL finishes its channel / buffer set-up over the first ~10 all-reduces of these tensors: untimed warm-up
    n_warm = max(warmup, 3) + (10 if dist is not None else 0)
    for _ in range(n_warm):
        step_dev()
    torch.cuda.synchronize()
    st = query_stats(model, rb_dev)
    native.Timers.enabled = True
    native.Timers.spans = []
    l0 = native.LAUNCHES["n"]
    ms_total = timed_steps(step_dev, steps, c.flush, dist)
    launches = native.LAUNCHES["n"] - l0
    spans = native.Timers.collect()
    native.Timers.enabled = False
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_steps(step_e2e, steps, c.flush, dist)
    # for information: the same 4096 rays split over the ranks = strong scaling of ONE batch
    strong = None
    if dist is not None and TRAIN_RAYS % world == 0:
        r_s = TRAIN_RAYS // world
        rb_s = to_device(host_bundle(cam, pix[:r_s]), RayBundle)
        gt_s = gt_dev[:r_s].contiguous()
        for _ in range(3):
            engine.step(rb_s, gt_s)
        ms_s = max_over_ranks([timed_steps(lambda: engine.step(rb_s, gt_s), steps, c.flush, dist)], dist)[0]
        strong = {"rays_per_step_total": TRAIN_RAYS, "rays_per_step_per_gpu": r_s, "ms_per_step": ms_s / steps,
                  "value": TRAIN_RAYS * steps / (ms_s * 1e-3), "unit": "rays/s", "scaling": "strong"}
    if engine._graphs:       # per-stage device times need eager launches: a few untimed steps outside the graph (diagnostic only)
        engine.use_graph = False
        for _ in range(2):
            step_dev()
        native.Timers.enabled = True
        native.Timers.spans = []
        for _ in range(3):
            step_dev()
        torch.cuda.synchronize()
        spans = native.Timers.collect()
        native.Timers.enabled = False
        engine.use_graph = True
        n_span = 3
    else:
        n_span = steps
    # gradient exchange + Adam + gradient reset on their own: a few more steps with an event pair around TrainEngine.update()
    # (after a barrier, so that the wait for the slowest rank is not billed to it)
    engine.timing = []
    for _ in range(5):
        if dist is not None:
            dist.barrier()
        engine.update()          # on zero gradients: same traffic, same kernels
    torch.cuda.synchronize()
    ar_ms = statistics.median(a.elapsed_time(b) for a, b in engine.timing)
    engine.timing = None
    ms_total, ms_e2e, ar_ms = max_over_ranks([ms_total, ms_e2e, ar_ms], dist)
    stage_ms = {k: sum(v) / n_span for k, v in spans.items()}
    field_ms = stage_ms.get("render_fwd", 0.0) + stage_ms.get("render_bwd", 0.0)     # the two one-call launchers (pnerf_render_train_*)
    flops = 3.0 * (FIELD_FLOP_ROW * st["M"] + COLOR_FLOP_SAMPLE * st["S"])      # fwd + dgrad + wgrad
    ach = flops / (field_ms * 1e-3) / 1e12 if field_ms > 0 else 0.0
    n_pts = sum(p.numel() for p in model.get_param_groups()["neural_points"] if p.requires_grad)
    n_mlp = sum(p.numel() for p in model.get_param_groups()["fields"] if p.requires_grad)
    h2d = host[1].numel() * 4 + 14 * 4 + gt_host.numel() * 4
    rays_all = TRAIN_RAYS * world * steps
    return {"metric": "train rays/s", "value": rays_all / (ms_total * 1e-3), "unit": "rays/s", "ms_per_step": ms_total / steps,
            "steps": steps, "warmup": n_warm, "rays_per_step_per_gpu": TRAIN_RAYS,
            "update_ms": ar_ms, "exchange": engine.exchange, "cuda_graph": bool(engine._graphs), "exchange_bytes": 4 * (n_pts + n_mlp) if dist is not None else 0,
            "update_is": "gradient exchange over the ranks + Adam on both groups + gradient reset (TrainEngine.update); at 1 GPU it is the "
                         "Adam pass alone, so the difference to the 1-GPU figure is the cost of the collective",
            "device_ms": {k: v for k, v in stage_ms.items()},
            "e2e": {"value": rays_all / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e / steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "valid_samples_S": st["S"], "neighbour_rows_M": st["M"], "rays_hit": st["rays_hit"],
            "strong_scaling_4096_total": strong,
            "roofline": {"kernel": "forward + backward launchers (selection, query, field networks fwd + dgrad + wgrad on tcgen05, compositing)", "bound": "tensor", "achieved": ach, "peak": c.peaks["tensor"],
                         "unit": "TFLOP/s", "frac": ach / c.peaks["tensor"], "flops_per_step": flops, "ms_per_step": field_ms},
            "workload": "training step fwd+bwd+all-reduce+Adam (fields 5e-4, neural points 2e-3), 4096 rays per rank drawn from the "
                        "rank's own view, 1M-point synthetic cloud, K=8, SR=80 (configs[2])"}


def bench_scannet(c, steps, warmup, n_points=3_000_000):
    """configs[3]: ~3 M points, 1296 x 968 image, scaled voxel 0.016, radius 0.032, P = 30, SR = 24
    (dev_scripts/w_scannet_etf/scene101_points.sh:24-37); ONE image split over the ranks by interleaved rows."""
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle
    from pointnerf2studio_b200.parallel import SharedHostImage, gather_interleaved_image, interleaved_rows
    from pointnerf2studio_b200.synth import make_camera, make_cloud
    dist, world, rank = c.dist, c.world, c.rank
    cloud = make_cloud(n_points, seed=1237, scaled_vsize=0.016, P=30, radii=(0.5, 0.72, 0.93))
    cfg = PointNerfConfig(precision=c.precision, vsize=[0.008] * 3, P=30, SR=24)
    model = PointNerf(cfg, state_dict=cloud.state_dict())
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in c.weights.items():
            own[k].copy_(v)
    model.eval()
    cam = make_camera(H=968, W=1296, focal=1170.0, azim_deg=30.0, elev_deg=20.0)      # the same view on every rank
    n_img = cam.H * cam.W
    rows = np.asarray(interleaved_rows(cam.H, rank, world))                          # rows rank, rank + N, ...: balanced work
    pix = (rows[:, None] * cam.W + np.arange(cam.W)[None]).reshape(-1)
    host = host_bundle(cam, pix)
    rb_dev = to_device(host, RayBundle)
    shared = SharedHostImage(cam.H, cam.W, rank, world, dist, tag=f"pnerf_bench_{os.environ.get('MASTER_PORT', '0')}")

    def step_dev():
        o = model.get_outputs_for_camera_ray_bundle(rb_dev)
        if dist is not None:
            gather_interleaved_image(o["coarse_raycolor"], cam.H, cam.W, dist)

    def step_e2e():
        rb = RayBundle.for_camera(host[1].cuda(non_blocking=True), cam.origin, cam.R_c2w, cam.near, cam.far)
        o = model.get_outputs_for_camera_ray_bundle(rb)
        shared.put_rows(o["coarse_raycolor"])          # this rank's rows straight into the one pinned host image: no collective
        torch.cuda.current_stream().synchronize()

    for _ in range(max(warmup, 1)):
        step_dev()
    ms_dev = timed_steps(step_dev, steps, c.flush, dist)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_steps(step_e2e, steps, c.flush, dist)
    ms_dev, ms_e2e = max_over_ranks([ms_dev, ms_e2e], dist)
    shared.close()
    del model
    torch.cuda.empty_cache()
    return {"metric": "render rays/s", "scaling": "strong", "value": n_img * steps / (ms_dev * 1e-3), "unit": "rays/s",
            "ms_per_step": ms_dev / steps, "steps": steps,
            "e2e": {"value": n_img * steps / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step_per_gpu": host[1].numel() * 4 + 56, "d2h_bytes_per_step_per_gpu": len(pix) * 12,
                    "how": "every rank uploads its rows' directions and copies its rendered rows into ONE shared pinned host image"},
            "n_points": int(cloud.xyz.shape[0]), "rays_per_image": n_img,
            "workload": "ScanNet-scale render: one 1296x968 image, rows interleaved over the ranks, device-timed with the pixel "
                        "all-gather; 3M-point synthetic cloud, K=8, SR=24, voxel 0.016, P=30 (configs[3])"}


def bench_stress(c, reps=3, n_points=10_000_000, n_rays=262144):
    """configs[4]: Tanks-and-Temples-scale stress -- ~10 M-point cloud, K = 16, SR = 80, vsize 0.002 x vscale 2, kernel 5^3 (3 layers,
    125 voxels), P = 10 (dev_scripts/w_tt_ft/truck_points.sh:53-63): the neighbour query and the aggregation (fused field kernels)
    timed separately on `n_rays` rays of a 1024x1024 view.  The cloud is generated on the GPU (synth.make_cloud_state_dict_torch)."""
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig, RayBundle, native
    from pointnerf2studio_b200.synth import make_camera, make_cloud_state_dict_torch
    sd, stats = make_cloud_state_dict_torch(n_points, seed=1239, device="cuda", scaled_vsize=0.004, P=10, radii=(0.45, 0.65, 0.85), kernel_size=(5, 5, 5))
    cfg = PointNerfConfig(K=16, SR=80, P=10, vsize=[0.002] * 3, kernel_size=[5, 5, 5], max_o=1600000, precision=c.precision)
    model = PointNerf(cfg, state_dict=sd).eval()
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in c.weights.items():
            own[k].copy_(v)
    cam = make_camera(H=1024, W=1024, focal=1422.0)
    pix = np.sort(np.random.default_rng(n_rays).choice(cam.H * cam.W, size=n_rays, replace=False))
    rb = RayBundle.for_camera(torch.from_numpy(cam.rays(pix)).cuda(), cam.origin, cam.R_c2w, cam.near, cam.far)
    with torch.no_grad():
        model.get_outputs(rb)                         # builds the grid, warms up
        st = query_stats(model, rb)
        native.Timers.enabled, native.Timers.spans = True, []
        for _ in range(reps):
            c.flush.add_(1.0)
            model.get_outputs(rb)
        torch.cuda.synchronize()
        sp = {k: sum(v) / len(v) for k, v in native.Timers.collect().items()}
        native.Timers.enabled = False
    q_bytes = 12.0 * st["filled"] + 4.0 * st["vis"] + 16.0 * st["cand"] + 4.0 * cfg.K * st["filled"]
    flops = FIELD_FLOP_ROW * st["M"] + COLOR_FLOP_SAMPLE * st["S"]
    out = {"workload": "Tanks-and-Temples-scale stress: ~10M-point cloud, K=16, SR=80, 5x5x5 kernel, voxel 0.004, P=10 (configs[4])",
           "rays": n_rays, "n_points": stats["n_points"], "cloud": stats, "filled_slots": st["filled"], "valid_samples_S": st["S"],
           "neighbour_rows_M": st["M"], "mean_voxels_visited": st["vis"] / max(st["filled"], 1), "mean_candidates": st["cand"] / max(st["filled"], 1),
           "ms": sp, "query_samples_per_s": st["filled"] / (sp["query"] * 1e-3),
           "query_hbm": {"achieved": q_bytes / (sp["query"] * 1e-3) / 1e9, "peak": c.peaks["hbm"], "unit": "GB/s",
                         "frac": q_bytes / (sp["query"] * 1e-3) / 1e9 / c.peaks["hbm"]},
           "aggregate_rows_per_s": st["M"] / (sp["field"] * 1e-3),
           "aggregate_tensor": {"achieved": flops / (sp["field"] * 1e-3) / 1e12, "peak": c.peaks["tensor"], "unit": "TFLOP/s",
                                "frac": flops / (sp["field"] * 1e-3) / 1e12 / c.peaks["tensor"]}}
    del model, sd
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--cpu-rays", type=int, default=131072, help="pixels per step of the CPU arm / parity leg")
    ap.add_argument("--train-steps", type=int, default=0, help="timed steps of the train block (default: max(--steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--with-scannet", action="store_true", help="add the configs[3] block at N = 1 too (always on at N > 1)")
    ap.add_argument("--no-scannet", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the configs[4] block (10M points, K=16; on by default at N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="train block: launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="gradient exchange of the train block at N > 1")
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
    from pointnerf2studio_b200 import PointNerf, PointNerfConfig
    precision = args.precision
    if precision is None:
        precision = "bf16" if importlib.util.find_spec("pointnerf2studio_b200.native_tc") is not None else "fp32"

    c = Ctx()
    c.dist, c.world, c.rank, c.precision = dist, world, rank, precision
    c.exchange, c.no_graph = args.exchange, args.no_graph
    c.peaks = load_peaks()
    c.weights = make_weights()
    c.flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")     # 256 MB > 126 MB L2
    train_steps = args.train_steps or max(args.steps, 20)

    if args.workload == "scannet":       # stand-alone configs[3] line
        sampler = ClockSampler(local_rank)
        sampler.start()
        blk = bench_scannet(c, args.steps, args.warmup, 3_000_000 if args.points == N_POINTS else args.points)
        clocks = sampler.stop()
        line = {"metric": blk["metric"], "value": blk["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": blk["workload"], "n_points": blk["n_points"], "rays_per_image": blk["rays_per_image"],
                           "l2": "256 MB flush write between timed steps (outside the event pairs)"},
                "e2e": blk["e2e"], "clocks": clocks}
        if rank == 0:
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    cloud, _ = make_scene(args.points)
    cfg = PointNerfConfig(precision=precision)    # plugin defaults: SR=80, K=8, P=12, vsize .004 x vscale 2, jitter 0.3
    model = PointNerf(cfg, state_dict=cloud.state_dict())
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in c.weights.items():
            own[k].copy_(v)
    c.model = model
    cam = view(rank)      # weak scaling: a different view per rank (training: the rank's own pixels of its view)
    dtype = "bf16" if precision == "bf16" else "f32"
    cores = os.cpu_count() or 1

    if args.workload == "train":         # stand-alone configs[2] line
        sampler = ClockSampler(local_rank)
        sampler.start()
        t = bench_train(c, cam, args.steps, args.warmup)
        clocks = sampler.stop()
        line = {"metric": "train rays/s", "value": t["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": t["warmup"],
                "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
                "data": "synthetic",
                "config": {"workload": t["workload"], "n_points": int(cloud.xyz.shape[0]), "rays_per_step_per_gpu": TRAIN_RAYS,
                           "l2": "256 MB flush write between timed steps (outside the event pairs)", "precision": precision},
                "e2e": t["e2e"], "gpu_launches": t["gpu_launches"], "clocks": clocks, "roofline": t["roofline"], "train": t}
        if rank == 0:
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- default: render (value / e2e / roofline) + parity + train [+ scannet]
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = bench_render(c, cam, args.steps, args.warmup)
    clocks = sampler.stop()
    st, stage_ms, R = r["stats"], r["stage_ms"], r["R"]
    rays_all = R * world * args.steps
    value = rays_all / (r["ms_total"] * 1e-3)
    e2e_value = rays_all / (r["ms_e2e"] * 1e-3)
    flops = FIELD_FLOP_ROW * st["M"] + COLOR_FLOP_SAMPLE * st["S"]
    field_ms = stage_ms.get("field", 0.0)
    ach_tf = flops / (field_ms * 1e-3) / 1e12 if field_ms > 0 else 0.0
    peak_tf = c.peaks["tensor"]
    q_bytes = 12.0 * st["filled"] + 4.0 * st["vis"] + 16.0 * st["cand"] + 4.0 * cfg.K * st["filled"]
    q_ms = stage_ms.get("query", 0.0)
    q_gbs = q_bytes / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0
    traffic, traffic_src = (load_traffic("field_tc_kernel", "render") if (args.points == N_POINTS and precision == "bf16") else (None, None))
    roofline = {"kernel": "field networks (gather + per-neighbour MLP + aggregation + colour MLP)", "bound": "tensor",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": c.peaks["src"] + " (sustained cuBLAS bf16)",
                "flops_per_launch": flops, "ms_per_launch": field_ms}
    stages = {"ms": stage_ms, "grid_build_ms": r["grid_build_ms"],
              "query_hbm": {"achieved": q_gbs, "peak": c.peaks["hbm"], "unit": "GB/s", "frac": q_gbs / c.peaks["hbm"],
                            "algorithmic_bytes": q_bytes, "mean_voxels_visited": st["vis"] / max(st["filled"], 1),
                            "mean_candidates": st["cand"] / max(st["filled"], 1)}}
    per_rank_ms = [None] * world
    if dist is not None:
        dist.all_gather_object(per_rank_ms, r["ms_this_rank"])
    else:
        per_rank_ms = [r["ms_this_rank"]]
    par = "one view per GPU (azimuth 30 + 45 * rank deg), cloud replicated" + ("; pixels of all views all-gathered inside the timed region"
                                                                                   if world > 1 else "")
    line = {"metric": "render rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": r["ms_total"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": "render 800x800 view, 1M-point synthetic cloud, K=8, SR=80, voxel 0.008 (configs[1])",
                       "n_points": int(cloud.xyz.shape[0]), "rays_per_step_per_gpu": R, "rays_hit": st["rays_hit"],
                       "filled_slots": st["filled"], "valid_samples_S": st["S"], "neighbour_rows_M": st["M"], "cloud": cloud.stats,
                       "l2": "256 MB flush write between timed steps (outside the event pairs)", "precision": precision,
                       "jitter": cfg.jitter, "parallelism": f"ray-sharded x{world}: {par}", "ms_per_step_by_rank": per_rank_ms},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": r["ms_e2e"] / args.steps},
            "gpu_launches": r["launches"], "clocks": clocks, "roofline": roofline, "stages": stages}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.cpu_rays
        cpu_arm(cloud, cam, c.weights, np.arange(256) * 97, cfg.SR, cfg.K)        # warm-up (builds nothing that persists)
        parity, dt = parity_leg(model, cloud, cam, c.weights, n, cfg.SR, cfg.K)
        line["parity"] = parity
        line["cpu_baseline"] = {"value": n / dt, "unit": "rays/s", "cores": cores, "kind": "port",
                                "sample": f"{n} pixels drawn uniformly from the same 800x800 view ({dt:.1f} s); C grid querier (grid "
                                          f"rebuilt per call like the reference) + torch fp32 field/compositing on {cores} threads, "
                                          "fed the GPU kernel's jittered t table so that its pixels double as the parity check"}
    def block(name, fn):
        """The extra blocks must not cost the line its headline: a failure is recorded in the block (on every rank alike)."""
        try:
            line[name] = fn()
        except Exception as e:      # noqa: BLE001
            import traceback
            line[name] = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-1500:]}

    if not args.no_train:
        block("train", lambda: bench_train(c, cam, train_steps, args.warmup))
    del model
    c.model = None
    torch.cuda.empty_cache()
    if world == 1 and not args.no_stress and args.points == N_POINTS:
        block("stress", lambda: bench_stress(c))
    if (world > 1 or args.with_scannet) and not args.no_scannet:
        block("scannet", lambda: bench_scannet(c, args.steps, args.warmup))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
