"""Ray-sharded data parallelism: one process per GPU, the point cloud / grid / MLP weights are
replicated, every rank draws its own ray batch, and the gradients are summed over ranks once per step.

Mirrors what the reference gets from `DDP(model, find_unused_parameters=True)` (studio_pipeline.py:48-53): an
average of every trainable parameter's gradient -- MLP weights (1.37 MB) and the dense neural-point tensors
(156 B per point).  One flat bucket per step: over NVSwitch every peer is one hop, so a single large all-reduce
(NVLS when available) beats DDP's 25 MB buckets; parameters that received no gradient on a rank contribute
zeros (DDP's find_unused_parameters behaviour).
"""
from __future__ import annotations

import os
import types
from typing import Iterable, List

import torch


def _as_i32(u: int) -> int:
    u &= 0xffffffff
    return u - (1 << 32) if u >= (1 << 31) else u


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the ray list owned by `rank` (row-block sharding of an image)."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    return lo, min(lo + per, n_rays)


LARGE = 1 << 20     # tensors with at least this many elements are reduced in place, the rest travel in one flat bucket


def allreduce_gradients(params: Iterable[torch.nn.Parameter], dist, average: bool = True) -> None:
    """In-place sum (or mean) of `.grad` over all ranks.  The dense neural-point gradients (embedding: 128 B per point) are
    reduced where they lie -- no flatten / unflatten copies of 156 B per point -- and the small MLP gradients share one flat
    bucket; all collectives are issued asynchronously and waited for together.  With NCCL the mean is taken by the
    collective itself (ReduceOp.AVG)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    world = dist.get_world_size()
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    use_avg = average and dist.get_backend() == "nccl"
    op = dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM
    big = [p for p in params if p.grad.numel() >= LARGE and p.grad.is_contiguous()]
    small = [p for p in params if not (p.grad.numel() >= LARGE and p.grad.is_contiguous())]
    work = [dist.all_reduce(p.grad, op=op, async_op=True) for p in big]
    flat = None
    if small:
        flat = torch.cat([p.grad.reshape(-1) for p in small])
        work.append(dist.all_reduce(flat, op=op, async_op=True))
    for w in work:
        w.wait()
    if average and not use_avg:
        for p in big:
            p.grad.div_(world)
        if flat is not None:
            flat.div_(world)
    if flat is not None:
        off = 0
        for p in small:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n


def gather_pixels(local_rgb: torch.Tensor, local_mask: torch.Tensor, n_rays: int, dist):
    """Render-side collective: all-gather the per-rank (R_r,3) pixels and (R_r,) masks of a row-block-sharded
    image into the full (R,3) / (R,) tensors on every rank."""
    world = dist.get_world_size()
    per = (n_rays + world - 1) // world
    dev = local_rgb.device
    pad_rgb = torch.zeros((per, 3), dtype=local_rgb.dtype, device=dev)
    pad_mask = torch.zeros((per,), dtype=local_mask.dtype, device=dev)
    pad_rgb[:local_rgb.shape[0]] = local_rgb
    pad_mask[:local_mask.shape[0]] = local_mask
    rgb = [torch.empty_like(pad_rgb) for _ in range(world)]
    mask = [torch.empty_like(pad_mask) for _ in range(world)]
    dist.all_gather(rgb, pad_rgb)
    dist.all_gather(mask, pad_mask)
    return torch.cat(rgb)[:n_rays], torch.cat(mask)[:n_rays]


def interleaved_rows(height: int, rank: int, world: int):
    """Image rows rank, rank + world, ... : rays that miss the scene are nearly free, so contiguous row blocks leave the
    ranks holding the object's rows with most of the work (82 % strong-scaling efficiency at 2 GPUs on the ScanNet-scale
    bench); interleaving balances them."""
    return list(range(rank, height, world))


def gather_interleaved_image(local_rgb: torch.Tensor, height: int, width: int, dist):
    """All-gather the (rows_r * W, 3) pixels of interleaved row shards into the full (H * W, 3) image on every rank."""
    world = dist.get_world_size()
    per = (height + world - 1) // world
    dev = local_rgb.device
    pad = torch.zeros((per * width, 3), dtype=local_rgb.dtype, device=dev)
    pad[:local_rgb.shape[0]] = local_rgb
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    img = torch.empty((height, width, 3), dtype=local_rgb.dtype, device=dev)
    for r in range(world):
        n = len(range(r, height, world))
        img[r::world] = parts[r][:n * width].view(n, width, 3)
    return img.view(height * width, 3)



class SharedHostImage:
    """One (H, W, 3) fp32 image in POSIX shared memory, page-locked in every rank's address space.  Each rank copies the
    interleaved rows it rendered (rank, rank + world, ...) straight from its GPU into the image with ONE strided copy on its
    current stream -- no collective, no staging copy, and nobody downloads pixels it did not render (an all-gather followed
    by a full-image download on every rank moved N x the image over the host links: 22.5 ms end to end against 4.9 ms of
    device time at 8 GPUs on the ScanNet-scale image).  `image` is valid on every rank after all ranks synchronised their
    streams (+ a barrier)."""

    def __init__(self, height: int, width: int, rank: int, world: int, dist=None, tag: str = "pnerf_image"):
        import ctypes as C
        import numpy as np
        from . import _lib
        self.H, self.W, self.rank, self.world = height, width, rank, world
        self.path = f"/dev/shm/{tag}_{height}x{width}.f32"
        nbytes = height * width * 3 * 4
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if dist is not None:
            dist.barrier()
        self._map = np.memmap(self.path, dtype=np.float32, mode="r+", shape=(height, width, 3))
        self.image = torch.from_numpy(self._map)
        self._lib = _lib.load()
        self._ptr = self.image.data_ptr()
        _lib.check(self._lib.pnerf_host_register(C.c_void_p(self._ptr), nbytes), "pnerf_host_register")
        self._registered = True
        self._dist = dist

    def put_rows(self, local_rgb: torch.Tensor):
        """local_rgb (rows_r * W, 3) fp32 on the device -> rows rank, rank + world, ... of the shared image (asynchronous)."""
        import ctypes as C
        from . import _lib
        n_rows = len(range(self.rank, self.H, self.world))
        assert local_rgb.is_cuda and local_rgb.is_contiguous() and local_rgb.dtype == torch.float32
        assert local_rgb.shape[0] == n_rows * self.W
        row_bytes = self.W * 12
        _lib.check(self._lib.pnerf_copy_rows_to_host(C.c_void_p(self._ptr + self.rank * row_bytes), self.world * row_bytes,
                                                     C.c_void_p(local_rgb.data_ptr()), row_bytes, row_bytes, n_rows,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pnerf_copy_rows_to_host")

    def close(self):
        import ctypes as C
        if self._registered:
            torch.cuda.synchronize()
            self._lib.pnerf_host_unregister(C.c_void_p(self._ptr))
            self._registered = False
        if self._dist is not None:
            self._dist.barrier()
        del self.image, self._map
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class TrainEngine:
    """One training step of the plugin on N ranks: forward -> losses -> backward -> gradient exchange -> Adam on both parameter
    groups (studio_config.py:33-48) with the reference's exponential decay (studio_utils.py:38-44).

    All trainable parameters of a rank live in ONE flat fp32 buffer P = [embedding N*32 | colour N*3 | dir N*3 | conf N | the 18
    MLP tensors] (the module's Parameters become views of it, names and shapes unchanged) and all gradients in a second flat
    buffer G with the same layout (`.grad` of every Parameter is a view of G; the backward kernels accumulate straight into it
    through PointNerf.set_grad_sink, so no gradient is ever copied, concatenated or re-allocated).  Gradient exchange:

      "p2p"   (default at world > 1 when torch's symmetric memory is available): P and G are peer-mapped on every rank and ONE
              kernel (pnerf_dp_adam_step) does reduce-scatter + Adam + all-gather over NVLink between two device-side barriers;
              the Adam moments are sharded (rank r keeps them for its 1/world slice only).  With NVLS multicast addresses
              (NVSwitch) the sum is one multimem.ld_reduce and the broadcast one multimem.st per 16 bytes.
      "nccl"  one in-place NCCL all-reduce per bucket -- the point gradients start on a side stream as soon as the backward has
              scattered them (under the weight-gradient GEMMs), the confidence + MLP bucket follows -- then the same fused Adam
              kernel runs replicated on every rank.  This is DDP's data movement.
      "local" world == 1.
    """

    def __init__(self, model, dist=None, exchange: str = "auto", lr_fields: float = 5e-4, lr_points: float = 2e-3,
                 lr_decay_exp: float = 0.1, lr_decay_iters: int = 1000000, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_graph: bool = False):
        from . import _lib
        self.model, self.dist = model, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.lr0 = (float(lr_points), float(lr_fields))
        self.decay = (float(lr_decay_exp), float(lr_decay_iters))
        self.betas, self.eps, self.steps = betas, float(eps), 0
        self.timing = None           # set to a list to collect (start, end) CUDA-event pairs around update()
        self.use_graph = bool(use_graph)
        # NVLS multimem forms: measured on B200 / NVSwitch, update() of the 157 MB flat buffers: 8 GPUs 0.45 ms against 0.55 ms with plain
        # peer loads / stores; 2 GPUs 0.50 against 0.32 ms (one peer: the switch adds a hop and saves nothing) -> on from 3 ranks up
        mc = os.environ.get("PNERF_DP_MULTICAST")
        self.use_multicast = (self.world > 2) if mc is None else (mc != "0")
        self._graphs = {}            # number of rays -> captured step
        self._lib = _lib.load()
        npnts = model.neural_points
        dev = npnts.points_xyz.device
        N = npnts.points_xyz.shape[0]
        pts = [npnts.points_embeding, npnts.points_color, npnts.points_dir, npnts.points_conf]
        mlp = model.mlp_param_list()
        assert [p.numel() for p in pts] == [N * 32, N * 3, N * 3, N]
        self.N, self.boundary = N, N * 39
        self.total = self.boundary + sum(p.numel() for p in mlp)
        W = self.world
        self.slice = (self.total + 4 * W - 1) // (4 * W) * 4
        padded = self.slice * W
        if exchange == "auto":
            exchange = "local" if W == 1 else ("p2p" if self._symm_ok() else "nccl")
        self.exchange = exchange
        self._hdl_p = self._hdl_g = None
        if exchange == "p2p":
            # symmetric memory is a collective set-up: every rank reports how far it got and all fall back to NCCL together
            def agree(ok: bool) -> bool:
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                return bool(flag.item())
            why = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                group = dist.group.WORLD.group_name
                self.P = symm_mem.empty(padded, dtype=torch.float32, device=dev)
                self.G = symm_mem.empty(padded, dtype=torch.float32, device=dev)
            except Exception as e:      # noqa: BLE001
                why = e
            if agree(why is None):
                try:
                    self._hdl_p, self._hdl_g = symm_mem.rendezvous(self.P, group), symm_mem.rendezvous(self.G, group)
                    self.P.zero_()
                except Exception as e:  # noqa: BLE001
                    why = e
            if not agree(why is None):
                import warnings
                warnings.warn(f"TrainEngine: peer-mapped buffers unavailable ({why!r}); falling back to the NCCL all-reduce route")
                exchange = self.exchange = "nccl"
                self._hdl_p = self._hdl_g = None
        if exchange != "p2p":
            self.P = torch.zeros(padded, dtype=torch.float32, device=dev)
            self.G = torch.empty(padded, dtype=torch.float32, device=dev)
        self.G.zero_()
        self.params = pts + mlp
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                view = self.P[off:off + n].view(p.shape)
                view.copy_(p.detach())
                p.data = view                     # same Parameter object (optimiser / checkpoint code keeps working), storage in P
                if p.requires_grad:
                    p.grad = self.G[off:off + n].view(p.shape)
                off += n
        assert off == self.total
        lo = self.rank * self.slice if exchange == "p2p" else 0
        hi = min(lo + self.slice, padded) if exchange == "p2p" else padded
        self.lo, self.hi = lo, hi
        self.m = torch.zeros(hi - lo, dtype=torch.float32, device=dev)
        self.v = torch.zeros(hi - lo, dtype=torch.float32, device=dev)
        self._ev = torch.cuda.Event()
        self._comm = torch.cuda.Stream(device=dev) if exchange == "nccl" else None
        self._side = torch.cuda.Stream(device=dev) if exchange == "p2p" else None
        # starting the bulk of the exchange under the weight-gradient GEMMs was measured and LOST (2 GPUs: 1.84 against 1.66 ms per
        # step): the exchange kernel and the spinning barrier take SMs from wgrad_tc_kernel (1 CTA per SM, all of TMEM).  Off by default.
        self.overlap = os.environ.get("PNERF_DP_OVERLAP", "0") != "0"
        self._backward_ran = False
        self._ev.record()
        model.set_grad_sink(self.G[:self.boundary], self.G[self.boundary:self.total], self._ev.cuda_event if exchange in ("nccl", "p2p") else 0)
        if exchange == "p2p":
            self._hdl_p.barrier()

    def _symm_ok(self) -> bool:
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
            return self.dist.get_backend() == "nccl"
        except Exception:
            return False

    def lrs(self):
        """LambdaLR semantics: optimiser step k (1-based) runs with lr0 * lambda(k - 1)."""
        f = pow(self.decay[0], self.steps / self.decay[1])
        return self.lr0[0] * f, self.lr0[1] * f

    def backward_and_update(self, loss: torch.Tensor, hyper_dev=None):
        loss.backward()
        self._backward_ran = True          # the event inside the backward was recorded for THIS step: the early exchange may wait on it
        self.update(hyper_dev)

    def _adam(self, a, lo, hi, lr_p, lr_f, hyper_dev):
        """One pnerf_dp_adam_step launch over the flat elements [lo, hi) of this rank's slice."""
        import ctypes as C
        from . import _lib
        if hi <= lo:
            return
        a.m, a.v = self.m.data_ptr() + 4 * (lo - self.lo), self.v.data_ptr() + 4 * (lo - self.lo)
        a.lo, a.hi, a.boundary, a.step = lo, hi, self.boundary, self.steps
        a.lr = (C.c_float * 2)(lr_p, lr_f)
        if hyper_dev is not None:
            a.hyper_dev = hyper_dev.data_ptr()
        _lib.check(self._lib.pnerf_dp_adam_step(C.byref(a), C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                                C.c_float(1.0 / self.world), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "pnerf_dp_adam_step")

    def update(self, hyper_dev=None):
        """Gradient exchange + Adam + gradient reset, all enqueued on the current stream (nothing blocks the host).
        hyper_dev: device tensor {lr_points / bc1, lr_fields / bc1, 1 / sqrt(bc2)} read by the kernel at run time (graph replay)."""
        from . import _lib, native
        lr_p, lr_f = self.lrs()
        self.steps += 1
        if self.timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
        a = _lib.DpAdam()
        W = self.world
        n_early = (self.N * 38) // 4 * 4           # embedding / colour / dir gradients: complete at the event recorded inside the backward
        if self.exchange == "p2p":
            for w in range(W):
                a.p[w], a.g[w] = self._hdl_p.buffer_ptrs[w], self._hdl_g.buffer_ptrs[w]
            a.world, a.rank = W, self.rank
            if self.use_multicast and self._hdl_p.multicast_ptr and self._hdl_g.multicast_ptr:
                a.mc_p, a.mc_g = self._hdl_p.multicast_ptr, self._hdl_g.multicast_ptr     # NVLS: reduce / broadcast inside the switch
            cur = torch.cuda.current_stream()
            mid = min(max(n_early, self.lo), self.hi)
            if self.overlap and self._backward_ran:
                # the bulk of the exchange (97 % of the bytes) starts on a side stream as soon as the backward has scattered the
                # point gradients, under its weight-gradient GEMMs; the confidence + MLP tail follows the whole backward
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._ev)
                    self._hdl_g.barrier(channel=1)
                    self._adam(a, self.lo, mid, lr_p, lr_f, hyper_dev)
                self._hdl_g.barrier(channel=0)            # every rank's backward is complete
                self._adam(a, mid, self.hi, lr_p, lr_f, hyper_dev)
                cur.wait_stream(self._side)
            else:
                self._hdl_g.barrier(channel=0)            # every rank's gradients are complete
                self._adam(a, self.lo, self.hi, lr_p, lr_f, hyper_dev)
            self._hdl_p.barrier(channel=0)                # every rank has read my gradients and written my parameters
        else:
            if self.exchange == "nccl":
                cur = torch.cuda.current_stream()
                with torch.cuda.stream(self._comm):
                    self._comm.wait_event(self._ev)
                    w1 = self.dist.all_reduce(self.G[:n_early], async_op=True)
                w2 = self.dist.all_reduce(self.G[n_early:self.total], async_op=True)
                w1.wait(); w2.wait()
                cur.wait_stream(self._comm)
            a.p[0], a.g[0] = self.P.data_ptr(), self.G.data_ptr()
            a.world, a.rank = 1, 0
            self._adam(a, self.lo, self.hi, lr_p, lr_f, hyper_dev)
        self._backward_ran = False
        self.G.zero_()
        native.LAUNCHES["n"] += 2
        torch.autograd.graph.increment_version(self.params)
        if self.timing is not None:
            t1.record()
            self.timing.append((t0, t1))

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def _hyper(self, step):
        """What pnerf_dp_adam_step derives on the host for 1-based optimiser step `step` (bias corrections in double, as torch)."""
        f = pow(self.decay[0], (step - 1) / self.decay[1])
        bc1, bc2 = 1.0 - self.betas[0] ** step, 1.0 - self.betas[1] ** step
        return [self.lr0[0] * f / bc1, self.lr0[1] * f / bc1, 1.0 / (bc2 ** 0.5)]

    def _consts(self, ray_bundle, step):
        """The 19 words a replay needs: camera, near / far, jitter seed (pnerf_camera.dev layout) + the Adam step scalars."""
        npnts = self.model.neural_points
        origin, R_c2w, near, far = npnts.camera_of(ray_bundle, with_near_far=True)
        npnts._jitter_calls += 1
        seed = (int(self.model.config.jitter_seed) << 32) | (npnts._jitter_calls & 0xffffffff)
        npnts._last_jitter = (near, far, float(self.model.config.jitter), seed)
        vals = [float(v) for v in origin] + [float(v) for v in R_c2w.reshape(-1)] + [float(near), float(far)]
        host = torch.tensor(vals + [0.0, 0.0] + self._hyper(step), dtype=torch.float32)
        iv = host.view(torch.int32)
        iv[14], iv[15] = _as_i32(seed & 0xffffffff), _as_i32(seed >> 32)
        return host

    def _capture(self, ray_bundle, image):
        from .model import RayBundle
        dev = self.P.device
        R = ray_bundle.directions.shape[0]
        c = types.SimpleNamespace()
        c.dirs = torch.empty((R, 3), dtype=torch.float32, device=dev)
        c.gt = torch.empty((R, 3), dtype=torch.float32, device=dev)
        c.consts = torch.zeros((19,), dtype=torch.float32, device=dev)
        origin, R_c2w, near, far = self.model.neural_points.camera_of(ray_bundle, with_near_far=True)
        cam = torch.zeros((14,), dtype=torch.float32, device=dev)
        c.bundle = RayBundle(origins=cam[0:3].view(1, 3).expand(R, 3), directions=c.dirs, nears=cam[12:13].view(1, 1).expand(R, 1),
                             fars=cam[13:14].view(1, 1).expand(R, 1),
                             metadata={"camrotc2w": cam[3:12].view(3, 3),
                                       "camera_host": {"origin": origin, "camrotc2w": R_c2w, "near": near, "far": far}})
        c.dirs.copy_(ray_bundle.directions)
        c.gt.copy_(image)
        self.model.set_step_consts(c.consts[:16])
        self.model._force_repack = True          # the bf16 weight pack is a launch of every step: it has to be inside the graph
        hyper = c.consts[16:19]

        def eager():
            out = self.model.get_outputs(c.bundle)
            loss = sum(self.model.get_loss_dict(out, {"image": c.gt}).values())
            self.backward_and_update(loss, hyper)
            return loss

        calls0, steps0 = self.model.neural_points._jitter_calls, self.steps
        # warm-up on a side stream (allocator pools, lazy module state), with learning rates of zero so that nothing moves
        c.consts.copy_(self._consts(ray_bundle, 1))
        c.consts[16:18] = 0.0
        keep_m, keep_v = self.m.clone(), self.v.clone()
        # the capture stream differs from the stream the parameters' AccumulateGrad nodes were created on: expected here
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                eager()
        torch.cuda.current_stream().wait_stream(side)
        from . import native
        c.graph = torch.cuda.CUDAGraph()
        n0 = native.LAUNCHES["n"]
        with torch.cuda.graph(c.graph):
            c.loss = eager()
        c.launches = native.LAUNCHES["n"] - n0          # kernels of one replay
        # warm-up and capture ran Adam's moment updates with real gradients: restore the state (the parameters did not move: lr 0)
        self.m.copy_(keep_m); self.v.copy_(keep_v)
        self.steps = steps0
        self.model.neural_points._jitter_calls = calls0
        self.model.set_step_consts(None)
        self.model._force_repack = False
        torch.cuda.synchronize()
        return c

    def step_graph(self, ray_bundle, image):
        R = ray_bundle.directions.shape[0]
        c = self._graphs.get(R)
        if c is None:
            c = self._graphs[R] = self._capture(ray_bundle, image)
        self.steps += 1
        c.consts.copy_(self._consts(ray_bundle, self.steps))      # pageable -> device: staged synchronously, safe to rebuild next step
        c.dirs.copy_(ray_bundle.directions, non_blocking=True)
        c.gt.copy_(image, non_blocking=True)
        c.graph.replay()
        torch.autograd.graph.increment_version(self.params)
        from . import native
        native.LAUNCHES["n"] += c.launches
        return c.loss

    def step(self, ray_bundle, image):
        """-> the (rank-local) loss tensor; nothing is read back.  With use_graph the whole step (forward, losses, backward,
        exchange, Adam) is ONE captured CUDA graph per batch size, replayed after three small copies (directions, ground truth,
        19 step constants): the host cost of a step drops from ~1.7 ms of Python to ~0.1 ms."""
        if (self.use_graph and self.exchange in ("local", "p2p") and float(self.model.config.jitter) > 0 and self.model.training
                and self.model.config.precision == "bf16"):
            return self.step_graph(ray_bundle, image)
        out = self.model.get_outputs(ray_bundle)
        loss_dict = self.model.get_loss_dict(out, {"image": image})
        loss = sum(loss_dict.values())
        self.backward_and_update(loss)
        return loss
