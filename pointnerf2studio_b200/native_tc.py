"""Torch-facing wrapper of the tensor-core field path (csrc/field_tc.cu): bf16 operands, fp32 accumulation in TMEM.

Host side only allocates, caches the packed bf16 weights per parameter version and launches; there is no eager or
fp32 fallback -- a missing symbol or a failed launch raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, native
from ._lib import check
from .native import LAUNCHES, MLP_PARAM_NAMES, Timers, _ptr, _stream, make_mlp, make_points

_PACK_CACHE = {}
MAX_SAMPLES_PER_LAUNCH = 1 << 23      # valid samples per field-kernel launch (4.3 GB of bf16 features); a memory knob only


def packed_weights(mlp_params, force: bool = False):
    """bf16 K-slab copy of the seven weight matrices, re-packed only when a parameter changed (force: always, into the SAME
    buffer -- the pack launch of a training step has to be part of a captured CUDA graph)."""
    lib = _lib.load()
    key = tuple((p.data_ptr(), p._version) for p in mlp_params)
    dev = mlp_params[0].device
    hit = _PACK_CACHE.get(dev)
    if hit is not None and hit[0] == key and not force:
        return hit[1], hit[2], hit[3]
    if hit is not None and force and tuple(k[0] for k in hit[0]) == tuple(k[0] for k in key):      # same storages: re-pack in place
        check(lib.pnerf_tc_pack_weights(C.byref(hit[2]), _ptr(hit[1]), _stream()), "pnerf_tc_pack_weights")
        LAUNCHES["n"] += 1
        _PACK_CACHE[dev] = (key, hit[1], hit[2], hit[3])
        return hit[1], hit[2], hit[3]
    names = [n for _, w, b in MLP_PARAM_NAMES for n in (w, b)]
    params = dict(zip(names, [p.detach().contiguous() for p in mlp_params]))
    mlp = make_mlp(params)
    buf = torch.empty(lib.pnerf_tc_wpack_bytes(), dtype=torch.uint8, device=dev)
    check(lib.pnerf_tc_pack_weights(C.byref(mlp), _ptr(buf), _stream()), "pnerf_tc_pack_weights")
    LAUNCHES["n"] += 1
    _PACK_CACHE[dev] = (key, buf, mlp, params)
    return buf, mlp, params


def class_rows(K: int):
    """Rows-per-sample classes for lists bucketed by neighbour count: the power of two >= K, then its halves down to 2."""
    kp = 2
    while kp < K:
        kp *= 2
    out = [kp]
    while out[-1] > 2:
        out.append(out[-1] // 2)
    return out


def compact_sample_classes(sample_valid, K):
    """Valid samples bucketed by neighbour count (pnerf_query's sample_valid holds the count; csrc/scan.cu): -> (ids, per-class counts on the host, rows per sample of each class).
    One host sync for the counts (the data-dependent sizes of the launches that follow)."""
    lib = _lib.load()
    n = sample_valid.numel()
    dev = sample_valid.device
    kps = class_rows(K)
    ids = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    cnt = torch.empty((len(kps),), dtype=torch.int32, device=dev)
    ws_bytes = lib.pnerf_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.pnerf_sample_compact_classes(_ptr(sample_valid, torch.uint8), n, K, len(kps), (C.c_int * len(kps))(*kps), _ptr(ids), _ptr(cnt), _ptr(ws),
                                           ws_bytes, _stream()), "pnerf_sample_compact_classes")
    LAUNCHES["n"] += 5 * len(kps)
    return ids, [int(v) for v in cnt.tolist()], kps


def field_forward_tc(cfg, q, dirs, pts, mlp, wpack, ids, counts, kps):
    """Inference: per-neighbour networks class by class (a class-c sample occupies kps[c] MMA rows), then the colour network."""
    lib = _lib.load()
    R, SR, K = q.sample_pidx.shape
    dev = dirs.device
    S = sum(counts)
    sigma = torch.zeros((R, SR), dtype=torch.float32, device=dev)
    rgb = torch.zeros((R, SR, 3), dtype=torch.float32, device=dev)
    # the aggregated features between the two kernels cost 512 B per valid sample: the sample list is walked in pieces so
    # that a dense scene (up to R * SR samples) cannot ask for more than MAX_SAMPLES_PER_LAUNCH * 512 B of workspace
    piece = min(S, MAX_SAMPLES_PER_LAUNCH)
    ws_bytes = lib.pnerf_field_tc_workspace_bytes(piece)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    bounds = [0]
    for c in counts:
        bounds.append(bounds[-1] + c)
    cam, mode = C.byref(cfg["camera"]), C.byref(cfg["mode"])
    with Timers.span("field"):
        for off in range(0, S, max(piece, 1)):
            n = min(piece, S - off)
            for ci, kp in enumerate(kps):                  # the part of class ci inside [off, off + n)
                lo, hi = max(bounds[ci], off), min(bounds[ci + 1], off + n)
                if hi <= lo:
                    continue
                check(lib.pnerf_field_forward_tc_part(C.byref(pts), cam, C.byref(mlp), _ptr(wpack), mode, _ptr(dirs), _ptr(q.sample_loc),
                                                      _ptr(q.sample_pidx), ids.data_ptr() + 4 * lo, hi - lo, kp, lo - off, SR, K, _ptr(sigma),
                                                      _ptr(ws), ws_bytes, _stream()), "pnerf_field_forward_tc_part")
                LAUNCHES["n"] += 1
            check(lib.pnerf_color_forward_tc(C.byref(pts), cam, C.byref(mlp), _ptr(wpack), mode, _ptr(dirs), ids.data_ptr() + 4 * off, n, SR,
                                             _ptr(rgb), _ptr(ws), ws_bytes, _stream()), "pnerf_color_forward_tc")
            LAUNCHES["n"] += 1
    return sigma, rgb


class _RenderTC(torch.autograd.Function):
    """Tensor-core training path: fused per-neighbour networks and colour network (operand tiles kept), step length + compositing
    as one autograd node; backward = composite backward + bf16 tcgen05 dgrad / wgrad GEMMs + scatter into the point tensors."""

    @staticmethod
    def forward(ctx, cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, *mlp_params):
        lib = _lib.load()
        wpack, mlp, params = packed_weights(mlp_params)
        pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
        ids, n_dev = native.compact_samples(q.sample_valid)
        S = int(n_dev.item())
        R, SR, K = q.sample_pidx.shape
        dev = dirs.device
        sigma = torch.zeros((R, SR), dtype=torch.float32, device=dev)
        rgb = torch.zeros((R, SR, 3), dtype=torch.float32, device=dev)
        ws_bytes = lib.pnerf_field_tc_train_workspace_bytes(S, K)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        with Timers.span("field"):
            check(lib.pnerf_field_forward_tc_train(C.byref(pts), C.byref(cfg["camera"]), C.byref(mlp), _ptr(wpack), C.byref(cfg["mode"]),
                                                   _ptr(dirs), _ptr(q.sample_loc), _ptr(q.sample_pidx), _ptr(ids), S, None, SR, K,
                                                   _ptr(sigma), _ptr(rgb), _ptr(ws), ws_bytes, _stream()), "pnerf_field_forward_tc_train")
        LAUNCHES["n"] += 7
        out = native.composite_forward(cfg, q, sigma, rgb)
        ctx.cfg, ctx.q, ctx.S, ctx.ids, ctx.ws = cfg, q, S, ids, ws
        ctx.keep = (dirs, xyz, Rw2c, embed, color, dirn, conf, params, mlp, sigma, rgb)
        ctx.shapes = [p.shape for p in mlp_params]
        cfg["last"] = {"sigma": sigma, "rgb": rgb, "n_samples": S}
        return out

    @staticmethod
    def backward(ctx, d_out):
        native.pin_stream()
        try:
            return _RenderTC._backward(ctx, d_out)
        finally:
            native.unpin_stream()

    @staticmethod
    def _backward(ctx, d_out):
        lib = _lib.load()
        cfg, q, S, ids, ws = ctx.cfg, ctx.q, ctx.S, ctx.ids, ctx.ws
        dirs, xyz, Rw2c, embed, color, dirn, conf, params, mlp, sigma, rgb = ctx.keep
        R, SR, K = q.sample_pidx.shape
        dev = dirs.device
        mode, cam = cfg["mode"], cfg["camera"]
        d_out = d_out.contiguous().float()
        d_sigma = torch.empty((R, SR), dtype=torch.float32, device=dev)
        d_rgb = torch.empty((R, SR, 3), dtype=torch.float32, device=dev)
        check(lib.pnerf_composite_backward(C.byref(cam), C.byref(mode), _ptr(q.sample_loc), _ptr(q.sample_valid), _ptr(sigma),
                                           _ptr(rgb), _ptr(d_out), R, SR, _ptr(d_sigma), _ptr(d_rgb), _stream()),
              "pnerf_composite_backward")
        need = ctx.needs_input_grad
        g_embed = torch.zeros_like(embed) if need[5] else None
        g_color = torch.zeros_like(color) if need[6] else None
        g_dir = torch.zeros_like(dirn) if need[7] else None
        g_conf = torch.zeros_like(conf) if (need[8] and mode.weight_conf) else None
        names = [n for _, w, b in MLP_PARAM_NAMES for n in (w, b)]
        flat = torch.zeros(sum(p.numel() for p in params.values()), dtype=torch.float32, device=dev)   # one memset for all 18 tensors
        grads, o = {}, 0
        for n in names:
            grads[n] = flat[o:o + params[n].numel()].view(params[n].shape)
            o += params[n].numel()
        pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
        gm = make_mlp(grads, _lib.MlpGrad)
        with Timers.span("field_bwd"):
            check(lib.pnerf_field_backward_tc(C.byref(pts), C.byref(cam), C.byref(mlp), C.byref(mode), _ptr(dirs), _ptr(q.sample_loc),
                                              _ptr(q.sample_pidx), _ptr(ids), S, None, SR, K, _ptr(d_sigma), _ptr(d_rgb), _ptr(rgb), _ptr(g_embed),
                                              _ptr(g_color), _ptr(g_dir), _ptr(g_conf), C.byref(gm), _ptr(ws), ws.numel(), None, _stream()),
                  "pnerf_field_backward_tc")
        LAUNCHES["n"] += 22
        mlp_grads = [grads[n].reshape(s) for n, s in zip(names, ctx.shapes)]
        return (None, None, None, None, None, g_embed, g_color, g_dir, g_conf, *mlp_grads)


def render_tc(cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, mlp_params):
    if torch.is_grad_enabled() and any(t.requires_grad for t in (embed, color, dirn, conf, *mlp_params)):
        return _RenderTC.apply(cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, *mlp_params)
    wpack, mlp, _ = packed_weights(mlp_params)
    pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
    with Timers.span("compact"):
        ids, counts, kps = compact_sample_classes(q.sample_valid, q.sample_pidx.shape[2])
    sigma, rgb = field_forward_tc(cfg, q, dirs, pts, mlp, wpack, ids, counts, kps)
    cfg["last"] = {"sigma": sigma, "rgb": rgb, "n_samples": sum(counts), "class_counts": counts, "class_rows": kps}
    return native.composite_forward(cfg, q, sigma, rgb)
