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


def packed_weights(mlp_params):
    """bf16 K-slab copy of the seven weight matrices, re-packed only when a parameter changed."""
    lib = _lib.load()
    key = tuple((p.data_ptr(), p._version) for p in mlp_params)
    dev = mlp_params[0].device
    hit = _PACK_CACHE.get(dev)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2], hit[3]
    names = [n for _, w, b in MLP_PARAM_NAMES for n in (w, b)]
    params = dict(zip(names, [p.detach().contiguous() for p in mlp_params]))
    mlp = make_mlp(params)
    buf = torch.empty(lib.pnerf_tc_wpack_bytes(), dtype=torch.uint8, device=dev)
    check(lib.pnerf_tc_pack_weights(C.byref(mlp), _ptr(buf), _stream()), "pnerf_tc_pack_weights")
    LAUNCHES["n"] += 1
    _PACK_CACHE[dev] = (key, buf, mlp, params)
    return buf, mlp, params


def field_forward_tc(cfg, q, dirs, pts, mlp, wpack, ids, S: int):
    lib = _lib.load()
    R, SR, K = q.sample_pidx.shape
    dev = dirs.device
    sigma = torch.zeros((R, SR), dtype=torch.float32, device=dev)
    rgb = torch.zeros((R, SR, 3), dtype=torch.float32, device=dev)
    ws_bytes = lib.pnerf_field_tc_workspace_bytes(S)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with Timers.span("field"):
        check(lib.pnerf_field_forward_tc(C.byref(pts), C.byref(cfg["camera"]), C.byref(mlp), _ptr(wpack), C.byref(cfg["mode"]),
                                         _ptr(dirs), _ptr(q.sample_loc), _ptr(q.sample_pidx), _ptr(ids), S, SR, K, _ptr(sigma),
                                         _ptr(rgb), _ptr(ws), ws_bytes, _stream()), "pnerf_field_forward_tc")
    LAUNCHES["n"] += 2
    return sigma, rgb


def render_tc(cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, mlp_params):
    if torch.is_grad_enabled() and any(t.requires_grad for t in (embed, color, dirn, conf, *mlp_params)):
        raise NotImplementedError("the tensor-core path is forward-only so far; train with PointNerfConfig(precision='fp32')")
    wpack, mlp, _ = packed_weights(mlp_params)
    pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
    ids, n_dev = native.compact_samples(q.sample_valid)
    S = int(n_dev.item())
    sigma, rgb = field_forward_tc(cfg, q, dirs, pts, mlp, wpack, ids, S)
    cfg["last"] = {"sigma": sigma, "rgb": rgb, "n_samples": S}
    return native.composite_forward(cfg, q, sigma, rgb)
