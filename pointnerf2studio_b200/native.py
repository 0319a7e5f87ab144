"""Torch-facing wrappers of the C ABI: device memory, streams and autograd plumbing only.

Everything that computes is a kernel in libpnerf_b200.so; torch allocates the buffers, supplies the
current stream and records the autograd edge.  No function here has a CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Camera, GridView, Mlp, MlpGrad, Mode, Points, check

LAUNCHES = {"n": 0}   # kernels launched through this module (bench.py reports it)


_STREAM = {"h": None}


def pin_stream():
    """Look the current torch stream up once per entry point (get_outputs / backward / loss): torch.cuda.current_stream()
    costs ~20 us and a step makes ~180 calls through this module."""
    _STREAM["h"] = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def unpin_stream():
    _STREAM["h"] = None


def _stream():
    return _STREAM["h"] if _STREAM["h"] is not None else C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor], dtype=None):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return C.c_void_p(t.data_ptr())


def _f3(x):
    return (C.c_float * 3)(*[float(v) for v in x])


# ---------------------------------------------------------------------------------------------- grid
@dataclass
class GridFrame:
    lo: np.ndarray     # (3,) f32   ranges_tensor[:3]
    hi: np.ndarray     # (3,) f32   ranges_tensor[3:]
    sv: np.ndarray     # (3,) f32   scaled_vsize
    dim: np.ndarray    # (3,) i32   scaled_vdim

    @property
    def cells(self):
        return int(self.dim[0]) * int(self.dim[1]) * int(self.dim[2])


def get_hyperparameters(xyz: torch.Tensor, vsize, vscale, kernel_size, ranges) -> GridFrame:
    """NeuralPoints.get_hyperparameters (studio_utils.py:115-127): the two reductions run in
    pnerf_bbox, the float64 dim arithmetic stays on the host exactly as numpy does it there."""
    lib = _lib.load()
    mm = torch.empty(6, dtype=torch.float32, device=xyz.device)
    pts = xyz.reshape(-1, 3)
    check(lib.pnerf_bbox(_ptr(pts, torch.float32), pts.shape[0], _ptr(mm), _stream()), "pnerf_bbox")
    LAUNCHES["n"] += 2
    mm = mm.cpu().numpy()
    mn, mx = mm[:3], mm[3:]
    if ranges is not None:
        r = np.asarray(ranges, dtype=np.float32)
        mn, mx = np.maximum(mn, r[:3]), np.minimum(mx, r[3:])
    vscale_i = np.asarray(vscale, dtype=np.int32)
    sv = (np.asarray(vsize, dtype=np.float64) * vscale_i).astype(np.float32)
    half = (sv.astype(np.float64) * np.asarray(kernel_size, dtype=np.int64) / 2).astype(np.float32)
    lo, hi = (mn - half).astype(np.float32), (mx + half).astype(np.float32)
    vdim = (hi - lo).astype(np.float32).astype(np.float64) / np.asarray(vsize, dtype=np.float64)
    dim = np.ceil(vdim / vscale_i).astype(np.int32)
    return GridFrame(lo=lo, hi=hi, sv=sv, dim=dim)


class VoxelGrid:
    """CSR voxel buckets + occupancy bitmask of one point-cloud version (grid.cu)."""

    def __init__(self, xyz: torch.Tensor, frame: GridFrame, P: int, query_size: Sequence[int]):
        lib = _lib.load()
        pts = xyz.detach().reshape(-1, 3).contiguous().float()
        n, G = pts.shape[0], frame.cells
        dev = pts.device
        self.frame, self.P, self.n = frame, int(P), n
        self.cell_start = torch.empty(G + 1, dtype=torch.int32, device=dev)
        self.recs = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
        self.occ_bits = torch.empty((G + 31) // 32, dtype=torch.int32, device=dev)
        ws_bytes = lib.pnerf_grid_workspace_bytes(n, G)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        lo, sv = _f3(frame.lo), _f3(frame.sv)
        dim = (C.c_int * 3)(*[int(v) for v in frame.dim])
        qs = (C.c_int * 3)(*[int(v) for v in query_size])
        check(lib.pnerf_grid_build(_ptr(pts), n, lo, sv, dim, int(P), qs, _ptr(self.cell_start), _ptr(self.recs),
                                   _ptr(self.occ_bits), _ptr(ws), ws_bytes, _stream()), "pnerf_grid_build")
        LAUNCHES["n"] += 12
        self._xyz = pts
        del ws
        v = GridView()
        v.lo, v.sv = lo, sv
        v.dim = dim
        v.cell_start, v.recs, v.occ_bits = self.cell_start.data_ptr(), self.recs.data_ptr(), self.occ_bits.data_ptr()
        self.view = v

    @property
    def n_kept(self) -> int:
        return int(self.cell_start[-1].item())

    def bytes(self) -> int:
        return self.cell_start.numel() * 4 + self.n * 16 + self.occ_bits.numel() * 4


@dataclass
class QueryResult:
    sample_loc: torch.Tensor      # (R,SR,3) f32, zeros where empty
    sample_cnt: torch.Tensor      # (R,) i32
    sample_pidx: torch.Tensor     # (R,SR,K) i32, -1 padded
    sample_valid: torch.Tensor    # (R,SR) u8
    stats: Optional[torch.Tensor] = None   # (2,) u64 as int64: voxel entries visited, candidates examined
    # hit-ray compaction: when ray_index is set, the tensors above hold only the R' rays whose selection found an occupied
    # position (row i = ray ray_index[i] of the R_total rays of the call), and `dirs` holds their directions
    ray_index: Optional[torch.Tensor] = None
    R_total: Optional[int] = None
    dirs: Optional[torch.Tensor] = None

    def dense(self) -> "QueryResult":
        """The (R_total, ...) tensors of the uncompacted call (what the reference's mask_raypos stage produces)."""
        if self.ray_index is None:
            return self
        R, (R2, SR, K) = self.R_total, self.sample_pidx.shape
        dev, idx = self.sample_pidx.device, self.ray_index.long()
        loc = torch.zeros((R, SR, 3), dtype=torch.float32, device=dev).index_copy_(0, idx, self.sample_loc)
        cnt = torch.zeros((R,), dtype=torch.int32, device=dev).index_copy_(0, idx, self.sample_cnt)
        pidx = torch.full((R, SR, K), -1, dtype=torch.int32, device=dev).index_copy_(0, idx, self.sample_pidx)
        valid = torch.zeros((R, SR), dtype=torch.uint8, device=dev).index_copy_(0, idx, self.sample_valid)
        return QueryResult(loc, cnt, pidx, valid, self.stats)


def coarse_t(near: float, far: float, jitter: float, seed: int, R: int, D: int, device, want_u: bool = False):
    """The (R,D) t mid-points (and uniforms) the jittered sample selection generates in registers -- checker hook."""
    lib = _lib.load()
    t = torch.empty((R, D), dtype=torch.float32, device=device)
    u = torch.empty((R, D), dtype=torch.float32, device=device) if want_u else None
    check(lib.pnerf_coarse_t(C.c_float(near), C.c_float(far), C.c_float(jitter), C.c_uint64(seed), R, D, _ptr(t), _ptr(u), _stream()),
          "pnerf_coarse_t")
    return (t, u) if want_u else t


def sample_and_query(grid: VoxelGrid, R: int, D: int, SR: int, K: int, kernel_size0: int, radius: float,
                     raypos: Optional[torch.Tensor] = None, origin=None, dirs: Optional[torch.Tensor] = None,
                     t_vals: Optional[torch.Tensor] = None, want_stats: bool = False, jitter_gen=None,
                     compact: bool = False, across_rays: Optional[bool] = None) -> QueryResult:
    """Rows G0/G2/Q: select the first SR occupied coarse positions per ray and query K neighbours each.
    Position source: `raypos` (R,D,3), or origin + dirs * `t_vals` ((D,) or (R,D)), or -- `jitter_gen` =
    (near, far, jitter, seed) -- jittered t generated inside the selection kernel.
    compact=True: the rays whose selection found nothing (5 of 6 for an object-centred view) are dropped right after the
    selection -- the reference's own R -> R' step (CU:381-391) -- and the query and everything after it run on R' rays
    (one host sync for R'); the result carries `ray_index`."""
    lib = _lib.load()
    dev = grid.cell_start.device
    loc = torch.empty((R, SR, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((R,), dtype=torch.int32, device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev) if want_stats else None
    fill = 0 if compact else 1
    t_stride = 0
    if raypos is None and jitter_gen is None:
        assert dirs is not None and t_vals is not None and origin is not None
        t_stride = 0 if t_vals.dim() == 1 else D
    with Timers.span("select"):
        if jitter_gen is not None:
            near, far, jitter, seed = jitter_gen
            check(lib.pnerf_sample_select_jitter(C.byref(grid.view), _f3(origin), _ptr(dirs, torch.float32), C.c_float(near),
                                                 C.c_float(far), C.c_float(jitter), C.c_uint64(int(seed)), R, D, SR, fill, _ptr(loc),
                                                 _ptr(cnt), _stream()), "pnerf_sample_select_jitter")
        else:
            check(lib.pnerf_sample_select(C.byref(grid.view), _ptr(raypos, torch.float32), _f3(origin) if origin is not None else None,
                                          _ptr(dirs, torch.float32), _ptr(t_vals, torch.float32), t_stride, R, D, SR, fill, _ptr(loc),
                                          _ptr(cnt), _stream()), "pnerf_sample_select")
    ray_index, R_total, dirs_c = None, None, None
    if compact:
        with Timers.span("compact"):
            ray_index = torch.empty((max(R, 1),), dtype=torch.int32, device=dev)
            n_dev = torch.empty((1,), dtype=torch.int32, device=dev)
            ws_bytes = lib.pnerf_scan_workspace_bytes(R)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.pnerf_hit_rays(_ptr(cnt), R, _ptr(ray_index), _ptr(n_dev), _ptr(ws), ws_bytes, _stream()), "pnerf_hit_rays")
            R2 = int(n_dev.item()) if R > 0 else 0
            loc2 = torch.empty((R2, SR, 3), dtype=torch.float32, device=dev)
            cnt2 = torch.empty((R2,), dtype=torch.int32, device=dev)
            dirs_c = torch.empty((R2, 3), dtype=torch.float32, device=dev) if dirs is not None else None
            check(lib.pnerf_gather_hit_rays(_ptr(ray_index), R2, SR, _ptr(loc), _ptr(cnt), _ptr(dirs, torch.float32), _ptr(loc2), _ptr(cnt2),
                                            _ptr(dirs_c), _stream()), "pnerf_gather_hit_rays")
        LAUNCHES["n"] += 5
        ray_index, R_total, loc, cnt, R = ray_index[:R2], R, loc2, cnt2, R2
    pidx = torch.empty((R, SR, K), dtype=torch.int32, device=dev)
    valid = torch.empty((R, SR), dtype=torch.uint8, device=dev)
    with Timers.span("query"):
        # the compacted hit-ray list of an image keeps neighbouring pixels next to each other: warps take one slot of 32 rays
        hint = compact if across_rays is None else bool(across_rays)
        check(lib.pnerf_query(C.byref(grid.view), _ptr(loc), _ptr(cnt), R, SR, K, int(kernel_size0), C.c_float(float(radius)),
                              _ptr(pidx), _ptr(valid), _ptr(stats), 1 if hint else 0, _stream()), "pnerf_query")
    LAUNCHES["n"] += 2
    return QueryResult(loc, cnt, pidx, valid, stats, ray_index, R_total, dirs_c)


def compact_rays(q: QueryResult):
    """The reference op's compact return value (query_worldcoords.cu:425-432).  Costs one host sync
    for the data-dependent R'' (the reference spends five)."""
    lib = _lib.load()
    R, SR, K = q.sample_pidx.shape
    dev = q.sample_pidx.device
    ray_mask = torch.empty((R,), dtype=torch.int8, device=dev)
    ray_index = torch.empty((max(R, 1),), dtype=torch.int32, device=dev)
    n_rays = torch.empty((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib.pnerf_scan_workspace_bytes(R)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.pnerf_ray_compact(_ptr(q.sample_valid), R, SR, _ptr(ray_mask), _ptr(ray_index), _ptr(n_rays), _ptr(ws),
                                ws_bytes, _stream()), "pnerf_ray_compact")
    R2 = int(n_rays.item()) if R > 0 else 0
    out_pidx = torch.empty((R2, SR, K), dtype=torch.int32, device=dev)
    out_loc = torch.empty((R2, SR, 3), dtype=torch.float32, device=dev)
    check(lib.pnerf_gather_rays(_ptr(ray_index), R2, SR, K, _ptr(q.sample_pidx), _ptr(q.sample_loc), _ptr(out_pidx),
                                _ptr(out_loc), _stream()), "pnerf_gather_rays")
    LAUNCHES["n"] += 5
    return out_pidx, out_loc, ray_mask, ray_index[:R2], n_rays


def compact_samples(sample_valid: torch.Tensor):
    """Ascending list of slot ids (r*SR+s) that have >= 1 neighbour, and their count (device)."""
    lib = _lib.load()
    n = sample_valid.numel()
    dev = sample_valid.device
    ids = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib.pnerf_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.pnerf_sample_compact(_ptr(sample_valid), n, _ptr(ids), _ptr(cnt), _ptr(ws), ws_bytes, _stream()),
          "pnerf_sample_compact")
    LAUNCHES["n"] += 5
    return ids, cnt


# ---------------------------------------------------------------------------------------------- structs
MLP_PARAM_NAMES = [("mlp_base.layers.0", "w1", "b1"), ("mlp_base.layers.1", "w2", "b2"),
                   ("mlp_head.layers.0", "w3", "b3"), ("mlp_head.layers.1", "w4", "b4"),
                   ("field_output_density.net", "wa", "ba"),
                   ("mlp_color.layers.0", "wc1", "bc1"), ("mlp_color.layers.1", "wc2", "bc2"),
                   ("mlp_color.layers.2", "wc3", "bc3"), ("field_output_color.net", "wc4", "bc4")]
MLP_SHAPES = {"w1": (256, 284), "w2": (256, 256), "w3": (256, 263), "w4": (256, 256), "wa": (1, 256),
              "wc1": (128, 280), "wc2": (128, 128), "wc3": (128, 128), "wc4": (3, 128)}


def make_mode(mode: str = "plugin", training: bool = True, bg=(1.0, 1.0, 1.0), vsize_z: float = 0.004) -> Mode:
    m = Mode()
    original = mode == "original"
    assert mode in ("plugin", "original")
    m.lrelu_slope = 0.01 if original else 0.1
    m.density_softplus = 1 if original else 0
    m.weight_conf = 1 if original else 0
    m.bg_mode = 1 if original else 0
    m.eval_clamp = 0 if (training or original) else 1
    m.bg = _f3(bg)
    m.vsize_z = float(np.float32(vsize_z))
    return m


_RW2C_HOST = {}


def _rw2c_host(Rw2c):
    """Host copy of points_Rw2c (a no-grad 3x3), fetched once per tensor version instead of one D2H sync per call.  The cache
    entry holds a reference to the tensor it was read from, so its storage cannot be freed and handed to another tensor while
    the entry is alive: (object identity, version) then identifies the contents."""
    hit = _RW2C_HOST.get("e")
    if hit is None or hit[0] is not Rw2c or hit[1] != Rw2c._version:
        hit = _RW2C_HOST["e"] = (Rw2c, Rw2c._version, [float(v) for v in Rw2c.detach().reshape(-1).cpu().tolist()])
    return hit[2]


def make_points(xyz, embed, color, dirn, conf, Rw2c) -> Points:
    p = Points()
    p.xyz, p.embed, p.color = xyz.data_ptr(), embed.data_ptr(), color.data_ptr()
    p.dir, p.conf = dirn.data_ptr(), conf.data_ptr()
    p.Rw2c = (C.c_float * 9)(*_rw2c_host(Rw2c))
    p.n = xyz.reshape(-1, 3).shape[0]
    for t in (xyz, embed, color, dirn, conf):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    return p


def make_camera(origin, R_c2w) -> Camera:
    c = Camera()
    c.origin = _f3(origin)
    c.R_c2w = (C.c_float * 9)(*[float(v) for v in np.asarray(R_c2w, dtype=np.float32).reshape(-1)])
    return c


def make_mlp(params: dict, cls=Mlp):
    m = cls()
    for _, w, b in MLP_PARAM_NAMES:
        for n in (w, b):
            t = params[n]
            assert t is None or (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32), n
            setattr(m, n, None if t is None else t.data_ptr())
    return m


# ---------------------------------------------------------------------------------------------- field + composite
F32_EVAL_CHUNK = 131072   # samples per field launch when no activations are kept (workspace = 61 KB per sample at K=8)


class Timers:
    """Optional CUDA-event timers around named stages (bench.py turns them on; off = zero cost)."""
    enabled = False
    spans = []   # (name, start_event, end_event)

    @classmethod
    def span(cls, name):
        return _Span(name) if cls.enabled else _NULL_SPAN

    @classmethod
    def collect(cls):
        """-> {name: [ms, ...]} (call after torch.cuda.synchronize())."""
        out = {}
        for name, a, b in cls.spans:
            out.setdefault(name, []).append(a.elapsed_time(b))
        cls.spans = []
        return out


class _Span:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()
        return self

    def __exit__(self, *exc):
        self.b.record()
        Timers.spans.append((self.name, self.a, self.b))
        return False


class _NullSpan:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_SPAN = _NullSpan()


def _names():
    return [n for _, w, b in MLP_PARAM_NAMES for n in (w, b)]


def field_forward_f32(cfg, q: QueryResult, dirs, pts: Points, mlp: Mlp, ids, S: int, keep_workspace: bool):
    """Launch the fp32 field networks over the compact sample list `ids[:S]`.  With keep_workspace the whole
    list goes in one launch and the activation workspace is returned for the backward pass; otherwise the
    list is processed in fixed-size chunks that reuse one workspace (full-image rendering)."""
    lib = _lib.load()
    R, SR, K = q.sample_pidx.shape
    dev = dirs.device
    mode, cam = cfg["mode"], cfg["camera"]
    sigma = torch.zeros((R, SR), dtype=torch.float32, device=dev)
    rgb = torch.zeros((R, SR, 3), dtype=torch.float32, device=dev)
    step = S if keep_workspace else min(S, F32_EVAL_CHUNK)
    ws_bytes = lib.pnerf_field_f32_workspace_bytes(step, K)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with Timers.span("field"):
        for s0 in range(0, S, max(step, 1)):
            n = min(step, S - s0)
            check(lib.pnerf_field_forward_f32(C.byref(pts), C.byref(cam), C.byref(mlp), C.byref(mode), _ptr(dirs),
                                              _ptr(q.sample_loc), _ptr(q.sample_pidx), C.c_void_p(ids.data_ptr() + 4 * s0), n, SR,
                                              K, _ptr(sigma), _ptr(rgb), _ptr(ws), ws_bytes, _stream()), "pnerf_field_forward_f32")
            LAUNCHES["n"] += 11
    return sigma, rgb, (ws if keep_workspace else None)


def composite_forward(cfg, q: QueryResult, sigma, rgb):
    lib = _lib.load()
    R, SR, K = q.sample_pidx.shape
    out = torch.empty((R, 3), dtype=torch.float32, device=sigma.device)
    with Timers.span("composite"):
        check(lib.pnerf_composite_forward(C.byref(cfg["camera"]), C.byref(cfg["mode"]), _ptr(q.sample_loc), _ptr(q.sample_valid),
                                          _ptr(sigma), _ptr(rgb), R, SR, _ptr(out), None, None, _stream()), "pnerf_composite_forward")
    LAUNCHES["n"] += 1
    return out


class _RenderF32(torch.autograd.Function):
    """fp32 path: field networks + step length + compositing as one autograd node."""

    @staticmethod
    def forward(ctx, cfg, q: QueryResult, dirs, xyz, Rw2c, embed, color, dirn, conf, *mlp_params):
        params = dict(zip(_names(), [p.detach().contiguous() for p in mlp_params]))
        pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
        mlp = make_mlp(params)
        ids, n_dev = compact_samples(q.sample_valid)
        S = int(n_dev.item())
        need_bwd = cfg.get("need_bwd", True) and any(ctx.needs_input_grad)
        sigma, rgb, ws = field_forward_f32(cfg, q, dirs, pts, mlp, ids, S, keep_workspace=need_bwd)
        out = composite_forward(cfg, q, sigma, rgb)
        ctx.cfg, ctx.q, ctx.S, ctx.ids, ctx.ws = cfg, q, S, ids, ws
        ctx.keep = (dirs, xyz, Rw2c, embed, color, dirn, conf, params, sigma, rgb)
        ctx.shapes = [p.shape for p in mlp_params]
        ctx.extra = {"sigma": sigma, "rgb": rgb, "n_samples": S}
        cfg["last"] = ctx.extra
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        cfg, q, S, ids, ws = ctx.cfg, ctx.q, ctx.S, ctx.ids, ctx.ws
        dirs, xyz, Rw2c, embed, color, dirn, conf, params, sigma, rgb = ctx.keep
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
        grads = {n: torch.zeros_like(p) for n, p in params.items()}
        pts = make_points(xyz.detach(), embed.detach(), color.detach(), dirn.detach(), conf.detach(), Rw2c)
        mlp, gm = make_mlp(params), make_mlp(grads, MlpGrad)
        ws_bytes = ws.numel()
        check(lib.pnerf_field_backward_f32(C.byref(pts), C.byref(cam), C.byref(mlp), C.byref(mode), _ptr(dirs), _ptr(q.sample_loc),
                                           _ptr(q.sample_pidx), _ptr(ids), S, SR, K, _ptr(d_sigma), _ptr(d_rgb), _ptr(g_embed),
                                           _ptr(g_color), _ptr(g_dir), _ptr(g_conf), C.byref(gm), _ptr(ws), ws_bytes, _stream()),
              "pnerf_field_backward_f32")
        LAUNCHES["n"] += 30
        mlp_grads = [grads[n].reshape(s) for n, s in zip(_names(), ctx.shapes)]
        return (None, None, None, None, None, g_embed, g_color, g_dir, g_conf, *mlp_grads)


def render_f32(cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, mlp_params):
    cfg["need_bwd"] = torch.is_grad_enabled()
    return _RenderF32.apply(cfg, q, dirs, xyz, Rw2c, embed, color, dirn, conf, *mlp_params)


def masked_mse(pred: torch.Tensor, image: torch.Tensor, ray_mask: torch.Tensor):
    """get_loss_dict's MSELoss over the masked rays + 1e-6 (studio_model.py:415-426): the `pnerf::masked_mse` custom op (ops.py)."""
    from . import ops  # noqa: F401  (registers the op)
    return torch.ops.pnerf.masked_mse(pred, image.contiguous(), ray_mask)[0]


def conf_loss(conf: torch.Tensor, q_pidx: torch.Tensor, ray_mask: torch.Tensor, n_rays: torch.Tensor, eps: float, weight: float):
    """Zero-one confidence loss of studio_model.py:288-292,427-429 with its analytic gradient: the `pnerf::conf_loss` custom op."""
    from . import ops  # noqa: F401
    return torch.ops.pnerf.conf_loss(conf, q_pidx, ray_mask, n_rays, float(eps), float(weight))[0]
