"""`torch.library` custom ops over the C ABI of libpnerf_b200.so ("thin C-ABI torch custom ops").

Every op is a `torch.library.custom_op` with a fake (meta) implementation -- so the boundary is traceable: shapes and dtypes
are known without running a kernel -- and, where the reference differentiates through the stage, a `register_autograd`
backward that launches the hand-written backward kernels.  The real implementations only allocate (torch owns memory and
streams) and forward raw pointers to ONE C call each; there is no eager / CPU fallback behind any of them.

  pnerf::render_train   rows G0 G2 Q P GA W E M1 A M2 D C F of SURVEY.md 8a, forward and backward, for one training batch
                        (what PointNerf.get_outputs SM:263-399 + torch autograd do for the reference).  No host sync: the
                        data-dependent counts (R'', S) stay on the device.
  pnerf::masked_mse     row L, MSE over the masked rays + 1e-6 (SM:415-426), forward and backward
  pnerf::conf_loss      row L, zero-one confidence term (SM:288-292,427-429) with its analytic gradient
  pnerf::sample_query   rows G0 G2 Q only (no gradient flows through them: SU:92-93,159)
  pnerf::adam_step      8f row 2: fused multi-tensor Adam (mutates its arguments)

Scalars travel as two flat lists so that a schema stays readable:
  fl (36 floats): grid lo[3] sv[3] | camera origin[3] R_c2w[9] | points_Rw2c[9] | near far jitter radius slope | bg[3] | vsize_z
  it (16 ints)  : grid dim[3] | seed_lo seed_hi | D SR K kernel_size0 | density_softplus weight_conf bg_mode eval_clamp |
                  t_stride | points_done_event (cudaEvent_t or 0) | workspace limit in MiB (0 = default)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, native
from ._lib import Camera, GridView, Mlp, MlpGrad, Mode, Points, RenderBuffers, check

N_FL, N_IT = 36, 16
(FL_LO, FL_SV, FL_ORIGIN, FL_RC2W, FL_RW2C, FL_NEAR, FL_FAR, FL_JITTER, FL_RADIUS, FL_SLOPE, FL_BG, FL_VSIZE_Z) = (0, 3, 6, 9, 18, 27, 28, 29, 30, 31, 32, 35)
(IT_DIM, IT_SEED_LO, IT_SEED_HI, IT_D, IT_SR, IT_K, IT_KS0, IT_SOFTPLUS, IT_WCONF, IT_BGMODE, IT_CLAMP, IT_TSTRIDE, IT_EVENT,
 IT_WS_LIMIT_MIB) = (0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
MLP_SHAPES = [(256, 284), (256,), (256, 256), (256,), (256, 263), (256,), (256, 256), (256,), (1, 256), (1,),
              (128, 280), (128,), (128, 128), (128,), (128, 128), (128,), (3, 128), (3,)]
MLP_NUMEL = [a[0] * (a[1] if len(a) > 1 else 1) for a in MLP_SHAPES]
# worst-case (R * SR samples) training workspace the sync-free path may ask for; above it the sample count is read back once
DEFAULT_WS_LIMIT_MIB = 40 * 1024


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _structs(fl, it, xyz, embed, color, dirn, conf, mlp, grid=None, step_consts=None):
    pts = Points()
    pts.xyz, pts.embed, pts.color, pts.dir, pts.conf = xyz.data_ptr(), embed.data_ptr(), color.data_ptr(), dirn.data_ptr(), conf.data_ptr()
    pts.Rw2c = (C.c_float * 9)(*fl[FL_RW2C:FL_RW2C + 9])
    pts.n = xyz.shape[0]
    cam = Camera()
    cam.origin = (C.c_float * 3)(*fl[FL_ORIGIN:FL_ORIGIN + 3])
    cam.R_c2w = (C.c_float * 9)(*fl[FL_RC2W:FL_RC2W + 9])
    if step_consts is not None:       # camera / near / far / seed read from device memory at run time (CUDA-graph replay)
        assert step_consts.is_cuda and step_consts.dtype == torch.float32 and step_consts.numel() >= 16 and step_consts.is_contiguous()
        cam.dev = step_consts.data_ptr()
    m = Mlp()
    for name, t in zip(_lib.MLP_FIELDS, mlp):
        setattr(m, name, t.data_ptr())
    mode = Mode()
    mode.lrelu_slope, mode.density_softplus, mode.weight_conf = fl[FL_SLOPE], it[IT_SOFTPLUS], it[IT_WCONF]
    mode.bg_mode, mode.eval_clamp = it[IT_BGMODE], it[IT_CLAMP]
    mode.bg = (C.c_float * 3)(*fl[FL_BG:FL_BG + 3])
    mode.vsize_z = fl[FL_VSIZE_Z]
    gv = None
    if grid is not None:
        gv = GridView()
        gv.lo = (C.c_float * 3)(*fl[FL_LO:FL_LO + 3])
        gv.sv = (C.c_float * 3)(*fl[FL_SV:FL_SV + 3])
        gv.dim = (C.c_int * 3)(*it[IT_DIM:IT_DIM + 3])
        gv.cell_start, gv.recs, gv.occ_bits = grid[0].data_ptr(), grid[1].data_ptr(), grid[2].data_ptr()
    return pts, cam, m, mode, gv


def _check_inputs(dirs, xyz, embed, color, dirn, conf, mlp):
    for t in (dirs, xyz, embed, color, dirn, conf, *mlp):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise ValueError("pnerf ops take contiguous fp32 CUDA tensors")
    if len(mlp) != 18 or any(tuple(t.shape) != s for t, s in zip(mlp, MLP_SHAPES)):
        raise ValueError("mlp: the 18 weight / bias tensors of the shipped network shape, in MLP_FIELDS order")


RenderOut = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]


@torch.library.custom_op("pnerf::render_train", mutates_args=())
def render_train(dirs: Tensor, xyz: Tensor, embed: Tensor, color: Tensor, dirn: Tensor, conf: Tensor, mlp: List[Tensor],
                 wpack: Tensor, cell_start: Tensor, recs: Tensor, occ_bits: Tensor, t_vals: Optional[Tensor],
                 sink_points: Optional[Tensor], sink_mlp: Optional[Tensor], step_consts: Optional[Tensor], fl: List[float],
                 it: List[int]) -> RenderOut:
    """-> (out_rgb (R,3), ray_mask (R) i8, n_rays (1) i32, sample_pidx (R,SR,K) i32, sample_loc (R,SR,3), sample_valid (R,SR) u8,
           sample_cnt (R) i32, sigma (R,SR), rgb (R,SR,3), sample_ids (R*SR) i32, n_samples (1) i32, workspace u8, ray_index (R) i32)

    sink_points / sink_mlp (optional): flat fp32 buffers [embed N*32 | color N*3 | dir N*3 | conf N] / the 18 MLP tensors back to
    back.  When given, the backward pass accumulates (+=) straight into them and returns no gradient for those inputs -- a
    data-parallel trainer aliases `.grad` to them so the collective runs in place (parallel.TrainEngine).
    step_consts (optional): 16 fp32 words on the device (pnerf_camera.dev layout: origin, R_c2w, near, far, jitter seed bits); the
    kernels then read the camera / near / far / seed from it at run time, so a captured CUDA graph serves every step."""
    lib = _lib.load()
    _check_inputs(dirs, xyz, embed, color, dirn, conf, mlp)
    R, dev = dirs.shape[0], dirs.device
    D, SR, K = it[IT_D], it[IT_SR], it[IT_K]
    slots = R * SR
    i32, f32 = torch.int32, torch.float32
    out_rgb = torch.empty((R, 3), dtype=f32, device=dev)
    ray_mask = torch.empty((R,), dtype=torch.int8, device=dev)
    n_rays = torch.empty((1,), dtype=i32, device=dev)
    sample_pidx = torch.empty((R, SR, K), dtype=i32, device=dev)
    sample_loc = torch.empty((R, SR, 3), dtype=f32, device=dev)
    sample_valid = torch.empty((R, SR), dtype=torch.uint8, device=dev)
    sample_cnt = torch.empty((R,), dtype=i32, device=dev)
    sigma = torch.empty((R, SR), dtype=f32, device=dev)
    rgb = torch.empty((R, SR, 3), dtype=f32, device=dev)
    sample_ids = torch.empty((max(slots, 1),), dtype=i32, device=dev)
    n_samples = torch.empty((1,), dtype=i32, device=dev)
    ray_index = torch.empty((max(R, 1),), dtype=i32, device=dev)
    if R == 0:
        n_rays.zero_(); n_samples.zero_()
        return (out_rgb, ray_mask, n_rays, sample_pidx, sample_loc, sample_valid, sample_cnt, sigma, rgb, sample_ids, n_samples,
                torch.empty((256,), dtype=torch.uint8, device=dev), ray_index)
    pts, cam, m, mode, gv = _structs(fl, it, xyz, embed, color, dirn, conf, mlp, (cell_start, recs, occ_bits), step_consts)
    scratch_bytes = lib.pnerf_render_train_scratch_bytes(R, SR)
    scratch = torch.empty((scratch_bytes,), dtype=torch.uint8, device=dev)
    b = RenderBuffers()
    b.sample_loc, b.sample_cnt, b.sample_pidx, b.sample_valid = sample_loc.data_ptr(), sample_cnt.data_ptr(), sample_pidx.data_ptr(), sample_valid.data_ptr()
    b.sample_ids, b.n_samples, b.sigma, b.rgb, b.out_rgb = sample_ids.data_ptr(), n_samples.data_ptr(), sigma.data_ptr(), rgb.data_ptr(), out_rgb.data_ptr()
    b.ray_mask, b.ray_index, b.n_rays = ray_mask.data_ptr(), ray_index.data_ptr(), n_rays.data_ptr()
    b.scratch, b.scratch_bytes = scratch.data_ptr(), scratch_bytes
    seed = (it[IT_SEED_HI] << 32) | (it[IT_SEED_LO] & 0xffffffff)
    st = _stream()

    def call(cap, phases):
        check(lib.pnerf_render_train_forward(C.byref(gv), C.byref(pts), C.byref(cam), C.byref(m), _p(wpack), C.byref(mode), _p(dirs), _p(t_vals),
                                             it[IT_TSTRIDE], C.c_float(fl[FL_NEAR]), C.c_float(fl[FL_FAR]), C.c_float(fl[FL_JITTER]),
                                             C.c_uint64(seed), R, D, SR, K, it[IT_KS0], C.c_float(fl[FL_RADIUS]), cap, phases, C.byref(b), st),
              "pnerf_render_train_forward")

    limit = (it[IT_WS_LIMIT_MIB] or DEFAULT_WS_LIMIT_MIB) << 20
    cap, ws_bytes = slots, lib.pnerf_field_tc_train_workspace_bytes(slots, K)
    with native.Timers.span("render_fwd"):
        if ws_bytes <= limit:                 # sync-free: the workspace covers every slot, the kernels clamp to the device-side count
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            b.workspace, b.workspace_bytes = ws.data_ptr(), ws_bytes
            call(cap, 3)
        else:                                 # a batch too large for that: read the sample count back once
            call(0, 1)
            cap = int(n_samples.item())
            ws_bytes = lib.pnerf_field_tc_train_workspace_bytes(cap, K)
            ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
            b.workspace, b.workspace_bytes = ws.data_ptr(), ws_bytes
            call(cap, 2)
    native.LAUNCHES["n"] += 22
    return out_rgb, ray_mask, n_rays, sample_pidx, sample_loc, sample_valid, sample_cnt, sigma, rgb, sample_ids, n_samples, ws, ray_index


@render_train.register_fake
def _(dirs, xyz, embed, color, dirn, conf, mlp, wpack, cell_start, recs, occ_bits, t_vals, sink_points, sink_mlp, step_consts, fl, it):
    R, SR, K = dirs.shape[0], it[IT_SR], it[IT_K]
    e = lambda shape, dt: dirs.new_empty(shape, dtype=dt)
    ws = torch.library.get_ctx().new_dynamic_size()
    return (e((R, 3), torch.float32), e((R,), torch.int8), e((1,), torch.int32), e((R, SR, K), torch.int32), e((R, SR, 3), torch.float32),
            e((R, SR), torch.uint8), e((R,), torch.int32), e((R, SR), torch.float32), e((R, SR, 3), torch.float32),
            e((max(R * SR, 1),), torch.int32), e((1,), torch.int32), e((ws,), torch.uint8), e((max(R, 1),), torch.int32))


def _render_train_setup(ctx, inputs, output):
    (dirs, xyz, embed, color, dirn, conf, mlp, wpack, cell_start, recs, occ_bits, t_vals, sink_points, sink_mlp, step_consts, fl, it) = inputs
    (out_rgb, ray_mask, n_rays, sample_pidx, sample_loc, sample_valid, sample_cnt, sigma, rgb, sample_ids, n_samples, ws, ray_index) = output
    ctx.save_for_backward(dirs, xyz, embed, color, dirn, conf, *mlp, sample_pidx, sample_loc, sample_valid, sample_ids, n_samples, sigma, rgb, ws)
    ctx.fl, ctx.it = list(fl), list(it)
    ctx.sinks = (sink_points, sink_mlp)
    ctx.step_consts = step_consts
    # autograd must NOT materialise zero gradients for the outputs nobody differentiates: one of them is the multi-GB workspace
    # (a zeros_like of it costs 4 ms per step), the others are index tensors
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(ray_mask, n_rays, sample_pidx, sample_loc, sample_valid, sample_cnt, sigma, rgb, sample_ids, n_samples, ws, ray_index)


def _render_train_backward(ctx, d_out, *unused):
    lib = _lib.load()
    saved = ctx.saved_tensors
    dirs, xyz, embed, color, dirn, conf = saved[:6]
    mlp = list(saved[6:24])
    sample_pidx, sample_loc, sample_valid, sample_ids, n_samples, sigma, rgb, ws = saved[24:]
    fl, it = ctx.fl, ctx.it
    sink_points, sink_mlp = ctx.sinks
    R, SR, K = dirs.shape[0], it[IT_SR], it[IT_K]
    n_in = 17
    if d_out is None or R == 0:
        return (None,) * 6 + ([None] * len(MLP_NUMEL),) + (None,) * (n_in - 7)
    dev = dirs.device
    N = xyz.shape[0]
    need = ctx.needs_input_grad      # per input; a list of flags for the List[Tensor] input
    d_out = d_out.contiguous().float()
    pts, cam, m, mode, _ = _structs(fl, it, xyz, embed, color, dirn, conf, mlp, None, ctx.step_consts)
    if sink_points is not None:
        assert sink_points.numel() == N * 39 and sink_points.is_contiguous() and sink_points.dtype == torch.float32
        g_embed, g_color, g_dir, g_conf = (sink_points[:N * 32], sink_points[N * 32:N * 35], sink_points[N * 35:N * 38], sink_points[N * 38:])
        if not need[2]: g_embed = None
        if not need[3]: g_color = None
        if not need[4]: g_dir = None
        if not (need[5] and it[IT_WCONF]): g_conf = None
        ret_pts = (None, None, None, None)
    else:
        g_embed = torch.zeros_like(embed) if need[2] else None
        g_color = torch.zeros_like(color) if need[3] else None
        g_dir = torch.zeros_like(dirn) if need[4] else None
        g_conf = torch.zeros_like(conf) if (need[5] and it[IT_WCONF]) else None
        ret_pts = (g_embed, g_color, g_dir, g_conf)
    flat = sink_mlp if sink_mlp is not None else torch.zeros((sum(MLP_NUMEL),), dtype=torch.float32, device=dev)
    gm = MlpGrad()
    views, o = [], 0
    for name, n, shape in zip(_lib.MLP_FIELDS, MLP_NUMEL, MLP_SHAPES):
        setattr(gm, name, flat.data_ptr() + 4 * o)
        if sink_mlp is None:
            views.append(flat[o:o + n].view(shape))
        o += n
    slots = R * SR
    limit = (it[IT_WS_LIMIT_MIB] or DEFAULT_WS_LIMIT_MIB) << 20
    cap = slots if lib.pnerf_field_tc_train_workspace_bytes(slots, K) <= limit else None
    if cap is None:        # the forward read the count back and sized the workspace from it
        cap = int(n_samples.item())
    scratch_bytes = lib.pnerf_render_train_scratch_bytes(R, SR)
    scratch = torch.empty((scratch_bytes,), dtype=torch.uint8, device=dev)
    b = RenderBuffers()
    b.sample_loc, b.sample_pidx, b.sample_valid = sample_loc.data_ptr(), sample_pidx.data_ptr(), sample_valid.data_ptr()
    b.sample_ids, b.n_samples, b.sigma, b.rgb = sample_ids.data_ptr(), n_samples.data_ptr(), sigma.data_ptr(), rgb.data_ptr()
    b.workspace, b.workspace_bytes, b.scratch, b.scratch_bytes = ws.data_ptr(), ws.numel(), scratch.data_ptr(), scratch_bytes
    ev = C.c_void_p(it[IT_EVENT]) if it[IT_EVENT] else None
    with native.Timers.span("render_bwd"):
        check(lib.pnerf_render_train_backward(C.byref(pts), C.byref(cam), C.byref(m), C.byref(mode), _p(dirs), _p(d_out), R, SR, K, cap, C.byref(b),
                                              _p(g_embed), _p(g_color), _p(g_dir), _p(g_conf), C.byref(gm), ev, _stream()),
              "pnerf_render_train_backward")
    native.LAUNCHES["n"] += 18
    mlp_ret = [None] * len(MLP_NUMEL) if sink_mlp is not None else views
    return (None, None, *ret_pts, mlp_ret, None, None, None, None, None, None, None, None, None, None)


render_train.register_autograd(_render_train_backward, setup_context=_render_train_setup)


# ---------------------------------------------------------------------------------------------- losses
@torch.library.custom_op("pnerf::masked_mse", mutates_args=())
def masked_mse(pred: Tensor, image: Tensor, ray_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """get_loss_dict's MSELoss over the masked rays + 1e-6 (studio_model.py:415-426): -> (loss 0-d, acc (3): sum sq, #masked, ticket)."""
    lib = _lib.load()
    pred_c, image_c = pred.contiguous(), image.contiguous()
    acc = torch.zeros((3,), dtype=torch.float32, device=pred.device)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    check(lib.pnerf_masked_mse_forward(_p(pred_c), _p(image_c), _p(ray_mask), pred_c.shape[0], _p(acc), _p(loss), _stream()),
          "pnerf_masked_mse_forward")
    native.LAUNCHES["n"] += 1
    return loss, acc


@masked_mse.register_fake
def _(pred, image, ray_mask):
    return pred.new_empty(()), pred.new_empty((3,))


def _mse_setup(ctx, inputs, output):
    pred, image, ray_mask = inputs
    ctx.save_for_backward(pred, image, ray_mask, output[1])
    ctx.mark_non_differentiable(output[1])
    ctx.set_materialize_grads(False)


def _mse_backward(ctx, d_loss, _d_acc):
    if d_loss is None:
        return None, None, None
    pred, image, ray_mask, acc = ctx.saved_tensors
    return torch.ops.pnerf.masked_mse_backward(pred, image, ray_mask, acc, d_loss), None, None


@torch.library.custom_op("pnerf::masked_mse_backward", mutates_args=())
def masked_mse_backward(pred: Tensor, image: Tensor, ray_mask: Tensor, acc: Tensor, d_loss: Tensor) -> Tensor:
    lib = _lib.load()
    pred_c, image_c = pred.contiguous(), image.contiguous()
    g = torch.empty_like(pred_c)
    d = d_loss.to(torch.float32).contiguous()
    check(lib.pnerf_masked_mse_backward(_p(pred_c), _p(image_c), _p(ray_mask), pred_c.shape[0], _p(acc), _p(d), _p(g), _stream()),
          "pnerf_masked_mse_backward")
    native.LAUNCHES["n"] += 1
    return g


@masked_mse_backward.register_fake
def _(pred, image, ray_mask, acc, d_loss):
    return torch.empty_like(pred)


masked_mse.register_autograd(_mse_backward, setup_context=_mse_setup)


@torch.library.custom_op("pnerf::conf_loss", mutates_args=())
def conf_loss(conf: Tensor, sample_pidx: Tensor, ray_mask: Tensor, n_rays: Tensor, eps: float, weight: float) -> Tuple[Tensor, Tensor]:
    """Zero-one confidence loss (studio_model.py:288-292,427-429) over all R''*SR*K slots: -> (loss 0-d, d loss / d conf (N,1))."""
    lib = _lib.load()
    R, SR, K = sample_pidx.shape
    loss = torch.zeros((), dtype=torch.float32, device=conf.device)
    g = torch.zeros_like(conf, memory_format=torch.contiguous_format)
    check(lib.pnerf_conf_loss(_p(conf.contiguous()), _p(sample_pidx), _p(ray_mask), R, SR, K, C.c_float(eps), C.c_float(weight), _p(n_rays),
                              _p(loss), _p(g), C.c_float(1.0), _stream()), "pnerf_conf_loss")
    native.LAUNCHES["n"] += 1
    return loss, g


@conf_loss.register_fake
def _(conf, sample_pidx, ray_mask, n_rays, eps, weight):
    return conf.new_empty(()), torch.empty_like(conf)


def _conf_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.mark_non_differentiable(output[1])
    ctx.set_materialize_grads(False)


def _conf_backward(ctx, d_loss, _d_g):
    if d_loss is None:
        return None, None, None, None, None, None
    (g,) = ctx.saved_tensors
    return g * d_loss, None, None, None, None, None


conf_loss.register_autograd(_conf_backward, setup_context=_conf_setup)


# ---------------------------------------------------------------------------------------------- query (no gradient)
@torch.library.custom_op("pnerf::sample_query", mutates_args=())
def sample_query(dirs: Tensor, cell_start: Tensor, recs: Tensor, occ_bits: Tensor, t_vals: Optional[Tensor], fl: List[float],
                 it: List[int]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Rows G0 / G2 / Q for R rays of one camera: -> (sample_loc (R,SR,3), sample_cnt (R), sample_pidx (R,SR,K), sample_valid (R,SR)).
    Positions: the (D,) / (R,D) table t_vals, or -- t_vals None -- jittered t generated in the kernel (fl near/far/jitter, it seed)."""
    lib = _lib.load()
    R, dev = dirs.shape[0], dirs.device
    D, SR, K = it[IT_D], it[IT_SR], it[IT_K]
    loc = torch.empty((R, SR, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((R,), dtype=torch.int32, device=dev)
    pidx = torch.empty((R, SR, K), dtype=torch.int32, device=dev)
    valid = torch.empty((R, SR), dtype=torch.uint8, device=dev)
    if R == 0:
        return loc, cnt, pidx, valid
    gv = GridView()
    gv.lo, gv.sv = (C.c_float * 3)(*fl[FL_LO:FL_LO + 3]), (C.c_float * 3)(*fl[FL_SV:FL_SV + 3])
    gv.dim = (C.c_int * 3)(*it[IT_DIM:IT_DIM + 3])
    gv.cell_start, gv.recs, gv.occ_bits = cell_start.data_ptr(), recs.data_ptr(), occ_bits.data_ptr()
    origin = (C.c_float * 3)(*fl[FL_ORIGIN:FL_ORIGIN + 3])
    st = _stream()
    if t_vals is None:
        seed = (it[IT_SEED_HI] << 32) | (it[IT_SEED_LO] & 0xffffffff)
        check(lib.pnerf_sample_select_jitter(C.byref(gv), origin, _p(dirs), C.c_float(fl[FL_NEAR]), C.c_float(fl[FL_FAR]), C.c_float(fl[FL_JITTER]),
                                             C.c_uint64(seed), R, D, SR, 1, _p(loc), _p(cnt), st), "pnerf_sample_select_jitter")
    else:
        check(lib.pnerf_sample_select(C.byref(gv), None, origin, _p(dirs), _p(t_vals), it[IT_TSTRIDE], R, D, SR, 1, _p(loc), _p(cnt), st),
              "pnerf_sample_select")
    check(lib.pnerf_query(C.byref(gv), _p(loc), _p(cnt), R, SR, K, it[IT_KS0], C.c_float(fl[FL_RADIUS]), _p(pidx), _p(valid), None, 0, st),
          "pnerf_query")
    native.LAUNCHES["n"] += 2
    return loc, cnt, pidx, valid


@sample_query.register_fake
def _(dirs, cell_start, recs, occ_bits, t_vals, fl, it):
    R, SR, K = dirs.shape[0], it[IT_SR], it[IT_K]
    return (dirs.new_empty((R, SR, 3)), dirs.new_empty((R,), dtype=torch.int32), dirs.new_empty((R, SR, K), dtype=torch.int32),
            dirs.new_empty((R, SR), dtype=torch.uint8))


# ---------------------------------------------------------------------------------------------- optimiser
@torch.library.custom_op("pnerf::adam_step", mutates_args=("params", "exp_avg", "exp_avg_sq"))
def adam_step(params: List[Tensor], grads: List[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor], steps: List[int],
              lrs: List[float], beta1: float, beta2: float, eps: float, grad_scale: float) -> None:
    """torch.optim.Adam semantics for up to 32 tensors per launch (csrc/optim.cu); `steps` are the 1-based per-tensor step counts."""
    lib = _lib.load()
    st = _stream()
    MAX = 32
    for i in range(0, len(params), MAX):
        n = min(MAX, len(params) - i)
        arr = (_lib.AdamSeg * n)()
        for j in range(n):
            p, g, m, v = params[i + j], grads[i + j], exp_avg[i + j], exp_avg_sq[i + j]
            arr[j].p, arr[j].g, arr[j].m, arr[j].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
            arr[j].n, arr[j].step, arr[j].lr = p.numel(), steps[i + j], lrs[i + j]
        check(lib.pnerf_adam_step(C.cast(arr, C.c_void_p), n, C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_float(grad_scale), st),
              "pnerf_adam_step")
        native.LAUNCHES["n"] += 1


def fl_it(frame, origin, R_c2w, Rw2c, near, far, jitter, radius, mode, D, SR, K, ks0, seed=0, t_stride=0, event=0, ws_limit_mib=0):
    """Pack the scalar arguments of the ops above (see the module docstring for the layout)."""
    fl = [float(v) for v in frame.lo] + [float(v) for v in frame.sv] + [float(v) for v in origin] + [float(v) for v in R_c2w.reshape(-1)]
    fl += [float(v) for v in Rw2c] + [float(near), float(far), float(jitter), float(radius), float(mode.lrelu_slope)]
    fl += [float(v) for v in mode.bg] + [float(mode.vsize_z)]
    it = [int(v) for v in frame.dim] + [int(seed) & 0xffffffff, (int(seed) >> 32) & 0xffffffff, int(D), int(SR), int(K), int(ks0),
                                         int(mode.density_softplus), int(mode.weight_conf), int(mode.bg_mode), int(mode.eval_clamp),
                                         int(t_stride), int(event), int(ws_limit_mib)]
    assert len(fl) == N_FL and len(it) == N_IT
    return fl, it
