"""Oracle (torch, CPU, fp32) for ray generation, gather, field networks, compositing and loss.
TEST INFRASTRUCTURE ONLY.

Restates rows G0, P, GA, W, E, M1, A, M2, D, C, F, L of SURVEY.md section 8a.  Reference files
(under /root/reference/pointnerf/):
  SM = nerfstudio/studio_model.py      SU = nerfstudio/studio_utils.py
  PA = models/aggregators/point_aggregators.py
  RM = models/rendering/diff_ray_marching.py
  NPV = models/neural_points_volumetric_model.py     NW = models/helpers/networks.py

Two modes share one code path:
  mode="plugin"   : what SM:263-399 computes (LeakyReLU 0.1, ReLU density, weights without
                    confidence, colour = sum(w c) + white (1 - sum w)).
  mode="original" : what PA:745-830 + NPV:271-279 + RM:495-541 compute (LeakyReLU 0.01,
                    Softplus(raw-1) density, weight * clamped confidence,
                    colour = sum(w c) + bg * T_end).
Gradients are torch autograd through this file (the reference has no hand-written backward).
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# LeakyReLU slope: the plugin builds nn.LeakyReLU(0.1) (SM:197-218); the original flow builds
# getattr(nn, "LeakyReLU")(inplace=True) (PA:286), i.e. torch's default slope 0.01 -- a fourth
# plugin/original delta, found by executing the reference (tests/golden/make_golden.py).
SLOPE = {"plugin": 0.1, "original": 0.01}


# ----------------------------------------------------------------------------- G0
def coarse_positions(origin, dirs, D, near, far, jitter=0.0, generator=None):
    """D jittered mid-points per ray between near and far (RM:292-336; called SU:166).

    origin (3,), dirs (R,3) -> raypos (R,D,3), t_mid (R,D)
    """
    dirs = dirs.float()
    R = dirs.shape[0]
    tau = torch.linspace(0, 1, D + 1).view(1, -1)
    edge = near * (1 - tau) + far * tau
    if jitter:
        u = torch.rand((1, R, D), generator=generator)[0]
    else:
        u = torch.full((R, D), 0.5)
    seg = (edge[..., 1:] - edge[..., :-1]) * (1 + jitter * (u - 0.5))
    end = torch.cumsum(seg, dim=1)
    end = near + torch.cat([torch.zeros(R, 1), end], dim=1)
    t_mid = (end[:, :-1] + end[:, 1:]) / 2
    raypos = origin.float().view(1, 1, 3) + dirs[:, None, :] * t_mid[:, :, None]
    return raypos, t_mid


# ----------------------------------------------------------------------------- E
def positional_encoding(x, F_, ori=False):
    """NW:176-191 == SU:58-68.  ori=False: per input dim d, per f: [sin, cos] interleaved
    (index (d*F+f)*2 + {0,1}).  ori=True: [x, sin block (d-major,f-minor), cos block]."""
    freq = (2 ** torch.arange(F_).float())
    p = (x[..., None] * freq).reshape(x.shape[:-1] + (F_ * x.shape[-1],))
    if ori:
        return torch.cat([x, torch.sin(p), torch.cos(p)], dim=-1)
    return torch.stack([torch.sin(p), torch.cos(p)], dim=-1).reshape(p.shape[:-1] + (p.shape[-1] * 2,))


# ----------------------------------------------------------------------------- P
def w2pers(xyz, R_c2w, origin):
    """cam = R_c2w^T (p - o); (x/z, y/z, z)  (SU:129-144)."""
    cam = ((xyz - origin)[..., :, None] * R_c2w).sum(-2)      # R^T (p-o), mul-then-sum like SU:131,140
    return torch.stack([cam[..., 0] / cam[..., 2], cam[..., 1] / cam[..., 2], cam[..., 2]], dim=-1)


# ----------------------------------------------------------------------------- weights container
@dataclass
class FieldWeights:
    """Parameter names follow the plugin (SM:193-221); `from_aggregator` maps the original-flow
    checkpoint keys (PA:274-346) one to one."""
    p: Dict[str, torch.Tensor] = _dc_field(default_factory=dict)

    NAMES = [
        ("mlp_base.layers.0", "block1.0", 256, 284), ("mlp_base.layers.1", "block1.2", 256, 256),
        ("mlp_head.layers.0", "block3.0", 256, 263), ("mlp_head.layers.1", "block3.2", 256, 256),
        ("field_output_density.net", "alpha_branch.0", 1, 256),
        ("mlp_color.layers.0", "color_branch.0", 128, 280), ("mlp_color.layers.1", "color_branch.2", 128, 128),
        ("mlp_color.layers.2", "color_branch.4", 128, 128), ("field_output_color.net", "color_branch.6", 3, 128),
    ]

    @classmethod
    def from_aggregator(cls, sd, prefix="aggregator."):
        out = {}
        for new, old, _, _ in cls.NAMES:
            out[new + ".weight"] = sd[prefix + old + ".weight"].detach().clone().float()
            out[new + ".bias"] = sd[prefix + old + ".bias"].detach().clone().float()
        return cls(out)

    @classmethod
    def random(cls, seed=0, scale=1.0):
        g = torch.Generator().manual_seed(seed)
        out = {}
        for new, _, o, i in cls.NAMES:
            bound = scale * (6.0 / (i + o)) ** 0.5
            out[new + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
            out[new + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * 0.05
        return cls(out)

    def requires_grad_(self, flag=True):
        for v in self.p.values():
            v.requires_grad_(flag)
        return self

    def lin(self, name, x):
        return F.linear(x, self.p[name + ".weight"], self.p[name + ".bias"])


# ----------------------------------------------------------------------------- GA
def gather(pidx, points, sample_loc_w, ray_dirs, R_c2w, origin, SR):
    """SU:190-209.  pidx (R2,SR,K) int; points dict(xyz (N,3), embed (N,32), color (N,3), dir (N,3),
    conf (N,1)); sample_loc_w (R2,SR,3); ray_dirs (R2,3) of the surviving rays."""
    mask = pidx >= 0
    idx = pidx.clamp(min=0).long()
    g = {
        "mask": mask,
        "xyz": points["xyz"][idx],
        "xyz_pers": w2pers(points["xyz"], R_c2w, origin)[idx],
        "embed": points["embed"][idx],
        "color": points["color"][idx],
        "dir": points["dir"][idx],
        "conf": points["conf"][idx],
        "loc_pers": w2pers(sample_loc_w, R_c2w, origin),
        "loc_w": sample_loc_w,
        "ray_dirs": ray_dirs[:, None, :].expand(-1, SR, -1),
    }
    return g


# ----------------------------------------------------------------------------- W
def dists_and_weights(g):
    """dists6 (SM:273-281 / PA:790-799) and normalised inverse-distance weights
    (SM:284-286,467-475 / PA:419-427,816-819)."""
    pp, sp = g["xyz_pers"], g["loc_pers"][:, :, None, :]
    d_pers = torch.stack([pp[..., 0] * pp[..., 2] - sp[..., 0] * sp[..., 2],
                          pp[..., 1] * pp[..., 2] - sp[..., 1] * sp[..., 2],
                          pp[..., 2] - sp[..., 2]], dim=-1)
    dists = torch.cat([g["xyz"] - g["loc_w"][:, :, None, :], d_pers], dim=-1)
    w = g["mask"] * (1.0 / torch.clamp(torch.norm(dists[..., :3], dim=-1), min=1e-6))
    w = w / torch.clamp(w.sum(-1, keepdim=True), min=1e-8)
    return dists, w


def stclamp(x, lo=1e-4, hi=1.0):
    """Straight-through clamp (SM:289-291 / PA:740-742)."""
    return x - (x - x.clamp(lo, hi)).detach()


# ----------------------------------------------------------------------------- M1, A, M2
def bf16_st(x):
    """Round to bf16 with a straight-through gradient: where the tensor-core kernels round an MMA operand."""
    return x + (x.to(torch.bfloat16).float() - x).detach()


def field_forward(g, W: FieldWeights, Rw2c, mode="plugin", freqs=(3, 5, 4), bf16=False):
    """Per-neighbour network, aggregation and colour network (SM:300-366 / PA:486-662).

    Returns decoded (R2,SR,4) = (sigma, rgb) zero at invalid samples, valid (R2,SR) bool,
    and extras for tests.

    bf16=True restates the rounding points of the tensor-core kernels (csrc/field_tc.cu): the inputs and weights
    of mlp_base / mlp_head and the aggregated feature F_s are rounded to bf16 (fp32 accumulation, straight-through
    gradients); the reference itself is fp32 everywhere.  It exists so that the backward kernels can be checked
    against autograd at the SAME activations: with LeakyReLU a unit whose pre-activation rounds across zero changes
    its gradient tenfold, which is a property of the bf16 forward, not an error of the backward."""
    rnd = bf16_st if bf16 else (lambda t: t)

    def lin_r(name, x):
        # the kernel carries the bias as a bf16 weight column multiplying a constant-1 operand column (fp32 accumulation)
        return F.linear(rnd(x), rnd(W.p[name + ".weight"]), rnd(W.p[name + ".bias"]))

    assert mode in ("plugin", "original")
    ff, fd, fv = freqs
    LRELU = SLOPE[mode]
    mask = g["mask"]
    R2, SR, K = mask.shape
    dists, w = dists_and_weights(g)
    conf_c = stclamp(g["conf"][..., 0])
    w_used = w * conf_c if mode == "original" else w                      # PA:826 vs SM:318
    valid = mask.any(-1)
    mflat, vflat = mask.reshape(-1), valid.reshape(-1)
    Rn = Rw2c.t()                                                         # SM:303
    v = g["ray_dirs"].reshape(-1, 3) @ Rn                                 # SM:304
    venc_full = positional_encoding(v, fv, ori=True)                      # SM:305
    v_ori, venc = venc_full[:, :3], venc_full[:, 3:]
    d6 = dists.reshape(-1, 6)[mflat].clone()
    d6 = torch.cat([d6[:, :3] @ Rn, d6[:, 3:]], dim=-1)                   # SM:312
    f = g["embed"].reshape(-1, g["embed"].shape[-1])[mflat]
    x284 = torch.cat([f, positional_encoding(f, ff), positional_encoding(d6, fd)], dim=-1)   # SM:313-317
    h = F.leaky_relu(lin_r("mlp_base.layers.0", x284), LRELU)
    h = F.leaky_relu(lin_r("mlp_base.layers.1", h), LRELU)               # SM:319
    col = g["color"].reshape(-1, 3)[mflat]
    dr = g["dir"].reshape(-1, 3)[mflat] @ Rn                              # SM:330
    vk = v_ori[:, None, :].expand(-1, K, -1).reshape(-1, 3)[mflat]       # SM:331-333
    x263 = torch.cat([h, col, dr - vk, (dr * vk).sum(-1, keepdim=True)], dim=-1)            # SM:325,334
    gfeat = F.leaky_relu(lin_r("mlp_head.layers.0", x263), LRELU)
    gfeat = F.leaky_relu(lin_r("mlp_head.layers.1", gfeat), LRELU)       # SM:335
    raw = W.lin("field_output_density.net", gfeat)
    alpha = F.softplus(raw - 1) if mode == "original" else F.relu(raw)   # PA:260-265 vs SM:221
    wk = w_used.reshape(R2 * SR, K, 1)
    a_hold = torch.zeros(R2 * SR * K, 1).index_put((mflat.nonzero()[:, 0],), alpha)          # SM:339-343
    sigma = (a_hold.view(R2 * SR, K, 1) * wk).sum(-2)[vflat]             # SM:344
    f_hold = torch.zeros(R2 * SR * K, gfeat.shape[-1]).index_put((mflat.nonzero()[:, 0],), gfeat)
    Fs = (f_hold.view(R2 * SR, K, -1) * wk).sum(-2)[vflat]               # SM:348-353
    def lin_c(name, x):   # colour kernel: bf16 operands, fp32 bias added in the epilogue
        return F.linear(rnd(x), rnd(W.p[name + ".weight"]), W.p[name + ".bias"])

    cin = torch.cat([Fs, venc[vflat]], dim=-1)                            # SM:356
    c = F.leaky_relu(lin_c("mlp_color.layers.0", cin), LRELU)
    c = F.leaky_relu(lin_c("mlp_color.layers.1", c), LRELU)
    c = F.leaky_relu(lin_c("mlp_color.layers.2", c), LRELU)
    rgb = torch.sigmoid(W.lin("field_output_color.net", c)) * (1 + 2 * 0.001) - 0.001        # SM:358-359
    dec = torch.zeros(R2 * SR, 4).index_put((vflat.nonzero()[:, 0],), torch.cat([sigma, rgb], dim=-1))
    extras = {"dists": dists, "weight": w, "weight_used": w_used, "conf_coefficient": conf_c,
              "x284": x284, "x263": x263, "g": gfeat, "Fs": Fs, "alpha_rows": alpha}
    return dec.view(R2, SR, 4), valid, extras


# ----------------------------------------------------------------------------- D
def ray_dist(loc_pers, valid, vsize_z, unit_clamp=True):
    """Step length per sample (SM:368-375 / NPV:271-279)."""
    z = torch.cummax(loc_pers[..., 2], dim=-1)[0]
    d = torch.cat([z[:, 1:] - z[:, :-1], torch.full((z.shape[0], 1), float(vsize_z))], dim=-1)
    m = d < 1e-8
    if unit_clamp:
        m = m | (d > 2 * vsize_z)
    m = m.float()
    d = d * (1.0 - m) + m * vsize_z
    return d * valid.float()


# ----------------------------------------------------------------------------- C
def composite(decoded, valid, delta, mode="plugin", bg=None, training=True):
    """Alpha compositing (SM:379-390 via nerfstudio RGBRenderer / RM:495-541).

    plugin  : C = sum w c + bg (1 - sum w); eval additionally nan_to_num(rgb) and clamp [0,1]
              (nerfstudio RGBRenderer.forward; from memory, unpinned).
    original: C = sum w c + bg * prod(1 - alpha + 1e-10)."""
    sigma = decoded[..., 0] * valid.float()
    opacity = 1 - torch.exp(-sigma * delta)
    T = torch.cumprod(1.0 - opacity + 1e-10, dim=-1)
    T_end = T[:, -1:]
    T = torch.cat([torch.ones_like(T[:, :1]), T[:, :-1]], dim=-1)
    bw = opacity * T
    rgb = decoded[..., 1:4]
    if bg is None:
        bg = torch.ones(3)
    if mode == "original":
        C = (rgb * bw[..., None]).sum(-2) + bg.view(1, 3) * T_end
    else:
        if not training:
            rgb = torch.nan_to_num(rgb)
        C = (rgb * bw[..., None]).sum(-2) + bg.view(1, 3) * (1.0 - bw.sum(-1, keepdim=True))
        if not training:
            C = C.clamp(0.0, 1.0)
    return C, bw, opacity, T_end


# ----------------------------------------------------------------------------- F
def fill_invalid(C, ray_mask, bg=None):
    """Scatter the R'' colours into (R,3) pre-filled with the background (SM:491-504)."""
    if bg is None:
        bg = torch.ones(3)
    out = bg.view(1, 3).repeat(len(ray_mask), 1)
    return out.index_put((torch.as_tensor(ray_mask).bool().nonzero()[:, 0],), C)


# ----------------------------------------------------------------------------- L
def loss(C_full, ray_mask, gt, conf_coefficient=None, zero_eps=1e-3, zero_one_w=1e-4):
    """SM:415-431: masked MSE + 1e-6 and, in training, the zero-one confidence term."""
    m = torch.as_tensor(ray_mask).bool()
    out = {"ray_masked_coarse_raycolor_loss": F.mse_loss(C_full[m], gt[m]) + 1e-6}
    if conf_coefficient is not None:
        val = conf_coefficient.clamp(zero_eps, 1 - zero_eps)
        out["conf_coefficient_loss"] = (torch.log(val) + torch.log(1 - val)).mean() * zero_one_w
    return out


# ----------------------------------------------------------------------------- whole path
def render(points, W: FieldWeights, origin, dirs, R_c2w, pidx, loc_w, ray_mask, vsize_z, SR,
           mode="plugin", training=True, bg=None, bf16=False):
    """Everything after the querier: gather -> field -> step length -> composite -> fill."""
    keep = torch.as_tensor(ray_mask).bool()
    g = gather(torch.as_tensor(pidx), points, torch.as_tensor(loc_w), dirs[keep], R_c2w, origin, SR)
    dec, valid, ex = field_forward(g, W, points["Rw2c"], mode=mode, bf16=bf16)
    delta = ray_dist(g["loc_pers"], valid, vsize_z)
    C, bw, opacity, T_end = composite(dec, valid, delta, mode=mode, bg=bg, training=training)
    out = {"coarse_raycolor": fill_invalid(C, ray_mask, bg), "ray_mask": torch.as_tensor(ray_mask),
           "decoded": dec, "valid": valid, "delta": delta, "blend_weight": bw, "C_valid": C, "opacity": opacity, "gather": g,
           "conf_coefficient": ex["conf_coefficient"], "extras": ex}
    return out


# ----------------------------------------------------------------------------- hole probing
def probe(points, g, extras, opacity, valid_rays=None):
    """models/neural_points_volumetric_model.py:331-362 restated.  g = gather(...) of the surviving rays, extras = the third
    return value of field_forward, opacity (R2,SR) = per-sample opacity of composite().  Returns the probe outputs by ray."""
    R2, SR, K = g["mask"].shape
    mx, ind = torch.max(opacity, dim=-1, keepdim=True)                        # NPV:335
    take = lambda t: torch.gather(t, 1, ind.view(R2, 1, *([1] * (t.dim() - 2))).expand(-1, 1, *t.shape[2:])).squeeze(1)
    loc = take(g["loc_w"])                                                   # (R2,3)
    w = take(extras["weight_used"] * extras["conf_coefficient"])[..., None]  # NPV:340 (R2,K,1)
    xyz = take(g["xyz"])                                                     # (R2,K,3) incl. point 0 at invalid slots
    far = torch.norm(xyz - loc[:, None, :], dim=-1).min(-1, keepdim=True)[0]  # NPV:344
    out = {"ray_max_shading_opacity": mx, "ray_max_sample_loc_w": loc, "ray_max_far_dist": far,
           "shading_avg_color": (take(g["color"]) * w).sum(-2), "shading_avg_dir": (take(g["dir"]) * w).sum(-2),
           "shading_avg_conf": (take(g["conf"]) * w).sum(-2), "shading_avg_embedding": (take(g["embed"]) * w).sum(-2)}
    return out

