"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE on the CPU.

Runs only in the authoring container (needs /root/reference).  What is executed from the
reference, unmodified, imported from where it lies (oracle/ref_import.py):
  * models/helpers/networks.py            positional_encoding          (NW:176-191)
  * models/rendering/diff_ray_marching.py near_far_linear_ray_generation (RM:292-336), ray_march (RM:495-541)
  * models/rendering/diff_render_func.py  radiance_render, alpha_blend  (RF:36-37,48-49)
  * models/aggregators/point_aggregators.py PointAggregator.forward     (PA:745-830) with the
    shipped trained weights mvsnet_checkpoints/init/dtu_dgt_d012_img0123_conf_agg2_32_dirclr20/
    best_net_ray_marching.pth (aggregator.* keys, strict load).
The neighbour indices fed to the aggregator come from the oracle's deterministic querier (the
reference querier is CUDA-only); the gather in between restates SU:190-209 with torch index_select
so that autograd puts the gradients on the (N, .) point tensors exactly as in the plugin.

    python tests/golden/make_golden.py      # rewrites the fixtures next to this file
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import grid_query as gq                       # noqa: E402
from oracle.ref_import import chair_opt, load_reference   # noqa: E402
from pointnerf2studio_b200.synth import make_camera, make_cloud  # noqa: E402

CKPT = "/root/reference/pointnerf/mvsnet_checkpoints/init/dtu_dgt_d012_img0123_conf_agg2_32_dirclr20/best_net_ray_marching.pth"

CFG = dict(vsize=[0.004] * 3, vscale=[2, 2, 2], kernel_size=[3, 3, 3], query_size=[3, 3, 3],
           ranges=[-1.2, -1.2, -1.2, 1.2, 1.2, 1.2], D=400, SR=12, K=8, P=12, near=2.0, far=6.0)


def scene():
    cloud = make_cloud(1800, seed=4242, radii=(0.03, 0.042), P=CFG["P"])
    cam = make_camera()
    c = cam.H // 2
    pix = np.array([(c - 15 + i) * cam.W + (c - 6 + j) for i in range(24) for j in range(24)])
    return cloud, cam, pix


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    ref = load_reference()
    sd = torch.load(CKPT, map_location="cpu")
    agg_sd = {k[len("aggregator."):]: v for k, v in sd.items() if k.startswith("aggregator.")}
    np.savez_compressed(os.path.join(HERE, "aggregator_weights.npz"), **{k: v.numpy() for k, v in agg_sd.items()})

    out = {}
    # ---- positional encoding (NW:176-191)
    g = torch.Generator().manual_seed(7)
    x6 = (torch.rand(5, 6, generator=g) - 0.5) * 0.05
    x32 = torch.randn(4, 32, generator=g) * 0.3
    x3 = torch.nn.functional.normalize(torch.randn(6, 3, generator=g), dim=-1)
    out["pe_x6"], out["pe_x32"], out["pe_x3"] = x6.numpy(), x32.numpy(), x3.numpy()
    out["pe_y6"] = ref.networks.positional_encoding(x6, 5).numpy()
    out["pe_y32"] = ref.networks.positional_encoding(x32, 3).numpy()
    out["pe_y3"] = ref.networks.positional_encoding(x3, 4, ori=True).numpy()

    # ---- scene, coarse positions (RM:292-336, jitter 0)
    cloud, cam, pix = scene()
    dirs = torch.from_numpy(cam.rays(pix))
    origin = torch.from_numpy(cam.origin)
    R_c2w = torch.from_numpy(cam.R_c2w)
    raypos, _, _, t_mid = ref.diff_ray_marching.near_far_linear_ray_generation(
        origin[None], dirs[None], CFG["D"], near=CFG["near"], far=CFG["far"], jitter=0.0)
    raypos = raypos[0]
    out["pix"] = pix.astype(np.int32)
    out["raypos_first4"] = raypos[:4].numpy()
    out["t_mid"] = t_mid[0, 0].numpy()

    # ---- oracle querier (deterministic restatement; not reference-executed)
    frame = gq.hyperparameters(cloud.xyz, CFG["vsize"], CFG["vscale"], CFG["kernel_size"], CFG["ranges"])
    radius = np.float32(4 * max(CFG["vsize"][0], CFG["vsize"][1]))
    pidx, loc_w, ray_mask = gq.woord_query_grid_point_index(
        raypos.numpy(), cloud.xyz, CFG["kernel_size"], CFG["query_size"], CFG["SR"], CFG["K"], frame, CFG["P"], radius)
    out["frame_lo"], out["frame_hi"], out["frame_dim"] = frame.lo, frame.hi, frame.dim
    out["pidx"] = pidx.astype(np.int16)
    out["loc_w"] = loc_w
    out["ray_mask"] = ray_mask

    # ---- gather exactly as SU:190-209 (index_select on the point tensors), points are leaves
    pts = {"embed": torch.from_numpy(cloud.embed)[None].clone().requires_grad_(True),
           "color": torch.from_numpy(cloud.color)[None].clone().requires_grad_(True),
           "dir": torch.from_numpy(cloud.dir)[None].clone().requires_grad_(True),
           "conf": torch.from_numpy(cloud.conf)[None].clone().requires_grad_(True)}
    xyz = torch.from_numpy(cloud.xyz)
    Rw2c = torch.from_numpy(cloud.Rw2c)
    pidx_t = torch.from_numpy(pidx)[None]
    B, R2, SR, K = pidx_t.shape
    mask = pidx_t >= 0
    flat = pidx_t.clamp(min=0).view(-1).long()
    loc_w_t = torch.from_numpy(loc_w)[None]

    def w2pers_loc(p):      # SU:137-144
        s = p - origin[None, None, :]
        c = torch.sum(s[..., None, :] * R_c2w.t()[None, None, None], dim=-1)
        return torch.stack([c[..., 0] / c[..., 2], c[..., 1] / c[..., 2], c[..., 2]], dim=-1)

    def w2pers(p):          # SU:129-135
        s = p[None] - origin[None, None, :]
        c = torch.sum(R_c2w[None, None] * s[:, :, :, None], dim=-2)
        return torch.stack([c[:, :, 0] / c[:, :, 2], c[:, :, 1] / c[:, :, 2], c[:, :, 2]], dim=-1)

    loc_pers = w2pers_loc(loc_w_t)
    cat = torch.cat([xyz[None], w2pers(xyz), pts["embed"]], dim=-1)
    s_emb = torch.index_select(cat, 1, flat).view(B, R2, SR, K, -1)
    s_col = torch.index_select(pts["color"], 1, flat).view(B, R2, SR, K, 3)
    s_dir = torch.index_select(pts["dir"], 1, flat).view(B, R2, SR, K, 3)
    s_conf = torch.index_select(pts["conf"], 1, flat).view(B, R2, SR, K, 1)
    ray_dirs = dirs[torch.from_numpy(ray_mask).bool()][None, :, None, :].expand(-1, -1, SR, -1).contiguous()
    vsize = np.asarray(CFG["vsize"], dtype=np.float64)

    # ---- the reference aggregator, shipped weights, original mode (PA:745-830)
    agg = ref.point_aggregators.PointAggregator(chair_opt())
    agg.load_state_dict(agg_sd, strict=True)
    decoded, ray_valid, weight, conf_c = agg(s_col, Rw2c, s_dir, s_conf, s_emb[..., 6:], s_emb[..., 3:6],
                                              s_emb[..., :3], mask, loc_pers, loc_w_t, ray_dirs, vsize, None)
    # ---- step length: NPV:271-279 (inline code of the reference's forward, raydist_mode_unit=1)
    rd = torch.cummax(loc_pers[..., 2], dim=-1)[0]
    rd = torch.cat([rd[..., 1:] - rd[..., :-1], torch.full((rd.shape[0], rd.shape[1], 1), vsize[2])], dim=-1)
    m = torch.logical_or(rd < 1e-8, rd > 2 * vsize[2]).to(torch.float32)
    rd = rd * (1.0 - m) + m * vsize[2]
    rd = rd * ray_valid.float()
    # ---- the reference compositor (RM:495-541), white background as in the NeRF-synthetic scripts
    bg = torch.ones(1, 3)
    ray_color, point_color, opacity, acc_T, blend_w, bg_T, _ = ref.diff_ray_marching.ray_march(
        rd, ray_valid, decoded, ref.diff_render_func.radiance_render, ref.diff_render_func.alpha_blend, bg)

    gen = torch.Generator().manual_seed(11)
    gt = torch.rand(ray_color.shape, generator=gen)
    val = torch.clamp(conf_c, 1e-3, 1 - 1e-3)
    loss = torch.nn.functional.mse_loss(ray_color, gt) + 1e-4 * torch.mean(torch.log(val) + torch.log(1 - val))
    loss.backward()

    out.update(decoded=decoded[0].detach().numpy(), ray_valid=ray_valid[0].numpy(), weight=weight[0].detach().numpy(),
               conf_coefficient=conf_c[0].detach().numpy(), ray_dist=rd[0].detach().numpy().astype(np.float32),
               ray_color=ray_color[0].detach().numpy(), opacity=opacity[0].detach().numpy(),
               blend_weight=blend_w[0, ..., 0].detach().numpy(), bg_T=bg_T[0].detach().numpy(),
               gt=gt[0].numpy(), loss=np.float64(loss.item()))
    for k, v in pts.items():
        out["grad_" + k] = v.grad[0].numpy()
    for k, p in agg.named_parameters():
        out["gradw_" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "pointnerf_golden.npz"), **out)
    print("rays", len(pix), "R''", R2, "valid samples", int(ray_valid.sum()), "rows", int(mask.sum()),
          "loss", loss.item(), "cloud", cloud.stats)
    for f in ("aggregator_weights.npz", "pointnerf_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
