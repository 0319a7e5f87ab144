"""The one-call training path (pnerf::render_train custom op, device-side sample count, gradient sink) and parallel.TrainEngine
(flat parameter / gradient buffers + pnerf_dp_adam_step) against the step-by-step path of round 1: the same kernels driven stage by
stage through autograd.Function with fresh gradient tensors and the per-group FusedAdam."""
import numpy as np
import pytest
import torch

from oracle import field as of
from test_gpu_parity import _bundle, _make_model, _scene

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    scale = float(b.abs().max())
    err = float((a - b).abs().max())
    assert err <= tol * scale + 1e-12, (what, err, scale)


@pytest.mark.parametrize("flow", ["plugin", "original"])
def test_fused_training_op_matches_the_staged_path(flow):
    from pointnerf2studio_b200 import native_tc
    s, cloud, cam, pix = _scene("config1")
    pix = pix[:512]
    W = of.FieldWeights.random(seed=11, scale=1.5)
    gt = torch.rand((len(pix), 3), generator=torch.Generator().manual_seed(1)).cuda()
    grads = {}
    for path in ("fused", "staged"):
        model = _make_model(cloud, "bf16", flow, SR=24, K=s["K"], P=s["P"], weights=W).train()
        rb = _bundle(cam, pix)
        if path == "staged":       # round 1's route: stage-by-stage launches through autograd.Function (host knows S)
            q, origin, R_c2w, dirs = model.neural_points.query(rb)
            from pointnerf2studio_b200 import native
            mode = native.make_mode(flow, training=True, bg=[1.0, 1.0, 1.0], vsize_z=0.004)
            cfg = {"mode": mode, "camera": native.make_camera(origin, R_c2w)}
            npnts = model.neural_points
            rgb = native_tc._RenderTC.apply(cfg, q, dirs, npnts.points_xyz, npnts.points_Rw2c, npnts.points_embeding.view(-1, 32),
                                            npnts.points_color.view(-1, 3), npnts.points_dir.view(-1, 3), npnts.points_conf.view(-1, 1),
                                            *model.mlp_param_list())
            _, _, ray_mask, _, n_rays = native.compact_rays(q)
            from pointnerf2studio_b200.model import ConfCoefficient
            out = {"coarse_raycolor": rgb, "ray_mask": ray_mask, "conf_coefficient": ConfCoefficient(npnts.points_conf, q.sample_pidx, ray_mask, n_rays)}
        else:
            out = model.get_outputs(rb)
        ld = model.get_loss_dict(out, {"image": gt})
        sum(ld.values()).backward()
        torch.cuda.synchronize()
        grads[path] = ({k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None},
                       out["coarse_raycolor"].detach().clone(), {k: v.detach().clone() for k, v in ld.items()})
    (ga, ca, la), (gb, cb, lb) = grads["fused"], grads["staged"]
    assert torch.equal(ca, cb)                                   # same forward kernels, same inputs: identical pixels
    for k in lb:
        assert abs(float(la[k]) - float(lb[k])) <= 1e-6 * max(1.0, abs(float(lb[k])))
    assert set(ga) == set(gb) and len(ga) >= 21
    for k in gb:                                                  # fp32 atomics reorder sums: not bit-equal
        _close(ga[k], gb[k], 2e-3, k)


def test_train_engine_tracks_the_per_group_optimisers():
    from pointnerf2studio_b200.optim import make_optimizers
    from pointnerf2studio_b200.parallel import TrainEngine
    s, cloud, cam, pix = _scene("config1")
    pix = pix[:512]
    W = of.FieldWeights.random(seed=12, scale=1.5)
    gt = torch.rand((len(pix), 3), generator=torch.Generator().manual_seed(2)).cuda()
    rb = _bundle(cam, pix)
    a = _make_model(cloud, "bf16", "plugin", SR=24, K=s["K"], P=s["P"], weights=W).train()
    b = _make_model(cloud, "bf16", "plugin", SR=24, K=s["K"], P=s["P"], weights=W).train()
    opts, scheds = make_optimizers(a, lr_fields=5e-3, lr_points=2e-2)
    eng = TrainEngine(b, None, lr_fields=5e-3, lr_points=2e-2)
    assert eng.exchange == "local"
    names = [n for n, _ in b.named_parameters()]
    assert names == [n for n, _ in a.named_parameters()]          # still the reference's parameter names / shapes
    assert b.neural_points.points_embeding.shape == (1, len(cloud.xyz), 32)
    first = {}
    for it in range(3):
        for p in a.parameters():
            p.grad = None
        out = a.get_outputs(rb)
        la = sum(a.get_loss_dict(out, {"image": gt}).values())
        la.backward()
        if it == 0:
            first = {n: p.grad.detach().clone() for n, p in a.named_parameters() if p.grad is not None}
            outb = b.get_outputs(rb)
            lb = sum(b.get_loss_dict(outb, {"image": gt}).values())
            lb.backward()
            for n, p in b.named_parameters():
                if p.requires_grad:
                    _close(p.grad, first[n], 2e-3, n)             # the sink holds what autograd would have produced
            eng.update()
        else:
            lb = eng.step(rb, gt)
        for k in opts:
            opts[k].step()
            scheds[k].step()
        assert abs(float(la.detach()) - float(lb.detach())) <= 2e-3 * max(abs(float(la.detach())), 1e-3), (it, float(la.detach()), float(lb.detach()))
    assert float(eng.G.abs().max()) == 0.0                        # gradients are reset for the next step
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for n in names:
        if not pa[n].requires_grad:
            continue
        d = (pa[n].detach() - pb[n].detach()).abs()
        # Adam normalises by |g|: an element whose tiny gradient changed sign under the atomics' reordering moves by up to 2 lr per step
        assert float(d.mean()) <= 2e-4 and float((d > 1e-3).float().mean()) <= 0.02, (n, float(d.mean()), float(d.max()))
    with torch.no_grad():
        a.eval(); b.eval()
        ca = a.get_outputs(rb)["coarse_raycolor"]
        cb = b.get_outputs(rb)["coarse_raycolor"]
    assert float((ca - cb).abs().max()) <= 2e-2


def test_conf_coefficient_behaves_as_the_reference_tensor():
    """outputs["conf_coefficient"] (SM:396-397) is the (1,R'',SR,K) gathered, straight-through-clamped confidence tensor for any
    torch consumer, although the loss kernel never materialises it."""
    s, cloud, cam, pix = _scene("config1")
    pix = pix[:256]
    model = _make_model(cloud, "bf16", "plugin", SR=24, K=s["K"], P=s["P"]).train()
    out = model.get_outputs(_bundle(cam, pix))
    cc = out["conf_coefficient"]
    n_hit = int(out["ray_mask"].sum())
    assert tuple(cc.shape) == (1, n_hit, 24, s["K"])
    eps = 1e-3
    ref = torch.clamp(cc, eps, 1 - eps)                             # SM:427 applied to the object itself
    loss_ref = 1e-4 * torch.mean(torch.log(ref) + torch.log(1 - ref))
    ld = model.get_loss_dict(out, {"image": torch.zeros((len(pix), 3)).cuda()})
    assert abs(float(loss_ref) - float(ld["conf_coefficient_loss"])) <= 1e-6 * abs(float(loss_ref)) + 1e-9


def test_graph_replay_tracks_the_eager_engine_over_changing_cameras():
    """TrainEngine(use_graph=True): the whole step captured once as a CUDA graph and replayed with a NEW camera / jitter seed /
    learning-rate schedule per step (19 words of device-side step constants) against the eager engine on the same batches."""
    from pointnerf2studio_b200 import RayBundle
    from pointnerf2studio_b200.parallel import TrainEngine
    from pointnerf2studio_b200.synth import make_camera
    s, cloud, cam0, pix = _scene("config1")
    pix = pix[:512]
    W = of.FieldWeights.random(seed=13, scale=1.5)
    gts = [torch.rand((len(pix), 3), generator=torch.Generator().manual_seed(20 + i)).cuda() for i in range(5)]
    cams = [make_camera(azim_deg=30.0 + 17.0 * i, elev_deg=20.0 + 3.0 * i) for i in range(5)]
    res = {}
    for mode in ("eager", "graph"):
        m = _make_model(cloud, "bf16", "plugin", SR=24, K=s["K"], P=s["P"], weights=W).train()
        m.config.jitter = 0.3                       # the in-kernel jitter (seed = call counter) is what a graph replays
        eng = TrainEngine(m, None, lr_fields=5e-3, lr_points=2e-2, lr_decay_iters=10, use_graph=(mode == "graph"))
        losses = []
        for i in range(5):
            rb = RayBundle.for_camera(torch.from_numpy(cams[i].rays(pix)).cuda(), cams[i].origin, cams[i].R_c2w, cams[i].near, cams[i].far)
            losses.append(float(eng.step(rb, gts[i]).detach()))
        torch.cuda.synchronize()
        if mode == "graph":
            assert len(eng._graphs) == 1 and eng.steps == 5
        with torch.no_grad():
            m.eval()
            img = m.get_outputs(RayBundle.for_camera(torch.from_numpy(cams[0].rays(pix)).cuda(), cams[0].origin, cams[0].R_c2w, 2.0, 6.0))
        res[mode] = (losses, {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}, img["coarse_raycolor"].clone(),
                     m.neural_points._jitter_calls)
    (la, pa, ia, ja), (lb, pb, ib, jb) = res["eager"], res["graph"]
    assert ja == jb                                  # the same jitter stream was consumed
    assert len(set(round(l, 5) for l in la)) == 5    # five different batches
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-3 * max(abs(x), 1e-3), (la, lb)
    for n in pa:
        d = (pa[n] - pb[n]).abs()
        assert float(d.mean()) <= 3e-4 and float((d > 2e-3).float().mean()) <= 0.02, (n, float(d.mean()), float(d.max()))
    assert float((ia - ib).abs().max()) <= 2e-2


def test_custom_ops_pass_opcheck():
    """torch.library.opcheck on the registered ops: the schema matches what the implementations do (nothing mutated, no aliasing) and
    the fake (meta) implementations produce the shapes / dtypes / devices of the real ones -- i.e. the boundary is traceable."""
    from pointnerf2studio_b200 import native, native_tc, ops
    s, cloud, cam, pix = _scene("config1")
    pix = pix[:256]
    m = _make_model(cloud, "bf16", "plugin", SR=24, K=s["K"], P=s["P"]).train()
    npnts, c = m.neural_points, m.config
    rb = _bundle(cam, pix)
    origin, R_c2w, near, far = npnts.camera_of(rb, with_near_far=True)
    grid = npnts.grid()
    mode = native.make_mode("plugin", training=True, bg=[1.0, 1.0, 1.0], vsize_z=0.004)
    t = npnts.coarse_t(len(pix), near, far, 0.0)
    fl, it = ops.fl_it(grid.frame, origin, R_c2w, native._rw2c_host(npnts.points_Rw2c), near, far, 0.0, float(npnts.radius_limit_np), mode,
                       c.z_depth_dim, c.SR, c.K, 3)
    params = [p.detach() for p in m.mlp_param_list()]
    wpack = native_tc.packed_weights(m.mlp_param_list())[0]
    dirs = rb.directions.contiguous()
    args = (dirs, npnts.points_xyz.detach(), npnts.points_embeding.detach().view(-1, 32), npnts.points_color.detach().view(-1, 3),
            npnts.points_dir.detach().view(-1, 3), npnts.points_conf.detach().view(-1, 1), params, wpack, grid.cell_start, grid.recs,
            grid.occ_bits, t, None, None, None, fl, it)
    checks = ("test_schema", "test_faketensor")
    torch.library.opcheck(torch.ops.pnerf.render_train.default, args, test_utils=checks)
    torch.library.opcheck(torch.ops.pnerf.sample_query.default, (dirs, grid.cell_start, grid.recs, grid.occ_bits, t, fl, it), test_utils=checks)
    out = torch.ops.pnerf.render_train(*args)
    pred, ray_mask, n_rays, pidx = out[0], out[1], out[2], out[3]
    img = torch.rand_like(pred)
    torch.library.opcheck(torch.ops.pnerf.masked_mse.default, (pred, img, ray_mask), test_utils=checks + ("test_autograd_registration",))
    torch.library.opcheck(torch.ops.pnerf.conf_loss.default, (npnts.points_conf.detach().view(-1, 1), pidx, ray_mask, n_rays, 1e-3, 1e-4),
                          test_utils=checks + ("test_autograd_registration",))
