"""The `pointnerf-original` registration and the Nerfstudio `Model` contract, executed against a stub `nerfstudio` package
(tests/nerfstudio_stub.py: Nerfstudio is not installable here).  Each test runs in a subprocess because the package decides at
import time whether `PointNerf` subclasses Nerfstudio's `Model`."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, timeout=900):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]))
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def test_method_specification_against_stub_nerfstudio():
    out = _run("""
        import nerfstudio_stub as S
        S.install()
        import importlib, tomllib
        from pointnerf2studio_b200 import model, nerfstudio_plugin as P
        assert model.HAVE_NERFSTUDIO and P.HAVE_NERFSTUDIO
        assert issubclass(model.PointNerf, S.Model) and issubclass(model.PointNerfConfig, S.ModelConfig)
        spec = P.pointnerf_original
        cfg = spec.config
        assert isinstance(spec, S.MethodSpecification) and cfg.method_name == "pointnerf-original"
        assert cfg.max_num_iterations == 200000 and cfg.steps_per_save == 25000 and cfg.steps_per_eval_batch == 1000
        assert cfg.steps_per_eval_image == 2000 and cfg.steps_per_eval_all_images == 100000
        pl = cfg.pipeline
        assert pl._target is P.PointNerfPipeline and pl.datamanager._target is P.PointNerfDataManager
        assert pl.datamanager.train_num_rays_per_batch == 4096 and pl.datamanager.eval_num_rays_per_batch == 4096
        assert pl.datamanager.random_image_idx is True and pl.datamanager.near_plane == 2.0 and pl.datamanager.far_plane == 6.0
        assert pl.model._target is model.PointNerf and pl.model.eval_num_rays_per_chunk == 2304
        assert pl.model.enable_collider and pl.model.collider_params == {"near_plane": 2.0, "far_plane": 6.0}
        assert set(cfg.optimizers) == {"fields", "neural_points"}
        assert cfg.optimizers["fields"]["optimizer"].lr == 0.0005 and cfg.optimizers["neural_points"]["optimizer"].lr == 0.002
        import torch
        w = torch.nn.Parameter(torch.zeros(2))
        opt = cfg.optimizers["fields"]["optimizer"].setup(params=[w])
        sched = cfg.optimizers["fields"]["scheduler"].setup().get_scheduler(optimizer=opt, lr_init=0.0005)
        for _ in range(1000):
            opt.step(); sched.step()
        assert abs(opt.param_groups[0]["lr"] - 0.0005 * 0.1 ** (1000 / 1e6)) < 1e-12
        # the entry point of pyproject.toml resolves to this very object
        ep = tomllib.load(open("pyproject.toml", "rb"))["project"]["entry-points"]["nerfstudio.method_configs"]["pointnerf2studio"]
        mod, attr = ep.split(":")
        assert getattr(importlib.import_module(mod), attr) is spec
        print("spec ok")
        """)
    assert "spec ok" in out


def test_model_keeps_working_without_nerfstudio():
    out = _run("""
        from pointnerf2studio_b200 import model, nerfstudio_plugin as P
        import torch
        assert not model.HAVE_NERFSTUDIO and not P.HAVE_NERFSTUDIO and not hasattr(P, "pointnerf_original")
        assert model.PointNerf.__mro__[1] is torch.nn.Module
        c = model.NearFarCollider(2.0, 6.0)
        rb = model.RayBundle(torch.zeros(5, 3), torch.zeros(5, 3), None, None, {})
        rb = c(rb)
        assert rb.nears.shape == (5, 1) and float(rb.fars[3]) == 6.0
        print("plain ok")
        """)
    assert "plain ok" in out


@pytest.mark.gpu
def test_pipeline_call_sequence_on_the_gpu(tmp_path):
    """Trainer.train_iteration's sequence (zero_grad_all -> get_train_loss_dict -> backward -> optimizer / scheduler steps), an eval
    batch and a full eval image through PointNerfPipeline / PointNerfDataManager / PointNerf, with the stub's torch.optim.Adam
    optimisers -- the kernels behind get_outputs / get_loss_dict / backward are the product's."""
    out = _run(f"""
        import nerfstudio_stub as S
        S.install()
        import functools, random, torch
        torch.manual_seed(0); random.seed(0)
        from pointnerf2studio_b200 import nerfstudio_plugin as P
        from pointnerf2studio_b200.synth import make_cloud
        cloud = make_cloud(50000, seed=1241, radii=(0.11, 0.16, 0.2), P=12)
        torch.save(cloud.state_dict(), r"{tmp_path}/0_net_ray_marching.pth")
        torch.save({{"total_steps": 0}}, r"{tmp_path}/0_states.pth")
        import pathlib
        spec = P._make_spec(path_point_cloud=pathlib.Path(r"{tmp_path}"))
        cams, imgs = S.synthetic_scene(3)
        dm = spec.config.pipeline.datamanager
        dm.stub_cameras, dm.stub_images = cams, imgs
        pipe = spec.config.pipeline.setup(device="cuda", test_mode="val", world_size=1, local_rank=0)
        assert isinstance(pipe, P.PointNerfPipeline) and isinstance(pipe.datamanager, P.PointNerfDataManager)
        model = pipe.model
        assert model.collider is not None and model.num_train_data == 3
        with torch.no_grad():          # a freshly initialised density head can sit entirely below its ReLU (sigma = 0, no gradient at all)
            model.field_output_density.net.bias.fill_(5.0)
        opts = S.Optimizers(spec.config.optimizers, pipe.get_param_groups())
        before = {{n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}}
        pipe.train()
        losses = []
        for step in range(3):
            opts.zero_grad_all()
            outputs, loss_dict, metrics = pipe.get_train_loss_dict(step)
            assert outputs["coarse_raycolor"].shape == (4096, 3) and outputs["ray_mask"].shape == (4096,) and outputs["ray_mask"].dtype == torch.int8
            assert set(loss_dict) == {{"ray_masked_coarse_raycolor_loss", "conf_coefficient_loss"}}
            loss = functools.reduce(torch.add, loss_dict.values())
            loss.backward()
            opts.optimizer_step_all()
            opts.scheduler_step_all(step)
            losses.append(float(loss))
        assert all(l == l and abs(l) < 1e3 for l in losses), losses
        moved = [n for n, p in model.named_parameters() if p.requires_grad and not torch.equal(p.detach(), before[n])]
        assert any(n.startswith("neural_points.points_embeding") for n in moved) and any(n.startswith("mlp_base") for n in moved), moved
        outputs, loss_dict, _ = pipe.get_eval_loss_dict(0)
        assert set(loss_dict) == {{"ray_masked_coarse_raycolor_loss"}}
        m, images = pipe.get_eval_image_metrics_and_images(0)
        assert images["img"].shape == (800, 1600, 3) and m["num_rays"] == 640000 and m["psnr"] > 0
        print("pipeline ok", losses, m["psnr"])
        """)
    assert "pipeline ok" in out
