"""`ns-train pointnerf-original` registration (reference: nerfstudio/studio_config.py:13-54, studio_datamanager.py:36-110,
studio_pipeline.py:16-53, pyproject.toml:20-21).  Nerfstudio is an optional dependency that is not installed in the build image;
everything the hot path needs lives in model.py and runs without it.  When it is importable this module defines, like the
reference,

  PointNerfDataManagerConfig / PointNerfDataManager   one camera per batch (random image, SD:66-73), `metadata["camrotc2w"]`
                                                       injected into every bundle (SD:79,100,108) -- plus the `camera_host`
                                                       hint, since the datamanager has the pose on the host anyway (no read-back)
  PointNerfPipeline                                    datamanager + model set-up, DDP wrap at world_size > 1 (SP:29-53)
  PointNerfSchedulerConfig / PointNerfScheduler        lr * 0.1 ** (step / 1e6) (SU:24-44)
  pointnerf_original                                   the MethodSpecification behind the entry point

with the reference's TrainerConfig values: 4096 train / eval rays per batch, eval chunk 2304, 200 k iterations, Adam 5e-4
("fields") and 2e-3 ("neural_points").  tests/test_nerfstudio_plugin.py executes all of it against a stub package.
"""
from __future__ import annotations

import random

METHOD_NAME = "pointnerf-original"
TRAINER_VALUES = dict(max_num_iterations=200000, steps_per_save=25000, steps_per_eval_batch=1000,
                      steps_per_eval_image=2000, steps_per_eval_all_images=100000,
                      train_num_rays_per_batch=4096, eval_num_rays_per_batch=4096, eval_num_rays_per_chunk=2304,
                      lr_fields=0.0005, lr_neural_points=0.002, lr_decay_exp=0.1, lr_decay_iters=1000000)


def lr_lambda(step, lr_decay_exp=0.1, lr_decay_iters=1000000):
    """PointNerfScheduler (studio_utils.py:38-44)."""
    return pow(lr_decay_exp, step / lr_decay_iters)


try:
    from nerfstudio.engine.optimizers import AdamOptimizerConfig
    from nerfstudio.engine.schedulers import Scheduler, SchedulerConfig
    from nerfstudio.engine.trainer import TrainerConfig
    from nerfstudio.pipelines.base_pipeline import Pipeline, VanillaPipeline, VanillaPipelineConfig
    from nerfstudio.plugins.types import MethodSpecification
    from nerfstudio.data.datamanagers.base_datamanager import VanillaDataManager, VanillaDataManagerConfig
    HAVE_NERFSTUDIO = True
except Exception:  # nerfstudio absent
    HAVE_NERFSTUDIO = False

if HAVE_NERFSTUDIO:
    import dataclasses
    import typing
    from typing import Dict, Tuple, Type

    import torch
    from torch.optim import lr_scheduler

    from .model import PointNerf, PointNerfConfig

    @dataclasses.dataclass
    class PointNerfSchedulerConfig(SchedulerConfig):
        """SU:24-31."""
        _target: Type = dataclasses.field(default_factory=lambda: PointNerfScheduler)
        lr_decay_iters: int = 1000000
        lr_decay_exp: float = 0.1

    class PointNerfScheduler(Scheduler):
        """SU:34-44."""
        config: PointNerfSchedulerConfig

        def get_scheduler(self, optimizer, lr_init):
            return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda s: lr_lambda(s, self.config.lr_decay_exp, self.config.lr_decay_iters))

    @dataclasses.dataclass
    class PointNerfDataManagerConfig(VanillaDataManagerConfig):
        """SD:36-43."""
        _target: Type = dataclasses.field(default_factory=lambda: PointNerfDataManager)
        random_image_idx: bool = True
        near_plane: float = 2.0
        far_plane: float = 6.0

    class PointNerfDataManager(VanillaDataManager):
        """SD:46-110: every batch holds the rays of ONE image, and the camera rotation travels in the bundle's metadata."""
        config: PointNerfDataManagerConfig

        def _one_image_batch(self, image_batch, count):
            n = image_batch["image_idx"].shape[0]
            image_idx = random.randint(0, n - 1) if self.config.random_image_idx else (count - 1) % n       # SD:66-69
            sel = torch.nonzero(image_batch["image_idx"] == image_idx).squeeze()
            return {"image_idx": torch.tensor(image_idx).unsqueeze(0), "image": image_batch["image"][sel].unsqueeze(0)}

        def _attach_camera(self, ray_bundle, cameras):
            c2w = cameras[ray_bundle.camera_indices.cpu()].camera_to_worlds[0][0]                           # SD:79
            ray_bundle.metadata["camrotc2w"] = c2w[0:3, 0:3].to(ray_bundle.origins.device)
            # the pose is host data here: hand it to the model so that it does not read ray 0 back from the device
            # (NearFarCollider fills nears / fars with the two planes of the model config, SM:169-171)
            ray_bundle.metadata["camera_host"] = {"origin": c2w[0:3, 3].cpu().numpy(), "camrotc2w": c2w[0:3, 0:3].cpu().numpy(),
                                                  "near": float(self.config.near_plane), "far": float(self.config.far_plane)}
            return ray_bundle

        def next_train(self, step: int) -> Tuple[typing.Any, Dict]:
            self.train_count += 1
            image_batch = self._one_image_batch(next(self.iter_train_image_dataloader), self.train_count)
            assert self.train_pixel_sampler is not None
            batch = self.train_pixel_sampler.sample(image_batch)
            ray_bundle = self.train_ray_generator(batch["indices"])
            return self._attach_camera(ray_bundle, self.train_dataset.cameras), batch

        def next_eval(self, step: int) -> Tuple[typing.Any, Dict]:
            self.eval_count += 1
            image_batch = self._one_image_batch(next(self.iter_eval_image_dataloader), self.train_count)    # SD:90: train_count, as upstream
            assert self.eval_pixel_sampler is not None
            batch = self.eval_pixel_sampler.sample(image_batch)
            ray_bundle = self.eval_ray_generator(batch["indices"])
            return self._attach_camera(ray_bundle, self.eval_dataset.cameras), batch

        def next_eval_image(self, step: int):
            for camera_ray_bundle, batch in self.eval_dataloader:
                assert camera_ray_bundle.camera_indices is not None
                image_idx = int(camera_ray_bundle.camera_indices[0, 0, 0])
                c2w = self.eval_dataset.cameras[image_idx].camera_to_worlds
                H, W = camera_ray_bundle.origins.shape[:2]
                # the reference expands the 3x3 to (800, 800, 9) (SD:108, image size hard-coded); any H x W works here
                camera_ray_bundle.metadata["camrotc2w"] = c2w[0:3, 0:3].reshape(1, 1, 9).expand(H, W, 9).to(camera_ray_bundle.origins.device)
                camera_ray_bundle.metadata["camera_host"] = {"origin": c2w[0:3, 3].cpu().numpy(), "camrotc2w": c2w[0:3, 0:3].cpu().numpy(),
                                                             "near": float(self.config.near_plane), "far": float(self.config.far_plane)}
                return image_idx, camera_ray_bundle, batch
            raise ValueError("No more eval images")

    class PointNerfPipeline(VanillaPipeline):
        """SP:16-53."""

        def __init__(self, config, device: str, test_mode="val", world_size: int = 1, local_rank: int = 0, grad_scaler=None):
            Pipeline.__init__(self)
            self.config = config
            self.test_mode = test_mode
            self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size, local_rank=local_rank)
            self.datamanager.to(device)
            assert self.datamanager.train_dataset is not None, "Missing input dataset"
            self._model = config.model.setup(scene_box=self.datamanager.train_dataset.scene_box,
                                             num_train_data=len(self.datamanager.train_dataset),
                                             cameras=self.datamanager.train_dataset.cameras, device=device)
            self.model.to(device)
            self.world_size = world_size
            if world_size > 1:
                from nerfstudio.pipelines.base_pipeline import DDP, dist
                self._model = DDP(self._model, device_ids=[local_rank], find_unused_parameters=True)     # SP:48-52
                dist.barrier(device_ids=[local_rank])

    def _make_spec(**model_overrides):
        sched = PointNerfSchedulerConfig(lr_decay_exp=0.1, lr_decay_iters=1000000)
        cfg = TrainerConfig(
            method_name=METHOD_NAME, experiment_name="pointnerf2studio",
            pipeline=VanillaPipelineConfig(
                _target=PointNerfPipeline,
                datamanager=PointNerfDataManagerConfig(_target=PointNerfDataManager, eval_num_rays_per_batch=4096,
                                                       train_num_rays_per_batch=4096),
                model=PointNerfConfig(_target=PointNerf, eval_num_rays_per_chunk=2304, **model_overrides)),
            max_num_iterations=200000, steps_per_save=25000, steps_per_eval_batch=1000, steps_per_eval_image=2000,
            steps_per_eval_all_images=100000,
            optimizers={"fields": {"optimizer": AdamOptimizerConfig(lr=0.0005), "scheduler": sched},
                        "neural_points": {"optimizer": AdamOptimizerConfig(lr=0.002), "scheduler": sched}})
        return MethodSpecification(config=cfg, description="Point-NeRF per-ray hot path on B200 (sm_100a).")

    pointnerf_original = _make_spec()
