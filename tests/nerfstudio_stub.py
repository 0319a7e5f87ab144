"""A minimal stand-in for the parts of Nerfstudio (>= 0.3, `pyproject.toml:14` of the reference) that the `pointnerf-original`
plugin touches -- TEST INFRASTRUCTURE ONLY.  Nerfstudio is neither vendored by the reference nor installable here (no network),
so the plugin module (`pointnerf2studio_b200/nerfstudio_plugin.py`) and the `Model` subclass are exercised against these stubs:
same class names, constructor contracts and call sequences as upstream (written from the upstream API, not copied), with
synthetic data behind the datamanager.  `install()` registers the modules in `sys.modules`; call it BEFORE importing
`pointnerf2studio_b200` (the package decides at import time whether it subclasses Nerfstudio's `Model`).
"""
from __future__ import annotations

import dataclasses
import sys
import types
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type

import torch
from torch import nn


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------------------------------------- configs
@dataclass
class InstantiateConfig:
    _target: Type = object

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)
    enable_collider: bool = True
    collider_params: Optional[Dict[str, float]] = field(default_factory=lambda: {"near_plane": 2.0, "far_plane": 6.0})
    loss_coefficients: Dict[str, float] = field(default_factory=lambda: {"rgb_loss_coarse": 1.0, "rgb_loss_fine": 1.0})
    eval_num_rays_per_chunk: int = 4096
    prompt: Optional[str] = None


# ---------------------------------------------------------------------------------------------- rays / cameras
@dataclass
class RayBundle:
    origins: torch.Tensor
    directions: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None
    camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None
    metadata: Dict[str, torch.Tensor] = field(default_factory=dict)
    times: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def __len__(self):
        n = 1
        for s in self.origins.shape[:-1]:
            n *= s
        return n

    def _map(self, fn, meta_fn=None):
        def g(t):
            return None if t is None else fn(t)
        md = {k: (meta_fn or fn)(v) if isinstance(v, torch.Tensor) else v for k, v in self.metadata.items()}
        return RayBundle(g(self.origins), g(self.directions), g(self.pixel_area), g(self.camera_indices), g(self.nears), g(self.fars),
                         md, g(self.times))

    def flatten(self):
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()._map(lambda t: t[start_idx:end_idx])

    def to(self, device):
        return self._map(lambda t: t.to(device))


class NearFarCollider(nn.Module):
    def __init__(self, near_plane: float, far_plane: float, **kwargs):
        super().__init__()
        self.near_plane, self.far_plane = near_plane, far_plane

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        ray_bundle.nears = ones * self.near_plane
        ray_bundle.fars = ones * self.far_plane
        return ray_bundle


class Cameras:
    """camera_to_worlds (..., 3, 4); indexing with an int / a tensor of indices like upstream's TensorDataclass."""

    def __init__(self, camera_to_worlds, fx, height, width):
        self.camera_to_worlds, self.fx, self.height, self.width = camera_to_worlds, float(fx), int(height), int(width)

    def __len__(self):
        return self.camera_to_worlds.shape[0]

    def __getitem__(self, idx):
        return Cameras(self.camera_to_worlds[idx], self.fx, self.height, self.width)

    def generate_rays(self, camera_index: int, coords: Optional[torch.Tensor] = None) -> RayBundle:
        """coords (R,2) = (y, x) pixel centres, or None for the whole image as an (H,W,.) bundle."""
        c2w = self.camera_to_worlds[camera_index]
        H, W = self.height, self.width
        if coords is None:
            ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
            coords = torch.stack([ys, xs], dim=-1).float() + 0.5
        y, x = coords[..., 0], coords[..., 1]
        d = torch.stack([(x - W / 2) / self.fx, (y - H / 2) / self.fx, torch.ones_like(x)], dim=-1)
        d = d / d.norm(dim=-1, keepdim=True)
        dirs = (d[..., None, :] * c2w[:3, :3]).sum(-1)
        origins = c2w[:3, 3].expand_as(dirs).contiguous()
        cam = torch.full(dirs.shape[:-1] + (1,), int(camera_index), dtype=torch.long)
        return RayBundle(origins=origins, directions=dirs.contiguous(), pixel_area=torch.ones_like(dirs[..., :1]), camera_indices=cam)


# ---------------------------------------------------------------------------------------------- model
class Model(nn.Module):
    config: ModelConfig

    def __init__(self, config: ModelConfig, scene_box=None, num_train_data: int = 0, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.render_aabb = None
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.populate_modules()
        self.callbacks = None
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self):
        return self.device_indicator_param.device

    def get_training_callbacks(self, training_callback_attributes) -> List:
        return []

    def populate_modules(self):
        if self.config.enable_collider:
            assert self.config.collider_params is not None
            self.collider = NearFarCollider(near_plane=self.config.collider_params["near_plane"],
                                            far_plane=self.config.collider_params["far_plane"])

    def forward(self, ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    def get_metrics_dict(self, outputs, batch) -> Dict[str, torch.Tensor]:
        return {}

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        num_rays_per_chunk = self.config.eval_num_rays_per_chunk
        image_height, image_width = camera_ray_bundle.origins.shape[:2]
        num_rays = len(camera_ray_bundle)
        outputs_lists: Dict[str, list] = {}
        for i in range(0, num_rays, num_rays_per_chunk):
            ray_bundle = camera_ray_bundle.get_row_major_sliced_ray_bundle(i, i + num_rays_per_chunk)
            outputs = self.forward(ray_bundle=ray_bundle)
            for name, out in outputs.items():
                if torch.is_tensor(out):
                    outputs_lists.setdefault(name, []).append(out)
        return {name: torch.cat(lst).view(image_height, image_width, -1) for name, lst in outputs_lists.items()}

    def load_model(self, loaded_state: Dict[str, Any]) -> None:
        self.load_state_dict({k.replace("module.", ""): v for k, v in loaded_state["model"].items()})

    def update_to_step(self, step: int) -> None:
        pass


# ---------------------------------------------------------------------------------------------- optimisers / schedulers / trainer
@dataclass
class OptimizerConfig(InstantiateConfig):
    _target: Type = torch.optim.Adam
    lr: float = 0.0005
    eps: float = 1e-08
    max_norm: Optional[float] = None

    def setup(self, params):
        kw = {k: v for k, v in vars(self).items() if k not in ("_target", "max_norm")}
        return self._target(params, **kw)


@dataclass
class AdamOptimizerConfig(OptimizerConfig):
    _target: Type = torch.optim.Adam
    weight_decay: float = 0


@dataclass
class SchedulerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Scheduler)


class Scheduler:
    config: SchedulerConfig

    def __init__(self, config: SchedulerConfig) -> None:
        self.config = config

    def get_scheduler(self, optimizer, lr_init: float):
        raise NotImplementedError


class Optimizers:
    def __init__(self, config: Dict[str, Any], param_groups: Dict[str, List[nn.Parameter]]) -> None:
        self.config, self.optimizers, self.schedulers, self.parameters = config, {}, {}, {}
        for name, params in param_groups.items():
            lr_init = config[name]["optimizer"].lr
            self.optimizers[name] = config[name]["optimizer"].setup(params=params)
            self.parameters[name] = params
            if config[name]["scheduler"]:
                self.schedulers[name] = config[name]["scheduler"].setup().get_scheduler(optimizer=self.optimizers[name], lr_init=lr_init)

    def zero_grad_all(self):
        for o in self.optimizers.values():
            o.zero_grad()

    def optimizer_step_all(self):
        for o in self.optimizers.values():
            o.step()

    def scheduler_step_all(self, step: int):
        for s in self.schedulers.values():
            s.step()


@dataclass
class TrainerConfig(InstantiateConfig):
    _target: Type = object
    method_name: Optional[str] = None
    experiment_name: Optional[str] = None
    pipeline: Any = None
    optimizers: Dict[str, Any] = field(default_factory=dict)
    max_num_iterations: int = 1000000
    steps_per_save: int = 1000
    steps_per_eval_batch: int = 500
    steps_per_eval_image: int = 500
    steps_per_eval_all_images: int = 25000
    mixed_precision: bool = False


@dataclass
class MethodSpecification:
    config: TrainerConfig
    description: str


# ---------------------------------------------------------------------------------------------- datamanager / pipeline
class _Dataset:
    def __init__(self, cameras: Cameras, images: torch.Tensor):
        self.cameras, self.images, self.scene_box = cameras, images, None

    def __len__(self):
        return len(self.cameras)


class _PixelSampler:
    def __init__(self, num_rays_per_batch: int, seed: int):
        self.n, self.g = num_rays_per_batch, torch.Generator().manual_seed(seed)

    def sample(self, image_batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        img = image_batch["image"]                          # (n_img, H, W, 3)
        n_img, H, W, _ = img.shape
        c = torch.randint(0, n_img, (self.n,), generator=self.g)
        y = torch.randint(0, H, (self.n,), generator=self.g)
        x = torch.randint(0, W, (self.n,), generator=self.g)
        idx = image_batch["image_idx"][c]
        return {"indices": torch.stack([idx, y, x], dim=-1), "image": img[c, y, x]}


class _RayGenerator:
    def __init__(self, cameras: Cameras, device):
        self.cameras, self.device = cameras, device

    def __call__(self, ray_indices: torch.Tensor) -> RayBundle:
        c = int(ray_indices[0, 0])
        assert bool((ray_indices[:, 0] == c).all())
        rb = self.cameras.generate_rays(c, ray_indices[:, 1:].float() + 0.5)
        return rb.to(self.device)


@dataclass
class VanillaDataManagerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: VanillaDataManager)
    train_num_rays_per_batch: int = 1024
    eval_num_rays_per_batch: int = 1024
    # stub-only: the synthetic scene that stands in for a dataparser
    stub_cameras: Any = None
    stub_images: Any = None


class VanillaDataManager(nn.Module):
    config: VanillaDataManagerConfig

    def __init__(self, config, device="cpu", test_mode="val", world_size: int = 1, local_rank: int = 0, **kwargs):
        super().__init__()
        self.config, self.device, self.world_size, self.local_rank, self.test_mode = config, device, world_size, local_rank, test_mode
        self.train_count = self.eval_count = 0
        cams, imgs = config.stub_cameras, config.stub_images
        self.train_dataset = self.eval_dataset = _Dataset(cams, imgs)
        batch = {"image_idx": torch.arange(len(cams)), "image": imgs}

        def forever():
            while True:
                yield batch
        self.iter_train_image_dataloader = forever()
        self.iter_eval_image_dataloader = forever()
        self.train_pixel_sampler = _PixelSampler(config.train_num_rays_per_batch, 1 + local_rank)
        self.eval_pixel_sampler = _PixelSampler(config.eval_num_rays_per_batch, 101 + local_rank)
        self.train_ray_generator = _RayGenerator(cams, device)
        self.eval_ray_generator = _RayGenerator(cams, device)
        self.eval_dataloader = [(cams.generate_rays(i).to(device), {"image": imgs[i]}) for i in range(min(len(cams), 2))]


@dataclass
class VanillaPipelineConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: VanillaPipeline)
    datamanager: Any = None
    model: Any = None


class Pipeline(nn.Module):
    @property
    def model(self):
        m = self._model
        return m.module if hasattr(m, "module") else m

    @property
    def device(self):
        return self.model.device


class VanillaPipeline(Pipeline):
    def get_train_loss_dict(self, step: int):
        ray_bundle, batch = self.datamanager.next_train(step)
        model_outputs = self._model(ray_bundle)
        metrics_dict = self.model.get_metrics_dict(model_outputs, batch)
        loss_dict = self.model.get_loss_dict(model_outputs, batch, metrics_dict)
        return model_outputs, loss_dict, metrics_dict

    def get_eval_loss_dict(self, step: int):
        self.eval()
        ray_bundle, batch = self.datamanager.next_eval(step)
        model_outputs = self.model(ray_bundle)
        metrics_dict = self.model.get_metrics_dict(model_outputs, batch)
        loss_dict = self.model.get_loss_dict(model_outputs, batch, metrics_dict)
        self.train()
        return model_outputs, loss_dict, metrics_dict

    def get_eval_image_metrics_and_images(self, step: int):
        self.eval()
        image_idx, camera_ray_bundle, batch = self.datamanager.next_eval_image(step)
        outputs = self.model.get_outputs_for_camera_ray_bundle(camera_ray_bundle)
        metrics_dict, images_dict = self.model.get_image_metrics_and_images(outputs, batch)
        metrics_dict["image_idx"] = image_idx
        metrics_dict["num_rays"] = len(camera_ray_bundle)
        self.train()
        return metrics_dict, images_dict

    def get_param_groups(self):
        return self.model.get_param_groups()


def install():
    """Register the stub package; returns the namespace of stub classes."""
    if "nerfstudio" in sys.modules and not getattr(sys.modules["nerfstudio"], "_pnerf_stub", False):
        raise RuntimeError("a real nerfstudio is importable: use it instead of the stub")
    ns = _mod("nerfstudio", _pnerf_stub=True)
    _mod("nerfstudio.cameras")
    _mod("nerfstudio.cameras.rays", RayBundle=RayBundle)
    _mod("nerfstudio.cameras.cameras", Cameras=Cameras)
    _mod("nerfstudio.configs")
    _mod("nerfstudio.configs.base_config", InstantiateConfig=InstantiateConfig)
    _mod("nerfstudio.models")
    _mod("nerfstudio.models.base_model", Model=Model, ModelConfig=ModelConfig)
    _mod("nerfstudio.model_components")
    _mod("nerfstudio.model_components.scene_colliders", NearFarCollider=NearFarCollider)
    _mod("nerfstudio.engine")
    _mod("nerfstudio.engine.optimizers", AdamOptimizerConfig=AdamOptimizerConfig, OptimizerConfig=OptimizerConfig, Optimizers=Optimizers)
    _mod("nerfstudio.engine.schedulers", SchedulerConfig=SchedulerConfig, Scheduler=Scheduler)
    _mod("nerfstudio.engine.trainer", TrainerConfig=TrainerConfig)
    _mod("nerfstudio.pipelines")
    from torch.nn.parallel import DistributedDataParallel as DDP
    import torch.distributed as dist
    _mod("nerfstudio.pipelines.base_pipeline", VanillaPipelineConfig=VanillaPipelineConfig, VanillaPipeline=VanillaPipeline,
         Pipeline=Pipeline, Model=Model, DDP=DDP, dist=dist)
    _mod("nerfstudio.plugins")
    _mod("nerfstudio.plugins.types", MethodSpecification=MethodSpecification)
    _mod("nerfstudio.data")
    _mod("nerfstudio.data.datamanagers")
    _mod("nerfstudio.data.datamanagers.base_datamanager", VanillaDataManager=VanillaDataManager, VanillaDataManagerConfig=VanillaDataManagerConfig)
    return ns


def synthetic_scene(n_cameras: int = 3, H: int = 800, W: int = 800, focal: float = 1111.1):
    """Cameras on the radius-4 sphere of pointnerf2studio_b200.synth.make_camera + random images."""
    from pointnerf2studio_b200.synth import make_camera
    c2w = []
    for i in range(n_cameras):
        cam = make_camera(H=H, W=W, focal=focal, azim_deg=30.0 + 50.0 * i, elev_deg=20.0)
        m = torch.zeros(3, 4)
        m[:, :3] = torch.from_numpy(cam.R_c2w)
        m[:, 3] = torch.from_numpy(cam.origin)
        c2w.append(m)
    images = torch.rand((n_cameras, H, W, 3), generator=torch.Generator().manual_seed(0))
    return Cameras(torch.stack(c2w), focal, H, W), images
