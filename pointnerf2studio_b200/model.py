"""Host-side mirror of the reference's plugin classes for the hot path.

Same names, argument meaning and outputs as
  pointnerf/nerfstudio/studio_model.py  (PointNerfConfig SM:61-118, PointNerf SM:122-504)
  pointnerf/nerfstudio/studio_utils.py  (NeuralPoints SU:71-209, PointNeRFEncoding SU:47-68)
but `get_outputs` runs on libpnerf_b200.so: grid built once per cloud version, sample selection +
neighbour query + field networks + compositing as CUDA kernels.  Host syncs per call: NONE on the bf16
training path (one custom op, the data-dependent counts stay on the device); the inference paths read
back the sizes of their data-dependent launches (R' after the hit-ray compaction of an image, the
per-class sample counts) and -- for a bundle without the `camera_host` hint -- ray 0's camera once.
Nerfstudio is not required: without it `PointNerf` is a plain nn.Module that accepts any object with
`origins (R,3)`, `directions (R,3)`, `nears/fars (R,1)`, `metadata["camrotc2w"]`; with it, it is a
nerfstudio `Model`.  The `ns-train pointnerf-original` registration lives in nerfstudio_plugin.py.
"""
from __future__ import annotations

import dataclasses
import glob
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch
from torch import nn
from torch.nn import Parameter

from . import native, ops  # noqa: F401  (ops registers the pnerf:: custom ops)

try:      # Nerfstudio (>= 0.3, the reference's only framework dependency) is optional: subclass its Model / ModelConfig when it is there
    from nerfstudio.models.base_model import Model as _ModelBase, ModelConfig as _ConfigBase
    HAVE_NERFSTUDIO = True
except Exception:
    HAVE_NERFSTUDIO = False
    _ModelBase = nn.Module

    @dataclass
    class _ConfigBase:
        pass


HIT_COMPACTION_MIN_RAYS = 32768       # bundles up to this size skip the hit-ray compaction (R -> R')
RENDER_RAYS_PER_LAUNCH = 1 << 21   # rays per query launch in full-image rendering (memory knob only)


@dataclass
class RayBundle:
    """Minimal stand-in for nerfstudio.cameras.rays.RayBundle (the fields SU:148-155 reads)."""
    origins: torch.Tensor
    directions: torch.Tensor
    nears: torch.Tensor
    fars: torch.Tensor
    metadata: Dict[str, torch.Tensor]

    def __len__(self):
        return self.origins.shape[0]

    @classmethod
    def for_camera(cls, directions: torch.Tensor, origin, camrotc2w, near: float, far: float) -> "RayBundle":
        """The rays of ONE camera -- all this path supports (origin, rotation, near and far are read from ray 0, SU:148-155).
        `directions` (R,3) is on the device; origin (3,), camrotc2w (3,3), near, far are host values.  origins / nears / fars
        are zero-copy (R,.) views of one 14-float upload instead of R copies, and the host values travel along as the
        `camera_host` hint, so nothing is read back.  A server uploads 12 B per ray instead of 32."""
        o = np.asarray(origin, dtype=np.float32).reshape(3)
        r = np.asarray(camrotc2w, dtype=np.float32).reshape(3, 3)
        R, dev = directions.shape[0], directions.device
        cam = torch.from_numpy(np.concatenate([o, r.reshape(-1), np.asarray([near, far], dtype=np.float32)])).to(dev, non_blocking=True)
        return cls(origins=cam[0:3].view(1, 3).expand(R, 3), directions=directions, nears=cam[12:13].view(1, 1).expand(R, 1),
                   fars=cam[13:14].view(1, 1).expand(R, 1),
                   metadata={"camrotc2w": cam[3:12].view(3, 3),
                             "camera_host": {"origin": o, "camrotc2w": r, "near": float(near), "far": float(far)}})


class NearFarCollider(nn.Module):
    """What nerfstudio's NearFarCollider does for the reference (Model.populate_modules, called at SM:171): every ray gets the
    two planes of `collider_params`.  Used only when Nerfstudio is absent; a bundle that already carries nears / fars keeps them."""

    def __init__(self, near_plane: float, far_plane: float):
        super().__init__()
        self.near_plane, self.far_plane = float(near_plane), float(far_plane)

    def forward(self, ray_bundle):
        if getattr(ray_bundle, "nears", None) is None or getattr(ray_bundle, "fars", None) is None:
            ones = torch.ones_like(ray_bundle.origins[..., 0:1])
            ray_bundle.nears, ray_bundle.fars = ones * self.near_plane, ones * self.far_plane
        return ray_bundle


@dataclass
class PointNerfConfig(_ConfigBase):
    """Every field of the reference's PointNerfConfig (SM:61-118), same names and defaults (a nerfstudio ModelConfig when
    Nerfstudio is importable)."""
    _target: Any = dataclasses.field(default_factory=lambda: PointNerf)
    enable_collider: bool = True                  # ModelConfig defaults the reference inherits: NearFarCollider(2.0, 6.0)
    collider_params: Optional[Dict[str, float]] = dataclasses.field(default_factory=lambda: {"near_plane": 2.0, "far_plane": 6.0})
    path_point_cloud: Optional[Path] = None
    eval_num_rays_per_chunk: int = 4096
    feat_grad: bool = True
    conf_grad: bool = True
    dir_grad: bool = True
    color_grad: bool = True
    num_pos_freqs: Optional[int] = 10
    num_viewdir_freqs: Optional[int] = 4
    num_feat_freqs: Optional[int] = 3
    num_dist_freqs: Optional[int] = 5
    agg_dist_pers: Optional[int] = 20
    point_features_dim: Optional[int] = 32
    point_color_mode: Optional[bool] = True
    point_dir_mode: Optional[bool] = True
    num_samples: int = 80
    use_biased_sampler: bool = False
    field_dim: int = 64
    num_mlp_base_layers: Optional[int] = 2
    num_mlp_head_layers: Optional[int] = 2
    num_color_layers: Optional[int] = 3
    num_alpha_layers: Optional[int] = 1
    hidden_size: int = 256
    hidden_size_color: int = 128
    apply_pnt_mask: bool = True
    act_super: bool = False
    axis_weight: List[float] = dataclasses.field(default_factory=lambda: [1., 1., 1.])
    kernel_size: List[int] = dataclasses.field(default_factory=lambda: [3, 3, 3])
    vscale: List[float] = dataclasses.field(default_factory=lambda: [2, 2, 2])
    vsize: List[float] = dataclasses.field(default_factory=lambda: [0.004, 0.004, 0.004])
    query_size: List[float] = dataclasses.field(default_factory=lambda: [3, 3, 3])
    ranges: List[float] = dataclasses.field(default_factory=lambda: [-1.200, -1.200, -1.200, 1.200, 1.200, 1.200])
    z_depth_dim: int = 400
    SR: int = 80
    K: int = 8
    max_o: int = 1000000
    P: int = 12
    NN: int = 2
    gpu_maxthr: int = 1024
    zero_epsilon: float = 1e-3
    zero_one_loss_weights: float = 0.0001
    loss_coefficients: Dict[str, float] = dataclasses.field(default_factory=dict)
    # ---- additions of this implementation (not in the reference)
    precision: str = "bf16"        # "fp32": SIMT exact-parity kernels; "bf16": tcgen05 tensor-core kernels
    flow: str = "plugin"           # "plugin" (SM math) or "original" (PointAggregator/ray_march math)
    jitter: float = 0.3            # SU:166 hard-codes 0.3 in train and eval
    jitter_seed: int = 0           # Philox key (high word) of the in-kernel jitter; the low word counts calls

    def __post_init__(self):
        if self.path_point_cloud is not None and not Path(self.path_point_cloud).exists():
            raise RuntimeError(f"PointCloud path {self.path_point_cloud} does not exist")
        unsupported = (self.num_viewdir_freqs != 4 or self.num_feat_freqs != 3 or self.num_dist_freqs != 5
                       or self.agg_dist_pers != 20 or self.point_features_dim != 32 or self.hidden_size != 256
                       or self.hidden_size_color != 128 or self.num_mlp_base_layers != 2 or self.num_mlp_head_layers != 2
                       or self.num_color_layers != 3 or not self.point_color_mode or not self.point_dir_mode
                       or not self.apply_pnt_mask or list(self.axis_weight) != [1., 1., 1.])
        if unsupported:
            raise NotImplementedError("the CUDA kernels are specialised to the reference's shipped network shape "
                                      "(SM:72-97 defaults); other shapes are out of scope (SURVEY.md section 8)")

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


def get_latest_epoch(resume_dir):
    """SM:55-59."""
    os.makedirs(resume_dir, exist_ok=True)
    str_epoch = [f.split("_")[0] for f in os.listdir(resume_dir) if f.endswith("_states.pth")]
    int_epoch = [int(i) for i in str_epoch]
    return None if len(int_epoch) == 0 else str_epoch[int_epoch.index(max(int_epoch))]


class PointNeRFEncoding(nn.Module):
    """SU:47-68 (== positional_encoding NW:176-191); kept for API parity, the kernels encode in registers."""

    def __init__(self, in_dim: int, num_frequencies: int, ori: bool = False):
        super().__init__()
        self.in_dim, self.num_frequencies, self.ori = in_dim, num_frequencies, ori

    def forward(self, in_tensor, covs=None):
        freq = (2 ** torch.arange(self.num_frequencies).float()).to(in_tensor.device)
        pts = (in_tensor[..., None] * freq).reshape(in_tensor.shape[:-1] + (self.num_frequencies * in_tensor.shape[-1],))
        if self.ori:
            return torch.cat([in_tensor, torch.sin(pts), torch.cos(pts)], dim=-1)
        return torch.stack([torch.sin(pts), torch.cos(pts)], dim=-1).reshape(pts.shape[:-1] + (pts.shape[-1] * 2,))


class MLP(nn.Module):
    """Parameter container with nerfstudio's MLP naming (`layers.{i}.weight/bias`)."""

    def __init__(self, in_dim, num_layers, layer_width):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, layer_width
        self.layers = nn.ModuleList([nn.Linear(in_dim if i == 0 else layer_width, layer_width) for i in range(num_layers)])

    def get_out_dim(self):
        return self.out_dim


class FieldHead(nn.Module):
    """Parameter container with nerfstudio's FieldHead naming (`net.weight/bias`)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.net = nn.Linear(in_dim, out_dim)


class NeuralPoints(nn.Module):
    """SU:71-209: owns the neural-point parameters and runs the querier.

    Parameter names and shapes follow the reference (`points_embeding` keeps its spelling):
    points_xyz (N,3) no grad, points_embeding (1,N,32), points_conf (1,N,1), points_dir (1,N,3),
    points_color (1,N,3), points_Rw2c (3,3) no grad.
    """

    def __init__(self, state_dict, device, config: PointNerfConfig):
        super().__init__()
        self.config = config
        self.device = torch.device(device)
        self.points_xyz = Parameter(state_dict["neural_points.xyz"].to(self.device).float().contiguous(), requires_grad=False)
        self.points_embeding = Parameter(state_dict["neural_points.points_embeding"].to(self.device).float().contiguous(),
                                         requires_grad=config.feat_grad)
        self.points_conf = Parameter(state_dict["neural_points.points_conf"].to(self.device).float().contiguous(),
                                     requires_grad=config.conf_grad)
        self.points_dir = Parameter(state_dict["neural_points.points_dir"].to(self.device).float().contiguous(),
                                    requires_grad=config.dir_grad)
        self.points_color = Parameter(state_dict["neural_points.points_color"].to(self.device).float().contiguous(),
                                      requires_grad=config.color_grad)
        self.points_Rw2c = Parameter(state_dict["neural_points.Rw2c"].to(self.device).float().contiguous(), requires_grad=False)
        if self.points_Rw2c.dim() != 2:
            raise NotImplementedError("per-point Rw2c (SU:207 second branch) is not produced by any shipped script")
        self.kernel_size = np.asarray(config.kernel_size, dtype=np.int32)
        self.query_size = np.asarray(config.query_size, dtype=np.int32)
        self.radius_limit_np = np.asarray(4 * max(config.vsize[0], config.vsize[1])).astype(np.float32)   # SU:110
        self.vscale_np = np.array(config.vscale, dtype=np.int32)
        self.scaled_vsize_np = (config.vsize * self.vscale_np).astype(np.float32)                          # SU:112
        self._grid = None
        self._grid_key = None
        self._jitter_calls = 0
        self._last_jitter = None

    # ---- grid cache: rebuilt only when the cloud changes (the reference rebuilds on every call, CU:314-365)
    def grid(self) -> native.VoxelGrid:
        key = (self.points_xyz.data_ptr(), self.points_xyz._version, tuple(self.points_xyz.shape))
        if self._grid is None or self._grid_key != key:
            frame = native.get_hyperparameters(self.points_xyz.detach(), self.config.vsize, self.config.vscale,
                                               self.config.kernel_size, self.config.ranges)
            self._grid = native.VoxelGrid(self.points_xyz.detach(), frame, self.config.P, self.config.query_size)
            self._grid_key = key
        return self._grid

    # ---- point-cloud surgery of the original flow (models/neural_points/neural_points.py:341-393); both replace the parameter
    # tensors, so the cached voxel grid is rebuilt on the next query and optimiser state must be re-created by the caller
    # (run/train_studio.py:677-683 does clean_optimizer / setup_optimizer around them)
    def _replace(self, xyz, embed, conf, dirn, color):
        c = self.config
        self.points_xyz = Parameter(xyz.contiguous(), requires_grad=False)
        self.points_embeding = Parameter(embed.contiguous(), requires_grad=c.feat_grad)
        self.points_conf = Parameter(conf.contiguous(), requires_grad=c.conf_grad)
        self.points_dir = Parameter(dirn.contiguous(), requires_grad=c.dir_grad)
        self.points_color = Parameter(color.contiguous(), requires_grad=c.color_grad)
        self._grid, self._grid_key = None, None

    @torch.no_grad()
    def prune(self, thresh: float) -> int:
        """NP:341-362: keep the points whose confidence is >= thresh.  Returns the number of pruned points."""
        mask = self.points_conf[0, :, 0] >= thresh
        n_pruned = int((~mask).sum())
        self._replace(self.points_xyz[mask], self.points_embeding[:, mask], self.points_conf[:, mask], self.points_dir[:, mask],
                      self.points_color[:, mask])
        return n_pruned

    @torch.no_grad()
    def grow_points(self, add_xyz, add_embedding, add_color, add_dir, add_conf) -> int:
        """NP:367-393: append points (N_add,3), (N_add,32), (N_add,3), (N_add,3), (N_add,1)."""
        dev = self.points_xyz.device
        cat = lambda a, b: torch.cat([a, b.to(dev).float()[None]], dim=1)
        self._replace(torch.cat([self.points_xyz, add_xyz.to(dev).float()], dim=0), cat(self.points_embeding, add_embedding),
                      cat(self.points_conf, add_conf), cat(self.points_dir, add_dir), cat(self.points_color, add_color))
        return int(add_xyz.shape[0])

    def get_hyperparameters(self, vsize_np, point_xyz_w_tensor, ranges=None):
        """SU:115-127, same return triple."""
        f = native.get_hyperparameters(point_xyz_w_tensor, vsize_np, self.config.vscale, self.config.kernel_size, ranges)
        ranges_tensor = torch.as_tensor(np.concatenate([f.lo, f.hi]), device=point_xyz_w_tensor.device)
        return ranges_tensor, vsize_np, f.dim

    @staticmethod
    def camera_of(ray_bundle, with_near_far=False):
        """SU:148-155: one camera per call; rotation and origin come from ray 0 (and near / far, SU:154-155).
        A bundle that carries the `camera_host` hint (RayBundle.for_camera, PointNerfDataManager, a server that just uploaded the
        rays) costs nothing; otherwise ray 0 is read back with ONE 14-float device->host copy and the result is remembered ON
        THE BUNDLE (its metadata dict), never keyed on device pointers: the caching allocator hands freed blocks back at the
        same address, so a pointer / version key can alias two different cameras."""
        hint = ray_bundle.metadata.get("camera_host")
        if hint is None or (with_near_far and not ("near" in hint and "far" in hint)):
            rot = ray_bundle.metadata["camrotc2w"]
            rot = rot[0].view(3, 3) if rot.shape[0] != 3 else rot
            parts = [ray_bundle.origins[0].detach().float().reshape(-1), rot.detach().float().reshape(-1),
                     ray_bundle.nears[0].detach().float().reshape(-1)[:1], ray_bundle.fars[0].detach().float().reshape(-1)[:1]]
            h = torch.cat(parts).cpu().numpy()
            hint = {"origin": h[:3].copy(), "camrotc2w": h[3:12].reshape(3, 3).copy(), "near": float(h[12]), "far": float(h[13])}
            ray_bundle.metadata["camera_host"] = hint
        o = np.asarray(hint["origin"], dtype=np.float32).reshape(3).copy()
        r = np.asarray(hint["camrotc2w"], dtype=np.float32).reshape(3, 3).copy()
        return (o, r, float(hint["near"]), float(hint["far"])) if with_near_far else (o, r)

    def coarse_t(self, R, near, far, jitter, generator=None):
        """t mid-points of near_far_linear_ray_generation (RM:312-329): (D,) when jitter == 0 else (R,D)."""
        D = self.config.z_depth_dim
        dev = self.device
        tau = torch.linspace(0, 1, D + 1, device=dev).view(1, -1)
        edge = near * (1 - tau) + far * tau
        seg = edge[..., 1:] - edge[..., :-1]
        if jitter:
            seg = seg * (1 + jitter * (torch.rand((1, R, D), device=dev, generator=generator)[0] - 0.5))
        else:
            seg = seg * (1 + jitter * (torch.full((1, D), 0.5, device=dev) - 0.5))
        end = torch.cumsum(seg, dim=1)
        end = near + torch.cat([torch.zeros((end.shape[0], 1), device=dev), end], dim=1)
        t_mid = (end[:, :-1] + end[:, 1:]) / 2
        return t_mid[0].contiguous() if not jitter else t_mid.contiguous()

    def query(self, ray_bundle, jitter=None, generator=None, want_stats=False, near=None, far=None, compact=False):
        """Rows G0/G2/Q on the cached grid.  Returns (QueryResult, origin, R_c2w, dirs); with compact=True the result and
        `dirs` cover only the rays whose selection found an occupied position (QueryResult.ray_index)."""
        cfg = self.config
        if near is None:
            origin, R_c2w, near, far = self.camera_of(ray_bundle, with_near_far=True)
        else:
            origin, R_c2w = self.camera_of(ray_bundle)
        dirs = ray_bundle.directions.to(self.device).float().contiguous()
        R = dirs.shape[0]
        jitter = cfg.jitter if jitter is None else jitter
        if jitter and generator is None:
            # jittered t generated inside the selection kernel (Philox keyed by a per-call seed): no (R,D) tensor at all
            self._jitter_calls += 1
            seed = (int(cfg.jitter_seed) << 32) | (self._jitter_calls & 0xffffffff)
            self._last_jitter = (near, far, float(jitter), seed)
            q = native.sample_and_query(self.grid(), R, cfg.z_depth_dim, cfg.SR, cfg.K, int(self.kernel_size[0]),
                                        float(self.radius_limit_np), origin=origin, dirs=dirs, want_stats=want_stats,
                                        jitter_gen=self._last_jitter, compact=compact)
            return q, origin, R_c2w, (q.dirs if compact else dirs)
        t = self.coarse_t(R, near, far, jitter, generator)     # torch path: jitter 0 (shared table) or an explicit generator
        q = native.sample_and_query(self.grid(), R, cfg.z_depth_dim, cfg.SR, cfg.K, int(self.kernel_size[0]),
                                    float(self.radius_limit_np), origin=origin, dirs=dirs, t_vals=t, want_stats=want_stats,
                                    compact=compact)
        return q, origin, R_c2w, (q.dirs if compact else dirs)

    def forward(self, ray_bundle):
        """Reference-shaped return value (SU:209): the 13 gathered tensors.  Compatibility API only --
        `PointNerf.get_outputs` never materialises these (the kernels gather in registers)."""
        cfg = self.config
        q, origin, R_c2w, dirs = self.query(ray_bundle)
        pidx, loc_w, ray_mask, ray_index, _ = native.compact_rays(q)
        pidx, loc_w, ray_mask = pidx[None], loc_w[None], ray_mask[None]
        o = torch.as_tensor(origin, device=self.device)[None]
        Rc = torch.as_tensor(R_c2w, device=self.device)[None]

        def pers(p):
            cam = torch.sum((p - o[:, None, :] if p.dim() == 3 else p - o[:, None, None, :])[..., :, None] * Rc[0], dim=-2)
            return torch.stack([cam[..., 0] / cam[..., 2], cam[..., 1] / cam[..., 2], cam[..., 2]], dim=-1)

        mask = pidx >= 0
        B, R2, SR, K = pidx.shape
        flat = pidx.clamp(min=0).view(-1).long()
        cat = torch.cat([self.points_xyz[None], pers(self.points_xyz[None]), self.points_embeding], dim=-1)
        emb = torch.index_select(cat, 1, flat).view(B, R2, SR, K, -1)
        col = torch.index_select(self.points_color, 1, flat).view(B, R2, SR, K, 3)
        dr = torch.index_select(self.points_dir, 1, flat).view(B, R2, SR, K, 3)
        cf = torch.index_select(self.points_conf, 1, flat).view(B, R2, SR, K, 1)
        ray_dirs = dirs[ray_index.long()][None, :, None, :].expand(-1, -1, SR, -1).contiguous()
        return (col, self.points_Rw2c, dr, emb[..., 6:], emb[..., 3:6], emb[..., :3], cf, pers(loc_w), loc_w, mask, ray_dirs,
                cfg.vsize, ray_mask)


class ConfCoefficient:
    """What `outputs["conf_coefficient"]` carries in training (SM:396-397): the straight-through-clamped confidences gathered at
    all (1,R'',SR,K) slots.  The loss kernel reads the indices directly, so the gathered tensor is only built when somebody asks
    for it: the object behaves as that tensor in any torch function (`torch.clamp(cc, ...)`, `torch.log(cc)`, `cc.shape`, ...) and
    materialises it on first use (one boolean-mask gather, i.e. one host sync, exactly what SU:194-203 costs the reference)."""

    def __init__(self, conf, pidx, ray_mask, n_rays):
        self.conf, self.pidx, self.ray_mask, self.n_rays = conf, pidx, ray_mask, n_rays
        self._t = None

    def materialize(self):
        if self._t is None:
            keep = self.ray_mask.bool()
            c = self.conf.reshape(-1)[self.pidx[keep].clamp(min=0).long()]
            self._t = (c - (c - c.clamp(1e-4, 1)).detach())[None]
        return self._t

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.materialize() if isinstance(a, ConfCoefficient) else a
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    def __getattr__(self, name):          # .shape, .clamp(...), .detach(), ... of the gathered tensor
        if name.startswith("__") or name in ("conf", "pidx", "ray_mask", "n_rays", "_t"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class PointNerf(_ModelBase):
    """SM:122-504 -- a nerfstudio `Model` when Nerfstudio is importable, otherwise an nn.Module with the same method set.
    `get_outputs(ray_bundle)` -> {"coarse_raycolor" (R,3), "ray_mask" (R,) int8, ["conf_coefficient"]};
    `get_param_groups()` -> {"fields", "neural_points"}; `get_loss_dict`; `get_outputs_for_camera_ray_bundle`;
    `get_image_metrics_and_images`; `get_metrics_dict`; `get_training_callbacks`."""
    config: PointNerfConfig

    def __init__(self, config: PointNerfConfig, scene_box=None, num_train_data: int = 0, cameras=None, state_dict=None,
                 device="cuda", **kwargs):
        if HAVE_NERFSTUDIO:
            super().__init__(config=config, scene_box=scene_box, num_train_data=num_train_data, **kwargs)   # calls populate_modules()
        else:
            nn.Module.__init__(self)
            self.config, self.scene_box, self.num_train_data, self.kwargs = config, scene_box, num_train_data, kwargs
            self.render_aabb, self.collider, self.callbacks = None, None, None
            self.populate_modules()
        self._point_initialized = False
        self.cameras = cameras
        self._device = torch.device(device)
        self._init_pointnerf(state_dict)
        self.to(self._device)

    @property
    def device(self):
        return self._device

    def _init_pointnerf(self, state_dict=None):
        """SM:147-166: latest `<iter>_net_ray_marching.pth` under path_point_cloud."""
        if state_dict is None:
            if self.config.path_point_cloud is None:
                raise RuntimeError("The point_cloud_path must be specified.")
            from .checkpoint import load_point_cloud_checkpoint
            state_dict = load_point_cloud_checkpoint(self.config.path_point_cloud)
        self._loaded_state = state_dict
        self.neural_points = NeuralPoints(state_dict, self._device, self.config)
        from .checkpoint import AGGREGATOR_MAP
        if any(k.startswith("aggregator.") for k in state_dict):     # original-flow checkpoint: reuse its MLPs
            own = dict(self.named_parameters())
            with torch.no_grad():
                for new, old in AGGREGATOR_MAP.items():
                    for s in ("weight", "bias"):
                        own[f"{new}.{s}"].copy_(state_dict[f"aggregator.{old}.{s}"])
        self._point_initialized = True

    def populate_modules(self):
        """SM:169-237: collider (through the base class, SM:171), encodings, networks, background.  The image-quality metric
        modules (PSNR / SSIM / LPIPS objects, SM:229-236) are evaluation tooling outside the hot path: get_image_metrics_and_images
        computes PSNR / RMSE directly and adds SSIM / LPIPS when torchmetrics is importable."""
        c = self.config
        if HAVE_NERFSTUDIO:
            super().populate_modules()
        elif c.enable_collider and c.collider_params is not None:
            self.collider = NearFarCollider(c.collider_params["near_plane"], c.collider_params["far_plane"])
        self.direction_encoding = PointNeRFEncoding(2, c.num_viewdir_freqs, ori=True)
        self.feature_encoding = PointNeRFEncoding(2, c.num_feat_freqs, ori=False)
        self.dists_encoding = PointNeRFEncoding(2, c.num_dist_freqs, ori=False)
        dist_dim = (4 if c.agg_dist_pers == 30 else 6) if c.agg_dist_pers > 9 else 3
        dist_xyz_dim = dist_dim if c.num_dist_freqs == 0 else 2 * abs(c.num_dist_freqs) * dist_dim
        mlp_in = 2 * c.num_feat_freqs * c.point_features_dim + dist_xyz_dim + c.point_features_dim
        self.mlp_base = MLP(mlp_in, c.num_mlp_base_layers, c.hidden_size)
        self.mlp_head = MLP(self.mlp_base.get_out_dim() + 3 + 4, c.num_mlp_head_layers, c.hidden_size)
        self.mlp_color = MLP(self.mlp_head.get_out_dim() + 2 * c.num_viewdir_freqs * 3, c.num_color_layers, c.hidden_size_color)
        self.field_output_color = FieldHead(self.mlp_color.get_out_dim(), 3)
        self.field_output_density = FieldHead(self.mlp_head.get_out_dim(), 1)
        self._background_color = torch.ones(3)

    def get_background_color(self):
        """SM:257-261."""
        return self._background_color

    def get_training_callbacks(self, training_callback_attributes=None):
        return []

    def get_metrics_dict(self, outputs, batch):
        return {}            # the reference does not override Model.get_metrics_dict

    def update_to_step(self, step: int) -> None:
        pass

    def mlp_param_list(self):
        """The 14 MLP tensors in the kernels' order; the Parameter objects live as long as the module, so the walk over
        named_parameters() (0.1 ms of a 3 ms training step) is done once."""
        cached = self.__dict__.get("_mlp_params")
        if cached is None:
            own = dict(self.named_parameters())
            cached = []
            for name, _, _ in native.MLP_PARAM_NAMES:
                cached += [own[name + ".weight"], own[name + ".bias"]]
            self.__dict__["_mlp_params"] = cached
        return cached

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        """SM:401-413."""
        named = list(self.named_parameters())
        return {"neural_points": [p for n, p in named if n.startswith("neural_points.points")],
                "fields": [p for n, p in named if not n.startswith("neural_points.points")]}

    def forward(self, ray_bundle):
        """nerfstudio Model.forward: the collider fills nears / fars, then get_outputs."""
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    def get_outputs(self, ray_bundle, generator=None):
        """SM:263-399."""
        native.pin_stream()
        try:
            return self._get_outputs(ray_bundle, generator)
        finally:
            native.unpin_stream()

    def set_step_consts(self, step_consts=None):
        """A device tensor of 16 fp32 words (pnerf_camera.dev layout) the training kernels read the camera / near / far / jitter
        seed from at run time, instead of the by-value launch parameters: what makes a captured CUDA graph of the training step
        replayable for every batch (parallel.TrainEngine writes it before each replay).  None switches it off."""
        self._step_consts = step_consts

    def set_grad_sink(self, sink_points=None, sink_mlp=None, points_done_event=0):
        """A data-parallel trainer (parallel.TrainEngine) hands the backward kernels the flat buffers the gradients are to be
        accumulated in (`.grad` of the parameters aliases them) and, optionally, a CUDA event to record as soon as the point
        gradients are complete.  None / 0 restores plain autograd behaviour (fresh gradient tensors per call)."""
        self._grad_sink = (sink_points, sink_mlp, int(points_done_event))

    def _get_outputs_train_fused(self, ray_bundle, generator=None):
        """Training step forward on the tensor-core path as ONE custom op (`pnerf::render_train`): selection, query, compaction,
        field networks, compositing and ray mask in one host call, nothing read back from the device."""
        from . import native_tc, ops
        c, npnts = self.config, self.neural_points
        origin, R_c2w, near, far = npnts.camera_of(ray_bundle, with_near_far=True)
        dirs = ray_bundle.directions.to(self._device).float().contiguous()
        R = dirs.shape[0]
        grid = npnts.grid()
        t_vals, t_stride, seed, jitter = None, 0, 0, float(c.jitter)
        if jitter and generator is None:
            npnts._jitter_calls += 1
            seed = (int(c.jitter_seed) << 32) | (npnts._jitter_calls & 0xffffffff)
            npnts._last_jitter = (near, far, jitter, seed)
        else:                          # an explicit generator or jitter 0: the torch t table (SU:166 semantics, replayable)
            t_vals = npnts.coarse_t(R, near, far, jitter, generator)
            t_stride = 0 if t_vals.dim() == 1 else c.z_depth_dim
        mode = native.make_mode(c.flow, training=self.training, bg=self._background_color.tolist(), vsize_z=c.vsize[2])
        params = self.mlp_param_list()
        wpack = native_tc.packed_weights(params, force=self.__dict__.get("_force_repack", False))[0]
        sink_points, sink_mlp, event = self.__dict__.get("_grad_sink", (None, None, 0))
        step_consts = self.__dict__.get("_step_consts")
        if step_consts is not None and t_vals is not None:
            raise RuntimeError("step constants on the device need the in-kernel jitter (jitter > 0, no generator)")
        fl, it = ops.fl_it(grid.frame, origin, R_c2w, native._rw2c_host(npnts.points_Rw2c), near, far, jitter, float(npnts.radius_limit_np),
                           mode, c.z_depth_dim, c.SR, c.K, int(npnts.kernel_size[0]), seed=seed, t_stride=t_stride, event=event)
        (rgb, ray_mask, n_rays, pidx, loc, valid, cnt, sigma, srgb, ids, n_samples, _ws, _ridx) = torch.ops.pnerf.render_train(
            dirs, npnts.points_xyz, npnts.points_embeding.view(-1, c.point_features_dim), npnts.points_color.view(-1, 3),
            npnts.points_dir.view(-1, 3), npnts.points_conf.view(-1, 1), [p for p in params], wpack, grid.cell_start, grid.recs,
            grid.occ_bits, t_vals, sink_points, sink_mlp, step_consts, fl, it)
        out = {"coarse_raycolor": rgb, "ray_mask": ray_mask,
               "conf_coefficient": ConfCoefficient(npnts.points_conf, pidx, ray_mask, n_rays)}
        self._last_query = native.QueryResult(loc, cnt, pidx, valid)
        self._last_render = {"sigma": sigma, "rgb": srgb, "n_samples": n_samples}
        return out

    def _get_outputs(self, ray_bundle, generator=None):
        c = self.config
        npnts = self.neural_points
        if (c.precision == "bf16" and self.training and torch.is_grad_enabled() and 0 < len(ray_bundle) <= HIT_COMPACTION_MIN_RAYS
                and any(p.requires_grad for p in self.parameters())):
            return self._get_outputs_train_fused(ray_bundle, generator)
        # dropping the rays without an occupied position (R -> R') costs a host sync and six small launches: worth it for an image,
        # not for a 4096-ray training batch, where those rays are nearly free in every kernel anyway
        q, origin, R_c2w, dirs = npnts.query(ray_bundle, generator=generator, compact=len(ray_bundle) > HIT_COMPACTION_MIN_RAYS)
        mode = native.make_mode(c.flow, training=self.training, bg=self._background_color.tolist(), vsize_z=c.vsize[2])
        cfg = {"mode": mode, "camera": native.make_camera(origin, R_c2w)}
        args = (cfg, q, dirs, npnts.points_xyz, npnts.points_Rw2c, npnts.points_embeding.view(-1, c.point_features_dim),
                npnts.points_color.view(-1, 3), npnts.points_dir.view(-1, 3), npnts.points_conf.view(-1, 1))
        if c.precision == "fp32":
            rgb = native.render_f32(*args, self.mlp_param_list())
        elif c.precision == "bf16":
            from . import native_tc
            rgb = native_tc.render_tc(*args, self.mlp_param_list())
        else:
            raise ValueError(c.precision)
        # rays dropped by the hit-ray compaction keep the background colour (fill_invalid, SM:491-504)
        R_total = q.R_total
        if q.ray_index is not None:
            src = self._background_color
            key = (id(src), src._version, rgb.device, rgb.dtype)
            hit = self.__dict__.get("_bg_dev")
            if hit is None or hit[0] != key:             # one tiny host->device copy per background, not per call
                hit = self.__dict__["_bg_dev"] = (key, src.to(device=self._device, dtype=rgb.dtype))
            rgb = hit[1].view(1, 3).expand(R_total, 3).index_copy(0, q.ray_index.long(), rgb)
        lib = native._lib.load()
        R, SR = q.sample_valid.shape
        ray_mask = torch.empty((R,), dtype=torch.int8, device=self._device)
        ray_index = torch.empty((max(R, 1),), dtype=torch.int32, device=self._device)
        n_rays = torch.empty((1,), dtype=torch.int32, device=self._device)
        ws_bytes = lib.pnerf_scan_workspace_bytes(R)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self._device)
        native.check(lib.pnerf_ray_compact(native._ptr(q.sample_valid), R, SR, native._ptr(ray_mask), native._ptr(ray_index),
                                           native._ptr(n_rays), native._ptr(ws), ws_bytes, native._stream()), "pnerf_ray_compact")
        native.LAUNCHES["n"] += 3
        if q.ray_index is not None:
            full_mask = torch.zeros((R_total,), dtype=torch.int8, device=self._device).index_copy_(0, q.ray_index.long(), ray_mask)
        else:
            full_mask = ray_mask
        out = {"coarse_raycolor": rgb, "ray_mask": full_mask}
        if self.training:
            out["conf_coefficient"] = ConfCoefficient(npnts.points_conf, q.sample_pidx, ray_mask, n_rays)
        self._last_query = q
        self._last_render = cfg.get("last")
        return out

    def last_query_dense(self):
        """The last call's query result over ALL its rays (tests / debugging): compacted rays scattered back."""
        return self._last_query.dense()

    def last_render_dense(self):
        """Per-slot sigma (R,SR) / rgb (R,SR,3) of the last call over ALL its rays, zero for rays without samples."""
        q, last = self._last_query, self._last_render
        if q.ray_index is None:
            return last
        idx = q.ray_index.long()
        SR = q.sample_valid.shape[1]
        sig = torch.zeros((q.R_total, SR), dtype=torch.float32, device=idx.device).index_copy_(0, idx, last["sigma"])
        rgb = torch.zeros((q.R_total, SR, 3), dtype=torch.float32, device=idx.device).index_copy_(0, idx, last["rgb"])
        return {"sigma": sig, "rgb": rgb, "n_samples": last["n_samples"]}

    @torch.no_grad()
    def probe(self, ray_bundle):
        """Hole probing of the original flow (models/neural_points_volumetric_model.py:331-362, consumed by probe_hole,
        run/train_studio.py:335-444): renders the rays and returns, besides `coarse_raycolor` / `ray_mask`, per ray the largest
        sample opacity, that sample's position, its distance to the nearest neighbour and the (weight x confidence)-averaged
        attributes of its neighbours -- all (R, .) tensors, zero for rays without neighbours."""
        import ctypes as C
        was_training = self.training
        self.eval()
        try:
            out = self.get_outputs(ray_bundle)
        finally:
            self.train(was_training)
        q, last, npnts, c = self._last_query, self._last_render, self.neural_points, self.config
        R2, SR, K = q.sample_pidx.shape
        dev = self._device
        lib = native._lib.load()
        origin, R_c2w = npnts.camera_of(ray_bundle)
        mode = native.make_mode(c.flow, training=False, bg=self._background_color.tolist(), vsize_z=c.vsize[2])
        cam = native.make_camera(origin, R_c2w)
        pts = native.make_points(npnts.points_xyz.detach(), npnts.points_embeding.detach().view(-1, c.point_features_dim),
                                 npnts.points_color.detach().view(-1, 3), npnts.points_dir.detach().view(-1, 3),
                                 npnts.points_conf.detach().view(-1, 1), npnts.points_Rw2c)
        shapes = {"ray_max_shading_opacity": 1, "ray_max_sample_loc_w": 3, "ray_max_far_dist": 1, "shading_avg_color": 3,
                  "shading_avg_dir": 3, "shading_avg_conf": 1, "shading_avg_embedding": c.point_features_dim}
        t = {k: torch.zeros((R2, n), dtype=torch.float32, device=dev) for k, n in shapes.items()}
        native.check(lib.pnerf_probe(C.byref(pts), C.byref(cam), C.byref(mode), native._ptr(q.sample_loc), native._ptr(q.sample_valid),
                                     native._ptr(last["sigma"]), native._ptr(q.sample_pidx), R2, SR, K,
                                     native._ptr(t["ray_max_shading_opacity"]), native._ptr(t["ray_max_sample_loc_w"]),
                                     native._ptr(t["ray_max_far_dist"]), native._ptr(t["shading_avg_color"]),
                                     native._ptr(t["shading_avg_dir"]), native._ptr(t["shading_avg_conf"]),
                                     native._ptr(t["shading_avg_embedding"]), native._stream()), "pnerf_probe")
        keep = out["ray_mask"].to(torch.float32)[:, None]
        for k, n in shapes.items():
            if q.ray_index is None:            # small bundle: no hit-ray compaction, the rows are the rays
                out[k] = t[k] * keep
            else:
                out[k] = torch.zeros((q.R_total, n), dtype=torch.float32, device=dev).index_copy_(0, q.ray_index.long(), t[k]) * keep
        return out

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, ray_bundle, chunk=None):
        """nerfstudio Model.get_outputs_for_camera_ray_bundle.  The reference slices the image into
        eval_num_rays_per_chunk = 2304 rays (SC:25) and rebuilds its grid for each slice; chunking does not change
        any pixel (rays are independent), so here `chunk` only bounds the rays per query launch (default: the whole
        image) and the field kernels walk the compact list of valid samples in fixed-size pieces.
        An image-shaped bundle (origins (H,W,3), what Nerfstudio's eval loop passes) gives (H,W,.) outputs; a flat one (R,.)."""
        chunk = chunk or RENDER_RAYS_PER_LAUNCH
        shape = tuple(ray_bundle.origins.shape[:-1])
        if len(shape) != 1:          # flatten row-major, metadata included (Model.get_row_major_sliced_ray_bundle upstream)
            flat = lambda t: t.reshape(-1, t.shape[-1]) if (torch.is_tensor(t) and tuple(t.shape[:len(shape)]) == shape) else t
            md = {k: flat(v) for k, v in ray_bundle.metadata.items()}
            ray_bundle = RayBundle(flat(ray_bundle.origins), flat(ray_bundle.directions), flat(ray_bundle.nears), flat(ray_bundle.fars), md)
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        R = len(ray_bundle)
        cols, masks = [], []
        for i in range(0, R, chunk):
            rb = RayBundle(ray_bundle.origins[i:i + chunk], ray_bundle.directions[i:i + chunk], ray_bundle.nears[i:i + chunk],
                           ray_bundle.fars[i:i + chunk], ray_bundle.metadata)
            o = self.get_outputs(rb)
            cols.append(o["coarse_raycolor"])
            masks.append(o["ray_mask"])
        out = {"coarse_raycolor": cols[0] if len(cols) == 1 else torch.cat(cols), "ray_mask": masks[0] if len(masks) == 1 else torch.cat(masks)}
        if len(shape) != 1:
            out = {k: v.view(*shape, -1) for k, v in out.items()}
        return out

    def get_image_metrics_and_images(self, outputs, batch):
        """SM:433-464: metrics of a full evaluation image + the side-by-side picture.  PSNR and RMSE are computed here; SSIM / LPIPS
        need torchmetrics (+ network weights), which the reference imports at module level and this build image lacks."""
        image = batch["image"].to(outputs["coarse_raycolor"].device)
        rgb = outputs["coarse_raycolor"].reshape(image.shape)
        combined = torch.cat([image, rgb], dim=1)
        mse = float(torch.mean((image - rgb) ** 2))
        metrics = {"psnr": float(10.0 * np.log10(1.0 / max(mse, 1e-20))), "rmse": float(np.sqrt(mse))}
        try:
            from torchmetrics.functional.image import structural_similarity_index_measure as ssim
            metrics["torchmetrics_ssim"] = float(ssim(torch.moveaxis(image, -1, 0)[None], torch.moveaxis(rgb, -1, 0)[None]))
        except Exception:
            pass
        return metrics, {"img": combined}

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        """SM:415-431: MSE over the masked rays + 1e-6 and, in training, the zero-one confidence term."""
        pred = outputs["coarse_raycolor"]
        image = batch["image"]
        if image.device != pred.device or image.dtype != pred.dtype:
            image = image.to(device=pred.device, dtype=pred.dtype)
        mask = outputs["ray_mask"]
        if pred.shape[0] > 0:      # MSELoss over the masked_select rows + 1e-6, one kernel each way
            mask8 = mask if (mask.dtype == torch.int8 and mask.is_contiguous()) else mask.to(torch.int8).contiguous()
            loss_dict = {"ray_masked_coarse_raycolor_loss": torch.ops.pnerf.masked_mse(pred, image.contiguous(), mask8)[0]}
        else:
            loss_dict = {"ray_masked_coarse_raycolor_loss": pred.sum() * float("nan")}      # empty bundle: MSELoss of nothing
        if self.training and "conf_coefficient" in outputs:
            h = outputs["conf_coefficient"]
            loss_dict["conf_coefficient_loss"] = torch.ops.pnerf.conf_loss(h.conf.view(-1, 1), h.pidx, h.ray_mask, h.n_rays,
                                                                           float(self.config.zero_epsilon),
                                                                           float(self.config.zero_one_loss_weights))[0]
        for k, v in self.config.loss_coefficients.items():                  # misc.scale_dict (SM:430)
            if k in loss_dict:
                loss_dict[k] = loss_dict[k] * v
        return loss_dict
