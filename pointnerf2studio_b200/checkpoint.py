"""Neural-point checkpoint layout of the reference (drop-in contract).

Reader: `<dir>/<iter>_net_ray_marching.pth` chosen through the sibling `<iter>_states.pth` names
(studio_model.py:55-59,147-163); a flat dict with neural_points.{xyz, points_embeding, points_conf,
points_dir, points_color, Rw2c} (studio_utils.py:84-90) and, optionally, the original flow's
aggregator.* MLP weights (models/base_model.py:85-102 writes them).  The plugin ignores aggregator.*;
this implementation maps them onto the plugin's module names so original-flow checkpoints render
without retraining (shapes match one to one).
"""
import glob
import os

import torch

POINT_KEYS = ["neural_points.xyz", "neural_points.points_embeding", "neural_points.points_conf",
              "neural_points.points_dir", "neural_points.points_color", "neural_points.Rw2c"]

AGGREGATOR_MAP = {
    "mlp_base.layers.0": "block1.0", "mlp_base.layers.1": "block1.2",
    "mlp_head.layers.0": "block3.0", "mlp_head.layers.1": "block3.2",
    "field_output_density.net": "alpha_branch.0",
    "mlp_color.layers.0": "color_branch.0", "mlp_color.layers.1": "color_branch.2",
    "mlp_color.layers.2": "color_branch.4", "field_output_color.net": "color_branch.6",
}


def latest_epoch(resume_dir):
    names = [f.split("_")[0] for f in os.listdir(resume_dir) if f.endswith("_states.pth")]
    ints = [int(i) for i in names]
    return None if not ints else names[ints.index(max(ints))]


def load_point_cloud_checkpoint(path, map_location="cpu"):
    path = str(path)
    if not os.path.exists(path):
        raise RuntimeError(f"Specified point_cloud path {path} does not exist")
    if not [n for n in glob.glob(path + "/*_net_ray_marching.pth") if os.path.isfile(n)]:
        raise RuntimeError(f"Cannot find any _net_ray_marching.pth in {path}")
    it = latest_epoch(path)
    f = os.path.join(path, f"{it}_net_ray_marching.pth")
    if not os.path.isfile(f):
        raise RuntimeError(f"cannot load {it}_net_ray_marching.pth")
    sd = torch.load(f, map_location=map_location)
    missing = [k for k in POINT_KEYS if k not in sd]
    if missing:
        raise RuntimeError(f"{f} lacks {missing}")
    return sd


def save_point_cloud_checkpoint(path, iteration, model, total_steps=0):
    """Writer in the original flow's format (models/base_model.py:85-102, run/train_studio.py:491-519)."""
    os.makedirs(path, exist_ok=True)
    npnts = model.neural_points
    sd = {"neural_points.xyz": npnts.points_xyz.detach().cpu(),
          "neural_points.points_embeding": npnts.points_embeding.detach().cpu(),
          "neural_points.points_conf": npnts.points_conf.detach().cpu(),
          "neural_points.points_dir": npnts.points_dir.detach().cpu(),
          "neural_points.points_color": npnts.points_color.detach().cpu(),
          "neural_points.Rw2c": npnts.points_Rw2c.detach().cpu()}
    own = dict(model.named_parameters())
    for new, old in AGGREGATOR_MAP.items():
        for s in ("weight", "bias"):
            sd[f"aggregator.{old}.{s}"] = own[f"{new}.{s}"].detach().cpu()
    torch.save(sd, os.path.join(path, f"{iteration}_net_ray_marching.pth"))
    torch.save({"epoch_count": 0, "total_steps": total_steps, "best_PSNR": 0.0, "best_iter": iteration},
               os.path.join(path, f"{iteration}_states.pth"))
