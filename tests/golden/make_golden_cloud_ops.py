"""Golden vectors for the point-cloud maintenance rows, produced by EXECUTING the reference's own source
on CPU.  Runs only where /root/reference exists (the authoring container); the output, tests/golden/cloud_ops_golden.npz, is committed.

  * construct_vox_points_closest (models/mvs/mvs_utils.py:537-561): the function's source is cut out of the file with `ast` and
    executed with pure-torch stand-ins for torch_scatter.scatter_mean / scatter_min (the package is not installed; the stand-ins
    implement its documented semantics: per-index mean; per-index minimum whose argument is the first minimal element).
  * probe_hole's candidate filter (run/train_studio.py:414-423) and bloat_inds (:447-455): those source lines are cut out by line
    number, `.cuda()` / device="cuda" are rewritten to the CPU, and executed on seeded synthetic probe maps.
"""
import ast
import os
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference/pointnerf"
HERE = os.path.dirname(os.path.abspath(__file__))


def scatter_mean(src, index, dim=0):
    n = int(index.max()) + 1
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=torch.float64).index_add_(0, index, src.double())
    cnt = torch.bincount(index, minlength=n).double()
    return (out / cnt.view(-1, *([1] * (src.dim() - 1)))).to(src.dtype)


def scatter_min(src, index, dim=0):
    n = int(index.max()) + 1
    best = torch.full((n,), float("inf"), dtype=src.dtype)
    arg = torch.full((n,), -1, dtype=torch.long)
    s, ix = src.numpy(), index.numpy()
    b, a = best.numpy(), arg.numpy()
    for i in range(len(s)):
        if s[i] < b[ix[i]]:
            b[ix[i]], a[ix[i]] = s[i], i
    return best, arg


def function_source(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def main():
    out = {}
    # ---- construct_vox_points_closest
    ns = {"torch": torch, "scatter_mean": scatter_mean, "scatter_min": scatter_min, "print": lambda *a, **k: None}
    exec(function_source(os.path.join(REF, "models/mvs/mvs_utils.py"), "construct_vox_points_closest"), ns)
    g = torch.Generator().manual_seed(77)
    for tag, n, res in (("a", 4000, 24), ("b", 20000, 64)):
        centres = torch.rand((40, 3), generator=g) * 2 - 1
        xyz = (centres[torch.randint(0, 40, (n,), generator=g)] + 0.08 * torch.randn((n, 3), generator=g)).float()
        cen, grid, amin = ns["construct_vox_points_closest"](xyz, res)
        out[f"vox_{tag}_xyz"], out[f"vox_{tag}_res"] = xyz.numpy(), np.int32(res)
        out[f"vox_{tag}_centroid"], out[f"vox_{tag}_grid"], out[f"vox_{tag}_min_idx"] = cen.numpy(), grid.numpy(), amin.numpy()
    # ---- probe_hole's filter: lines 417-426 of run/train_studio.py + bloat_inds
    lines = open(os.path.join(REF, "run/train_studio.py")).read().split("\n")
    bloat = function_source(os.path.join(REF, "run/train_studio.py"), "bloat_inds").replace(".cuda()", "")
    first = next(i for i, ln in enumerate(lines) if ln.strip().startswith("miss_ray_mask = (prob_maps"))
    last = next(i for i, ln in enumerate(lines) if ln.strip().startswith('neighboring_miss_mask = (prob_maps["ray_mask"].squeeze'))
    print("probe_hole filter lines (1-based):", first + 1, "-", last + 1)
    body = textwrap.dedent("\n".join(lines[first:last + 1]))        # miss_ray_mask ... neighboring_miss_mask
    assert body.lstrip().startswith("miss_ray_mask") and "opacity_thresh" in body and last - first < 16, body
    for tag, far_thresh in (("nofar", -1.0), ("far", 0.012)):
        H, W = 40, 56
        g = torch.Generator().manual_seed(5)
        ray_mask = (torch.rand((H, W, 1), generator=g) > 0.35).float()
        bg = torch.tensor([[1.0, 1.0, 1.0]])
        gt = torch.rand((H, W, 3), generator=g)
        gt[torch.rand((H, W), generator=g) > 0.6] = 1.0                      # background pixels
        color = (gt + 0.12 * torch.randn((H, W, 3), generator=g)).clamp(0, 1)
        prob_maps = {"ray_mask": ray_mask, "coarse_raycolor": color, "ray_max_far_dist": torch.rand((H, W, 1), generator=g) * 0.03,
                     "ray_max_shading_opacity": torch.rand((H, W, 1), generator=g)}
        edge_mask = torch.rand((H * W,), generator=g) > 0.1
        ns = {"torch": torch, "prob_maps": prob_maps, "gt_image": gt, "bg": bg, "edge_mask": edge_mask, "height": H, "width": W,
              "opt": types.SimpleNamespace(far_thresh=far_thresh), "opacity_thresh": 0.7}
        exec(bloat, ns)
        exec(body, ns)
        out[f"probe_{tag}_ray_mask"] = ray_mask[..., 0].numpy().astype(np.int8)
        out[f"probe_{tag}_gt"], out[f"probe_{tag}_color"] = gt.numpy(), color.numpy()
        out[f"probe_{tag}_far_dist"] = prob_maps["ray_max_far_dist"][..., 0].numpy()
        out[f"probe_{tag}_opacity"] = prob_maps["ray_max_shading_opacity"][..., 0].numpy()
        out[f"probe_{tag}_edge"] = edge_mask.reshape(H, W).numpy()
        out[f"probe_{tag}_far_thresh"] = np.float32(far_thresh)
        out[f"probe_{tag}_keep"] = ns["neighboring_miss_mask"].numpy()
        assert 0 < out[f"probe_{tag}_keep"].sum() < H * W
    # ---- the probe outputs (models/neural_points_volumetric_model.py:334-355): the body of `if weight is not None:` executed on
    # seeded tensors of the shapes run_network_models hands it (B = 1)
    lines = open(os.path.join(REF, "models/neural_points_volumetric_model.py")).read().split("\n")
    first = next(i for i, ln in enumerate(lines) if 'output["ray_max_shading_opacity"], opacity_ind = torch.max(' in ln)
    last = next(i for i, ln in enumerate(lines) if 'output["shading_avg_embedding"] = torch.sum(sampled_embedding * weight' in ln)
    print("probe output lines (1-based):", first + 1, "-", last + 1)
    body = textwrap.dedent("\n".join(lines[first:last + 1]))
    g = torch.Generator().manual_seed(11)
    R, SR, K = 37, 12, 8
    rnd = lambda *sh: torch.rand(sh, generator=g)
    ns = {"torch": torch, "output": {"coarse_point_opacity": rnd(1, R, SR)}, "sample_loc_w": rnd(1, R, SR, 3) * 2 - 1,
          "weight": rnd(1, R, SR, K), "conf_coefficient": rnd(1, R, SR, K), "sampled_xyz": rnd(1, R, SR, K, 3) * 2 - 1,
          "sampled_color": rnd(1, R, SR, K, 3), "sampled_dir": rnd(1, R, SR, K, 3) * 2 - 1, "sampled_conf": rnd(1, R, SR, K, 1),
          "sampled_embedding": rnd(1, R, SR, K, 32)}
    ns["output"]["coarse_point_opacity"][0, 5, 3] = ns["output"]["coarse_point_opacity"][0, 5, 7] = 2.0      # a tie: torch.max keeps the first
    inputs = {k: v.clone() for k, v in ns.items() if torch.is_tensor(v)}
    inputs["coarse_point_opacity"] = ns["output"]["coarse_point_opacity"].clone()
    exec(body, ns)
    for k, v in inputs.items():
        out[f"probeout_in_{k}"] = v.numpy()
    for k in ("ray_max_shading_opacity", "ray_max_sample_loc_w", "ray_max_far_dist", "shading_avg_color", "shading_avg_dir",
              "shading_avg_conf", "shading_avg_embedding"):
        out[f"probeout_{k}"] = ns["output"][k].numpy()
    np.savez_compressed(os.path.join(HERE, "cloud_ops_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
