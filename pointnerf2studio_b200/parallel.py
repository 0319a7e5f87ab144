"""Ray-sharded data parallelism: one process per GPU, the point cloud / grid / MLP weights are
replicated, every rank draws its own ray batch, and the gradients are summed over ranks once per step.

Mirrors what the reference gets from `DDP(model, find_unused_parameters=True)` (studio_pipeline.py:48-53): an
average of every trainable parameter's gradient -- MLP weights (1.37 MB) and the dense neural-point tensors
(156 B per point).  One flat bucket per step: over NVSwitch every peer is one hop, so a single large all-reduce
(NVLS when available) beats DDP's 25 MB buckets; parameters that received no gradient on a rank contribute
zeros (DDP's find_unused_parameters behaviour).
"""
from __future__ import annotations

from typing import Iterable, List

import torch


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the ray list owned by `rank` (row-block sharding of an image)."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    return lo, min(lo + per, n_rays)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], dist, average: bool = True) -> None:
    """In-place sum (or mean) of `.grad` over all ranks through one flat fp32 bucket."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    world = dist.get_world_size()
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat)
    if average:
        flat.div_(world)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


def gather_pixels(local_rgb: torch.Tensor, local_mask: torch.Tensor, n_rays: int, dist):
    """Render-side collective: all-gather the per-rank (R_r,3) pixels and (R_r,) masks of a row-block-sharded
    image into the full (R,3) / (R,) tensors on every rank."""
    world = dist.get_world_size()
    per = (n_rays + world - 1) // world
    dev = local_rgb.device
    pad_rgb = torch.zeros((per, 3), dtype=local_rgb.dtype, device=dev)
    pad_mask = torch.zeros((per,), dtype=local_mask.dtype, device=dev)
    pad_rgb[:local_rgb.shape[0]] = local_rgb
    pad_mask[:local_mask.shape[0]] = local_mask
    rgb = [torch.empty_like(pad_rgb) for _ in range(world)]
    mask = [torch.empty_like(pad_mask) for _ in range(world)]
    dist.all_gather(rgb, pad_rgb)
    dist.all_gather(mask, pad_mask)
    return torch.cat(rgb)[:n_rays], torch.cat(mask)[:n_rays]
