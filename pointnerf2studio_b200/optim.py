"""Fused Adam over the plugin's parameter groups (csrc/optim.cu): torch.optim.Adam semantics, one launch per group.

A `torch.optim.Optimizer` subclass, so `torch.optim.lr_scheduler.LambdaLR` with the reference's decay
(`nerfstudio_plugin.lr_lambda`, studio_utils.py:38-44) drives it unchanged.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AdamSeg, check

MAX_SEGS = 32


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        assert closure is None
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for group in self.param_groups:
            segs = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["step"] = 0
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                assert p.is_cuda and p.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
                segs.append((p, g, st["exp_avg"], st["exp_avg_sq"], st["step"]))
            b1, b2 = group["betas"]
            for i in range(0, len(segs), MAX_SEGS):
                chunk = segs[i:i + MAX_SEGS]
                arr = (AdamSeg * len(chunk))()
                for j, (p, g, m, v, t) in enumerate(chunk):
                    arr[j].p, arr[j].g, arr[j].m, arr[j].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    arr[j].n, arr[j].step, arr[j].lr = p.numel(), t, float(group["lr"])
                check(lib.pnerf_adam_step(C.cast(arr, C.c_void_p), len(chunk), C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]),
                                          C.c_float(grad_scale), stream), "pnerf_adam_step")
            # the kernel writes the parameters through raw pointers: tell torch (and every cache keyed on `_version`, e.g. the
            # bf16 weight pack of native_tc.packed_weights) that they changed
            if segs:
                torch.autograd.graph.increment_version([s[0] for s in segs])
        return None


def make_optimizers(model, lr_fields=5e-4, lr_points=2e-3, lr_decay_exp=0.1, lr_decay_iters=1000000):
    """The plugin's two optimisers + schedulers (studio_config.py:33-48): {"fields", "neural_points"}."""
    from .nerfstudio_plugin import lr_lambda
    groups = model.get_param_groups()
    opts = {"fields": FusedAdam([p for p in groups["fields"] if p.requires_grad], lr=lr_fields),
            "neural_points": FusedAdam([p for p in groups["neural_points"] if p.requires_grad], lr=lr_points)}
    scheds = {k: torch.optim.lr_scheduler.LambdaLR(o, lr_lambda=lambda s: lr_lambda(s, lr_decay_exp, lr_decay_iters))
              for k, o in opts.items()}
    return opts, scheds
