#!/usr/bin/env python
"""Run csrc/tc_microbench.cu on all SMs and print per-SM rates (cycles are SM clocks)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointnerf2studio_b200 import _lib


def run(which, iters, param, src):
    lib = _lib.load()
    out = torch.zeros(148 * 9, dtype=torch.int64, device="cuda")
    sink = torch.zeros(1, device="cuda")
    _lib.check(lib.pnerf_tc_microbench(which, iters, param, src.data_ptr(), out.data_ptr(), sink.data_ptr(), None), "microbench")
    torch.cuda.synchronize()
    o = out.cpu().double()
    if which in (3, 4):
        return o[:148].mean().item(), o[148:].sum().item() / 148
    o = o[:148]
    return o.mean().item(), o.max().item()


def main():
    src = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    for N in (256, 128):
        run(0, 2000, N, src)
        mean, mx = run(0, 20000, N, src)
        print(f"mma M=128 N={N} K=16 (smem K-slab operands): {mean / 20000:.1f} clk/MMA (max SM {mx / 20000:.1f}); nominal {128 * N / 256}")
    for lanes in (1, 2, 4, 8):
        for b in (16384, 8192, 4096):
            run(1, 2000, b | (lanes << 20), src)
            mean, mx = run(1, 20000, b | (lanes << 20), src)
            print(f"bulk ring 8 x {b} B, {lanes} producer lane(s): {b * 20000 / mean:.1f} B/clk/SM mean, {b * 20000 / mx:.1f} slowest SM, "
                  f"{mean / 20000:.0f} clk/chunk")
    for which, what in ((3, "16-byte shared stores"), (4, "tcgen05.ld of the other accumulator")):
        for w in (0, 1, 2, 4, 7):
            run(which, 2000, w, src)
            mean, ops = run(which, 20000, w, src)
            extra = f", {ops * 16 * 512 / mean:.1f} B/clk/SM stored" if which == 3 else f", {ops * 4096 / mean:.1f} B/clk/SM read"
            print(f"mma N=256 with {w} warps doing {what}: {mean / 20000:.1f} clk/MMA{extra}")
    for n in (1, 2, 4, 8, 16):
        run(5, 2000, n, src)
        mean, _ = run(5, 20000, n, src)
        print(f"mma N=256 with a tcgen05.commit every {n} MMAs: {mean / 20000:.1f} clk/MMA")
    for w in (4, 8):
        run(2, 200, w, src)
        mean, mx = run(2, 2000, w, src)
        print(f"tcgen05.ld 32x32b.x32, {w} warps x (32 lanes x 256 cols): {w * 32 * 256 * 4 * 2000 / mean:.1f} B/clk/SM, {mean / 2000 / 8:.0f} clk per 32-col load+wait")


if __name__ == "__main__":
    main()
