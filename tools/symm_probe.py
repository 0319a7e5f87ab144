#!/usr/bin/env python
"""Probe: does torch's symmetric memory (peer-mapped buffers + NVLS multicast) work on this box?
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 40 * 1024 * 1024     # 160 MB of fp32
    try:
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    except Exception as e:
        print(f"rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
        dist.destroy_process_group()
        return
    print(f"rank {rank}: buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} multicast_ptr {hex(hdl.multicast_ptr)} "
          f"signal_pads {len(hdl.signal_pad_ptrs)} world {hdl.world_size}", flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    print(f"rank {rank}: peer[0] = {float(peer[0])} (expect {(rank + 1) % world + 1})", flush=True)
    # peer read bandwidth: sum of the peer's buffer into a local one
    out = torch.empty_like(t)
    for _ in range(3):
        out.copy_(peer)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out.copy_(peer)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"rank {rank}: peer -> local copy of {n * 4 / 1e6:.0f} MB: {ms:.3f} ms = {n * 4 / ms / 1e6:.0f} GB/s", flush=True)
    # NCCL all-reduce of the same size for comparison
    g = torch.ones(n, device=dev)
    for _ in range(5):
        dist.all_reduce(g)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        dist.all_reduce(g)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"rank {rank}: NCCL all-reduce {n * 4 / 1e6:.0f} MB: {ms:.3f} ms", flush=True)
    hdl.barrier()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
