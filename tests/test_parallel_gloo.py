"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo processes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointnerf2studio_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.zeros(5, 3))
    b = torch.nn.Parameter(torch.zeros(7))
    c = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    a.grad = torch.full((5, 3), float(rank + 1))
    if rank == 0:
        b.grad = torch.arange(7.0)                 # rank 1 has no grad for b (unused parameter)
    parallel.allreduce_gradients([a, b, c], dist)
    # the in-place path of large tensors (the dense neural-point gradients): same result as the flat bucket
    d = torch.nn.Parameter(torch.zeros(6, 4))
    e = torch.nn.Parameter(torch.zeros(3))
    d.grad = torch.full((6, 4), float(10 * (rank + 1)))
    e.grad = torch.full((3,), float(rank))
    parallel.LARGE = 8
    parallel.allreduce_gradients([d, e], dist)
    # render side: 10 rays split in row blocks
    lo, hi = parallel.shard_rays(10, rank, world)
    rgb = torch.arange(lo, hi, dtype=torch.float32)[:, None].expand(-1, 3).contiguous()
    mask = torch.ones(hi - lo, dtype=torch.int8) * (rank + 1)
    full_rgb, full_mask = parallel.gather_pixels(rgb, mask, 10, dist)
    # interleaved row shards of a 5 x 4 image
    rows = parallel.interleaved_rows(5, rank, world)
    local = torch.tensor([[float(r * 4 + c)] * 3 for r in rows for c in range(4)])
    img = parallel.gather_interleaved_image(local, 5, 4, dist)
    if rank == 0:
        torch.save({"img": img, "a": a.grad, "b": b.grad, "d": d.grad, "e": e.grad, "rgb": full_rgb, "mask": full_mask}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_and_pixel_gather_world2(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    torch.testing.assert_close(r["a"], torch.full((5, 3), 1.5))
    torch.testing.assert_close(r["b"], torch.arange(7.0) / 2)
    torch.testing.assert_close(r["d"], torch.full((6, 4), 15.0))
    torch.testing.assert_close(r["e"], torch.full((3,), 0.5))
    torch.testing.assert_close(r["rgb"][:, 0], torch.arange(10.0))
    assert r["mask"].tolist() == [1] * 5 + [2] * 5
    torch.testing.assert_close(r["img"][:, 0], torch.arange(20.0))


def test_shard_rays_covers_everything():
    for n in (0, 1, 7, 640000):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_rays(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
