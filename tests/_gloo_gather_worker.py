"""Worker of tests/test_tiles.py: one rank of a world_size-N gloo group (CPU).  Every rank holds the slot-major
accumulator of its tiles, the FrameGather of the package gathers them on rank 0, rank 0 un-permutes and checks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    width, height = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = ge.load_package()
    from ada_ray_tracer_b200 import tiles
    rng = np.random.default_rng(42)                                   # the same "rendered frame" on every rank
    image = rng.uniform(0, 8, (height, width, 4)).astype(np.float32)
    n = tiles.local_slots(width, height, world)
    local = torch.from_numpy(tiles.scatter_image(image, rank, world))
    g = tiles.FrameGather(n, 4, "cpu", torch.float32)
    for _ in range(2):                                                # buffers are reused across passes
        allb = g.gather(local)
    ok = True
    if rank == 0:
        frame = tiles.unpermute(allb.numpy(), width, height, world)
        ok = np.array_equal(frame.view(np.uint32), image.view(np.uint32))
    else:
        ok = allb is None
    # every pixel is owned by exactly one rank
    x, y, act = tiles.slot_to_pixel(np.arange(n), rank, world, width, height)
    cover = torch.zeros(height * width, dtype=torch.int32)
    cover[torch.from_numpy((y[act] * width + x[act]).astype(np.int64))] += 1
    dist.all_reduce(cover)
    ok = ok and bool((cover == 1).all())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 3)


if __name__ == "__main__":
    main()
