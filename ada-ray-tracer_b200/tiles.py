"""Tile partition of the image over the GPUs of one box and the frame gather.

The reference scales samples over Ada tasks inside one process (ray_tracer.adb:264-277); here pixels are split instead:
the image is cut into 32x32-pixel tiles, tile t belongs to rank t % world, every rank keeps the whole scene and renders
only its tiles into a *slot-major* accumulator (one float4 per slot):

    slot q of rank r:  tile_local = q // 1024,  tile = tile_local * world + r,  w = (q % 1024) // 32,  l = q % 32
                       ty = tile // tiles_x,  tx = (tile % tiles_x - 3 * ty) mod tiles_x      (rows are rotated by 3 tiles each, so that
                                                                                                a rank does not own whole tile columns)
                       x = tx * 32 + (w % 4) * 8 + (l % 8)
                       y = ty * 32 + (w // 4) * 4 + (l // 8)

(32 consecutive slots = one 8x4 pixel block, so a warp's accumulator traffic is one 512-byte line).  Every rank owns
ceil(ntiles / world) tiles' worth of slots -- the padding slots of the last ranks are inactive -- so the buffers gathered
to rank 0 all have the same size.  There is NO communication while rendering; the only exchange step is the gather of
the accumulators at the end of a Render_Pass, after which rank 0 un-permutes and resolves on the device
(resolve_kernel in csrc/render_kernels.cu, gcore_resolve_gathered_device).  Per-pixel results do not depend on `world`
(the RNG is keyed by pixel and sample index), so the gathered frame is bit-identical to the 1-GPU frame.

This module holds the host-side logic only: the index math (numpy; used for planning, tests and as the executable
statement of the layout) and the torch.distributed plumbing of the gather (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np

TILE = 32
TILE_SLOTS = TILE * TILE
TILE_SKEW = 3                      # GC_TILE_SKEW of csrc/gcore_internal.h


def tile_grid(width, height):
    tiles_x = (width + TILE - 1) // TILE
    tiles_y = (height + TILE - 1) // TILE
    return tiles_x, tiles_y


def local_slots(width, height, world):
    """accumulator slots per rank (identical on every rank)"""
    tiles_x, tiles_y = tile_grid(width, height)
    ntiles = tiles_x * tiles_y
    return ((ntiles + world - 1) // world) * TILE_SLOTS


def slot_to_pixel(q, rank, world, width, height):
    """-> (x, y, active) for slot indices q (array) of `rank`"""
    q = np.asarray(q, np.int64)
    tiles_x, tiles_y = tile_grid(width, height)
    tile = (q // TILE_SLOTS) * world + rank
    r = q % TILE_SLOTS
    w, l = r // 32, r % 32
    ty = tile // tiles_x
    tx = (tile % tiles_x - TILE_SKEW * ty) % tiles_x
    x = tx * TILE + (w % 4) * 8 + (l % 8)
    y = ty * TILE + (w // 4) * 4 + (l // 8)
    active = (tile < tiles_x * tiles_y) & (x < width) & (y < height)
    return x, y, active


def pixel_to_slot(x, y, world, width, height):
    """-> (rank, slot) for pixel coordinates (arrays)"""
    x = np.asarray(x, np.int64); y = np.asarray(y, np.int64)
    tiles_x, _ = tile_grid(width, height)
    tile = (y // TILE) * tiles_x + (x // TILE + TILE_SKEW * (y // TILE)) % tiles_x
    lx, ly = x % TILE, y % TILE
    w = (ly // 4) * 4 + (lx // 8)
    l = (ly % 4) * 8 + (lx % 8)
    return tile % world, (tile // world) * TILE_SLOTS + w * 32 + l


def scatter_image(image, rank, world):
    """row-major image[H, W, C] -> the slot-major buffer of `rank` (inactive slots are zero); tests / planning only"""
    h, w = image.shape[:2]
    n = local_slots(w, h, world)
    x, y, active = slot_to_pixel(np.arange(n), rank, world, w, h)
    out = np.zeros((n,) + image.shape[2:], image.dtype)
    out[active] = image[y[active], x[active]]
    return out


def unpermute(all_slots, width, height, world):
    """`world` slot-major buffers laid out back to back -> row-major [H, W, C]; the numpy statement of what
    resolve_kernel does on the device for the gathered accumulators"""
    n = local_slots(width, height, world)
    all_slots = np.asarray(all_slots)
    all_slots = all_slots.reshape(world, n, all_slots.size // (world * n))
    yy, xx = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    rank, slot = pixel_to_slot(xx, yy, world, width, height)
    return all_slots[rank, slot]


class FrameGather:
    """Gathers the per-rank slot-major accumulators on rank `dst` with torch.distributed (NCCL / gloo).

    gather(local) returns, on `dst`, one tensor holding the `world` buffers back to back (rank r at r * nslots) -- the layout
    gcore_resolve_gathered_device expects -- and None elsewhere.  Buffers are allocated once."""

    def __init__(self, nslots, channels=4, device="cpu", dtype=None, group=None, dst=0):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.dst = dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nslots = nslots
        dtype = dtype or torch.float32
        self.all = None
        self.views = None
        if self.rank == dst:
            self.all = torch.empty((self.world, nslots, channels), dtype=dtype, device=device)
            self.views = [self.all[r] for r in range(self.world)]

    def gather(self, local):
        if self.world == 1:
            self.all[0].copy_(local.reshape(self.all[0].shape))
            return self.all
        self.dist.gather(local.reshape(self.nslots, -1), self.views if self.rank == self.dst else None, dst=self.dst, group=self.group)
        return self.all
