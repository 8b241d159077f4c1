"""Tile partition + frame gather: index math on the CPU, the gather protocol on a world_size-2/3 gloo group,
and the tiled renderer against the plain one on the GPU."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiles(pkg):
    return importlib.import_module(pkg.__name__ + ".tiles")


@pytest.mark.parametrize("w,h", [(1024, 1024), (200, 136), (33, 65), (1, 1)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_slot_pixel_maps_are_inverse_and_cover_the_image(pkg, w, h, world):
    t = _tiles(pkg)
    n = t.local_slots(w, h, world)
    assert n % 1024 == 0 and n * world >= w * h
    seen = np.zeros((h, w), np.int32)
    for rank in range(world):
        x, y, act = t.slot_to_pixel(np.arange(n), rank, world, w, h)
        seen[y[act], x[act]] += 1
        r2, q2 = t.pixel_to_slot(x[act], y[act], world, w, h)
        assert np.all(r2 == rank) and np.array_equal(q2, np.arange(n)[act])
    assert np.all(seen == 1)
    # 32 consecutive slots are one 8x4 pixel block
    x, y, act = t.slot_to_pixel(np.arange(32), 0, world, w, h)
    assert x.max() - x.min() <= 7 and y.max() - y.min() <= 3


def test_unpermute_inverts_scatter(pkg):
    t = _tiles(pkg)
    rng = np.random.default_rng(1)
    img = rng.normal(size=(70, 90, 4)).astype(np.float32)
    for world in (1, 2, 5):
        allb = np.stack([t.scatter_image(img, r, world) for r in range(world)])
        assert np.array_equal(t.unpermute(allb, 90, 70, world), img)


@pytest.mark.parametrize("world,w,h", [(2, 200, 136), (3, 96, 64)])
def test_gather_protocol_gloo(world, w, h):
    """the N>1 path on CPU: torch.distributed (gloo), world_size 2 and 3"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "_gloo_gather_worker.py"), str(w), str(h)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_tiled_renderer_world1_equals_plain(gscene, pkg):
    import torch
    w, h = 160, 96
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 123
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    ref_hdr = rt.Render_Pass(read_hdr=True).copy(); ref_ldr = rt.screen_buffer.copy()
    tr = pkg.Tiled_Ray_Tracer(rt, 0)
    tr.Resize_Viewport(w, h)
    host = torch.empty(w * h, dtype=torch.int32).pin_memory()
    tr.Render_Pass(host_ldr=host, want_hdr=True)
    pkg.synchronize()
    assert np.array_equal(host.numpy().view(np.uint32).reshape(h, w), ref_ldr)
    got = tr.d_hdr.cpu().numpy().reshape(h, w, 4)
    assert np.array_equal(got[..., :3].view(np.uint32), ref_hdr[..., :3].view(np.uint32))
