"""torchrun worker of tests/test_gpu_multigpu.py: every rank renders its tiles with Tiled_Ray_Tracer, rank 0 also renders the
whole frame alone (world 1) and saves both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    workload, out = sys.argv[1], sys.argv[2]
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 2; rt.seed = 31337
    tr = pkg.Tiled_Ray_Tracer(rt, local)
    scene = pkg.Scene(); scene.Init()
    if workload == "lbvh":
        import importlib
        syn = importlib.import_module(pkg.__name__ + ".synthetic")
        pos, nrm, tris = syn.cube_sphere(24, radius=0.9, center=(0.3, 1.2, 2.6))
        scene.set_mesh(pos, nrm, tris, mesh_mode=pkg.MESH_NEAREST)
    w, h = 232, 168
    rt.Init_Render(pkg.PT_MIS)
    tr.Resize_Viewport(w, h)
    host = torch.empty(w * h, dtype=torch.int32).pin_memory() if rank == 0 else None
    tr.Render_Pass(host_ldr=host, want_hdr=True)
    pkg.synchronize(); torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        tiled_hdr = tr.d_hdr.cpu().numpy().reshape(h, w, 4)[..., :3].copy()
        tiled_ldr = host.numpy().view(np.uint32).reshape(h, w).copy()
        pkg.set_partition(0, 1)
        rt.Resize_Viewport(w, h)
        ref = rt.Render_Pass(read_hdr=True)
        np.savez(out, tiled_hdr=tiled_hdr, tiled_ldr=tiled_ldr, ref_hdr=ref[..., :3], ref_ldr=rt.screen_buffer)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
