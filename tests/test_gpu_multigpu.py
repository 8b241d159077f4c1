"""Multi-GPU rendering, both ways the library offers it -- skipped on boxes with one GPU:

  * ONE host process driving G devices behind the C ABI (gcore_set_gpu_count): what an Ada caller of gcore_render gets; the
    accumulators meet on device 0 through a grouped ncclSend/ncclRecv.  Replayed from plain C too (replay_test_adb.c).
  * one process per GPU (torchrun) with torch.distributed/NCCL for the gather: Tiled_Ray_Tracer, what bench.py --gpus N times.

Either way every pixel must carry the bits of the 1-GPU render."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy(); b[np.isnan(a)] = 0x7FC00000
    return b


needs2 = pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")


@needs2
@pytest.mark.parametrize("scene", ["pyramid2", "lbvh"])
def test_single_process_multi_gpu_bit_identical(pkg, ob, scene):
    import importlib.util
    pkg.set_partition(0, 1); pkg.set_gpu_count(1)
    gs = pkg.Scene(); gs.Init()
    if scene == "lbvh":
        spec = importlib.util.spec_from_file_location("syn", os.path.join(os.path.dirname(pkg.__file__), "synthetic.py"))
        syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)
        pos, nrm, tris = syn.cube_sphere(24, radius=0.9, center=(0.3, 1.2, 2.6))
        gs.set_mesh(pos, nrm, tris, mesh_mode=pkg.MESH_NEAREST)
    w, h = 200, 136
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 3; rt.seed = 4711
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(w, h)
    ref = rt.Render_Pass(read_hdr=True).copy(); ref_ldr = rt.screen_buffer.copy(); ref_rays = int(rt.last_stats.rays)
    for g in sorted({2, min(_ngpus(), 3), _ngpus()}):
        pkg.set_gpu_count(g)
        assert pkg.get_gpu_count() == g
        rt.Resize_Viewport(w, h)
        got = rt.Render_Pass(read_hdr=True)                       # gcore_render_units on g devices + gcore_read_frame (NCCL gather)
        assert np.array_equal(_bits(got[..., :3]), _bits(ref[..., :3])), "gpu_count=%d" % g
        assert np.array_equal(rt.screen_buffer, ref_ldr) and int(rt.last_stats.rays) == ref_rays
        ms = pkg.device_kernel_ms(g)
        assert np.all(ms > 0)
        # progressive second pass, checkpoint / resume across device counts
        rt.Render_Pass()
        assert rt.GetSPP() == 24
    # the one-shot call too
    hdr = np.zeros((h, w, 4), np.float32); ldr = np.zeros((h, w), np.uint32)
    st = rt.render(w, h, 3, hdr, ldr)
    assert np.array_equal(_bits(hdr[..., :3]), _bits(ref[..., :3])) and np.array_equal(ldr, ref_ldr) and st.rays == ref_rays
    pkg.set_gpu_count(1)
    gs.Init()


@needs2
def test_c_harness_with_all_gpus_writes_the_same_bmp(pkg, tmp_path):
    exe = str(tmp_path / "replay_test_adb")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tests", "c_harness", "replay_test_adb.c"), "-ldl"], check=True)
    out = []
    for g in (1, _ngpus()):
        bmp = str(tmp_path / ("g%d.bmp" % g))
        r = subprocess.run([exe, pkg.LIB_PATH, os.path.join(ROOT, "data", "pyramid2.vsgf"), bmp, "320", "200", "4", "2", str(g)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        out.append((open(bmp, "rb").read(), r.stdout.strip().splitlines()[-1].split()[:2]))     # (NCCL may print its version banner first)
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]      # same bytes, same spp and ray count


@needs2
def test_exclusive_partition_modes(pkg):
    pkg.set_partition(0, 1); pkg.set_gpu_count(2)
    assert pkg.lib().gcore_set_partition(0, 2) == -3             # GCORE_E_STATE
    pkg.set_gpu_count(1)
    pkg.set_partition(1, 2)
    assert pkg.lib().gcore_set_gpu_count(2) == -3
    pkg.set_partition(0, 1)


@needs2
@pytest.mark.parametrize("workload", ["pyramid2", "lbvh"])
def test_torchrun_tiled_ray_tracer_nccl(pkg, workload, tmp_path):
    """the real multi-process path: torchrun, one rank per GPU, Tiled_Ray_Tracer + dist.gather over NCCL, against a 1-GPU render"""
    n = min(_ngpus(), 4)
    out = str(tmp_path / "frame.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29711", os.path.join(ROOT, "tests", "_nccl_tiled_worker.py"), workload, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    z = np.load(out)
    assert np.array_equal(_bits(z["tiled_hdr"]), _bits(z["ref_hdr"])) and np.array_equal(z["tiled_ldr"], z["ref_ldr"])
