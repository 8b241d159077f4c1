#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (see DESIGN.md "Measurement").

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload C4_1024|C2|C3|C4|C1]

One "step" = one Render_Pass of the hot path (ray generation, scene queries, NEE + shadow rays, BSDF sampling, per-pixel
accumulation) over the whole image: `units` DoPass units of 4 fixed sub-pixel samples each.  The default workload is the
one BASELINE.json's metric is quoted on: pyramid2 Cornell scene, PT_MIS, depth 8, 1024x1024, 256 spp per step.

Printed JSON line (rank 0):
  value   Mrays/s, whole job, scene and accumulators resident in HBM, device-timed (CUDA events on the library's stream,
          max over ranks); N>1: image tiles split round-robin over the ranks, slot-major accumulators gathered to rank 0
          with NCCL and resolved there, inside the timed region.
  e2e     same metric through the reference-facing C-ABI call gcore_render() with HOST buffers: scene description
          re-sent (host->device) and the resolved frame read back (device->host) every step.
  roofline / cpu_baseline as specified in the task; the CPU leg times the oracle (a C++ restatement of the Ada
          renderer; neither GNAT nor Embree exists here) on the host cores with all threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

WORKLOADS = {
    # name: (width, height, render_type, aa, max_depth, units per step, description)
    "C4_1024": (1024, 1024, "PT_MIS", 1, 8, 64, "pyramid2 Cornell scene, PT_MIS depth 8, AA 4 fixed sub-pixels, 1024x1024, 256 spp per step"),
    "C4_1024_eighth": (1024, 128, "PT_MIS", 1, 8, 64, "scaling analysis: one eighth of the C4_1024 image (rows 0..127) on one GPU, 256 spp per step"),
    "C1": (1024, 768, "PT_MIS", 1, 8, 28, "reference default: 1024x768 PT_MIS depth 8, 28 units (112 spp) per Render_Pass"),
    "C2": (1024, 1024, "RT_DEBUG", 0, 1, 1, "primary-ray cast only (Debug_Ray_Tracing), 1024x1024, 1 spp (pixel centres)"),
    "C3": (1024, 1024, "PT_SHADOW", 1, 1, 4, "direct lighting with shadow rays, 1024x1024, 16 spp"),
    "C4": (2048, 2048, "PT_MIS", 1, 8, 64, "multi-bounce PT_MIS depth 8, 2048x2048, 256 spp per step"),
    "C5_small": (960, 540, "PT_MIS", 1, 8, 4, "C5 scene at 960x540, 16 spp per step (profiling size)"),
    "C5": (3840, 2160, "PT_MIS", 1, 8, 16, "synthetic 10 002 828-triangle cube-sphere (GPU LBVH) + 64 analytic spheres in the Cornell box, PT_MIS depth 8, 3840x2160, 64 spp per step"),
}
SEED = 0x5EED0004
FLOP_QUERY_MESH, FLOP_QUERY_NOMESH, FLOP_BOUNCE = 564.0, 140.0, 200.0      # SURVEY 8d, counted from the reference source


def kernel_source_hash():
    """sha256 over the device-side sources of the library (every kernel, every header they include, the compile flags; not gcore_api.cu, the
    host side of the C ABI): ties recorded ncu figures (profiles/r02_traffic.json) to the kernels they were taken on"""
    import glob
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "ada-ray-tracer_b200", "csrc")
    for f in sorted([os.path.join(csrc, "render_kernels.cu"), os.path.join(csrc, "lbvh_build.cu"), os.path.join(csrc, "Makefile")]
                    + glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.h"))):
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def recorded_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture -- only if it was taken on THIS source"""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tj.get("kernel_source_sha16") == kernel_source_hash() and workload in tj.get("workloads", {}):
            return tj["workloads"][workload]["dram_bytes_per_launch"], None
        return None, "profiles/r02_traffic.json was captured on other kernel sources (%s): not reported" % tj.get("kernel_source_sha16")
    except Exception:
        return None, "no ncu capture recorded"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0, ngpus=1):
        self.p = None
        self.gpu_index = gpu_index
        self.ngpus = ngpus

    def start(self):
        try:
            ids = ",".join(str(self.gpu_index + k) for k in range(self.ngpus))
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", ids], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ""
        sm, mx, pw, reasons = [], [], [], set()
        per_gpu = {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                per_gpu.setdefault(f[0], []).append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        res = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
               "samples": len(sm), "reasons": sorted(reasons)}
        if len(per_gpu) > 1:
            res["sm_mhz_per_gpu"] = [float(np.median(v)) for _, v in sorted(per_gpu.items(), key=lambda kv: int(kv[0]))]
        return res


_ORACLE_SCENES = {}


def oracle_scene(workload):
    """the CPU oracle's scene for a workload (the C5 scenes carry the 10 002 828-triangle mesh + 64 spheres; the oracle then
    uses its own CPU BVH -- a binned-SAH restatement, not Embree)"""
    ob = ge.load_oracle()
    key = "C5" if workload.startswith("C5") else "cornell"
    if key not in _ORACLE_SCENES:
        sc = ob.OracleScene()
        if key == "C5":
            import importlib
            syn = importlib.import_module(ge.load_package().__name__ + ".synthetic")
            sp, _box = syn.c5_analytic()
            pos, nrm, tris = syn.cube_sphere(913)
            sc.set_spheres(sp); sc.set_mesh(pos, nrm, tris)
            # build the CPU BVH outside any timed region (it is built lazily by the first query)
            sc.render(ob.make_params(8, 8, ob.PT_MIS, passes=1, mesh_mode=ob.MESH_NEAREST))
        _ORACLE_SCENES[key] = sc
    return _ORACLE_SCENES[key]


def oracle_run(width, height, rtype, aa, depth, units, threads, seed=SEED, first_sample=0, workload="C4_1024", rows=None):
    ob = ge.load_oracle()
    sc = oracle_scene(workload)
    mesh_mode = ob.MESH_NEAREST if workload.startswith("C5") else ob.MESH_REFERENCE_BF
    y0, y1 = rows if rows else (0, 0)
    p = ob.make_params(width, height, getattr(ob, rtype), aa=aa, max_depth=depth, passes=units, seed=seed,
                       first_sample=first_sample, threads=threads, mesh_mode=mesh_mode, y0=y0, y1=y1)
    t0 = time.perf_counter()
    _, rays = sc.render(p)
    return rays, time.perf_counter() - t0


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (kind "port": the Ada sources
    cannot be compiled here), all host threads, a bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    width, height, rtype, aa, depth, units, desc = wl
    cores = host_cores()
    sample_units = min(8, units)                        # up to 32 spp of the full image per step (~1-2 s of all host cores)
    if args.workload == "C5":                           # 8.3 Mpixel x 10 M triangles on the CPU: one unit (4 spp) of the full image per step
        sample_units = 1
    for _ in range(max(args.warmup, 1)):
        oracle_run(width, height, rtype, aa, depth, sample_units, cores, workload=args.workload)
    rays_tot, t_tot = 0, 0.0
    for k in range(args.steps):
        rays, dt = oracle_run(width, height, rtype, aa, depth, sample_units, cores, first_sample=4 * sample_units * k, workload=args.workload)
        rays_tot += rays; t_tot += dt
    v = rays_tot / t_tot / 1e6
    sample = "%dx%d %s depth %d, %d unit(s) (%d spp) of the full image per step" % (width, height, rtype, depth, sample_units,
                                                                                   sample_units * (4 if aa else 1))
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sample": sample, "note": "CPU oracle port of the Ada renderer (C++ -O2, no FMA); not GNAT, not Embree: neither exists in this image"
                       + ("; the 10 M-triangle mesh goes through the oracle's own CPU BVH" if args.workload.startswith("C5") else "")},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_device(pkg, tr, rt, args, workload, world, rank, local_rank, barrier, flush, stream, steps, warmup, sample_clocks=True):
    """the `value` leg of one workload: device-resident, CUDA events on the library's stream, max over ranks.
    -> dict of raw measurements (all ranks return the same numbers)"""
    import torch
    import torch.distributed as dist
    # one untimed counting pass (statistics level 2): the share of scene queries that enter the reference's mesh bounding box
    # (564 instead of 140 algorithmic flop); the timed passes run with the cheap counters
    pkg.set_stats_level(2)
    cc0 = pkg.get_counters()
    tr.Render_Pass()
    barrier()
    cc1 = pkg.get_counters()
    mesh_share = torch.tensor([cc1[1] - cc0[1], cc1[0] - cc0[0]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mesh_share, op=dist.ReduceOp.SUM)
    mesh_share = float(mesh_share[0] / torch.clamp(mesh_share[1], min=1.0))
    pkg.set_stats_level(1)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            flush.fill_(1); tr.Render_Pass()
    barrier()
    c0 = pkg.get_counters()
    bv0 = pkg.get_bvh_counters()
    sampler = ClockSampler(local_rank, world) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
    launches = 0
    barrier()
    wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(steps):
            flush.fill_(k & 255)                         # L2 flush between timed iterations (outside the events)
            launches += tr.Render_Pass(events=ev[k])     # render kernels, NCCL gather of the accumulators, resolve on rank 0
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    c1 = pkg.get_counters()
    bv1 = pkg.get_bvh_counters()
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    rays_local, mesh_local, bounce_local, _unused = (c1[i] - c0[i] for i in range(4))
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([rays_local, mesh_local, bounce_local, launches, bv1[0] - bv0[0], bv1[1] - bv0[1], bv1[2] - bv0[2]], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, dtype=torch.float64, device="cuda"); per_rank[rank] = kern_ms / steps
    per_rank_rays = torch.zeros(world, dtype=torch.float64, device="cuda"); per_rank_rays[rank] = rays_local / steps
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM); dist.all_reduce(per_rank, op=dist.ReduceOp.SUM); dist.all_reduce(per_rank_rays, op=dist.ReduceOp.SUM)
    rays, _mesh_stage_rays, bounces, launches_all, bvh_nodes, bvh_tris, bvh_trav = (float(x) for x in cnt)
    return {"step_ms": float(t[0]), "kern_ms": float(t[1]), "rays": rays, "bounces": bounces, "mesh_q": mesh_share * rays, "launches": launches_all,
            "bvh_nodes": bvh_nodes, "bvh_tris": bvh_tris, "bvh_trav": bvh_trav, "wall": wall, "clocks": clocks,
            "kernel_ms_per_rank": [float(x) for x in per_rank], "rays_per_rank": [float(x) for x in per_rank_rays]}


def lbvh_roofline(m, steps, world, hbm_bytes, peaks, rtype, achieved_tf, peak_tf, traffic, traffic_note):
    """LBVH scene: the mesh traversal is the bandwidth-shaped part (SURVEY 8d ii).  `achieved` uses SURVEY's algorithmic bytes -- 64 B per
    node visited + 48 B per triangle slot tested (both counted by the kernel) + 32 B ray in + 16 B hit out per traversal + the colour
    buffer -- against the measured HBM copy rate; `layout_bytes` is what this library's own layout moves for the same visits (32-B
    quantised nodes)."""
    kern_s = m["kern_ms"] * 1e-3 / steps
    trav_bytes = 64.0 * m["bvh_nodes"] + 48.0 * m["bvh_tris"] + 48.0 * m["bvh_trav"]
    layout_bytes = 32.0 * m["bvh_nodes"] + 48.0 * m["bvh_tris"] + 48.0 * m["bvh_trav"]
    bytes_per_launch = (trav_bytes / steps + hbm_bytes * world) / world
    gbs = bytes_per_launch / kern_s / 1e9
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
            "note": "algorithmic bytes (SURVEY 8d): 64 B x LBVH nodes visited + 48 B x triangle slots tested (kernel counters) + ray/hit + colour buffer",
            "layout_bytes_per_launch": (layout_bytes / steps + hbm_bytes * world) / world, "algorithmic_bytes_per_launch": bytes_per_launch,
            "kernel": "render_bvh_kernel<%s> + fold_kernel" % rtype, "kernel_ms_per_launch": m["kern_ms"] / steps,
            "nodes_per_traversal": m["bvh_nodes"] / max(m["bvh_trav"], 1.0), "tris_per_traversal": m["bvh_tris"] / max(m["bvh_trav"], 1.0),
            "traversals_per_query": m["bvh_trav"] / max(m["rays"], 1.0),
            "issue_view": {"achieved_TFLOPs": achieved_tf, "peak_TFLOPs": peak_tf, "frac": achieved_tf / peak_tf}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4_1024", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C5_small (LBVH scene) block of the default line")
    ap.add_argument("--numerics", default="exact", choices=["exact", "fast"], help="exact: the bit-exact product build (default); fast: the opt-in FMA / approximate-division build")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    width, height, rtype, aa, depth, units, desc = wl
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render core has no CPU path")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")           # host-side barrier for the phase in which rank 0 drives every GPU itself

    pkg = ge.load_package()
    pkg.select_numerics(args.numerics)
    L = pkg.lib()
    rt = pkg.Ray_Tracer()
    rt.Threads_Num = units; rt.Anti_Aliasing_On = bool(aa); rt.Max_Trace_Depth = depth; rt.seed = SEED
    tr = pkg.Tiled_Ray_Tracer(rt, local_rank)            # set_device + tile partition (rank, world) + the library's stream
    scene = pkg.Scene(); scene.Init()
    syn = None
    if args.workload.startswith("C5"):
        import importlib
        syn = importlib.import_module(pkg.__name__ + ".synthetic")
        syn.install_c5(scene)
    rt.Init_Render(getattr(pkg, rtype))
    tr.Resize_Viewport(width, height)
    stream = tr.stream
    npx = width * height
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- value: device-resident, CUDA events
    m = measure_device(pkg, tr, rt, args, args.workload, world, rank, local_rank, barrier, flush, stream, args.steps, args.warmup)
    step_ms, kern_ms, rays, bounces, mesh_q = m["step_ms"], m["kern_ms"], m["rays"], m["bounces"], m["mesh_q"]
    value = rays / (step_ms * 1e-3) / 1e6

    # ---------------------------------------------------------------- e2e: the reference-facing C-ABI call gcore_render() with HOST
    # buffers, every step: scene description host -> device, fresh accumulator, render, gather, resolve, frame device -> pinned host
    # memory.  ONE host process, like the Ada caller: for N > 1 rank 0 drives all N GPUs itself through gcore_set_gpu_count(N) (the
    # accumulators meet on device 0 over NCCL inside the library) while the other ranks wait at a host-side barrier.
    cfg = rt._cfg()
    h2d = 12 * 1024 * world if not args.workload.startswith("C5") else 0     # sizeof(DeviceScene) per device: the scene block is re-sent each step (C5: the LBVH stays resident)
    d2h = 4 * npx + 48 * world
    e2e_value, e2e_ms = None, None
    if world > 1:
        barrier()
        pkg.set_partition(0, 1)                          # leave the multi-process partition: every rank back to a plain 1-GPU state
        dist.barrier(group=cpu_group)
    if rank == 0:
        ldr_host = torch.empty(npx, dtype=torch.int32).pin_memory()
        if world > 1:
            pkg.set_gpu_count(world)
        e2e_rays, e2e_s = 0.0, 0.0
        st = pkg.RenderStats()
        for k in range(2 + args.steps):
            t0 = time.perf_counter()
            if not args.workload.startswith("C5"):
                scene.Init()                            # scene description from host data -> device(s) (dirty => re-upload)
            rc = L.gcore_render(C.byref(cfg), width, height, units, None, C.c_void_p(ldr_host.data_ptr()), C.byref(st))
            if rc != 0:
                raise RuntimeError(pkg.last_error())
            dt = time.perf_counter() - t0
            if k >= 2:
                e2e_s += dt; e2e_rays += float(st.rays)
        e2e_value = e2e_rays / e2e_s / 1e6
        e2e_ms = 1e3 * e2e_s / args.steps
        e2e_dev_ms = [float(x) for x in pkg.device_kernel_ms(world)]
        if world > 1:
            pkg.set_gpu_count(1)
    if world > 1:
        dist.barrier(group=cpu_group)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        clocks = m["clocks"]
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        # algorithmic FP32 work of the dominant kernel (render_kernel): per launch, this rank's share
        flop = mesh_q * FLOP_QUERY_MESH + (rays - mesh_q) * FLOP_QUERY_NOMESH + bounces * FLOP_BOUNCE
        flop_per_launch = flop / args.steps / world
        kern_s = kern_ms * 1e-3 / args.steps
        achieved = flop_per_launch / kern_s / 1e12
        peak = sms * 128 * sm_mhz * 1e6 / 1e12          # one non-FMA FP32 op per lane per cycle at the clock seen under load
        traffic, traffic_note = recorded_traffic(args.workload) if world == 1 else (None, "single-GPU captures only")
        colbuf_bytes = units * 16.0 * npx / world          # (pass, pixel) colours, written once by the wavefront kernel and read once by the fold
        hbm_bytes = (2 * units + 2) * 16.0 * npx / world   # per-pass colours written + read back by the fold, accumulator read + write
        roofline = {"bound": "fp32_issue", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": traffic_note,
                    "note": "analytic scene lives in constant memory: HBM traffic is only the per-pass colour buffer, so the bound "
                            "is FP32 instruction issue (SURVEY 8d), not hbm/tensor; achieved = algorithmic flops of the reference "
                            "(564 / 140 per scene query, 200 per bounce) / kernel time; peak = SMs x 128 lanes x SM clock sampled "
                            "during the run (no measured FP32 peak in MEASURED_PEAKS.json; -fmad=false => 1 flop per lane-instruction)",
                    "kernel": "render_wave_kernel<%s> + fold_kernel" % rtype, "kernel_ms_per_launch": kern_ms / args.steps,
                    "algorithmic_flop_per_launch": flop_per_launch,
                    "hbm_view": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_GBs": hbm_bytes / kern_s / 1e9, "peak_GBs": peaks.get("hbm_gbs", 6650.0),
                                 "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"}}
        if args.workload.startswith("C5"):
            roofline = lbvh_roofline(m, args.steps, world, hbm_bytes, peaks, rtype, achieved, peak, traffic, traffic_note)
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            r1, t1 = oracle_run(width, height, rtype, aa, depth, 1, cores, workload=args.workload)          # also warms the page cache
            n_units = int(min(max(round(12.0 / max(t1, 1e-3)), 1), units))
            r2, t2 = (r1, t1) if (n_units == 1 and t1 > 8.0) else oracle_run(width, height, rtype, aa, depth, n_units, cores, workload=args.workload)
            cpu_baseline = {"value": r2 / t2 / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                            "sample": "%d of the step's %d units (%d spp) of the full %dx%d image, %.1f s%s" % (
                                n_units, units, n_units * (4 if aa else 1), width, height, t2,
                                "; CPU BVH = the oracle's binned-SAH tree, single-ray traversal (not Embree)" if args.workload.startswith("C5") else "")}
        if cpu_baseline is not None and args.workload == "C1":
            # the reference's own scheme once, to show what its global lock costs: Threads_Num tasks x the full image, one lock per pixel add
            ob = ge.load_oracle()
            t0 = time.perf_counter()
            _acc, r3 = ob.render_task_lock(oracle_scene(args.workload), ob.make_params(width, height, getattr(ob, rtype), aa=aa, max_depth=depth, passes=1, seed=SEED), units)
            t3 = time.perf_counter() - t0
            cpu_baseline["reference_tasking"] = {"value": r3 / t3 / 1e6, "unit": "Mrays/s", "tasks": units,
                                                 "note": "Threads_Num tasks each rendering the full image, one process-wide lock per pixel add (ray_tracer.adb:142-194, integrators.adb:50-52)"}
        # ------------------------------------------------------------ secondary: a short LBVH-scene block, so that a driver-run record of the
        # default workload also carries evidence for the traversal path (C5 scene at 960x540, 16 spp per step; not part of `value`)
        secondary = None
        if world == 1 and args.workload == "C4_1024" and not args.no_secondary:
            try:
                import importlib
                syn = importlib.import_module(pkg.__name__ + ".synthetic")
                w2, h2, rtype2, aa2, depth2, units2, desc2 = WORKLOADS["C5_small"]
                syn.install_c5(scene)
                rt.Threads_Num = units2; rt.Anti_Aliasing_On = bool(aa2); rt.Max_Trace_Depth = depth2
                rt.Init_Render(getattr(pkg, rtype2)); tr.Resize_Viewport(w2, h2)
                m2 = measure_device(pkg, tr, rt, args, "C5_small", 1, 0, local_rank, barrier, flush, stream, 3, 3, sample_clocks=False)
                flop2 = m2["mesh_q"] * FLOP_QUERY_MESH + (m2["rays"] - m2["mesh_q"]) * FLOP_QUERY_NOMESH + m2["bounces"] * FLOP_BOUNCE
                hbm2 = (2 * units2 + 2) * 16.0 * w2 * h2
                t2n, t2note = recorded_traffic("C5_small")
                rl2 = lbvh_roofline(m2, 3, 1, hbm2, peaks, rtype2, flop2 / 3 / (m2["kern_ms"] * 1e-3 / 3) / 1e12, peak, t2n, t2note)
                info = pkg.GeometryCore().bvh_info()
                secondary = {"workload": desc2, "ms_per_step": m2["step_ms"] / 3, "value": m2["rays"] / (m2["step_ms"] * 1e-3) / 1e6, "unit": "Mrays/s",
                             "rays_per_step": m2["rays"] / 3, "bounces_per_step": m2["bounces"] / 3, "lbvh_build_ms": info["build_ms"], "lbvh_bytes": info["bytes"],
                             "lbvh_triangles": info["ntris"], "roofline": rl2}
                scene.Init()
            except Exception as e:                       # the secondary block must never cost the headline line
                secondary = {"error": repr(e)}
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "scene": ("scene.adb Cornell box + sphere light + 64 Philox spheres + 10 002 828-triangle displaced cube-sphere (synthetic.py, SURVEY 8d C5)"
                                                     if args.workload.startswith("C5") else "scene.adb Cornell box + 2 spheres + sphere light + data/pyramid2.vsgf"),
                           "seed": SEED, "parallelism": "tiles32x32_rr%d" % world,
                           "l2": "256 MiB buffer written between timed steps (outside the per-step CUDA events); inputs are < L2",
                           "numerics": ("bit-exact mode (-fmad=false, IEEE div/sqrt), identical to the CPU oracle" if args.numerics == "exact" else
                                        "NUMERICS=fast build (FMA contraction, approximate div/sqrt, FTZ): Monte-Carlo tolerance of SURVEY 8d, not bit-exact"),
                           "value_path": "one process per GPU (torchrun), torch.distributed/NCCL gather of the accumulators" if world > 1 else "one process, one GPU",
                           "e2e_path": "ONE host process: gcore_render() with host buffers; N > 1: gcore_set_gpu_count(N), the library drives every GPU and gathers over NCCL"},
                "msamples_per_s": (npx * units * (4 if aa else 1)) / (step_ms / args.steps * 1e-3) / 1e6,
                "rays_per_step": rays / args.steps, "bounces_per_step": bounces / args.steps, "mesh_queries_per_step": mesh_q / args.steps,
                "colbuf_bytes_per_launch": colbuf_bytes, "kernel_ms_per_rank": m["kernel_ms_per_rank"], "rays_per_step_per_rank": m["rays_per_rank"], "kernel_source_sha16": kernel_source_hash(),
                "wall_s": m["wall"], "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                        "kernel_ms_per_device": e2e_dev_ms},
                "gpu_launches": int(m["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline, "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
