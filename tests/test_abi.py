"""The C-ABI library loads and exports every symbol include/gcore.h declares (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "gcore.h")).read()
    return sorted(set(re.findall(r"GCORE_API\s+[\w\s\*]+?\b(gcore_\w+)\s*\(", text)))


def test_header_declares_the_legacy_seam():
    syms = _declared_symbols()
    for s in ("gcore_init_and_clear", "gcore_destroy", "gcore_add_mesh_3f", "gcore_instance_meshes",
              "gcore_commit_scene", "gcore_closest_hit"):          # cpp/embree_connect.cpp:51,60,69,147,196,241
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libgcore_b200.so does not export %s" % s
    assert sorted(pkg.EXPORTS) == syms


def test_no_oracle_or_cpu_fallback_linked(pkg):
    out = subprocess.run(["nm", "-D", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "ao_" not in out                                         # nothing of oracle/ is linked into the product
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(e.startswith("gcore_") for e in exported), exported  # only the C ABI is visible


def test_pod_layouts(pkg):
    assert pkg.hitcpp_dt.itemsize == 36                             # embree_connect.cpp:186-194
    assert pkg.hit_dt.itemsize == 44
    assert pkg.sphere_dt.itemsize == 20
    assert pkg.cornell_dt.itemsize == 120
    assert pkg.material_dt.itemsize == 36
    assert pkg.light_dt.itemsize == 72
    assert pkg.camera_dt.itemsize == 100
    assert C.sizeof(pkg.RenderConfig) == 28
    assert C.sizeof(pkg.RenderStats) == 56


def test_error_reporting_without_device(pkg):
    """Without a GPU a compute call must fail loudly (no fallback); with a GPU it must succeed."""
    import torch
    L = pkg.lib()
    hits = np.zeros(4, pkg.hit_dt)
    rc = L.gcore_primary_hits(2, 2, hits.ctypes.data_as(C.c_void_p))
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc == -1
        code, text = pkg.last_error()
        assert code == -1 and "no CUDA device" in text
