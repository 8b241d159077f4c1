"""ada/ holds what a maintainer of the reference adds to link the Ada renderer against libgcore_b200.so.  No GNAT exists in
this image, so nothing here is compiled; what CAN be checked is that the pieces are whole: the patch for ray_tracer.adb applies cleanly
to the reference's file (when /root/reference is present) and yields a body that still defines every subprogram of the unchanged spec,
every Import of the binding names a symbol the library exports with a matching argument count, and the units are bracketed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADA = os.path.join(ROOT, "ada")
REF = "/root/reference"


def _read(path):
    return open(path, "rb").read().decode("latin-1")


def test_binding_imports_exported_symbols_with_matching_arity(pkg):
    spec = _read(os.path.join(ADA, "gcore_b200.ads"))
    header = _read(os.path.join(ROOT, "include", "gcore.h"))
    imports = re.findall(r'pragma Import \(C, (\w+), "(\w+)"\)', spec)
    assert len(imports) >= 18
    L = pkg.lib()
    for ada_name, c_name in imports:
        assert ada_name == c_name and hasattr(L, c_name), c_name
        m = re.search(r"GCORE_API\s+[\w\s\*]+?\b%s\s*\(([^;]*?)\)\s*;" % c_name, header, re.S)
        assert m, "gcore.h does not declare %s" % c_name
        c_args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        d = re.search(r"(?:function|procedure)\s+%s\s*(\(([^)]*)\))?" % ada_name, spec, re.S)
        ada_args = 0
        if d and d.group(2):
            for group in d.group(2).split(";"):
                names = group.split(":")[0]
                ada_args += len([n for n in names.split(",") if n.strip()])
        assert ada_args == len(c_args), "%s: %d Ada parameters, %d C parameters" % (c_name, ada_args, len(c_args))


@pytest.mark.parametrize("unit,name", [("scene_b200.adb", "Scene"), ("gcore_b200.ads", "GCore_B200")])
def test_units_are_bracketed(unit, name):
    text = _read(os.path.join(ADA, unit))
    code = "\n".join(l.split("--")[0] for l in text.splitlines())
    assert re.search(r"package\s+(body\s+)?%s\s+is" % name, code) and re.search(r"end\s+%s\s*;" % name, code)
    assert code.count("(") == code.count(")")
    if unit.endswith(".adb"):
        # every subprogram of scene.ads has a body, and the helper the first version forgot to define is there
        for sub in ("Init_Host_Copy", "Init", "Destroy", "Find_Closest_Hit", "Material_At", "Light_At", "Camera_At"):
            assert re.search(r"(procedure|function)\s+%s\b\s*(\([^)]*\))?[^;]*?\bis\b" % sub, code, re.S), sub
            assert re.search(r"end\s+%s\s*;" % sub, code), sub


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ray_tracer.adb")) or shutil.which("patch") is None, reason="needs the reference tree and patch(1)")
def test_ray_tracer_patch_applies_to_the_reference(tmp_path):
    dst = tmp_path / "ray_tracer.adb"
    shutil.copy(os.path.join(REF, "ray_tracer.adb"), dst)
    r = subprocess.run(["patch", str(dst), os.path.join(ADA, "ray_tracer.adb.patch")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    body = _read(str(dst))
    code = "\n".join(l.split("--")[0] for l in body.splitlines())
    spec = _read(os.path.join(REF, "ray_tracer.ads"))
    for sub in re.findall(r"^\s*(?:procedure|function)\s+(\w+)", spec, re.M):
        if sub == "delete":
            continue
        assert re.search(r"(procedure|function)\s+%s\b" % sub, code), "the patched body lost %s" % sub
    assert "gcore_render_units" in code and "gcore_read_frame" in code and "gcore_resize_viewport" in code and "gcore_get_spp" in code
    assert re.search(r"package body Ray_Tracer is", code) and re.search(r"end Ray_Tracer;", code)
    assert "new Row_Major" in code                                   # the staging frame is on the heap, not on the stack
