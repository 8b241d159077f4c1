"""Race hunt without compute-sanitizer (closed on this pool): the same renders repeated, every HDR accumulator compared bit for bit
with the first run.  A data race in the shared-memory queues of the wavefront kernel would show up as a differing pixel."""
import ctypes as C, hashlib, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib(); pkg.set_device(0)
sc = pkg.Scene(); sc.Init()
bad = 0
for rtype, w, h, units, reps in ((pkg.PT_MIS, 1024, 1024, 8, 12), (pkg.PT_SHADOW, 777, 333, 5, 12), (pkg.PT_STUPID, 512, 512, 6, 8)):
    ref = None
    for k in range(reps):
        rt = pkg.Ray_Tracer(); rt.Threads_Num = units; rt.seed = 31337
        rt.Init_Render(rtype); rt.Resize_Viewport(w, h)
        hdr = rt.Render_Pass(read_hdr=True)
        dig = hashlib.sha256(hdr.tobytes()).hexdigest()
        if ref is None: ref = dig
        elif dig != ref: bad += 1
    print("render_type %d %dx%d x%d units: %d runs, digest %s, mismatches so far %d" % (rtype, w, h, units, reps, ref[:16], bad), flush=True)
# LBVH scene (time-sliced traversal)
syn = importlib.import_module(pkg.__name__ + ".synthetic")
pos, nrm, tris = syn.cube_sphere(200, radius=0.9, center=(0.3, 1.2, 2.6))
sc.set_mesh(pos, nrm, tris, mesh_mode=pkg.MESH_NEAREST)
ref = None
for k in range(10):
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 4; rt.seed = 4242
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(800, 600)
    dig = hashlib.sha256(rt.Render_Pass(read_hdr=True).tobytes()).hexdigest()
    if ref is None: ref = dig
    elif dig != ref: bad += 1
print("LBVH scene (480 k triangles) 800x600 x4 units: 10 runs, digest %s, mismatches %d" % (ref[:16], bad), flush=True)
print("SOAK", "OK" if bad == 0 else "FAILED")
sys.exit(0 if bad == 0 else 1)
