"""How much of the C5 step is the brute-force loop over the 65 analytic spheres?  Renders the C5 scene at 1920x1080, 16 spp with
all 65 spheres and with the light sphere only."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib(); pkg.set_device(0)
syn = importlib.import_module(pkg.__name__ + ".synthetic")
sc = pkg.Scene(); sc.Init()
pos, nrm, tris = syn.install_c5(sc)
sp, box = syn.c5_analytic()
for tag, spheres in (("65 spheres", sp), ("light sphere only", sp[-1:]), ("8 spheres", sp[-8:])):
    sc.set_analytic(spheres, box)
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 4; rt.seed = 0x5EED0005
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(1920, 1080)
    st = pkg.RenderStats()
    for k in range(3):
        L.gcore_render_units(4, C.byref(st))
    print("%-18s kernel_ms %8.2f rays %.3e -> %.1f Mrays/s" % (tag, st.kernel_ms, st.rays, st.rays / st.kernel_ms / 1e3), flush=True)
