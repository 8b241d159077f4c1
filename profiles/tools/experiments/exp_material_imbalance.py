import sys, time, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib()
pkg.set_device(0)
def run(tag, mod=None):
    sc = pkg.Scene(); sc.Init()
    if mod:
        mats = np.zeros(11, pkg.material_dt); mats["light"] = -1
        def setm(i,t,c,c2=(0,0,0),p=0,l=-1): mats[i]=(t,c,c2,p,l)
        setm(0, pkg.MAT_GLASS,(.75,.75,.75),(.85,.85,.85),1.75); setm(1,pkg.MAT_LAMBERT,(.5,.5,.5)); setm(2,pkg.MAT_LAMBERT,(.25,.5,0)); setm(3,pkg.MAT_LAMBERT,(.5,0,0))
        setm(4,pkg.MAT_LIGHT,(0,0,0),l=0); setm(5,pkg.MAT_MIRROR,(.75,.75,.75)); setm(8,pkg.MAT_PHONG,(.75,.75,.75),p=80.0); mats[9]=mats[1]; mats[10]=mats[1]
        if mod == "nophong": setm(8,pkg.MAT_LAMBERT,(.75,.75,.75))
        if mod == "alllambert": setm(8,pkg.MAT_LAMBERT,(.75,.75,.75)); setm(0,pkg.MAT_LAMBERT,(.75,.75,.75))
        sc.set_materials(mats)
    rt = pkg.Ray_Tracer(); rt.Threads_Num = 64; rt.seed = 0x5EED0004
    rt.Init_Render(pkg.PT_MIS); rt.Resize_Viewport(1024, 1024)
    st = pkg.RenderStats()
    for k in range(4):
        L.gcore_render_units(64, C.byref(st))
    print(tag, "kernel_ms %.2f rays %.3e bounces %.3e -> %.1f Mrays/s" % (st.kernel_ms, st.rays, st.bounces, st.rays / st.kernel_ms / 1e3))
run("baseline"); run("same-materials-explicit", "same"); run("phong->lambert", "nophong"); run("all lambert", "alllambert")
