"""The NUMERICS=fast experiment build (csrc/Makefile: FMA contraction, approximate division / square root, FTZ) measured against the
bit-exact product build -- VERDICT r01 asked for a measured answer on a tolerance build.  SURVEY 8d's acceptance for shaded images:
image-mean relative error < 1 % per channel, RMSE of 8x8-block means < 3 % of the image mean, >= 99.9 % of the pixels within 5 sigma.

MEASURED (B200, 1024 / 4096 spp): the fast build is 18 % faster (78.4 -> 63.2 ms per step) and does NOT meet that acceptance: image
means move by up to 1.6 % (blue channel of C3) and 8x8-block means by 5.4 % (C4, PT_MIS) / 48 % (C3, direct light only) RMS of the image mean -- a systematic shift, not noise.  Cause: Compute_Shadow
compares the hit on the emitter sphere with maxDist * (1 - 1e-6) (ray_tracer.adb:119-128); both come out of different binary32
expression chains and agree to a few ulp, so the share of light samples that count as "shadowed by the emitter itself" (5-30 %,
DESIGN section 2) is a property of the rounding mode.  FMA contraction changes it.  The fast build therefore stays an experiment
(`make fast`, bench.py --numerics fast), it is not a product mode; the product library is the exact build.  This test pins the measured
deviation (bounds a little above the measurements) so that the numbers in DESIGN.md stay true, and checks what must hold in any
numerics mode: primary hit records (ids, hit / miss) identical, images finite and close."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.select_numerics("fast"); pkg.set_device(0)
sc = pkg.Scene(); sc.Init()
out = {}
for name, rtype, depth, w, h, batches, units in %(cases)r:
    rt = pkg.Ray_Tracer(); rt.Threads_Num = units; rt.Max_Trace_Depth = depth; rt.seed = 0x5EED0003
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    frames = []
    for k in range(batches):
        frames.append(rt.Render_Pass(read_hdr=True)[..., :3].copy())
    out[name] = np.stack(frames)
    out[name + "_hits"] = sc.primary_hits(w, h)
np.savez(%(out)r, **out)
"""

# (name, integrator, depth, width, height, batches, DoPass units per batch): C3 = 1024 spp, C4 = 4096 spp -- the block-RMSE criterion is one
# for converged images (two 128-spp renders whose paths part at the first flipped shadow decision differ by ~9 % per 8x8 block)
CASES = [("C3", "PT_SHADOW", 1, 192, 192, 32, 8), ("C4", "PT_MIS", 8, 160, 160, 32, 32)]


@pytest.fixture(scope="module")
def fast_frames(tmp_path_factory):
    import __graft_entry__ as ge
    fast = os.path.join(ge.PKG_DIR, "libgcore_b200_fast.so")
    if not os.path.exists(fast):
        subprocess.run(["make", "-s", "-C", os.path.join(ge.PKG_DIR, "csrc"), "fast"], check=True)
    out = str(tmp_path_factory.mktemp("fast") / "fast.npz")
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "cases": CASES, "out": out}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out)


@pytest.mark.parametrize("name,rtype,depth,w,h,batches,units", CASES)
def test_fast_build_measured_deviation(gscene, pkg, fast_frames, name, rtype, depth, w, h, batches, units):
    rt = pkg.Ray_Tracer(); rt.Threads_Num = units; rt.Max_Trace_Depth = depth; rt.seed = 0x5EED0003
    rt.Init_Render(getattr(pkg, rtype)); rt.Resize_Viewport(w, h)
    acc = np.stack([rt.Render_Pass(read_hdr=True)[..., :3].copy() for _ in range(batches)]).astype(np.float64)
    fast = fast_frames[name].astype(np.float64)
    spp = 4 * units * batches
    exact_img, fast_img = acc[-1] / spp, fast[-1] / spp
    # (i) image mean per channel
    me, mf = exact_img.reshape(-1, 3).mean(0), fast_img.reshape(-1, 3).mean(0)
    # (measured: 0.5 / 0.8 / 1.6 % on C3 -- SURVEY asks 1 %: not met in the blue channel, see the module docstring)
    assert np.all(np.abs(mf - me) <= 0.025 * np.maximum(me, 1e-6)), (me, mf)
    # (ii) 8x8 block means
    be = exact_img.reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3)); bf = fast_img.reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    block_rmse = np.sqrt(((bf - be) ** 2).mean()) / exact_img.mean()
    # SURVEY asks 3 %: not met -- 5.4 % on C4 (PT_MIS: BSDF-sampled light hits compensate part of the changed shadowing), 48 % on C3
    # (direct light only, a dark image: whole regions change their share of emitter-self-shadowed samples)
    assert block_rmse < (0.08 if name == "C4" else 0.60), block_rmse
    # (iii) per pixel within 5 sigma of the exact estimate (sigma of the mean from the per-pass colours of the exact render)
    per_batch = np.diff(np.concatenate([np.zeros_like(acc[:1]), acc]), axis=0) / (4.0 * units)
    sigma = per_batch.std(0, ddof=1) / np.sqrt(batches)
    ok = np.abs(fast_img - exact_img) <= 5.0 * np.sqrt(2.0) * sigma + 1e-4 * np.abs(exact_img) + 1e-6     # both images carry the noise
    assert ok.all(-1).mean() >= (0.98 if name == "C4" else 0.5), ok.all(-1).mean()      # SURVEY asks 99.9 %: not met on C3 (regions of changed shadowing)
    print("fast vs exact %s: mean rel %s  block rmse %.4f  within 5 sigma %.5f" % (name, np.abs(mf - me) / np.maximum(me, 1e-6), block_rmse, ok.all(-1).mean()))
    # the same Philox streams: a path that meets no flipped decision differs by rounding only -- the first pass of most pixels does
    rel = np.abs(fast[0] - acc[0]) / np.maximum(np.abs(acc[0]), 1e-3)
    assert np.median(rel) < 1e-3, np.median(rel)
    # primary rays: hit / miss, primitive type and ids identical (documented exception: exact ties)
    g = gscene.primary_hits(w, h); f = fast_frames[name + "_hits"]
    for fld in ("is_hit", "prim_type", "prim_index", "matId", "mat"):
        assert (g[fld] != f[fld]).mean() < 1e-4, fld
