"""Bring-up helper: runs one small render per library variant (B200PT_LIB) in a subprocess and reports which ones crash."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import faulthandler, sys; faulthandler.enable()
sys.path.insert(0, %r)
import numpy as np, mitsuba3_b200 as mb
d = mb.cornell_box(); d["sensor"]["film"].update(width=32, height=32, rfilter={"type": "box"})
sc = mb.load_dict(d)
print("scene ok", flush=True)
img = mb.render(sc, spp=4, seed=0)
print("render ok", float(img.mean()), flush=True)
from mitsuba3_b200.integrators import PRBIntegrator
g = PRBIntegrator(max_depth=4).render_backward(sc, np.full(sc.film_shape, 1e-3, np.float32), seed=1, spp=4)
print("backward ok", {k: float(np.abs(v).sum()) for k, v in g.items()}, flush=True)
''' % ROOT
vd = os.path.join(ROOT, "mitsuba3_b200", "lib", "variants")
for name in sorted(os.listdir(vd)):
    for env_extra in ({}, {"B200PT_TRACE_COOP": "0"}):
        env = dict(os.environ, B200PT_LIB=os.path.join(vd, name), **env_extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("=====", name, env_extra, "rc", r.returncode); print(r.stdout[-400:]); print(r.stderr[-1200:])
