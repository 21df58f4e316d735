"""Small mixed workload for compute-sanitizer (not a pytest module):

    compute-sanitizer --tool memcheck|racecheck|initcheck python tests/sanitizer_smoke.py

Covers primal / PRB backward / PRB forward over diffuse, conductor, dielectric, principled, rough,
plastic, envmap + area light, constant environment with hide_emitters, and the global-memory BVH path
(heightfield). Round 1: 0 errors / 0 hazards with all three tools (profiles/r01_sanitizer.md)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, mitsuba3_b200 as mb
from conftest import cbox, env_scene, rough_cbox, materials_cbox
from mitsuba3_b200.integrators import PRBIntegrator, PathIntegrator
for name, d in [("cbox", cbox(res=24, spp=4, max_depth=5)), ("env", env_scene(res=24, spp=4, area_light=True)),
                ("rough", rough_cbox(res=24, spp=4, max_depth=5)), ("mat", materials_cbox(res=24, spp=4, max_depth=5)),
                ("const", env_scene(kind="constant", res=24, spp=4, hide=True))]:
    sc = mb.load_dict(d)
    img = mb.render(sc, spp=4, seed=1)
    gi = np.full(sc.film_shape, 1e-2, np.float32)
    g = PRBIntegrator(max_depth=4).render_backward(sc, gi, seed=1, spp=4)
    k = [k for k in g if "value" in k][0]
    f = PRBIntegrator(max_depth=4).render_forward(sc, {k: np.ones_like(g[k])}, seed=1, spp=4)
    print(name, float(img.mean()), float(f.mean()), len(g))
hf = mb.cornell_box_heightfield(24); hf["sensor"]["film"].update(width=32, height=32)
sc = mb.load_dict(hf); print("hf", float(mb.render(sc, spp=2, seed=0).mean()))
print("SANITY_DONE")
