#!/usr/bin/env python3
"""Golden fixtures from the reference's AD variant (llvm_ad_rgb) and its own AD test references.

Run in the build container with the reference runtime importable (oracle/build_ref.sh; the LLVM backend
starts through oracle/llvm_shim):

    python oracle/run_ref.py tests/golden/gen_golden_ad.py

Writes (committed; nothing reads /root/reference at test time):
  jit_renders.npz   llvm_ad_rgb renders of the Cornell box (path + prb, box + gaussian) at given seeds. The CUDA path
                    reproduces the JIT variants' sampler streams, so these compare PER PIXEL at equal seeds.
  prb_grads.npz     `prb` render_backward gradients and render_forward images of tests/golden/ad_scenes.py
                    (diffuse / textured wall / principled, rough conductor, rough dielectric, plastic parameters),
                    seeded, from the reference's own AD -- the adjoint's reference pin.
  ad_test_refs.npz  the reference's own test images resources/data/tests/integrators/test_<config>_image_{primal,fwd}_ref.exr
                    for the in-scope configs of src/integrators/tests/test_ad_integrators.py:227-334 (read with mi.Bitmap),
                    with the thresholds of those configs.
"""
import os
import sys

import numpy as np

import mitsuba as mi

mi.set_variant("llvm_ad_rgb")
import drjit as dr

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ad_scenes


class F:
    cbox = staticmethod(mi.cornell_box)
    T = staticmethod(lambda: mi.ScalarTransform4f())
    bitmap = staticmethod(lambda arr, **kw: dict(type="bitmap", bitmap=mi.Bitmap(arr), **kw))


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrs.items()})


def gen_jit_renders():
    out = {}
    for (integ, res, rf, spp, md, seed) in [("path", 32, "box", 16, 8, 0), ("path", 32, "gaussian", 8, 8, 3), ("path", 64, "box", 4, -1, 2),
                                            ("prb", 32, "box", 16, 6, 1), ("prb", 32, "gaussian", 8, 4, 5), ("path", 48, "gaussian", 32, 8, 9)]:
        d = ad_scenes.cbox_diffuse(F, res=res, rfilter=rf, spp=spp, max_depth=md)
        d["integrator"] = {"type": integ, "max_depth": md}
        scene = mi.load_dict(d, optimize=False)
        img = scene.integrator().render(scene, seed=seed, spp=spp)
        out[f"cbox_{integ}_{res}_{rf}_spp{spp}_d{md}_seed{seed}"] = np.array(img, np.float32)
    for name in ("materials", "ptrans", "texwall"):
        d = ad_scenes.SCENES[name](F)
        scene = mi.load_dict(d, optimize=False)
        out[f"{name}_prb_primal"] = np.array(scene.integrator().render(scene, seed=ad_scenes.SEED, spp=ad_scenes.SPP), np.float32)
    save("jit_renders.npz", **out)


def gen_prb_grads():
    out = {}
    g_in = ad_scenes.grad_in_image(32)
    out["grad_in"] = g_in
    for name, build in ad_scenes.SCENES.items():
        scene = mi.load_dict(build(F), optimize=False)
        integ = scene.integrator()
        params = mi.traverse(scene)
        keys = ad_scenes.KEYS[name]
        for rk, _ in keys:
            dr.enable_grad(params[rk])
        params.update()
        integ.render_backward(scene, params, mi.TensorXf(g_in), seed=ad_scenes.SEED, spp=ad_scenes.SPP)
        for rk, ok in keys:
            g = np.array(dr.grad(params[rk]), np.float32)
            out[f"{name}|{ok}|grad"] = g
            out[f"{name}|{ok}|value"] = np.array(params[rk], np.float32)
            print(name, ok, g.reshape(-1)[:6])
        # forward mode for the first key, tangent = 1 on every entry: the idiom of test_ad_integrators.py:1338-1349
        scene = mi.load_dict(build(F), optimize=False)
        params = mi.traverse(scene)
        rk, ok = keys[0]
        theta = mi.Float(0.0)
        dr.enable_grad(theta)
        params[rk] = type(params[rk])(params[rk]) + theta
        params.update()
        dr.forward(theta, dr.ADFlag.ClearEdges)
        img = scene.integrator().render_forward(scene, params=theta, seed=ad_scenes.SEED, spp=ad_scenes.SPP)
        out[f"{name}|{ok}|fwd"] = np.array(dr.detach(img), np.float32)
    save("prb_grads.npz", **out)


def gen_ad_test_refs():
    base = "/root/reference/resources/data/tests/integrators"
    out = {}
    # name: (error_mean_threshold, error_max_threshold, error_mean_threshold_bwd, max_depth)  test_ad_integrators.py:227-334
    cfgs = {"diffuse_albedo": (0.015, 0.25, 0.0005, 2), "diffuse_albedo_g_i": (0.04, 0.4, 0.0005, 3),
            "area_light_radiance": (0.02, 0.4, 0.0005, 2), "directly_visible_area_light_radiance": (0.02, 0.2, 0.02, 2),
            "constant_emitter_radiance": None}
    for name, th in cfgs.items():
        for kind in ("primal", "fwd"):
            fn = os.path.join(base, f"test_{name}_image_{kind}_ref.exr")
            if not os.path.exists(fn):
                print("missing", fn); continue
            out[f"{name}_{kind}"] = np.array(mi.TensorXf(mi.Bitmap(fn)), np.float32).astype(np.float16 if False else np.float32)
        if th is not None:
            out[f"{name}_thresholds"] = np.array(th, np.float32)
    save("ad_test_refs.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["jit", "grads", "refs"]
    if "jit" in which:
        gen_jit_renders()
    if "grads" in which:
        gen_prb_grads()
    if "refs" in which:
        gen_ad_test_refs()
