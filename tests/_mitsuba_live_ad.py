"""Runs inside a subprocess whose environment points at the reference runtime (oracle.ref_env), variant llvm_ad_rgb:
the optimisation loop of a Mitsuba user -- mi.render(scene, params) with the REGISTERED b200_prb integrator, dr.backward
of a loss, gradients arriving in dr.grad(params[key]) through RBIntegrator-style render_backward + dr.accum_grad, one
mi.ad.Adam step -- against the same loop with the reference's own `prb` integrator on the CPU at equal seeds."""
import sys

import numpy as np

import mitsuba as mi

mi.set_variant("llvm_ad_rgb")
import drjit as dr
from mitsuba3_b200 import mitsuba_plugin as plug


def cbox(integrator, res=32, spp=16):
    d = mi.cornell_box()
    d["sensor"]["film"].update(width=res, height=res, rfilter={"type": "box"})
    d["sensor"]["sampler"]["sample_count"] = spp
    d["integrator"] = integrator
    return d


def grads(integ, keys, spp=16):
    scene = mi.load_dict(cbox({"type": integ, "max_depth": 4}))
    params = mi.traverse(scene)
    for k in keys:
        dr.enable_grad(params[k])
    params.update()
    img = mi.render(scene, params, spp=spp, seed=3)
    w = mi.TensorXf(np.linspace(0.5, 1.5, 32 * 32 * 3, dtype=np.float32).reshape(32, 32, 3))
    loss = dr.mean(img * w)
    dr.backward(loss)
    return scene, params, np.array(img), {k: np.array(dr.grad(params[k])).reshape(-1) for k in keys}


def main():
    plug.register(mi)
    keys = ["red.reflectance.value", "white.reflectance.value", "light.emitter.radiance.value"]
    _, _, img_ref, g_ref = grads("prb", keys)
    scene, params, img, g = grads("b200_prb", keys)
    rel = np.abs(img - img_ref) / np.maximum(np.abs(img_ref), 1e-2)
    assert (rel.max(axis=2) > 1e-3).mean() < 0.01, rel.max()
    for k in keys:
        err = np.abs(g[k] - g_ref[k]).max() / np.abs(g_ref[k]).max()
        print(k, g[k], g_ref[k], err)
        assert err < 5e-3, (k, g[k], g_ref[k])
    # one optimiser step driven by those gradients changes what the plugin renders next
    opt = mi.ad.Adam(lr=0.05)
    key = keys[0]
    opt[key] = params[key]
    params.update(opt)
    img0 = mi.render(scene, params, spp=16, seed=5)
    dr.backward(dr.mean(img0))
    opt.step()
    params.update(opt)
    img1 = np.array(mi.render(scene, params, spp=16, seed=5))
    assert np.abs(img1 - np.array(img0)).max() > 1e-4
    maps = open("/proc/self/maps").read()
    assert "libb200pt.so" in maps and "libmitsuba.so" in maps
    print("LIVE_AD_OK")


if __name__ == "__main__":
    import os, traceback
    try:
        main()
        sys.stdout.flush()
        os._exit(0)
    except BaseException:
        traceback.print_exc(file=sys.stdout)
        sys.stdout.flush()
        os._exit(1)
