import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def built():
    """Native code is built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


def cbox(res=32, rfilter="box", spp=16, max_depth=8, **film):
    import mitsuba3_b200 as mb
    d = mb.cornell_box()
    d["sensor"]["film"].update(width=res, height=res, rfilter={"type": rfilter}, **film)
    d["sensor"]["sampler"]["sample_count"] = spp
    d["integrator"]["max_depth"] = max_depth
    return d


def materials_cbox(res=32, rfilter="box", spp=16, max_depth=8):
    """Cornell box with conductor / dielectric / principled / twosided materials
    (same scene as gen_golden.py:materials)."""
    import mitsuba3_b200 as mb
    d = cbox(res, rfilter, spp, max_depth)
    d["mirror"] = {"type": "conductor", "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}}
    d["glass"] = {"type": "dielectric", "int_ior": "bk7", "ext_ior": "air"}
    d["pr"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.94, 0.271, 0.361]}, "roughness": 0.3,
               "metallic": 0.2, "specular": 0.5, "clearcoat": 0.5, "clearcoat_gloss": 0.6, "sheen": 0.3}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "glass"}
    # lifted off the floor (coincident faces = backend-specific tie in the reference, see gen_golden.py)
    d["small-box"]["to_world"] = mb.Transform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["large-box"]["bsdf"] = {"type": "ref", "id": "mirror"}
    d["back"]["bsdf"] = {"type": "ref", "id": "pr"}
    d["floor"]["bsdf"] = {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.5, 0.5, 0.5]}}}
    return d


def compare_images(img, ref, rtol=1e-3, floor=1e-2, max_bad_frac=0.005, max_mean_rel=1e-3):
    """Per-pixel comparison at equal sampler seeds.

    fp32 evaluation-order differences (libm sincos vs CUDA sincosf, 1-ulp
    divisions) perturb a path's value by ~1e-6 relative; very rarely they flip a
    discrete decision (russian roulette, lobe choice, edge hit/miss) of a single
    sample. So: all but `max_bad_frac` of the pixels must agree within `rtol`
    (relative to max(|ref|, floor)), and the whole-image relative L2 error must
    stay below `max_mean_rel`."""
    img = np.asarray(img, np.float64); ref = np.asarray(ref, np.float64)
    assert img.shape == ref.shape
    assert np.isfinite(img).all()
    err = np.abs(img - ref) / np.maximum(np.abs(ref), floor)
    bad = (err.max(axis=-1) > rtol).mean()
    l2 = np.sqrt(((img - ref) ** 2).sum() / max((ref ** 2).sum(), 1e-30))
    assert bad <= max_bad_frac, f"{bad:.5f} of the pixels differ by more than {rtol} (allowed {max_bad_frac})"
    assert l2 <= max_mean_rel, f"relative L2 error {l2:.3e} > {max_mean_rel}"
    return bad, l2


def weighted_emitter_cbox(res=32, rfilter="box", spp=16, max_depth=6):
    """multi_emitter_cbox with sampling weights 0.5 / 2 / 1 (gen_golden.py:weighted_emitters)."""
    d = multi_emitter_cbox(res, rfilter, spp, max_depth)
    d["light"]["emitter"]["sampling_weight"] = 0.5
    d["cube-light"]["emitter"]["sampling_weight"] = 2.0
    d["side-light"]["emitter"]["sampling_weight"] = 1.0
    return d


def multi_emitter_cbox(res=32, rfilter="box", spp=16, max_depth=6):
    """Cornell box with three emitters (same scene as gen_golden.py:multi_emitter)."""
    import mitsuba3_b200 as mb
    T = mb.Transform4f
    d = cbox(res, rfilter, spp, max_depth)
    d["cube-light"] = {"type": "cube", "to_world": T().translate([-0.5, 0.2, 0.3]).rotate([0, 1, 0], 30).scale(0.08),
                       "bsdf": {"type": "ref", "id": "white"},
                       "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [2.0, 8.0, 3.0]}}}
    d["side-light"] = {"type": "rectangle", "to_world": T().translate([0.98, -0.3, 0.2]).rotate([0, 1, 0], -90).scale([0.15, 0.25, 1]),
                       "bsdf": {"type": "ref", "id": "white"},
                       "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [6.0, 2.0, 9.0]}}}
    return d


def env_scene(kind="envmap", res=32, spp=16, max_depth=6, area_light=False, hide=False, integrator="path", img=None,
              T=None, bitmap=lambda a: a):
    """Floor + principled cube + mirror cube under an environment emitter (same scene as
    gen_golden.py:env_scene; the lat-long image comes from the fixture tests/golden/env.npz).
    `T` / `bitmap`: transform class and image wrapper (mi.ScalarTransform4f / mi.Bitmap for a live Mitsuba)."""
    if T is None:
        import mitsuba3_b200 as mb
        T = mb.Transform4f
    if img is None:
        img = golden("env.npz")["image"]
    d = {"type": "scene",
         "integrator": {"type": integrator, "max_depth": max_depth, "hide_emitters": hide},
         "sensor": {"type": "perspective", "fov": 45, "near_clip": 0.01, "far_clip": 100,
                    "to_world": T().look_at(origin=[2.5, 1.6, 3.2], target=[0, 0.3, 0], up=[0, 1, 0]),
                    "film": {"type": "hdrfilm", "width": res, "height": res, "rfilter": {"type": "box"}, "pixel_format": "rgb"},
                    "sampler": {"type": "independent", "sample_count": spp}},
         "grey": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.5, 0.5, 0.5]}},
         "pr": {"type": "principled", "base_color": {"type": "rgb", "value": [0.8, 0.3, 0.2]}, "roughness": 0.35, "metallic": 0.6,
                "specular": 0.5, "clearcoat": 0.3, "clearcoat_gloss": 0.7},
         "mirror": {"type": "conductor", "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}},
         "floor": {"type": "rectangle", "to_world": T().rotate([1, 0, 0], -90).scale(3.0), "bsdf": {"type": "ref", "id": "grey"}},
         "cube-a": {"type": "cube", "to_world": T().translate([-0.6, 0.4, 0.1]).rotate([0, 1, 0], 25).scale(0.4), "bsdf": {"type": "ref", "id": "pr"}},
         "cube-b": {"type": "cube", "to_world": T().translate([0.7, 0.3, -0.4]).rotate([0, 1, 0], -35).scale(0.3), "bsdf": {"type": "ref", "id": "mirror"}}}
    if area_light:
        d["lamp"] = {"type": "rectangle", "to_world": T().translate([0.0, 1.8, 0.0]).rotate([1, 0, 0], 90).scale(0.3),
                     "bsdf": {"type": "ref", "id": "grey"},
                     "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [10.0, 9.0, 8.0]}}}
    if kind == "envmap":
        d["sky"] = {"type": "envmap", "bitmap": bitmap(img), "scale": 1.5,
                    "to_world": T().rotate([0, 1, 0], 40).rotate([1, 0, 0], 10)}
    else:
        d["sky"] = {"type": "constant", "radiance": {"type": "rgb", "value": [0.9, 1.1, 1.4]}}
    return d


# rough conductor / dielectric (same specs as gen_golden.py:ROUGH_SPECS)
ROUGH_SPECS = {
    "roughconductor_beckmann": {"type": "roughconductor", "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}},
    "roughconductor_beckmann_rough": {"type": "roughconductor", "alpha": 0.5, "eta": {"type": "rgb", "value": [1.6, 0.9, 0.5]},
                                      "k": {"type": "rgb", "value": [2.9, 2.0, 1.6]}},
    "roughconductor_ggx_aniso": {"type": "roughconductor", "distribution": "ggx", "alpha_u": 0.05, "alpha_v": 0.3,
                                 "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]},
                                 "specular_reflectance": {"type": "rgb", "value": [0.9, 0.8, 0.7]}},
    "roughdielectric_beckmann": {"type": "roughdielectric", "alpha": 0.2},
    "roughdielectric_ggx_aniso_tinted": {"type": "roughdielectric", "distribution": "ggx", "alpha_u": 0.1, "alpha_v": 0.4,
                                         "int_ior": "water", "ext_ior": "air",
                                         "specular_reflectance": {"type": "rgb", "value": [0.9, 0.95, 1.0]},
                                         "specular_transmittance": {"type": "rgb", "value": [0.8, 0.9, 0.7]}},
    "roughconductor_ggx_tinted": {"type": "roughconductor", "distribution": "ggx", "alpha": 0.15,
                                  "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]},
                                  "specular_reflectance": {"type": "rgb", "value": [0.9, 0.8, 0.7]}},
    "plastic_default": {"type": "plastic"},
    "plastic_tinted_nonlinear": {"type": "plastic", "int_ior": 1.9, "diffuse_reflectance": {"type": "rgb", "value": [0.1, 0.27, 0.36]},
                                 "specular_reflectance": {"type": "rgb", "value": [0.9, 0.7, 0.8]}, "nonlinear": True},
    "twosided_roughconductor_ggx": {"type": "twosided", "bsdf": {"type": "roughconductor", "distribution": "ggx", "alpha": 0.15,
                                                                   "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}}},
}


def rough_cbox(res=32, rfilter="box", spp=16, max_depth=8):
    """Cornell box with rough conductor / rough dielectric boxes and back wall (gen_golden.py:rough_materials)."""
    import copy
    import mitsuba3_b200 as mb
    d = cbox(res, rfilter, spp, max_depth)
    d["rc"] = copy.deepcopy(ROUGH_SPECS["roughconductor_ggx_tinted"])
    d["rd"] = copy.deepcopy(ROUGH_SPECS["roughdielectric_beckmann"])
    d["rb"] = copy.deepcopy(ROUGH_SPECS["roughconductor_beckmann_rough"])
    d["small-box"]["bsdf"] = {"type": "ref", "id": "rd"}
    d["small-box"]["to_world"] = mb.Transform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["large-box"]["bsdf"] = {"type": "ref", "id": "rc"}
    d["back"]["bsdf"] = {"type": "ref", "id": "rb"}
    d["pl"] = copy.deepcopy(ROUGH_SPECS["plastic_tinted_nonlinear"])
    d["floor"]["bsdf"] = {"type": "ref", "id": "pl"}
    return d


def textured_cbox(res=32, rfilter="box", spp=16, max_depth=6):
    """Cornell box with a bitmap back wall, a checkerboard floor and a principled box with checkerboard
    roughness (gen_golden.py:textured; the bitmap comes from tests/golden/textured_renders.npz)."""
    d = cbox(res, rfilter, spp, max_depth)
    tex = golden("textured_renders.npz")["tex"]
    d["tex-wall"] = {"type": "diffuse", "reflectance": {"type": "bitmap", "data": tex, "raw": True, "filter_type": "bilinear", "wrap_mode": "clamp"}}
    d["checker-floor"] = {"type": "diffuse", "reflectance": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.2, 0.1]},
                                                             "color1": {"type": "rgb", "value": [0.1, 0.3, 0.9]}, "to_uv": [[4, 0, 0], [0, 6, 0], [0, 0, 1]]}}
    d["pr-checker"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.6, 0.6, 0.6]}, "metallic": 0.5,
                       "roughness": {"type": "checkerboard", "color0": 0.15, "color1": 0.7, "to_uv": [[3, 0, 0], [0, 3, 0], [0, 0, 1]]}}
    d["back"]["bsdf"] = {"type": "ref", "id": "tex-wall"}
    d["floor"]["bsdf"] = {"type": "ref", "id": "checker-floor"}
    d["large-box"]["bsdf"] = {"type": "ref", "id": "pr-checker"}
    return d


def smooth_mesh_scene(res=32, spp=16, max_depth=5):
    """env_scene floor + envmap + area light with a smooth-shaded UV sphere (vertex normals and texture
    coordinates from tests/golden/smooth_mesh.npz; gen_golden.py:gen_smooth)."""
    g = golden("smooth_mesh.npz")
    d = env_scene(res=res, spp=spp, max_depth=max_depth, area_light=True)
    del d["cube-a"], d["cube-b"]
    d["ball-mat"] = {"type": "principled", "roughness": 0.3, "metallic": 0.3, "clearcoat": 0.5,
                     "base_color": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.2, 0.1]},
                                    "color1": {"type": "rgb", "value": [0.1, 0.3, 0.9]}, "to_uv": [[8, 0, 0], [0, 4, 0], [0, 0, 1]]}}
    d["ball"] = {"type": "mesh", "positions": g["positions"], "normals": g["normals"], "texcoords": g["texcoords"], "faces": g["faces"],
                 "bsdf": {"type": "ref", "id": "ball-mat"}}
    return d


TANGENT_MATERIALS = {
    "aniso_principled": {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.6, 0.2]}, "roughness": 0.35,
                         "anisotropic": 0.8, "metallic": 0.9, "specular": 0.5},
    "aniso_roughconductor": {"type": "roughconductor", "distribution": "ggx", "alpha_u": 0.05, "alpha_v": 0.3,
                             "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}},
}


def tangent_mesh_scene(mat="aniso_principled", tag="", res=32, spp=16, max_depth=5):
    """env_scene floor + envmap + area light with the UV sphere of tests/golden/tangent_mesh.npz: packed vertex records whose
    frame slot holds frame_encode(normal, tangent) and faces with the FaceUVFlipped bit, as the reference's loader packs them
    for an anisotropic BSDF (gen_golden_tangent.py); `tag` "_m" = the u-mirrored copy (every face uv-flipped)."""
    import mitsuba3_b200 as mb
    g = golden("tangent_mesh.npz")
    d = env_scene(res=res, spp=spp, max_depth=max_depth, area_light=True)
    del d["cube-a"], d["cube-b"]
    d["ball-mat"] = TANGENT_MATERIALS[mat]
    d["ball"] = {"type": "mesh", "packed_vertices": g["packed_vertices" + tag], "faces": g["faces" + tag], "layout": int(g["layout"]),
                 "bsdf": {"type": "ref", "id": "ball-mat"}}
    assert int(g["layout"]) & mb.abi.LAYOUT_TANGENTS
    # Scene::emitters() of the reference build that wrote this fixture lists the environment map BEFORE the area light (the
    # order decides which emitter a sample picks, scene.cpp:248-271; the live extractor reads it off the scene, the host mirror
    # follows the order of the dictionary): the fixture records it
    order = [str(x) for x in g["emitter_order"]]
    if order.index("envmap") < order.index("area"):
        sky = d.pop("sky")
        d = {k: v for kk, vv in d.items() for k, v in (([("sky", sky)] if kk == "lamp" else []) + [(kk, vv)])}
    return d


def principled_glass_cbox(res=32, rfilter="box", spp=16, max_depth=8):
    """Cornell box with a transmissive principled box and a sheen / flatness wall (gen_golden.py:principled_glass)."""
    import mitsuba3_b200 as mb
    d = cbox(res, rfilter, spp, max_depth)
    d["pglass"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.95, 1.0]}, "roughness": 0.1,
                   "spec_trans": 0.9, "eta": 1.45}
    d["cloth"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.3, 0.5, 0.2]}, "roughness": 0.8, "sheen": 0.8,
                  "sheen_tint": 0.5, "flatness": 0.4, "spec_tint": 0.3, "specular": 0.3}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "pglass"}
    d["small-box"]["to_world"] = mb.Transform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["green-wall"]["bsdf"] = {"type": "ref", "id": "cloth"}
    return d


def bsdf_known_answers(query):
    """src/bsdfs/tests/test_diffuse.py:14-36 (test02_eval_pdf) and test_twosided.py:29-45 (test02_pdf): `query(spec, q)` evaluates
    BSDF `spec` on rows q = wi (3), wo (3), uv (2), sample1, sample2 (2) and returns value (3), pdf, ..."""
    theta = np.arange(20) / 19.0 * (np.pi / 2)
    wo = np.stack([np.sin(theta), np.zeros(20), np.cos(theta)], 1)
    q = np.concatenate([np.tile([0, 0, 1], (20, 1)), wo, np.zeros((20, 5))], 1).astype(np.float32)
    out = query({"type": "diffuse"}, q)
    assert np.allclose(out[:, 3], wo[:, 2] / np.pi, rtol=1e-5, atol=1e-7)                  # pdf = cos / pi
    assert np.allclose(out[:, 0], 0.5 * wo[:, 2] / np.pi, rtol=1e-5, atol=1e-7)            # eval = 0.5 cos / pi (default reflectance)
    q2 = np.array([[0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0]], np.float32)
    out = query({"type": "twosided", "bsdf": {"type": "diffuse"}}, q2)
    assert np.isclose(out[0, 3], 1 / np.pi, rtol=1e-6) and out[1, 3] == 0.0
