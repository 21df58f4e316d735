"""Scene definitions shared by tests/golden/gen_golden_ad.py (run against the unmodified reference,
llvm_ad_rgb) and tests/test_prb_reference.py (run against the CUDA path / the oracle).

Every builder takes a factory `F` so that the SAME dictionary is produced for both sides:
  F.cbox()                 -> mi.cornell_box() / mitsuba3_b200.cornell_box()
  F.T()                    -> mi.ScalarTransform4f() / mitsuba3_b200.Transform4f()
  F.bitmap(arr, **props)   -> {"type": "bitmap", "bitmap": mi.Bitmap(arr), ...} / {"type": "bitmap", "data": arr, ...}
  F.uv(a, b)               -> 3x3 to_uv scale

`KEYS[name]` lists the differentiated parameters as (reference key, our key) pairs.
"""
import numpy as np


def _film(d, res, rfilter, spp):
    d["sensor"]["film"]["width"] = res
    d["sensor"]["film"]["height"] = res
    d["sensor"]["film"]["rfilter"] = {"type": rfilter}
    d["sensor"]["sampler"]["sample_count"] = spp


def wall_texture(n=8):
    return (0.15 + 0.7 * np.random.default_rng(21).random((n, n, 3))).astype(np.float32)


def cbox_diffuse(F, res=32, rfilter="box", spp=16, max_depth=4):
    d = F.cbox()
    _film(d, res, rfilter, spp)
    d["integrator"] = {"type": "prb", "max_depth": max_depth}
    return d


def cbox_textured_wall(F, res=32, rfilter="box", spp=16, max_depth=4, n=8):
    """BASELINE.json configs[2] in small: the back wall's albedo is an n x n x 3 bilinear bitmap (clamp, raw)."""
    d = cbox_diffuse(F, res, rfilter, spp, max_depth)
    d["tex-wall"] = {"type": "diffuse", "reflectance": F.bitmap(wall_texture(n), raw=True, filter_type="bilinear", wrap_mode="clamp")}
    d["back"]["bsdf"] = {"type": "ref", "id": "tex-wall"}
    return d


def cbox_materials(F, res=32, rfilter="box", spp=16, max_depth=4):
    """Principled back wall, rough conductor large box, rough dielectric small box, plastic floor."""
    d = cbox_diffuse(F, res, rfilter, spp, max_depth)
    d["pr"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.7, 0.3, 0.4]}, "roughness": 0.4, "metallic": 0.3,
               "specular": 0.6, "spec_tint": 0.2, "clearcoat": 0.4, "clearcoat_gloss": 0.5, "sheen": 0.3, "sheen_tint": 0.4, "flatness": 0.2}
    d["rc"] = {"type": "roughconductor", "distribution": "ggx", "alpha": 0.25, "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]},
               "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}, "specular_reflectance": {"type": "rgb", "value": [0.9, 0.8, 0.7]}}
    d["rd"] = {"type": "roughdielectric", "distribution": "beckmann", "alpha": 0.3, "int_ior": 1.5, "ext_ior": 1.0,
               "specular_reflectance": {"type": "rgb", "value": [0.9, 0.9, 0.8]}, "specular_transmittance": {"type": "rgb", "value": [0.8, 0.9, 0.9]}}
    d["pl"] = {"type": "plastic", "diffuse_reflectance": {"type": "rgb", "value": [0.3, 0.5, 0.7]}, "nonlinear": True, "int_ior": 1.49}
    d["back"]["bsdf"] = {"type": "ref", "id": "pr"}
    d["large-box"]["bsdf"] = {"type": "ref", "id": "rc"}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "rd"}
    d["small-box"]["to_world"] = F.T().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)   # off the floor plane
    d["floor"]["bsdf"] = {"type": "ref", "id": "pl"}
    return d


def cbox_principled_trans(F, res=32, rfilter="box", spp=16, max_depth=5):
    """Transmissive principled small box (spec_trans lobe) -- derivative through refraction."""
    d = cbox_diffuse(F, res, rfilter, spp, max_depth)
    d["pglass"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.8, 0.95]}, "roughness": 0.25, "spec_trans": 0.8, "eta": 1.45}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "pglass"}
    d["small-box"]["to_world"] = F.T().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    return d


SCENES = {
    "diffuse": cbox_diffuse,
    "diffuse_gauss": lambda F: cbox_diffuse(F, rfilter="gaussian"),
    "texwall": cbox_textured_wall,
    "texwall_gauss": lambda F: cbox_textured_wall(F, rfilter="gaussian"),
    "materials": cbox_materials,
    "ptrans": cbox_principled_trans,
}

def _pairs(ref_prefix, our_prefix, names):
    return [(f"{ref_prefix}.{n}", f"{our_prefix}.{n}") for n in names]


_SAME = lambda *names: [(n, n) for n in names]
_PR = ["base_color.value", "roughness.value", "metallic.value", "spec_tint.value", "clearcoat.value", "clearcoat_gloss.value",
       "sheen.value", "sheen_tint.value", "flatness.value"]

# (reference key, our key): the reference names a referenced BSDF's parameters after the shape that uses it
KEYS = {
    "diffuse": _SAME("white.reflectance.value", "green.reflectance.value", "red.reflectance.value", "light.emitter.radiance.value"),
    "diffuse_gauss": _SAME("white.reflectance.value", "red.reflectance.value", "light.emitter.radiance.value"),
    "texwall": [("back.bsdf.reflectance.data", "tex-wall.reflectance.data")] + _SAME("white.reflectance.value"),
    "texwall_gauss": [("back.bsdf.reflectance.data", "tex-wall.reflectance.data")],
    "materials": _pairs("back.bsdf", "pr", _PR)
                 + _pairs("large-box.bsdf", "rc", ["alpha.value", "eta.value", "k.value", "specular_reflectance.value"])
                 + _pairs("small-box.bsdf", "rd", ["alpha.value", "specular_reflectance.value", "specular_transmittance.value"])
                 + [("floor.bsdf.diffuse_reflectance.value", "pl.diffuse_reflectance.value")] + _SAME("white.reflectance.value"),
    "ptrans": _pairs("small-box.bsdf", "pglass", ["base_color.value", "roughness.value", "spec_trans.value"]),
}

SEED, SPP = 7, 16


def grad_in_image(res=32):
    """Deterministic, smooth, sign-changing adjoint image (so that cancellations are exercised too)."""
    y, x = np.mgrid[0:res, 0:res].astype(np.float32) / res
    g = np.stack([np.sin(5 * x + 1) + 0.6, np.cos(4 * y) * (x + 0.2), 0.8 - x * y], axis=-1).astype(np.float32)
    return (g / (res * res)).astype(np.float32)
