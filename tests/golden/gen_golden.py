#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container, with the reference's Python package importable
(variant scalar_rgb; e.g. `source <reference build>/setpath.sh`):

    python tests/golden/gen_golden.py

The reference cannot travel to the GPU box, so the (small) outputs are
committed: tests/golden/*.npz. Nothing in tests/, smoke() or bench.py imports
mitsuba at run time.

What is recorded (all from mitsuba scalar_rgb, reference v3.9.1):
  rng.npz          sample_tea_32 / PCG32 known answers
  cbox_scene.npz   mi.cornell_box(): packed meshes, BSDF/emitter values, sensor matrices
  cbox_rays.npz    camera rays, Scene.ray_intersect SI records, ray_test, emitter sampling
  bsdf_tables.npz  BSDF eval/pdf/sample tables (diffuse, conductor, dielectric, principled)
  cbox_renders.npz scalar_rgb renders (single 32x32 / 64x64 block; box + gaussian), several seeds
  materials_renders.npz  same for a Cornell box with conductor / dielectric / principled boxes
  rough.npz        roughconductor / roughdielectric eval_pdf_sample tables (Beckmann + GGX) and Cornell renders
  env.npz          envmap / constant emitter tables (sample_direction, eval, pdf_direction) and renders
"""
import os
import sys

import numpy as np

import mitsuba as mi
import drjit as dr

mi.set_variant("scalar_rgb")
OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(1234)


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrs.items()})


# --------------------------------------------------------------------------- rng
def gen_rng():
    pairs = np.array([[0, 0], [0, 5], [1, 1], [123456, 7], [0xffffffff, 0xdeadbeef], [42, 2 ** 31]], np.uint64)
    tea = np.array([mi.sample_tea_32(int(a), int(b)) for a, b in pairs], np.uint64)
    tea_f = np.array([mi.sample_tea_float32(1, 1, 4), mi.sample_tea_float32(1, 2, 4), mi.sample_tea_float32(7, 9, 4)], np.float32)
    seqs = []
    inits = [(mi.PCG32().state, 0)]  # placeholder to document defaults
    cases = [(0x853c49e6748fea9b, 0xda3e39cb94b95bdb), (42, 54), (1822236360, 2351406596), (5, 0xda3e39cb94b95bdb)]
    for st, sq in cases:
        r = mi.PCG32(initstate=st, initseq=sq)
        u = [int(r.next_uint32()) for _ in range(8)]
        r = mi.PCG32(initstate=st, initseq=sq)
        f = [float(r.next_float32()) for _ in range(8)]
        seqs.append((u, f))
    # independent sampler, scalar seeding (test_independent.py:22-34)
    sampler = mi.load_dict({"type": "independent"})
    sampler.seed(7)
    samp = [float(sampler.next_1d()) for _ in range(6)]
    save("rng.npz", tea_in=pairs, tea_out=tea, tea_float=tea_f,
         pcg_cases=np.array(cases, np.uint64), pcg_u32=np.array([s[0] for s in seqs], np.uint64),
         pcg_f32=np.array([s[1] for s in seqs], np.float32), sampler_seed7=np.array(samp, np.float32))


# --------------------------------------------------------------------------- scene dump
def cbox_dict(res=32, rfilter="box", spp=16, max_depth=8, block=True, extra=None):
    d = mi.cornell_box()
    d["sensor"]["film"]["width"] = res
    d["sensor"]["film"]["height"] = res
    d["sensor"]["film"]["rfilter"] = {"type": rfilter}
    d["sensor"]["sampler"]["sample_count"] = spp
    d["integrator"] = {"type": "path", "max_depth": max_depth}
    if block:
        d["integrator"]["block_size"] = res   # one spiral block -> block_id 0 (integrator.cpp:203-215)
    if extra:
        extra(d)
    return d


def dump_shapes(scene):
    out = {}
    for i, s in enumerate(scene.shapes()):
        # unmerged load (optimize=False) keeps one mesh per shape
        v = np.array(s.packed_vertices(), np.float32).reshape(-1, 8)
        f = np.array(s.faces(), np.uint32).reshape(-1, 3)
        out[f"shape{i}_id"] = np.array(s.id())
        out[f"shape{i}_vertices"] = v
        out[f"shape{i}_faces"] = f
        out[f"shape{i}_is_emitter"] = np.array(s.is_emitter())
        out[f"shape{i}_bsdf"] = np.array(s.bsdf().id())
        out[f"shape{i}_area"] = np.array(s.surface_area(), np.float32)
    out["n_shapes"] = np.array(len(scene.shapes()))
    return out


def gen_scene():
    d = cbox_dict(res=256, rfilter="gaussian", spp=64, block=False)
    scene = mi.load_dict(d, optimize=False)
    sensor = scene.sensors()[0]
    film = sensor.film()
    params = mi.traverse(scene)
    proj = mi.perspective_projection(film.size(), film.crop_size(), film.crop_offset(),
                                     params["sensor.x_fov"], params["sensor.near_clip"], params["sensor.far_clip"])
    s2c = np.array(proj.inverse().matrix, np.float32)
    out = dump_shapes(scene)
    out.update(sample_to_camera=s2c, to_world=np.array(params["sensor.to_world"].matrix, np.float32),
               x_fov=np.array(params["sensor.x_fov"], np.float32),
               near_clip=np.array(params["sensor.near_clip"], np.float32),
               far_clip=np.array(params["sensor.far_clip"], np.float32))
    for k in ("white", "green", "red"):
        out[f"{k}_reflectance"] = np.array(params[f"{k}.reflectance.value"], np.float32)
    out["light_radiance"] = np.array(params["light.emitter.radiance.value"], np.float32)
    out["light_to_world"] = np.array(params["light.to_world"].matrix, np.float32)
    # 32x32 sensor matrix as used by the render fixtures
    for res in (32, 64):
        sc = mi.load_dict(cbox_dict(res=res), optimize=False)
        f2 = sc.sensors()[0].film()
        p2 = mi.traverse(sc)
        pr = mi.perspective_projection(f2.size(), f2.crop_size(), f2.crop_offset(), p2["sensor.x_fov"], p2["sensor.near_clip"], p2["sensor.far_clip"])
        out[f"sample_to_camera_{res}"] = np.array(pr.inverse().matrix, np.float32)
    save("cbox_scene.npz", **out)
    return scene


# --------------------------------------------------------------------------- rays / SI
def gen_rays(scene):
    sensor = scene.sensors()[0]
    n = 256
    pos = rng.random((n, 2)).astype(np.float32)
    cam = np.zeros((n, 7), np.float32)
    for i in range(n):
        ray, w = sensor.sample_ray_differential(0.0, 0.5, mi.Point2f(float(pos[i, 0]), float(pos[i, 1])), mi.Point2f(0.5, 0.5))
        cam[i] = [*ray.o, *ray.d, ray.maxt]
    # random rays inside the box (origins in [-0.9,0.9]^3, random directions), plus camera rays
    m = 768
    o = (rng.random((m, 3)) * 1.8 - 0.9).astype(np.float32)
    dd = rng.normal(size=(m, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    rays = np.concatenate([cam, np.concatenate([o, dd.astype(np.float32), np.full((m, 1), np.float32(3.4028235e38))], axis=1)], axis=0).astype(np.float32)
    rays[-64:, 6] = (rng.random(64) * 1.5).astype(np.float32)     # finite maxt -> some misses
    si_rec = np.zeros((rays.shape[0], 24), np.float32)
    shape_ids = [s.id() for s in scene.shapes()]
    occl = np.zeros(rays.shape[0], np.uint8)
    for i, r in enumerate(rays):
        ray = mi.Ray3f(mi.Point3f(*map(float, r[0:3])), mi.Vector3f(*map(float, r[3:6])), float(r[6]), 0.0, [])
        si = scene.ray_intersect(ray)
        occl[i] = bool(scene.ray_test(ray))
        if si.is_valid():
            sidx = shape_ids.index(si.shape.id())
            si_rec[i] = [si.t, *si.p, *si.n, *si.sh_frame.n, *si.sh_frame.s, *si.sh_frame.t, *si.uv, *si.wi, sidx, si.prim_index, 0]
        else:
            si_rec[i, 0] = np.inf; si_rec[i, 21] = -1
            si_rec[i, 18:21] = -r[3:6]
    # emitter sampling from reference points (Scene.sample_emitter_direction, test_visibility off/on)
    k = 256
    refp = (rng.random((k, 3)) * 1.8 - 0.9).astype(np.float32)
    smp = rng.random((k, 2)).astype(np.float32)
    em = np.zeros((k, 16), np.float32)
    for i in range(k):
        it = dr.zeros(mi.Interaction3f)
        it.p = mi.Point3f(*map(float, refp[i])); it.n = mi.Normal3f(0, 0, 0); it.t = 1.0
        ds, w = scene.sample_emitter_direction(it, mi.Point2f(float(smp[i, 0]), float(smp[i, 1])), False)
        ds_v, w_v = scene.sample_emitter_direction(it, mi.Point2f(float(smp[i, 0]), float(smp[i, 1])), True)
        pdf = scene.pdf_emitter_direction(it, ds)
        em[i] = [*ds.p, *ds.n, ds.pdf, *ds.d, ds.dist, *w, float(np.any(np.array(w_v) != 0) or ds.pdf == 0), pdf]
    save("cbox_rays.npz", cam_pos=pos, cam_rays=cam, rays=rays, si=si_rec, occluded=occl,
         em_ref=refp, em_sample=smp, em_out=em, shape_ids=np.array(shape_ids))


# --------------------------------------------------------------------------- BSDF tables
def bsdf_table(bsdf_dict, n=64, seed=7, transmissive=False, uv_range=(0.0, 1.0)):
    r = np.random.default_rng(seed)
    bsdf = mi.load_dict(bsdf_dict)
    q = np.zeros((n, 11), np.float32)
    out = np.zeros((n, 14), np.float32)
    ctx = mi.BSDFContext()
    for i in range(n):
        def hemi(flip):
            v = r.normal(size=3); v /= np.linalg.norm(v)
            v[2] = abs(v[2]) * (-1 if flip else 1)
            return v
        wi = hemi(transmissive and r.random() < 0.3)
        wo = hemi(transmissive and r.random() < 0.5)
        uv = uv_range[0] + (uv_range[1] - uv_range[0]) * r.random(2); s1 = r.random(); s2 = r.random(2)
        q[i] = [*wi, *wo, *uv, s1, *s2]
        q32 = q[i]
        si = dr.zeros(mi.SurfaceInteraction3f)
        si.wi = mi.Vector3f(*map(float, q32[0:3])); si.uv = mi.Point2f(float(q32[6]), float(q32[7]))
        si.sh_frame = mi.Frame3f(mi.Vector3f(0, 0, 1)); si.n = mi.Normal3f(0, 0, 1); si.t = 1.0
        si.p = mi.Point3f(0, 0, 0)
        wo_v = mi.Vector3f(*map(float, q32[3:6]))
        val, pdf, bs, weight = bsdf.eval_pdf_sample(ctx, si, wo_v, float(q32[8]), mi.Point2f(float(q32[9]), float(q32[10])))
        st = np.array([int(bs.sampled_type)], np.uint32).view(np.float32)[0]
        out[i] = [*val, pdf, *bs.wo, bs.pdf, bs.eta, st, *weight, bs.sampled_component]
    return q, out


def gen_bsdfs():
    specs = {
        "diffuse": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.2, 0.5, 0.8]}},
        "conductor_none": {"type": "conductor", "material": "none"},
        "conductor_rgb": {"type": "conductor", "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]},
                          "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]},
                          "specular_reflectance": {"type": "rgb", "value": [0.9, 0.8, 0.7]}},
        "dielectric_bk7": {"type": "dielectric"},
        "dielectric_water_tinted": {"type": "dielectric", "int_ior": "water", "ext_ior": "air",
                                    "specular_reflectance": {"type": "rgb", "value": [0.9, 0.95, 1.0]},
                                    "specular_transmittance": {"type": "rgb", "value": [0.8, 0.9, 0.7]}},
        "twosided_diffuse": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.3, 0.6, 0.1]}}},
        "principled_default": {"type": "principled"},
        "principled_rough_metal": {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.6, 0.2]},
                                   "metallic": 0.8, "roughness": 0.35, "specular": 0.5},
        "principled_full": {"type": "principled", "base_color": {"type": "rgb", "value": [0.7, 0.1, 0.1]},
                            "roughness": 0.15, "anisotropic": 0.5, "metallic": 0.1, "spec_trans": 0.8, "eta": 1.33,
                            "spec_tint": 0.4, "sheen": 0.9, "sheen_tint": 0.2, "flatness": 0.23,
                            "clearcoat": 0.9, "clearcoat_gloss": 0.5},
        "principled_matpreview": {"type": "principled", "base_color": {"type": "rgb", "value": [0.94, 0.271, 0.361]},
                                  "roughness": 0.3, "metallic": 0.0, "specular": 0.5},
        "diffuse_checker": {"type": "diffuse", "reflectance": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.2, 0.1]},
                                                              "color1": {"type": "rgb", "value": [0.1, 0.3, 0.9]},
                                                              "to_uv": mi.ScalarTransform3f().scale([4, 6])}},
        "principled_checker_rough": {"type": "principled", "base_color": {"type": "rgb", "value": [0.6, 0.6, 0.6]},
                                     "roughness": {"type": "checkerboard", "color0": 0.15, "color1": 0.7,
                                                   "to_uv": mi.ScalarTransform3f().scale([3, 3])}, "metallic": 0.5},
        "principled_coat_sheen": {"type": "principled", "base_color": {"type": "rgb", "value": [0.2, 0.4, 0.9]},
                                  "roughness": 0.6, "clearcoat": 1.0, "clearcoat_gloss": 0.8, "sheen": 0.5, "sheen_tint": 0.7,
                                  "spec_tint": 0.3, "flatness": 0.5},
    }
    # bitmap textures (bitmap.cpp:496-519, texture_impl.h:87-205): raw float data, every wrap / filter mode,
    # uv outside [0,1] to exercise the wrapping
    tex = (0.1 + 0.8 * np.random.default_rng(99).random((5, 7, 3))).astype(np.float32)
    out = {"bitmap_data": tex}
    for wrap in ("repeat", "mirror", "clamp"):
        for filt in ("bilinear", "nearest"):
            specs[f"diffuse_bitmap_{wrap}_{filt}"] = {"type": "diffuse", "reflectance": {
                "type": "bitmap", "bitmap": mi.Bitmap(tex), "raw": True, "wrap_mode": wrap, "filter_type": filt}}
    for name, spec in specs.items():
        trans = name.startswith("dielectric") or name in ("principled_full", "twosided_diffuse")
        q, o = bsdf_table(spec, transmissive=trans, uv_range=(-1.6, 2.6) if "bitmap" in name else (0.0, 1.0))
        out[name + "_in"] = q; out[name + "_out"] = o
    save("bsdf_tables.npz", **out)


# --------------------------------------------------------------------------- renders
def render(d, seed, spp):
    scene = mi.load_dict(d, optimize=False)
    img = mi.render(scene, seed=seed, spp=spp)
    return np.array(img, np.float32)


def gen_renders():
    out = {}
    for (res, rf, spp, md, seed) in [(32, "box", 16, 8, 0), (32, "box", 8, 8, 3), (32, "box", 32, 3, 1),
                                      (32, "gaussian", 8, 8, 0), (64, "box", 4, 8, 2), (32, "box", 64, 1, 0),
                                      (32, "box", 8, -1, 5)]:
        key = f"cbox_{res}_{rf}_spp{spp}_d{md}_seed{seed}"
        out[key] = render(cbox_dict(res=res, rfilter=rf, spp=spp, max_depth=md), seed, spp)
    # statistics anchors: high-spp mean image (multi-block, default seeds) for z-tests
    d = cbox_dict(res=64, rfilter="box", spp=2048, max_depth=8, block=False)
    out["cbox_64_box_ref2048"] = render(d, 0, 2048)
    d = cbox_dict(res=64, rfilter="gaussian", spp=1024, max_depth=8, block=False)
    out["cbox_64_gaussian_ref1024"] = render(d, 0, 1024)
    save("cbox_renders.npz", **out)


def materials(d):
    d["mirror"] = {"type": "conductor", "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}}
    d["glass"] = {"type": "dielectric", "int_ior": "bk7", "ext_ior": "air"}
    d["pr"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.94, 0.271, 0.361]}, "roughness": 0.3,
               "metallic": 0.2, "specular": 0.5, "clearcoat": 0.5, "clearcoat_gloss": 0.6, "sheen": 0.3}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "glass"}
    # lifted off the floor: the stock small box has its bottom face exactly IN the floor plane, and which
    # of two coincident surfaces a ray "hits first" is backend-specific in the reference itself
    # (kd-tree / Embree / OptiX) -- it only matters once the box is transparent
    d["small-box"]["to_world"] = mi.ScalarTransform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["large-box"]["bsdf"] = {"type": "ref", "id": "mirror"}
    d["back"]["bsdf"] = {"type": "ref", "id": "pr"}
    d["floor"]["bsdf"] = {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.5, 0.5, 0.5]}}}


def principled_glass(d):
    """Transmissive principled small box (spec_trans lobe, eta 1.45) + sheen / flatness on a wall."""
    d["pglass"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.95, 1.0]}, "roughness": 0.1,
                   "spec_trans": 0.9, "eta": 1.45}
    d["cloth"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.3, 0.5, 0.2]}, "roughness": 0.8, "sheen": 0.8,
                  "sheen_tint": 0.5, "flatness": 0.4, "spec_tint": 0.3, "specular": 0.3}
    d["small-box"]["bsdf"] = {"type": "ref", "id": "pglass"}
    d["small-box"]["to_world"] = mi.ScalarTransform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["green-wall"]["bsdf"] = {"type": "ref", "id": "cloth"}


def gen_material_renders():
    out = {}
    for (res, rf, spp, md, seed) in [(32, "box", 16, 8, 1)]:
        out[f"pglass_{res}_{rf}_spp{spp}_d{md}_seed{seed}"] = render(cbox_dict(res=res, rfilter=rf, spp=spp, max_depth=md, extra=principled_glass), seed, spp)
    for (res, rf, spp, md, seed) in [(32, "box", 16, 8, 0), (32, "box", 8, 12, 4)]:
        out[f"mat_{res}_{rf}_spp{spp}_d{md}_seed{seed}"] = render(cbox_dict(res=res, rfilter=rf, spp=spp, max_depth=md, extra=materials), seed, spp)
    d = cbox_dict(res=64, rfilter="box", spp=2048, max_depth=8, block=False, extra=materials)
    out["mat_64_box_ref2048"] = render(d, 0, 2048)
    save("materials_renders.npz", **out)


def multi_emitter(d):
    """Second and third light: a small emissive cube (mesh area sampling, mesh.cpp:1662-1712) and a
    second rectangle -> uniform emitter selection with sample reuse (scene.cpp:248-271)."""
    T = mi.ScalarTransform4f
    d["cube-light"] = {"type": "cube", "to_world": T().translate([-0.5, 0.2, 0.3]).rotate([0, 1, 0], 30).scale(0.08),
                       "bsdf": {"type": "ref", "id": "white"},
                       "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [2.0, 8.0, 3.0]}}}
    d["side-light"] = {"type": "rectangle", "to_world": T().translate([0.98, -0.3, 0.2]).rotate([0, 1, 0], -90).scale([0.15, 0.25, 1]),
                       "bsdf": {"type": "ref", "id": "white"},
                       "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [6.0, 2.0, 9.0]}}}
    return d


def weighted_emitters(d):
    """Non-uniform emitter selection: Scene::m_emitter_distr (scene.cpp:120-140,257-260,378-389)."""
    multi_emitter(d)
    d["light"]["emitter"]["sampling_weight"] = 0.5
    d["cube-light"]["emitter"]["sampling_weight"] = 2.0
    d["side-light"]["emitter"]["sampling_weight"] = 1.0
    return d


def gen_multi_emitter():
    out = {}
    for (res, spp, md, seed) in [(32, 16, 6, 0), (32, 8, 3, 2)]:
        out[f"multi_{res}_box_spp{spp}_d{md}_seed{seed}"] = render(cbox_dict(res=res, rfilter="box", spp=spp, max_depth=md, extra=multi_emitter), seed, spp)
    out["weighted_32_box_spp16_d6_seed1"] = render(cbox_dict(res=32, rfilter="box", spp=16, max_depth=6, extra=weighted_emitters), 1, 16)
    # hide_emitters: directly visible area lights are skipped (skip_area_emitters, integrator.cpp:96-123, path.cpp:177-191)
    def hide(d):
        multi_emitter(d); d["integrator"]["hide_emitters"] = True
    out["multi_hide_32_box_spp8_d4_seed6"] = render(cbox_dict(res=32, rfilter="box", spp=8, max_depth=4, extra=hide), 6, 8)
    save("multi_emitter_renders.npz", **out)


# --------------------------------------------------------------------------- rough conductor / dielectric
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


def rough_materials(d):
    # (isotropic models only: anisotropic BSDFs make the reference's meshes pack tangent frames, mesh.cpp:2417-2429)
    d["rc"] = ROUGH_SPECS["roughconductor_ggx_tinted"]
    d["rd"] = ROUGH_SPECS["roughdielectric_beckmann"]
    d["rb"] = ROUGH_SPECS["roughconductor_beckmann_rough"]
    d["small-box"]["bsdf"] = {"type": "ref", "id": "rd"}
    d["small-box"]["to_world"] = mi.ScalarTransform4f().translate([0.335, -0.65, 0.38]).rotate([0, 1, 0], -17).scale(0.3)
    d["large-box"]["bsdf"] = {"type": "ref", "id": "rc"}
    d["back"]["bsdf"] = {"type": "ref", "id": "rb"}
    d["pl"] = ROUGH_SPECS["plastic_tinted_nonlinear"]
    d["floor"]["bsdf"] = {"type": "ref", "id": "pl"}


def gen_rough():
    out = {}
    for name, spec in ROUGH_SPECS.items():
        q, o = bsdf_table(spec, n=96, seed=11, transmissive=("dielectric" in name or "twosided" in name))
        out[name + "_in"] = q; out[name + "_out"] = o
    for (res, spp, md, seed) in [(32, 16, 8, 0), (32, 8, 5, 3)]:
        out[f"rough_{res}_box_spp{spp}_d{md}_seed{seed}"] = render(cbox_dict(res=res, rfilter="box", spp=spp, max_depth=md, extra=rough_materials), seed, spp)
    save("rough.npz", **out)


# --------------------------------------------------------------------------- textured surfaces in a render
def textured(d, tex, bitmap, to_uv3):
    """uv interpolation of rectangles / cubes (mesh.cpp:2380-2392) feeding bitmap and checkerboard
    textures on diffuse and principled slots."""
    d["tex-wall"] = {"type": "diffuse", "reflectance": {"type": "bitmap", "bitmap": bitmap(tex), "raw": True,
                                                        "filter_type": "bilinear", "wrap_mode": "clamp"}}
    d["checker-floor"] = {"type": "diffuse", "reflectance": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.2, 0.1]},
                                                             "color1": {"type": "rgb", "value": [0.1, 0.3, 0.9]}, "to_uv": to_uv3(4, 6)}}
    d["pr-checker"] = {"type": "principled", "base_color": {"type": "rgb", "value": [0.6, 0.6, 0.6]}, "metallic": 0.5,
                       "roughness": {"type": "checkerboard", "color0": 0.15, "color1": 0.7, "to_uv": to_uv3(3, 3)}}
    d["back"]["bsdf"] = {"type": "ref", "id": "tex-wall"}
    d["floor"]["bsdf"] = {"type": "ref", "id": "checker-floor"}
    d["large-box"]["bsdf"] = {"type": "ref", "id": "pr-checker"}
    return d


def gen_textured():
    tex = (0.1 + 0.8 * np.random.default_rng(5).random((6, 9, 3))).astype(np.float32)
    out = {"tex": tex}
    ext = lambda d: textured(d, tex, mi.Bitmap, lambda a, b: mi.ScalarTransform3f().scale([a, b]))
    for (spp, md, seed) in [(16, 6, 0), (8, 3, 5)]:
        out[f"textured_32_box_spp{spp}_d{md}_seed{seed}"] = render(cbox_dict(res=32, rfilter="box", spp=spp, max_depth=md, extra=ext), seed, spp)
    save("textured_renders.npz", **out)


# --------------------------------------------------------------------------- environment emitters
def env_image(w=16, h=8):
    """Small synthetic lat-long sky: gradient + a bright 'sun' blob + a dim ground (float32, linear RGB)."""
    y, x = np.meshgrid((np.arange(h, dtype=np.float32) + 0.5) / h, (np.arange(w, dtype=np.float32) + 0.5) / w, indexing="ij")
    sky = np.stack([0.3 + 0.5 * (1 - y), 0.4 + 0.5 * (1 - y), 0.6 + 0.6 * (1 - y)], -1)
    ground = np.stack([0.25 * np.ones_like(y), 0.2 * np.ones_like(y), 0.15 * np.ones_like(y)], -1)
    img = np.where((y < 0.55)[..., None], sky, ground)
    sun = 40.0 * np.exp(-(((x - 0.3) / 0.07) ** 2 + ((y - 0.3) / 0.1) ** 2))
    img = img + sun[..., None] * np.array([1.0, 0.9, 0.7], np.float32)
    return np.ascontiguousarray(img, np.float32)


def env_scene(T, img, bitmap, kind="envmap", res=32, spp=16, max_depth=6, area_light=False, hide=False, integrator="path"):
    """Floor + principled cube + mirror cube under an environment emitter (+ optionally an area light
    listed BEFORE the environment: emitter order = child order, scene.cpp:45-61)."""
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


def gen_env():
    img = env_image()
    T = mi.ScalarTransform4f
    out = {"image": img}
    for kind in ("envmap", "constant"):
        scene = mi.load_dict(env_scene(T, img, mi.Bitmap, kind=kind), optimize=False)
        em = [e for e in scene.emitters() if e.is_environment()][0]
        n = 192
        u = rng.random((n, 2)).astype(np.float32)
        u[0] = [0.0, 0.0]; u[1] = [1.0, 1.0]; u[2] = [0.5, 0.0]
        pts = (rng.random((n, 3)).astype(np.float32) - 0.5) * np.array([4, 2, 4], np.float32)
        pts[3] = [40.0, 10.0, -30.0]                      # outside the bounding sphere
        rec = np.zeros((n, 16), np.float32)
        for i in range(n):
            it = mi.Interaction3f()
            it.p = mi.Point3f(*pts[i].tolist()); it.t = 1.0
            ds, w = em.sample_direction(it, mi.Point2f(*u[i].tolist()))
            si = mi.SurfaceInteraction3f()
            si.wi = -ds.d
            ev = em.eval(si)
            pdf = em.pdf_direction(it, ds)
            rec[i] = [ds.d[0], ds.d[1], ds.d[2], ds.pdf, ds.dist, ds.uv[0], ds.uv[1], w[0], w[1], w[2], ev[0], ev[1], ev[2], pdf, 0, 0]
        # eval / pdf on arbitrary directions (not produced by the warp)
        dirs = rng.normal(size=(n, 3)).astype(np.float32)
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        dirs[0] = [0, 1, 0]; dirs[1] = [0, -1, 0]; dirs[2] = [0, 0, 1]; dirs[3] = [0, 0, -1]; dirs[4] = [1, 0, 0]
        rec2 = np.zeros((n, 4), np.float32)
        for i in range(n):
            si = mi.SurfaceInteraction3f(); si.wi = mi.Vector3f(*(-dirs[i]).tolist())
            ds = mi.DirectionSample3f(); ds.d = mi.Vector3f(*dirs[i].tolist())
            ev = em.eval(si)
            rec2[i] = [ev[0], ev[1], ev[2], em.pdf_direction(mi.Interaction3f(), ds)]
        out[f"{kind}_u"] = u; out[f"{kind}_p"] = pts; out[f"{kind}_sample"] = rec
        out[f"{kind}_dirs"] = dirs; out[f"{kind}_eval"] = rec2
    for (key, kw, spp, seed) in [("env_32_spp16_d6_seed0", dict(), 16, 0),
                                 ("env_32_spp8_d3_seed1_hide", dict(max_depth=3, hide=True), 8, 1),
                                 ("env_area_32_spp16_d6_seed2", dict(area_light=True), 16, 2),
                                 ("const_32_spp16_d6_seed0", dict(kind="constant"), 16, 0),
                                 ("const_area_32_spp8_d4_seed3", dict(kind="constant", area_light=True, max_depth=4), 8, 3)]:
        d = env_scene(T, img, mi.Bitmap, spp=spp, **kw)
        d["sensor"]["film"]["rfilter"] = {"type": "box"}
        d["sensor"]["sampler"]["sample_count"] = spp
        out[key] = render(block_one(d, 32), seed, spp)
    d = env_scene(T, img, mi.Bitmap, res=64, spp=1024, area_light=True)
    out["env_area_64_ref1024"] = render(d, 0, 1024)
    save("env.npz", **out)


# --------------------------------------------------------------------------- smooth-shaded mesh (vertex normals + texcoords)
def write_ply(path, pos, nrm, uv, faces):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property float nx\nproperty float ny\nproperty float nz\nproperty float u\nproperty float v\n"
                "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(pos), len(faces)))
        for p, n, t in zip(pos, nrm, uv):
            f.write("%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n" % (*p, *n, *t))
        for tri in faces:
            f.write("3 %d %d %d\n" % tuple(tri))


def gen_smooth():
    """UV sphere with interpolated shading normals and texture coordinates (mesh.cpp:2300-2392), principled
    with a checkerboard base colour, on the env_scene floor under the envmap (+ area light)."""
    import tempfile
    n_theta, n_phi, radius, center = 12, 24, 0.6, np.array([0.0, 0.6, 0.0])
    th = (np.arange(n_theta + 1, dtype=np.float64) / n_theta) * np.pi
    ph = (np.arange(n_phi + 1, dtype=np.float64) / n_phi) * 2 * np.pi
    T_, P_ = np.meshgrid(th, ph, indexing="ij")
    nrm = np.stack([np.sin(T_) * np.cos(P_), np.cos(T_), np.sin(T_) * np.sin(P_)], -1).reshape(-1, 3)
    pos = (nrm * radius + center).astype(np.float32); nrm = nrm.astype(np.float32)
    uv = np.stack([P_ / (2 * np.pi), T_ / np.pi], -1).reshape(-1, 2).astype(np.float32)
    idx = lambda i, j: i * (n_phi + 1) + j
    faces = []
    for i in range(n_theta):
        j = np.arange(n_phi)
        a, b, c, d_ = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
        if i > 0: faces.append(np.stack([a, d_, c], -1))
        if i < n_theta - 1: faces.append(np.stack([a, c, b], -1))
    faces = np.concatenate(faces, 0).astype(np.uint32)
    ply = os.path.join(tempfile.gettempdir(), "b200pt_sphere.ply")
    write_ply(ply, pos, nrm, uv, faces)
    img = env_image()
    T = mi.ScalarTransform4f
    out = {"positions": pos, "normals": nrm, "texcoords": uv, "faces": faces}
    for (key, spp, md, seed) in [("smooth_32_spp16_d5_seed0", 16, 5, 0), ("smooth_32_spp8_d3_seed4", 8, 3, 4)]:
        d = env_scene(T, img, mi.Bitmap, spp=spp, max_depth=md, area_light=True)
        del d["cube-a"], d["cube-b"]
        d["ball-mat"] = {"type": "principled", "roughness": 0.3, "metallic": 0.3, "clearcoat": 0.5,
                         "base_color": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.2, 0.1]},
                                        "color1": {"type": "rgb", "value": [0.1, 0.3, 0.9]}, "to_uv": mi.ScalarTransform3f().scale([8, 4])}}
        d["ball"] = {"type": "ply", "filename": ply, "bsdf": {"type": "ref", "id": "ball-mat"}}
        out[key] = render(block_one(d, 32), seed, spp)
    save("smooth_mesh.npz", **out)


def block_one(d, res):
    """Single-block render (so that the per-pixel sampler streams follow integrator.cpp:209-222 for one block)."""
    d["integrator"]["block_size"] = res
    return d


if __name__ == "__main__":
    what = sys.argv[1:] or ["rng", "scene", "rays", "bsdfs", "renders", "materials", "multi", "env", "rough", "textured", "smooth"]
    if "rng" in what:
        gen_rng()
    scene = None
    if "scene" in what or "rays" in what:
        scene = gen_scene()
    if "rays" in what:
        gen_rays(scene)
    if "bsdfs" in what:
        gen_bsdfs()
    if "renders" in what:
        gen_renders()
    if "materials" in what:
        gen_material_renders()
    if "multi" in what:
        gen_multi_emitter()
    if "env" in what:
        gen_env()
    if "rough" in what:
        gen_rough()
    if "textured" in what:
        gen_textured()
    if "smooth" in what:
        gen_smooth()
