"""Runs inside a subprocess whose environment points at the reference runtime (oracle.ref_env): the NATIVE integrator plugin
(native/b200_path_native.cpp, built into oracle/_ref/mitsuba_build/plugins by native/build_shim.sh) is instantiated by the
unmodified reference's PluginManager, asked to render with B200PT_SHIM_DUMP set, and the scene description it built through
the C++ API is compared with what the Python extractor (mitsuba3_b200/mitsuba_plugin.py) builds through the bindings.
Without a GPU the same plugin's real render must fail with the library's error, through Mitsuba's exception path."""
import os
import sys
import tempfile

import numpy as np

import mitsuba as mi

variant = sys.argv[1]
mi.set_variant(variant)
from mitsuba3_b200 import mitsuba_plugin as plug  # noqa: E402
from mitsuba3_b200 import abi  # noqa: E402


def read_dump(path):
    out, raw = {}, open(path, "rb").read()
    pos = 0
    while pos < len(raw):
        nl = raw.index(b"\n", pos)
        name, dtype, count = raw[pos:nl].decode().split()
        count = int(count)
        dt = {"f32": np.float32, "u32": np.uint32, "i32": np.int32}[dtype]
        out[name] = np.frombuffer(raw[nl + 1: nl + 1 + 4 * count], dt).copy()
        pos = nl + 1 + 4 * count
    return out


def main():
    d = mi.cornell_box()
    d["sensor"]["film"].update(width=48, height=32, crop_width=40, crop_height=24, crop_offset_x=3, crop_offset_y=5)
    d["sensor"]["sampler"]["seed"] = 9
    d["floor"]["bsdf"] = {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.3, 0.4, 0.5]}}}
    d["integrator"] = {"type": "b200_path_native", "max_depth": 5, "rr_depth": 3}
    sc = mi.load_dict(d)
    integ = sc.integrator()
    assert "B200PathNative" in str(integ) and "max_depth = 5" in str(integ)
    dump = os.path.join(tempfile.gettempdir(), f"b200pt_shim_{variant}.bin")
    os.environ["B200PT_SHIM_DUMP"] = dump
    img = np.array(mi.render(sc, spp=4))
    assert img.shape == (24, 40, 3) and not img.any()
    del os.environ["B200PT_SHIM_DUMP"]
    g = read_dump(dump)
    host = plug.extract_scene(mi, sc)

    n_tex, n_bsdf, n_shape, n_em = [int(v) for v in g["counts"]]
    assert n_shape == len(host.shapes) and n_em == len(host.emitters) and n_bsdf == len(host.bsdfs)
    bsdf_of = lambda tab, i: tab[i]
    for i, sh in enumerate(host.shapes):
        nv, nf, layout, bsdf, emitter, sampling = [int(v) for v in g[f"shape{i}"]]
        assert (nv, nf, layout, emitter, sampling) == (sh.vertices.shape[0], sh.faces.shape[0], sh.layout, sh.emitter, sh.sampling), (i, sh.id)
        assert np.array_equal(g[f"shape{i}.vertices"].reshape(-1, 8), sh.vertices), sh.id              # packed records: bit for bit
        assert np.array_equal(g[f"shape{i}.faces"].reshape(-1, 4), sh.faces), sh.id
        # same material: type, twosided and reflectance value (the two extractors number BSDFs / textures in their own order)
        nb, hb = g[f"bsdf{bsdf}"], host.bsdfs[sh.bsdf]
        assert (int(nb[0]), int(nb[1])) == (hb.type, int(hb.twosided))
        assert np.array_equal(g[f"tex{int(nb[2])}.value"], host.textures[hb.tex[abi.SLOT_REFLECTANCE]].value[:3])
        if sh.sampling == abi.SAMPLING_RECTANGLE:
            rect = g[f"shape{i}.rect"]
            assert np.allclose(rect[:16].reshape(4, 4), sh.to_world, rtol=1e-6, atol=1e-7)
            assert np.allclose(rect[16:19], sh.frame_n, atol=1e-6) and np.isclose(rect[19], sh.inv_area, rtol=1e-6)
    for k, em in enumerate(host.emitters):
        shape, rad, typ = [int(v) for v in g[f"emitter{k}"]]
        assert (shape, typ) == (em.shape, em.type)
        assert np.array_equal(g[f"tex{rad}.value"], host.textures[em.radiance_tex].value[:3])
        assert np.isclose(g[f"emitter{k}.weight"][0], em.sampling_weight)
    se = host.sensor
    assert np.allclose(g["sensor.sample_to_camera"].reshape(4, 4), se.sample_to_camera, rtol=1e-6, atol=1e-7)
    assert np.allclose(g["sensor.to_world"].reshape(4, 4), se.to_world, rtol=1e-6, atol=1e-7)
    assert np.allclose(g["sensor.clips_stddev"], [se.near_clip, se.far_clip, se.rfilter_stddev], rtol=1e-6)
    assert [int(v) for v in g["sensor.ints"]] == [*se.film_size, *se.crop_size, *se.crop_offset, se.rfilter, se.base_seed]
    # the real path: dlopen(libb200pt.so) -> b200pt_scene_create; on a box without a GPU the library's error comes back as a
    # Mitsuba exception, on a GPU box the frame matches the Python plugin's at the same seed
    os.environ["B200PT_ROOT"] = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    has_gpu = abi.load().b200pt_device_count() > 0
    if not has_gpu:
        try:
            mi.render(sc, spp=4)
            raise AssertionError("rendered without a GPU")
        except RuntimeError as e:
            assert "no CPU fallback" in str(e) or "custom differentiable operation" in str(e), e
    else:
        plug.register(mi)
        a = np.array(mi.render(sc, spp=16, seed=2))
        d2 = dict(d); d2["integrator"] = {"type": "b200_path", "max_depth": 5, "rr_depth": 3}
        b = np.array(mi.render(mi.load_dict(d2), spp=16, seed=2))
        # same description, same seeds: equal up to the order of the fp32 atomics of the gaussian splat
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-4, atol=1e-6), np.abs(a - b).max()
    print("NATIVE_SHIM_OK", variant, "gpu" if has_gpu else "cpu")


if __name__ == "__main__":
    import traceback
    try:
        main(); sys.stdout.flush(); os._exit(0)
    except BaseException:
        traceback.print_exc(file=sys.stdout); sys.stdout.flush(); os._exit(1)
