"""Registration of the B200 path tracer as integrator plugins of a live Mitsuba 3.

    import mitsuba as mi; mi.set_variant("llvm_ad_rgb")        # or scalar_rgb / cuda_ad_rgb
    import mitsuba3_b200.mitsuba_plugin as b200; b200.register(mi)
    scene = mi.load_dict({... "integrator": {"type": "b200_path", "max_depth": 8} ...})
    img = mi.render(scene, spp=256)                              # -> libb200pt.so

The plugin classes derive from ``mi.SamplingIntegrator`` (trampoline
src/render/python/integrator_v.cpp:59-132) and are registered with
``mi.register_integrator`` (src/render/python/scene_v.cpp:166-171), i.e. the
reference's own Python plugin route. ``render`` extracts the
scene from the live objects through public bindings only (``scene.shapes()``,
``Mesh.packed_vertices()/faces()``, ``mi.traverse``), hands POD buffers to the C
ABI and wraps the result as ``mi.TensorXf``. ``render_backward`` accumulates
into ``dr.grad`` of the attached parameters with ``dr.accum_grad``.

The host Mitsuba is never imported by this package on its own: ``register`` is
given the module. Extraction is variant-agnostic (host arrays in / out).
"""
from __future__ import annotations

import numpy as np

from . import abi
from .scene import (BsdfData, EmitterData, Scene, SensorData, ShapeData, TextureData, f32)
from .transform import _cross, _normalize, _sqnorm, Transform4f

_BSDF_CLASS = {"SmoothDiffuse": abi.BSDF_DIFFUSE, "SmoothConductor": abi.BSDF_CONDUCTOR,
               "SmoothDielectric": abi.BSDF_DIELECTRIC, "Principled": abi.BSDF_PRINCIPLED,
               "RoughConductor": abi.BSDF_CONDUCTOR, "RoughDielectric": abi.BSDF_DIELECTRIC,
               "SmoothPlastic": abi.BSDF_PLASTIC}


def _np(x, dtype=np.float32):
    return np.array(x, dtype=dtype)


def _f(x) -> float:
    """A scalar parameter as a Python float: JIT variants hand out width-1 Dr.Jit arrays where scalar variants hand out floats."""
    return float(np.array(x, dtype=np.float64).reshape(-1)[0])


# ---- state that neither mi.traverse nor to_string expose: recovered by probing the live object ------------------------
def _wrap_index(i, n, wrap):
    """dr::tex_wrap (drjit/texture_impl.h:85-104) for integer texel coordinates."""
    i = np.asarray(i, np.int64)
    if wrap == abi.WRAP_CLAMP:
        return np.clip(i, 0, n - 1)
    shift = np.where(i < 0, i + 1, i)
    div = np.trunc(shift / n).astype(np.int64)          # C++ integer division truncates
    mod = i - div * n
    mod = np.where(mod < 0, mod + n, mod)
    if wrap == abi.WRAP_MIRROR:
        mod = np.where(((div & 1) == 0) ^ (i < 0), mod, n - 1 - mod)
    return mod


def _bitmap_eval_host(data, uv, wrap, filt):
    """BitmapTexture::eval on raw float data (bitmap.cpp:496-519 + dr::Texture::eval, texture_impl.h:156-205), float64 --
    only used to tell the wrap / filter modes of a live texture apart."""
    h, w = data.shape[:2]
    x, y = uv[:, 0] * w, uv[:, 1] * h
    if filt == abi.FILTER_NEAREST:
        ix, iy = _wrap_index(np.floor(x), w, wrap), _wrap_index(np.floor(y), h, wrap)
        return data[iy, ix].astype(np.float64)
    x, y = x - 0.5, y - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    wx, wy = (x - x0)[:, None], (y - y0)[:, None]
    ix0, ix1 = _wrap_index(x0, w, wrap), _wrap_index(x0 + 1, w, wrap)
    iy0, iy1 = _wrap_index(y0, h, wrap), _wrap_index(y0 + 1, h, wrap)
    d = data.astype(np.float64)
    return (d[iy0, ix0] * (1 - wx) + d[iy0, ix1] * wx) * (1 - wy) + (d[iy1, ix0] * (1 - wx) + d[iy1, ix1] * wx) * wy


def _probe_bitmap_modes(mi, tex_obj, data, to_uv):
    """`wrap_mode` and `filter_type` of a BitmapTexture are construction-time properties that appear neither among its
    traversed parameters nor in its string form (bitmap.cpp:842-850): evaluate the live texture at positions outside the
    unit square and off the texel centres and keep the combination of modes that reproduces every value. Ambiguity
    (a constant image, say) is harmless -- every surviving combination then renders the same values at the probes --
    but no surviving combination is an error, not a guess."""
    rs = np.random.RandomState(7)
    uv = np.concatenate([rs.uniform(-1.3, 2.3, (24, 2)), rs.uniform(0.02, 0.98, (8, 2))]).astype(np.float32)
    uvt = uv.astype(np.float64) @ np.asarray(to_uv, np.float64)[:2, :2].T + np.asarray(to_uv, np.float64)[:2, 2]
    got = np.zeros((uv.shape[0], data.shape[2]))
    for k in range(uv.shape[0]):
        si = mi.SurfaceInteraction3f()
        si.uv = mi.Point2f(float(uv[k, 0]), float(uv[k, 1]))
        si.t = 1.0
        if data.shape[2] == 1:
            got[k, 0] = float(np.array(tex_obj.eval_1(si)).reshape(-1)[0])
        else:
            got[k] = np.array(tex_obj.eval(si), dtype=np.float64).reshape(-1)[:data.shape[2]]
    scale = max(1e-6, float(np.abs(data).max()))
    fits = [(wrap, filt) for wrap in (abi.WRAP_REPEAT, abi.WRAP_CLAMP, abi.WRAP_MIRROR) for filt in (abi.FILTER_BILINEAR, abi.FILTER_NEAREST)
            if np.abs(_bitmap_eval_host(data, uvt, wrap, filt) - got).max() <= 2e-4 * scale]
    if not fits:
        raise NotImplementedError("bitmap texture: its values match none of the repeat/clamp/mirror x bilinear/nearest modes "
                                  "(sRGB-encoded 8-bit storage, spectral upsampling or a cubic filter are outside the hot-path scope)")
    return fits[0]


def _probe_envmap_mis_compensation(mi, em, data, to_world):
    """`mis_compensation` (envmap.cpp:196,498-517) is not exposed either. It subtracts the mean luminance from the sampling
    density, so the RATIO of pdf_direction between two texel centres of one image row tells it apart (the row's
    sin(theta) and the normalisation cancel): lum_a / lum_b without, max(lum_a - mean, 0) / max(lum_b - mean, 0) with it."""
    d = np.asarray(data, np.float64)                           # without the halo columns
    lum = d @ np.array([0.212671, 0.715160, 0.072169])
    h, w = lum.shape
    mean, mn = lum.mean(), lum.min()
    if mean - mn <= 0.01 * mean:
        return False                                           # the emitter disables the compensation itself (envmap.cpp:515)
    best = None
    for y in range(1, h - 1):
        a, b = int(np.argmax(lum[y])), int(np.argmin(lum[y]))
        la, lb = lum[y, a], lum[y, b]
        if la <= 0 or la == lb:
            continue
        plain, comp = lb / la, max(lb - mean, 0.0) / max(la - mean, 1e-30) if la > mean else None
        if comp is None or abs(plain - comp) < 0.05:
            continue
        if best is None or abs(plain - comp) > best[0]:
            best = (abs(plain - comp), y, a, b, plain, comp)
    if best is None:
        raise NotImplementedError("envmap: cannot tell whether mis_compensation is set from this image (no row with two texels of "
                                  "sufficiently different luminance)")
    _, y, a, b, plain, comp = best
    tw = np.asarray(to_world, np.float64)[:3, :3]

    def pdf_at(x):
        # vertex (x, y) of the sampling grid (envmap.cpp:366-384,436-447,519-527: columns are texel-centred in phi, rows
        # align-corners in theta) -> direction d = (sin phi sin theta, cos theta, -cos phi sin theta)
        theta, phi = y / (h - 1) * np.pi, (x + 0.5) / w * 2 * np.pi
        dl = np.array([np.sin(phi) * np.sin(theta), np.cos(theta), -np.cos(phi) * np.sin(theta)])
        dw = tw @ dl
        it = mi.Interaction3f(); it.p = mi.Point3f(0, 0, 0); it.t = 0.0
        ds = mi.DirectionSample3f(); ds.d = mi.Vector3f(float(dw[0]), float(dw[1]), float(dw[2]))
        return float(np.array(em.pdf_direction(it, ds)).reshape(-1)[0])
    pa, pb = pdf_at(a), pdf_at(b)
    if not pa > 0:
        raise NotImplementedError("envmap: pdf_direction probe returned zero at the brightest texel of a row")
    r = pb / pa
    return abs(r - comp) < abs(r - plain)


class _Extractor:
    def __init__(self, mi, scene, sensor):
        self.mi, self.scene_mi = mi, scene
        self.out = Scene()
        self.bsdf_index = {}
        self.sensor_mi = scene.sensors()[sensor] if isinstance(sensor, int) else sensor

    # ---- textures from a traversed parameter set ---------------------------------
    def _texture(self, params, prefix, key, name, channels, default=None, differentiable=None):
        """`prefix.key.value` (constant) or `prefix.key.data` (bitmap tensor)."""
        k_val, k_data = f"{prefix}{key}.value", f"{prefix}{key}.data"
        t = TextureData(name=name, channels=channels, differentiable=differentiable)
        if k_val in params:
            v = _np(params[k_val]).reshape(-1)
            t.value = np.full(3, v[0], f32) if v.size == 1 else v[:3].astype(f32)
            if v.size == 1:
                t.channels = 1 if channels == 1 else 3
            t.name = name + ".value"
        elif k_data in params:
            d = _np(params[k_data])
            if d.ndim == 2:
                d = d[:, :, None]
            t.kind, t.data, t.channels = abi.TEX_BITMAP, np.ascontiguousarray(d), d.shape[2]
            t.name = name + ".data"
            if f"{prefix}{key}.to_uv" in params:
                t.to_uv = _np(params[f"{prefix}{key}.to_uv"].matrix).reshape(3, 3)
            # wrap_mode / filter_type: from the live texture object that owns the parameter (SceneParameters.properties)
            node = params.properties[k_data][2] if hasattr(params, "properties") else None
            if node is None or not hasattr(node, "eval"):
                raise NotImplementedError(f"bitmap texture {name}: the texture object is not reachable, its wrap / filter modes are unknown")
            t.wrap, t.filter = _probe_bitmap_modes(self.mi, node, t.data, t.to_uv if t.to_uv is not None else np.eye(3, dtype=f32))
        elif default is not None:
            t.value = np.full(3, default, f32); t.name = name + ".value"
        else:
            return -1
        self.out.textures.append(t)
        return len(self.out.textures) - 1

    # ---- BSDFs ----------------------------------------------------------------------
    def bsdf(self, b, name):
        key = b.id() if b.id() and not b.id().startswith("_unnamed") else name
        if key in self.bsdf_index:
            return self.bsdf_index[key]
        mi = self.mi
        params = mi.traverse(b)
        cls = b.class_name()
        prefix, twosided = "", False
        if cls in ("TwoSidedBRDF", "TwoSided"):
            twosided, prefix = True, "brdf_0."
            desc = str(b)
            import re
            m = re.search(r"brdf\[0\]\s*=\s*(\w+)\[", desc)      # twosided.cpp to_string: the nested plugin's class name
            if not m:
                raise NotImplementedError("twosided: cannot read the nested BRDF's class from its string form")
            cls = m.group(1)
        else:
            desc = str(b)
        if cls not in _BSDF_CLASS:
            raise NotImplementedError(f"BSDF class {cls!r} is outside the hot-path scope (SURVEY.md 8(a))")
        d = BsdfData(id=key, type=_BSDF_CLASS[cls], twosided=twosided)
        T = lambda k, ch, default=None: self._texture(params, prefix, k, f"{key}.{prefix}{k}", ch, default)
        if d.type == abi.BSDF_DIFFUSE:
            d.tex[abi.SLOT_REFLECTANCE] = T("reflectance", 3, 0.5)
        elif d.type == abi.BSDF_CONDUCTOR:
            d.tex[abi.SLOT_ETA], d.tex[abi.SLOT_K] = T("eta", 3, 0.0), T("k", 3, 1.0)
            if cls == "RoughConductor":
                d.tex[abi.SLOT_SPEC_REFL] = T("specular_reflectance", 3)
                self._microfacet(d, desc, T, abi.SLOT_ALPHA_U, abi.SLOT_ALPHA_V)
            else:
                d.tex[abi.SLOT_SPEC_REFL] = T("specular_reflectance", 3, 1.0)
        elif d.type == abi.BSDF_DIELECTRIC:
            d.eta = _f(params[f"{prefix}eta"])
            d.tex[abi.SLOT_D_SPEC_REFL] = T("specular_reflectance", 3)
            d.tex[abi.SLOT_D_SPEC_TRANS] = T("specular_transmittance", 3)
            if cls == "RoughDielectric":
                self._microfacet(d, desc, T, abi.SLOT_D_ALPHA_U, abi.SLOT_D_ALPHA_V)
        elif d.type == abi.BSDF_PLASTIC:
            import re
            d.eta = _f(params[f"{prefix}eta"])
            d.tex[abi.SLOT_PL_DIFFUSE] = T("diffuse_reflectance", 3, 0.5)
            d.tex[abi.SLOT_PL_SPEC_REFL] = T("specular_reflectance", 3)
            # derived members (plastic.cpp:196-208) as the plugin reports them (to_string, plastic.cpp:382-398)
            num = lambda k: float(re.search(k + r"\s*=\s*\[?([-+0-9.eE]+)", desc).group(1))
            d.plastic_fdr_int, d.plastic_spec_weight = num("fdr_int"), num("specular_sampling_weight")
            if int(num("nonlinear")):
                d.flags |= abi.M_NONLINEAR
        else:
            flags = 0
            slots = [("base_color", abi.SLOT_P_BASE_COLOR, 3, 0), ("roughness", abi.SLOT_P_ROUGHNESS, 1, 0),
                     ("anisotropic", abi.SLOT_P_ANISOTROPIC, 1, abi.P_HAS_ANISOTROPIC), ("metallic", abi.SLOT_P_METALLIC, 1, abi.P_HAS_METALLIC),
                     ("spec_trans", abi.SLOT_P_SPEC_TRANS, 1, abi.P_HAS_SPEC_TRANS), ("spec_tint", abi.SLOT_P_SPEC_TINT, 1, abi.P_HAS_SPEC_TINT),
                     ("sheen", abi.SLOT_P_SHEEN, 1, abi.P_HAS_SHEEN), ("sheen_tint", abi.SLOT_P_SHEEN_TINT, 1, abi.P_HAS_SHEEN_TINT),
                     ("flatness", abi.SLOT_P_FLATNESS, 1, abi.P_HAS_FLATNESS), ("clearcoat", abi.SLOT_P_CLEARCOAT, 1, abi.P_HAS_CLEARCOAT),
                     ("clearcoat_gloss", abi.SLOT_P_CLEARCOAT_GLOSS, 1, 0)]
            for k, slot, ch, flag in slots:
                d.tex[slot] = T(k, ch, 0.5 if k in ("base_color", "roughness") else 0.0)
                t = self.out.textures[d.tex[slot]]
                if flag and (t.kind == abi.TEX_BITMAP or float(t.value[0]) != 0.0):
                    flags |= flag           # m_has_* (principledhelpers.h get_flag): a zero constant == lobe off
            if f"{prefix}eta" in params:
                d.eta = _f(params[f"{prefix}eta"]); flags |= abi.P_ETA_SPECULAR
            else:
                spec = f32(_f(params[f"{prefix}specular"]))
                d.eta = float(f32(2) * (f32(1) / (f32(1) - np.sqrt(f32(0.08) * spec, dtype=f32))) - f32(1))
            d.spec_srate = _f(params[f"{prefix}main_specular_sampling_rate"])
            d.clearcoat_srate = _f(params[f"{prefix}clearcoat_sampling_rate"])
            d.diff_refl_srate = _f(params[f"{prefix}diffuse_reflectance_sampling_rate"])
            d.flags = flags
        self.out.bsdfs.append(d)
        self.bsdf_index[key] = len(self.out.bsdfs) - 1
        return self.bsdf_index[key]

    @staticmethod
    def _microfacet(d, desc, T, slot_u, slot_v):
        """alpha / alpha_u, alpha_v come from mi.traverse; the distribution type and `sample_visible` are
        only visible in the plugin's string form (roughconductor.cpp:531-550, roughdielectric.cpp:612-640)."""
        import re
        m = re.search(r"distribution\s*=\s*(\w+)", desc)
        distr = m.group(1).lower() if m else "beckmann"
        if distr not in ("beckmann", "ggx"):
            raise NotImplementedError(f"microfacet distribution {distr!r}")
        m = re.search(r"sample_visible\s*=\s*(\w+)", desc)
        if m and m.group(1).lower() in ("0", "false"):
            raise NotImplementedError("sample_visible=false is outside the hot-path scope")
        d.flags |= abi.M_ROUGH | (abi.M_GGX if distr == "ggx" else 0)
        au = T("alpha", 1)
        if au >= 0:
            d.tex[slot_u] = d.tex[slot_v] = au
        else:
            d.tex[slot_u], d.tex[slot_v] = T("alpha_u", 1), T("alpha_v", 1)      # anisotropic: its meshes pack tangent frames (shapes())

    # ---- shapes / emitters ---------------------------------------------------------------
    def shapes(self):
        mi = self.mi
        for i, s in enumerate(self.scene_mi.shapes()):
            if not s.is_mesh():
                raise NotImplementedError("only triangle meshes are on the hot path (SURVEY.md 2, row 9b)")
            sid = s.id() or f"shape_{i}"
            verts = _np(s.packed_vertices()).reshape(-1, 8)
            faces3 = _np(s.faces(), np.uint32).reshape(-1, 3)
            faces = np.concatenate([faces3, np.zeros((faces3.shape[0], 1), np.uint32)], axis=1)
            layout = (abi.LAYOUT_NORMALS if s.has_normals() else 0) | (abi.LAYOUT_TEXCOORDS if s.has_texcoords() else 0)
            if getattr(s, "packs_tangent", lambda: False)():
                # the frame slot of the packed records holds frame_encode(normal, tangent) (mesh_utils.h:74-118); the per-face
                # FaceUVFlipped bit is not exposed by the bindings: it is the sign of the face's uv determinant, with the
                # reference's own arithmetic (mesh.cpp:634-644: fmsub(duv0.x, duv1.y, duv0.y * duv1.x) < 0)
                layout |= abi.LAYOUT_TANGENTS
                uv = verts[:, 6:8]
                duv0, duv1 = uv[faces3[:, 1]] - uv[faces3[:, 0]], uv[faces3[:, 2]] - uv[faces3[:, 0]]
                c = (duv0[:, 1] * duv1[:, 0]).astype(f32)                              # rounded product
                det = duv0[:, 0].astype(np.float64) * duv1[:, 1].astype(np.float64) - c.astype(np.float64)   # exact product, one subtraction: the sign of the fused result
                faces[:, 3] = np.where(det < 0, np.uint32(0x80000000), np.uint32(0))
            sh = ShapeData(id=sid, vertices=verts, faces=faces, layout=layout, bsdf=self.bsdf(s.bsdf(), f"{sid}.bsdf"))
            if s.is_emitter():
                ep = mi.traverse(s.emitter())
                rad = self._texture(ep, "", "radiance", f"{sid}.emitter.radiance", 3)
                if rad < 0 or self.out.textures[rad].kind != abi.TEX_CONST:
                    raise NotImplementedError("only uniform area lights are on the hot path")
                # the emitter list follows Scene::emitters() (scene.cpp:45-61): its order decides which
                # emitter a sample picks (scene.cpp:248-271)
                sh.emitter = self._emitter_slot(s.emitter())
                self.out.emitters[sh.emitter] = EmitterData(
                    shape=len(self.out.shapes), radiance_tex=rad,
                    sampling_weight=_f(ep["sampling_weight"]) if "sampling_weight" in ep else 1.0)
                sp = mi.traverse(s)
                if "to_world" in sp and verts.shape[0] == 4 and faces.shape[0] == 2:
                    # Rectangle: sampled by its parameterisation (rectangle.cpp:159-172)
                    tw = Transform4f(_np(sp["to_world"].matrix).reshape(4, 4))
                    sh.sampling, sh.to_world = abi.SAMPLING_RECTANGLE, tw.matrix.copy()
                    sh.frame_n = _normalize(tw.normal([0, 0, 1])).astype(f32)
                    area = np.sqrt(_sqnorm(_cross(tw.vector([2, 0, 0]), tw.vector([0, 2, 0]))), dtype=f32)
                    sh.inv_area = float(f32(1) / area)
                else:
                    sh.sampling = abi.SAMPLING_MESH
            self.out.shapes.append(sh)

    # ---- sensor / film ---------------------------------------------------------------------
    def sensor(self):
        mi, se = self.mi, self.sensor_mi
        film = se.film()
        p = mi.traverse(se)
        size, crop, off = [int(v) for v in film.size()], [int(v) for v in film.crop_size()], [int(v) for v in film.crop_offset()]
        proj = mi.perspective_projection(film.size(), film.crop_size(), film.crop_offset(), p["x_fov"], p["near_clip"], p["far_clip"])
        rf = film.rfilter()
        if rf.is_box_filter():
            rfilter, stddev = abi.RFILTER_BOX, 0.0
        elif rf.class_name() == "GaussianFilter":
            rfilter, stddev = abi.RFILTER_GAUSSIAN, float(rf.radius()) / 4.0
        else:
            raise NotImplementedError(f"rfilter {rf.class_name()} is outside the hot-path scope")
        if film.sample_border():
            raise NotImplementedError("sample_border is outside the hot-path scope")
        sampler = se.sampler()
        if sampler.class_name() != "IndependentSampler":
            raise NotImplementedError(f"sampler {sampler.class_name()}: only `independent` (PCG32 streams, sampler.cpp:129-148) is on the hot path")
        import re
        m = re.search(r"base_seed\s*=\s*(\d+)", str(sampler))
        if not m:
            raise NotImplementedError("independent sampler: cannot read base_seed from its string form")
        self.out.sensor = SensorData(
            sample_to_camera=_np(proj.inverse().matrix).reshape(4, 4), to_world=_np(p["to_world"].matrix).reshape(4, 4),   # (JIT variants: width-1 arrays)
            near_clip=_f(p["near_clip"]), far_clip=_f(p["far_clip"]), film_size=tuple(size), crop_size=tuple(crop),
            crop_offset=tuple(off), rfilter=rfilter, rfilter_stddev=stddev, base_seed=int(m.group(1)) & 0xffffffff,
            sample_count=int(se.sampler().sample_count()), x_fov=_f(p["x_fov"]))

    # ---- emitters ----------------------------------------------------------------------------
    def _emitter_slot(self, em) -> int:
        for k, x in enumerate(self._ems):
            if x is em or x == em:
                return k
        raise RuntimeError("emitter of a shape is not in Scene::emitters()")

    def environment(self):
        """`constant` / `envmap` (constant.cpp, envmap.cpp) from their traversed parameters. The envmap
        `data` tensor carries the two periodic halo columns (envmap.cpp:155-192); they are dropped
        here, the library rebuilds them."""
        mi = self.mi
        for k, em in enumerate(self._ems):
            if not em.is_environment():
                continue
            ep = mi.traverse(em)
            eid = em.id() or f"emitter_{k}"
            if "data" in ep:
                data = _np(ep["data"]).reshape(tuple(int(v) for v in ep["data"].shape))
                if data.shape[2] != 3:
                    raise NotImplementedError("spectral environment maps are outside the hot-path scope")
                tw = Transform4f(_np(ep["to_world"].matrix).reshape(4, 4), _np(ep["to_world"].inverse_transpose).reshape(4, 4))
                t = TextureData(name=f"{eid}.data", channels=3)
                t.kind, t.data = abi.TEX_BITMAP, np.ascontiguousarray(data[:, 1:-1, :], f32)
                t.wrap, t.filter = abi.WRAP_CLAMP, abi.FILTER_BILINEAR
                self.out.textures.append(t)
                self.out.emitters[k] = EmitterData(
                    shape=-1, radiance_tex=len(self.out.textures) - 1, type=abi.EMITTER_ENVMAP,
                    sampling_weight=_f(ep["sampling_weight"]) if "sampling_weight" in ep else 1.0,
                    env_scale=float(_np(ep["scale"]).reshape(-1)[0]),
                    env_mis_compensation=_probe_envmap_mis_compensation(mi, em, t.data, tw.matrix),
                    to_world=tw.matrix.copy(), to_world_inv=np.ascontiguousarray(tw.inverse_transpose.T, f32))
            elif any(key.startswith("radiance") for key in ep.keys()):
                rad = self._texture(ep, "", "radiance", f"{eid}.radiance", 3)
                if rad < 0 or self.out.textures[rad].kind != abi.TEX_CONST:
                    raise NotImplementedError("constant emitter: expected a uniform radiance")
                self.out.emitters[k] = EmitterData(shape=-1, radiance_tex=rad, type=abi.EMITTER_CONSTANT,
                                                   sampling_weight=_f(ep["sampling_weight"]) if "sampling_weight" in ep else 1.0)
            else:
                raise NotImplementedError(f"environment emitter {em.class_name()} is outside the hot-path scope")

    def run(self) -> Scene:
        self._ems = list(self.scene_mi.emitters())
        self.out.emitters = [None] * len(self._ems)
        self.shapes()
        self.environment()
        if any(e is None for e in self.out.emitters):
            raise NotImplementedError("only area lights, `constant` and `envmap` emitters are on the hot path (SURVEY.md 8)")
        self.sensor()
        return self.out


def extract_scene(mi, scene, sensor=0) -> Scene:
    """Live ``mi.Scene`` -> host ``Scene`` (POD arrays) through public bindings only."""
    return _Extractor(mi, scene, sensor).run()


def register(mi):
    """Register ``b200_path`` and ``b200_prb`` for the CURRENT variant of ``mi``."""
    import drjit as dr
    from .integrators import PathIntegrator, PRBIntegrator, update_params

    class _Base(mi.SamplingIntegrator):
        _impl_cls = PathIntegrator

        def __init__(self, props):
            super().__init__(props)
            kw = {}
            for k in ("max_depth", "rr_depth", "hide_emitters"):
                if props.has_property(k):
                    kw[k] = props[k]
            self._impl = self._impl_cls(kw)
            self._cache = {}

        def _host_scene(self, scene, sensor):
            # the entry holds the scene object itself (mi.Scene has no weak references): its id() cannot be reused
            # by another scene while the entry lives
            key = (id(scene), sensor if isinstance(sensor, int) else id(sensor))
            ent = self._cache.get(key)
            if ent is not None and ent[1] is scene and ent[3] != self._fingerprint(ent[0], ent[2]):
                ent = None                                   # a non-texture parameter changed (vertex positions, to_world, eta, fov ...): extract again
            if ent is None or ent[1] is not scene:
                host = extract_scene(mi, scene, sensor)
                params = mi.traverse(scene)                  # the parameter map is built once per scene
                ent = (host, scene, params, self._fingerprint(host, params))
                self._cache[key] = ent
            self._params_of = ent[2]
            return ent[0]

        @staticmethod
        def _fingerprint(host, params):
            """Checksums of every traversed parameter that is NOT one of the textures `_sync_params` uploads: what they feed
            (meshes, transforms, indices of refraction, the sensor) is baked into the extracted scene, so a change means a new
            extraction (the reference rebuilds through parameters_changed, scene.cpp:517-540)."""
            tex = set(host.parameters().keys())
            out = []
            for k in params.keys():
                if k in tex:
                    continue
                try:
                    v = params[k]
                    v = getattr(v, "matrix", v)
                    a = np.asarray(v, dtype=np.float64).reshape(-1)
                    out.append((k, a.size, float(a.sum()), float(np.dot(a, a))))
                except Exception:
                    out.append((k, str(params[k])[:64]))
            return tuple(out)

        def _sync_params(self, host, scene):
            """Push the parameter values that CHANGED since the last call (optimiser steps) to the device scene."""
            params = self._params_of
            names = host.parameters()
            vals = {}
            env_tex = {host.textures[e.radiance_tex].name for e in host.emitters if e.type == abi.EMITTER_ENVMAP}
            for name, ti in names.items():
                if name in params:
                    v = np.array(params[name], np.float32)
                    if name in env_tex:      # the plugin's `data` tensor carries two halo columns (envmap.cpp:155-192)
                        v = np.ascontiguousarray(v.reshape(tuple(int(n) for n in params[name].shape))[:, 1:-1, :])
                    cur = host.textures[ti].array()
                    if v.size == cur.size and np.array_equal(v.reshape(-1), np.asarray(cur, np.float32).reshape(-1)):
                        continue             # unchanged: no upload, no envmap warp rebuild
                    vals[name] = v
            if vals:
                update_params(host, vals)

        def render(self, scene, sensor=0, seed=0, spp=0, develop=True, evaluate=True):
            host = self._host_scene(scene, sensor)
            self._sync_params(host, scene)
            img = self._impl.render(host, seed=int(seed), spp=int(spp))
            return mi.TensorXf(img)

        def to_string(self):
            return repr(self._impl).replace(type(self._impl).__name__, "B200" + type(self._impl).__name__)

    class B200Path(_Base):
        _impl_cls = PathIntegrator

    class B200PRB(_Base):
        _impl_cls = PRBIntegrator

        def render_backward(self, scene, params, grad_in, sensor=0, seed=0, spp=0):
            host = self._host_scene(scene, sensor)
            self._sync_params(host, scene)
            self._refuse_uncovered(host, params)
            grads = self._impl.render_backward(host, np.array(grad_in, np.float32), seed=int(seed), spp=int(spp))
            env_tex = {host.textures[e.radiance_tex].name for e in host.emitters if e.type == abi.EMITTER_ENVMAP}
            for name, g in grads.items():
                if name in env_tex:          # gradients of the halo columns are routed to the real ones (envmap.cpp:228-246)
                    g = np.pad(g, ((0, 0), (1, 1), (0, 0)))
                if name in params and dr.grad_enabled(params[name]):
                    dr.accum_grad(params[name], type(params[name])(g))     # opt.step() reads dr.grad (drjit/opt.py:451)

        @staticmethod
        def _refuse_uncovered(host, params):
            # a parameter the caller tracks (dr.enable_grad) whose derivative the adjoint lacks: fail, never a silent zero
            from .integrators import device_scene
            device_scene(host)          # resolves TextureData.differentiable
            names = host.parameters()
            bad = [k for k in names if k in params and dr.grad_enabled(params[k]) and not host.textures[names[k]].differentiable]
            if bad:
                raise NotImplementedError(f"b200_prb: no derivative implemented for {bad}")

        def render_forward(self, scene, params, sensor=0, seed=0, spp=0):
            # tangents = the gradients the caller attached with dr.set_grad (common.py:536-539)
            host = self._host_scene(scene, sensor)
            self._sync_params(host, scene)
            self._refuse_uncovered(host, params)
            names = host.parameters()
            env_tex = {host.textures[e.radiance_tex].name for e in host.emitters if e.type == abi.EMITTER_ENVMAP}
            tangents = {}
            for k in names:
                if k in params and host.textures[names[k]].differentiable and dr.grad_enabled(params[k]):
                    t = np.array(dr.grad(params[k]), np.float32)
                    if k in env_tex:         # the leaf's halo entries are overwritten from the real columns (envmap.cpp:228-246)
                        t = np.ascontiguousarray(t.reshape(tuple(int(n) for n in params[k].shape))[:, 1:-1, :])
                    tangents[k] = t
            return mi.TensorXf(self._impl.render_forward(host, tangents, seed=int(seed), spp=int(spp)))

    mi.register_integrator("b200_path", lambda props: B200Path(props))
    mi.register_integrator("b200_prb", lambda props: B200PRB(props))
    return B200Path, B200PRB
