"""ctypes front-end of the CPU oracle (oracle/libpt_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by mitsuba3_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mitsuba3_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")
_lib = None


class OrcStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("bounces", C.c_uint64), ("shadow_rays", C.c_uint64)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("pt_oracle.c", "pt_oracle_principled.inc", "pt_oracle_env.inc", "pt_oracle_rough.inc", "Makefile")]
    src.append(os.path.join(_HERE, "..", "include", "b200pt.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        vp, f32p = C.c_void_p, C.POINTER(C.c_float)
        lib.orc_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(vp)]
        lib.orc_scene_destroy.argtypes = [vp]; lib.orc_scene_destroy.restype = None
        lib.orc_scene_update_texture.argtypes = [vp, C.c_uint32, f32p, C.c_size_t]
        lib.orc_render.argtypes = [vp, C.POINTER(abi.RenderParams), C.c_int, f32p, f32p, C.POINTER(OrcStats)]
        lib.orc_render_backward.argtypes = [vp, C.POINTER(abi.RenderParams), f32p]
        lib.orc_grad_zero.argtypes = [vp]
        lib.orc_render_forward.argtypes = [vp, C.POINTER(abi.RenderParams), f32p]
        lib.orc_tangent_write.argtypes = [vp, C.c_uint32, f32p, C.c_size_t]
        lib.orc_tangent_zero.argtypes = [vp]
        lib.orc_grad_read.argtypes = [vp, C.c_uint32, f32p, C.c_size_t]
        lib.orc_ray_intersect.argtypes = [vp, C.c_uint32, f32p, f32p, f32p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        lib.orc_ray_test.argtypes = [vp, C.c_uint32, f32p, C.POINTER(C.c_uint8)]
        lib.orc_surface_interaction.argtypes = [vp, C.c_uint32, f32p, f32p]
        lib.orc_bsdf_eval_pdf_sample.argtypes = [vp, C.c_uint32, C.c_uint32, f32p, f32p]
        lib.orc_camera_rays.argtypes = [vp, C.c_uint32, f32p, f32p]
        lib.orc_env_query.argtypes = [vp, C.c_uint32, f32p, f32p]
        lib.orc_microfacet_query.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint32, f32p, f32p, f32p]
        lib.orc_microfacet_query.restype = None
        lib.orc_h2d_query.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p]
        lib.orc_h2d_query.restype = None
        lib.orc_rfilter_eval.argtypes = [vp, C.c_float]; lib.orc_rfilter_eval.restype = C.c_float
        lib.orc_tea32.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
        lib.orc_pcg32_floats.argtypes = [C.c_uint64, C.c_uint64, C.c_int, f32p]
        lib.orc_pcg32_uints.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
        lib.orc_fresnel_conductor.argtypes = [C.c_float] * 3; lib.orc_fresnel_conductor.restype = C.c_float
        lib.orc_fresnel.argtypes = [C.c_float, C.c_float, f32p]
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_params(scene, spp=None, seed=0, max_depth=None, rr_depth=None, hide_emitters=None,
                shard_rank=0, shard_count=1, tile_size=32, prb=None, chunk_lanes=0):
    it = scene.integrator
    p = abi.RenderParams()
    p.seed = seed
    p.spp = scene.sensor.sample_count if not spp else spp
    p.max_depth = it["max_depth"] if max_depth is None else max_depth
    p.rr_depth = it["rr_depth"] if rr_depth is None else rr_depth
    p.hide_emitters = int(it["hide_emitters"] if hide_emitters is None else hide_emitters)
    p.shard_rank, p.shard_count, p.tile_size, p.chunk_lanes = shard_rank, shard_count, tile_size, chunk_lanes
    p.prb = int(it["type"] == "prb") if prb is None else int(prb)
    return p


class OracleScene:
    def __init__(self, scene):
        self.lib = load()
        self.scene = scene
        desc, keep = scene.build_desc()
        h = C.c_void_p()
        rc = self.lib.orc_scene_create(C.byref(desc), C.byref(h))
        if rc:
            raise RuntimeError("orc_scene_create failed")
        self.h = h
        del keep

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_scene_destroy(self.h)
            self.h = None

    def render(self, spp=None, seed=0, mode=0, return_film=False, return_stats=False, **kw):
        H, W, _ = self.scene.film_shape
        p = make_params(self.scene, spp=spp, seed=seed, **kw)
        out = np.zeros((H, W, 3), np.float32)
        film = np.zeros((H, W, 4), np.float32)
        st = OrcStats()
        self.lib.orc_render(self.h, C.byref(p), mode, _fp(out), _fp(film), C.byref(st))
        res = [out]
        if return_film:
            res.append(film)
        if return_stats:
            res.append({"samples": st.samples, "bounces": st.bounces, "shadow_rays": st.shadow_rays})
        return res[0] if len(res) == 1 else tuple(res)

    def render_backward(self, grad_in, spp=None, seed=0, **kw):
        p = make_params(self.scene, spp=spp, seed=seed, prb=1, **kw)
        g = np.ascontiguousarray(grad_in, np.float32)
        self.lib.orc_render_backward(self.h, C.byref(p), _fp(g))

    def render_forward(self, tangents, spp=None, seed=0, **kw):
        """Forward-mode derivative image for the parameter tangents {texture index: array}."""
        H, W, _ = self.scene.film_shape
        p = make_params(self.scene, spp=spp, seed=seed, prb=1, **kw)
        self.lib.orc_tangent_zero(self.h)
        for tex, v in tangents.items():
            d = np.ascontiguousarray(v, np.float32).reshape(-1)
            assert self.lib.orc_tangent_write(self.h, tex, _fp(d), d.size) == 0
        out = np.zeros((H, W, 3), np.float32)
        self.lib.orc_render_forward(self.h, C.byref(p), _fp(out))
        return out

    def grad_zero(self):
        self.lib.orc_grad_zero(self.h)

    def grad(self, tex):
        t = self.scene.textures[tex]
        out = np.zeros(t.size, np.float32)
        rc = self.lib.orc_grad_read(self.h, tex, _fp(out), out.size)
        assert rc == 0
        return out.reshape(t.array().shape)

    def update_texture(self, tex, data):
        d = np.ascontiguousarray(data, np.float32).reshape(-1)
        rc = self.lib.orc_scene_update_texture(self.h, tex, _fp(d), d.size)
        assert rc == 0

    def ray_intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        n = rays.shape[0]
        t = np.zeros(n, np.float32); uv = np.zeros((n, 2), np.float32)
        prim = np.zeros(n, np.uint32); shape = np.zeros(n, np.int32)
        self.lib.orc_ray_intersect(self.h, n, _fp(rays), _fp(t), _fp(uv),
                                   prim.ctypes.data_as(C.POINTER(C.c_uint32)), shape.ctypes.data_as(C.POINTER(C.c_int32)))
        return t, uv, prim, shape

    def ray_test(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        hit = np.zeros(rays.shape[0], np.uint8)
        self.lib.orc_ray_test(self.h, rays.shape[0], _fp(rays), hit.ctypes.data_as(C.POINTER(C.c_uint8)))
        return hit.astype(bool)

    def surface_interaction(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        out = np.zeros((rays.shape[0], 24), np.float32)
        self.lib.orc_surface_interaction(self.h, rays.shape[0], _fp(rays), _fp(out))
        return out

    def bsdf_eval_pdf_sample(self, bsdf, q):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 11)
        out = np.zeros((q.shape[0], 14), np.float32)
        rc = self.lib.orc_bsdf_eval_pdf_sample(self.h, bsdf, q.shape[0], _fp(q), _fp(out))
        assert rc == 0
        return out

    def env_query(self, q):
        """q: (n, 8) = ref point, sample, direction -> (n, 20), see orc_env_query."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 8)
        out = np.zeros((q.shape[0], 20), np.float32)
        rc = self.lib.orc_env_query(self.h, q.shape[0], _fp(q), _fp(out))
        assert rc == 0
        return out

    def camera_rays(self, pos):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        out = np.zeros((pos.shape[0], 7), np.float32)
        self.lib.orc_camera_rays(self.h, pos.shape[0], _fp(pos), _fp(out))
        return out


def microfacet_query(is_ggx, alpha_u, alpha_v, v, m=(0.0, 0.0, 1.0)):
    """(D(v), smith_g1(v, m)) of MicrofacetDistribution(type, alpha_u, alpha_v) for directions v (n, 3)."""
    v = np.ascontiguousarray(v, np.float32).reshape(-1, 3); mm = np.asarray(m, np.float32)
    out = np.zeros((v.shape[0], 2), np.float32)
    load().orc_microfacet_query(int(is_ggx), float(alpha_u), float(alpha_v), v.shape[0], _fp(v), _fp(mm), _fp(out))
    return out[:, 0], out[:, 1]


def h2d_query(data, samples):
    """Hierarchical2D(data): rows = (sampled x, sampled y, pdf, eval at the input position)."""
    d = np.ascontiguousarray(data, np.float32); q = np.ascontiguousarray(samples, np.float32).reshape(-1, 2)
    out = np.zeros((q.shape[0], 4), np.float32)
    load().orc_h2d_query(_fp(d), d.shape[1], d.shape[0], q.shape[0], _fp(q), _fp(out))
    return out


def tea32(v0, v1, rounds=4):
    out = (C.c_uint32 * 2)()
    load().orc_tea32(v0, v1, rounds, out)
    return int(out[0]), int(out[1])


def pcg32_floats(initstate, initseq, n):
    out = np.zeros(n, np.float32)
    load().orc_pcg32_floats(initstate, initseq, n, _fp(out))
    return out


def pcg32_uints(initstate, initseq, n):
    out = np.zeros(n, np.uint32)
    load().orc_pcg32_uints(initstate, initseq, n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out
