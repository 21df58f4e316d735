"""Host-side mirror of the reference's integrator interface for the hot path.

``PathIntegrator`` / ``PRBIntegrator`` expose the virtuals of
``mi.SamplingIntegrator`` / ``mi.ad.integrators.common.RBIntegrator`` that
``mi.render()`` drives (src/render/python/integrator_v.cpp:59-132,178-283;
src/python/python/util.py:344-395,396-560):

    render(scene, sensor=0, seed=0, spp=0, develop=True, evaluate=True) -> (H, W, 3)
    render_backward(scene, params, grad_in, sensor=0, seed=0, spp=0)   -> None

All compute goes through the C ABI of ``libb200pt.so`` (include/b200pt.h); this
module only marshals host buffers. Property names and their validation follow
integrator.cpp:26-29,130-146,539-550 (``max_depth``, ``rr_depth``,
``hide_emitters``) and common.py:31 (PRB default ``max_depth = 6``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .scene import Scene


class DeviceScene:
    """Owns a ``b200pt_scene`` handle (device-resident geometry, BVH, textures)."""

    def __init__(self, scene: Scene, device: int = 0):
        self.lib = abi.load()
        self.scene = scene
        self.device = device
        desc, keep = scene.build_desc()
        h = C.c_void_p()
        abi.check(self.lib.b200pt_scene_create(C.byref(desc), device, C.byref(h)), self.lib)
        self.h = h
        del keep

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200pt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters ------------------------------------------------------------
    def update_texture(self, tex: int, data) -> None:
        d = np.ascontiguousarray(data, np.float32).reshape(-1)
        abi.check(self.lib.b200pt_scene_update_texture(self.h, tex, d.ctypes.data_as(C.POINTER(C.c_float)), d.size), self.lib)

    def grad_zero(self) -> None:
        abi.check(self.lib.b200pt_grad_zero(self.h), self.lib)

    def grad(self, tex: int) -> np.ndarray:
        t = self.scene.textures[tex]
        out = np.zeros(t.size, np.float32)
        abi.check(self.lib.b200pt_grad_read(self.h, tex, out.ctypes.data_as(C.POINTER(C.c_float)), out.size), self.lib)
        return out.reshape(t.array().shape)

    def grad_device_view(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        abi.check(self.lib.b200pt_grad_device_view(self.h, C.byref(ptr), C.byref(n)), self.lib)
        return ptr.value, n.value

    def stats(self) -> dict:
        st = abi.Stats()
        abi.check(self.lib.b200pt_get_stats(self.h, C.byref(st)), self.lib)
        return {k: getattr(st, k) for k, _ in abi.Stats._fields_}

    # -- operators (parity tests) ---------------------------------------------------
    def ray_intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        n = rays.shape[0]
        t = np.zeros(n, np.float32); uv = np.zeros((n, 2), np.float32)
        prim = np.zeros(n, np.uint32); shape = np.zeros(n, np.int32)
        f = C.POINTER(C.c_float)
        abi.check(self.lib.b200pt_ray_intersect(self.h, n, rays.ctypes.data_as(f), t.ctypes.data_as(f), uv.ctypes.data_as(f),
                                                prim.ctypes.data_as(C.POINTER(C.c_uint32)), shape.ctypes.data_as(C.POINTER(C.c_int32))), self.lib)
        return t, uv, prim, shape

    def ray_test(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        hit = np.zeros(rays.shape[0], np.uint8)
        abi.check(self.lib.b200pt_ray_test(self.h, rays.shape[0], rays.ctypes.data_as(C.POINTER(C.c_float)),
                                           hit.ctypes.data_as(C.POINTER(C.c_uint8))), self.lib)
        return hit.astype(bool)

    def bsdf_eval_pdf_sample(self, bsdf: int, q):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 11)
        out = np.zeros((q.shape[0], 14), np.float32)
        f = C.POINTER(C.c_float)
        abi.check(self.lib.b200pt_bsdf_eval_pdf_sample(self.h, bsdf, q.shape[0], q.ctypes.data_as(f), out.ctypes.data_as(f)), self.lib)
        return out

    def env_query(self, q):
        """Environment emitter tables (``b200pt_env_query``): q (n, 8) = ref point, sample, direction."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 8)
        out = np.zeros((q.shape[0], 20), np.float32)
        f = C.POINTER(C.c_float)
        abi.check(self.lib.b200pt_env_query(self.h, q.shape[0], q.ctypes.data_as(f), out.ctypes.data_as(f)), self.lib)
        return out


def device_scene(scene: Scene, device: int = 0) -> DeviceScene:
    """Cached device scene of a host scene (created on first use)."""
    ds = scene._handle
    if ds is None or ds.h is None or ds.device != device:
        ds = DeviceScene(scene, device)
        scene._handle = ds
    return ds


class PathIntegrator:
    """Drop-in for the ``path`` plugin (src/integrators/path.cpp)."""

    prb = False
    default_max_depth = -1

    def __init__(self, props: dict | None = None, **kw):
        props = dict(props or {}); props.update(kw)
        max_depth = int(props.pop("max_depth", self.default_max_depth))
        if max_depth < 0 and max_depth != -1:
            raise RuntimeError("\"max_depth\" must be set to -1 (infinite) or a value >= 0")
        rr_depth = int(props.pop("rr_depth", 5))
        if rr_depth <= 0:
            raise RuntimeError("\"rr_depth\" must be set to a value greater than zero!")
        self.max_depth, self.rr_depth = max_depth, rr_depth
        self.hide_emitters = bool(props.pop("hide_emitters", False))
        self.chunk_lanes = int(props.pop("chunk_lanes", 0))
        props.pop("type", None)
        for k in ("block_size", "samples_per_pass", "timeout"):
            props.pop(k, None)      # accepted and ignored (scalar-variant scheduling knobs)
        if props:
            raise RuntimeError(f"unreferenced properties: {sorted(props)}")   # Properties' unqueried-key error

    @classmethod
    def from_scene(cls, scene: Scene, **kw):
        it = scene.integrator
        base = dict(max_depth=it["max_depth"], rr_depth=it["rr_depth"], hide_emitters=it["hide_emitters"])
        base.update(kw)
        return cls(base)

    # -- helpers ------------------------------------------------------------------
    def params(self, scene: Scene, seed: int, spp: int, shard=(0, 1), tile_size: int = 32) -> abi.RenderParams:
        p = abi.RenderParams()
        p.seed = int(seed) & 0xffffffff
        p.spp = int(spp) if spp else scene.sensor.sample_count
        p.max_depth, p.rr_depth, p.hide_emitters = self.max_depth, self.rr_depth, int(self.hide_emitters)
        p.shard_rank, p.shard_count, p.tile_size = shard[0], shard[1], tile_size
        p.chunk_lanes = self.chunk_lanes
        p.prb = int(self.prb)
        return p

    def aov_names(self):
        return []

    # -- SamplingIntegrator::render (integrator.cpp:151-396) -------------------------
    def render(self, scene: Scene, sensor=0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True,
               device: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        """`out`: optional preallocated float32 (H, W, 3) host array for the image (e.g. the numpy view of a pinned buffer: the
        device-to-host copy is then a plain asynchronous DMA instead of a staged copy into pageable memory)."""
        if not develop:
            raise NotImplementedError("develop=False: use render_accumulate() to obtain the raw film block")
        ds = device_scene(scene, device)
        p = self.params(scene, seed, spp)
        if out is None:
            out = np.empty(scene.film_shape, np.float32)
        elif out.dtype != np.float32 or tuple(out.shape) != tuple(scene.film_shape) or not out.flags.c_contiguous:
            raise ValueError("`out` must be a C-contiguous float32 array of the film's shape")
        abi.check(ds.lib.b200pt_render(ds.h, C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float))), ds.lib)
        return out

    def __repr__(self):
        return f"{type(self).__name__}[\n  max_depth = {self.max_depth},\n  rr_depth = {self.rr_depth}\n]"


class PRBIntegrator(PathIntegrator):
    """Drop-in for the ``prb`` plugin (src/python/python/ad/integrators/prb.py)."""

    prb = True
    default_max_depth = 6      # common.py:31

    # -- RBIntegrator.render_forward (common.py:560-623) -------------------------------
    def render_forward(self, scene: Scene, params: dict, sensor=0, seed: int = 0, spp: int = 0, device: int = 0) -> np.ndarray:
        """Forward-mode derivative of the image: ``params`` maps parameter names to their tangents
        (what ``dr.set_grad`` attaches in the reference); returns d(image) as an (H, W, 3) array."""
        ds = device_scene(scene, device)
        names = scene.parameters()
        f = C.POINTER(C.c_float)
        abi.check(ds.lib.b200pt_tangent_zero(ds.h), ds.lib)
        for k, v in params.items():
            i = names[k]
            if not scene.textures[i].differentiable:
                raise RuntimeError(f"parameter {k!r} is not differentiable")
            d = np.ascontiguousarray(v, np.float32).reshape(-1)
            abi.check(ds.lib.b200pt_tangent_write(ds.h, i, d.ctypes.data_as(f), d.size), ds.lib)
        p = self.params(scene, seed, spp)
        out = np.empty(scene.film_shape, np.float32)
        abi.check(ds.lib.b200pt_render_forward(ds.h, C.byref(p), out.ctypes.data_as(f)), ds.lib)
        return out

    # -- RBIntegrator.render_backward (common.py:625-783) ------------------------------
    def render_backward(self, scene: Scene, grad_in, params=None, sensor=0, seed: int = 0, spp: int = 0, device: int = 0,
                        zero: bool = True) -> dict:
        """Accumulates dLoss/dparam for every differentiable parameter. Returns
        ``{parameter name: gradient array}`` (the plugin adds these to
        ``dr.grad(params[name])``)."""
        ds = device_scene(scene, device)
        g = np.ascontiguousarray(grad_in, np.float32)
        if g.shape != scene.film_shape:
            raise RuntimeError(f"grad_in has shape {g.shape}, expected {scene.film_shape}")
        if zero:
            ds.grad_zero()
        p = self.params(scene, seed, spp)
        abi.check(ds.lib.b200pt_render_backward(ds.h, C.byref(p), g.ctypes.data_as(C.POINTER(C.c_float))), ds.lib)
        names = scene.parameters() if params is None else {k: scene.parameters()[k] for k in params}
        if params is not None:
            bad = [k for k, i in names.items() if not scene.textures[i].differentiable]
            if bad:      # never a silent zero: the adjoint has no derivative for these slots
                raise RuntimeError(f"no gradient available for {bad}: the parameter is marked non-differentiable or sits in a "
                                   f"BSDF slot whose derivative the PRB adjoint does not implement")
        return {k: ds.grad(i) for k, i in names.items() if scene.textures[i].differentiable}


def make_integrator(scene: Scene, integrator=None):
    if integrator is not None:
        return integrator
    cls = PRBIntegrator if scene.integrator["type"] == "prb" else PathIntegrator
    return cls.from_scene(scene)


def render(scene: Scene, params=None, sensor=0, integrator=None, seed: int = 0, seed_grad: int = 0, spp: int = 0,
           spp_grad: int = 0, device: int = 0) -> np.ndarray:
    """``mi.render`` (util.py:396-560) for the hot path: primal image as an (H, W, 3) array.
    For gradients use :func:`render_torch` (autograd) or the integrator's ``render_backward``."""
    return make_integrator(scene, integrator).render(scene, sensor, seed=seed, spp=spp, device=device)


def update_params(scene: Scene, values: dict, device: int = 0) -> None:
    """``params.update()``: push new parameter values to the host scene and the device copy."""
    names = scene.parameters()
    for k, v in values.items():
        i = names[k]
        t = scene.textures[i]
        v = np.asarray(v, np.float32)
        if t.kind == abi.TEX_BITMAP:
            t.data = np.ascontiguousarray(v.reshape(t.data.shape))
        elif t.kind == abi.TEX_CHECKERBOARD:
            v = v.reshape(2, -1)
            t.value[: t.channels], t.value1[: t.channels] = v[0, : t.channels], v[1, : t.channels]
        else:
            t.value[: t.channels] = v.reshape(-1)[: t.channels]
        if scene._handle is not None and scene._handle.h is not None:
            scene._handle.update_texture(i, t.array())


def update_vertices(scene: Scene, shape_id: str, packed_vertices) -> None:
    """Geometry update (``params['<shape>.vertex_positions'] = ...; params.update()`` in the reference,
    scene.cpp:517-540): replaces the packed (V, 8) records of one mesh. With a live device scene the new vertices are
    uploaded and the BVH is REFITTED on the device (``b200pt_scene_update_vertices``: same topology, boxes recomputed
    bottom-up); shapes sampled as emitters (host-built sampling tables) drop the device scene instead, which the next
    render re-creates. The mesh topology must stay the same."""
    idx = next((i for i, s in enumerate(scene.shapes) if s.id == shape_id), None)
    if idx is None:
        raise KeyError(shape_id)
    sh = scene.shapes[idx]
    v = np.ascontiguousarray(packed_vertices, np.float32).reshape(-1, 8)
    if v.shape != sh.vertices.shape:
        raise ValueError(f"expected packed vertices of shape {sh.vertices.shape}, got {v.shape}")
    if sh.sampling == abi.SAMPLING_RECTANGLE:
        raise NotImplementedError("rectangle emitters are sampled through their to_world transform; re-create the shape instead")
    sh.vertices = v
    ds = scene._handle
    if ds is not None and ds.h is not None:
        if sh.sampling == abi.SAMPLING_NONE:
            abi.check(ds.lib.b200pt_scene_update_vertices(ds.h, idx, v.ctypes.data_as(C.POINTER(C.c_float)), v.shape[0]), ds.lib)
        else:
            ds.close()
            scene._handle = None


def render_torch(scene: Scene, params: dict, integrator=None, seed: int = 0, seed_grad=None, spp: int = 0, spp_grad: int = 0,
                 device: int = 0):
    """Differentiable ``mi.render(scene, params, ...)``: ``params`` maps parameter
    names to torch tensors (``requires_grad``); the returned image participates in
    autograd like the reference's ``_RenderOp`` (util.py:344-395): forward = primal
    render with ``seed``, backward = ``render_backward`` with ``seed_grad``
    (default: ``sample_tea_32(seed, 1)[0]``, util.py:505-507)."""
    import torch

    integ = make_integrator(scene, integrator)
    names = list(params.keys())
    if seed_grad is None:
        from .rng import sample_tea_32
        seed_grad = sample_tea_32(seed, 1)[0]

    class _RenderOp(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *tensors):
            update_params(scene, {k: t.detach().cpu().numpy() for k, t in zip(names, tensors)}, device)
            img = integ.render(scene, seed=seed, spp=spp, device=device)
            return torch.from_numpy(img)

        @staticmethod
        def backward(ctx, grad_out):
            prb = integ if isinstance(integ, PRBIntegrator) else PRBIntegrator(max_depth=integ.max_depth, rr_depth=integ.rr_depth,
                                                                                hide_emitters=integ.hide_emitters)
            grads = prb.render_backward(scene, grad_out.detach().cpu().numpy(), params=names, seed=seed_grad,
                                        spp=spp_grad or spp, device=device)
            return tuple(torch.from_numpy(grads[k]).reshape(params[k].shape).to(params[k].dtype) if k in grads else None for k in names)

    return _RenderOp.apply(*[params[k] for k in names])
