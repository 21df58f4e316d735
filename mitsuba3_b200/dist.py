"""Multi-GPU rendering: pixel-tile sharding + one all-reduce of the film (and of
the parameter gradients in the PRB path) -- SURVEY.md 8(e).

One process per GPU (``torch.distributed``, NCCL over NVLink). Every rank holds
the full scene and a full-frame raw film block ``(H, W, 4)`` = (R, G, B, weight)
because a non-box reconstruction filter splats across tile borders; rank r
renders the pixel tiles dealt to it by ``tile_owner`` using the GLOBAL lane index
``pixel * spp + s`` for its sampler streams, so the image does not depend on the
number of GPUs. The single collective is ``all_reduce(SUM)`` of the raw block
before the (non-linear) weight division of ``HDRFilm::develop``; there is no
exchange inside the bounce loop. The reference has no distributed mode at all
(one device per process), so this layer has no reference counterpart.
"""
from __future__ import annotations

import ctypes as C



def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def tile_stride(world_size):
    """Row stride of the tile deal: the smallest value >= N // 2 + 1 that is coprime with N (N = 2: 1)."""
    import math
    if world_size <= 2:
        return 1
    st = world_size // 2 + 1
    while math.gcd(st, world_size) != 1:
        st += 1
    return st


def tile_owner(x, y, width, tile_size, world_size):
    """Rank that renders pixel (x, y): tile (tx, ty) goes to rank (tx + ty * stride) % N with a stride coprime
    with N -- a diagonal deal (a checkerboard for N = 2) in which every rank meets every tile column and row
    (same rule as ensure_pix_ids in csrc/api.cu). ``width`` is kept for signature compatibility."""
    del width
    return ((x // tile_size) + (y // tile_size) * tile_stride(world_size)) % world_size


def all_reduce_film(film):
    """SUM-reduce the raw film block (torch tensor, any device) over all ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    return film


def develop(film):
    """HDRFilm::develop (hdrfilm.cpp:393) on a raw block tensor: rgb / (w == 0 ? 1 : w)."""
    import torch
    w = film[..., 3:4]
    return film[..., :3] / torch.where(w == 0, torch.ones_like(w), w)


def render_distributed(scene, integrator=None, seed: int = 0, spp: int = 0, tile_size: int = 32, device=None):
    """Render ``scene`` cooperatively on all ranks; every rank returns the full image
    as a CUDA tensor (H, W, 3). Timing-critical callers keep the result on the device.
    The tensor is a buffer owned by the device scene: it is overwritten by the next call (clone to keep it).
    Nothing here waits for the device: kernels, the all-reduce and develop are enqueued on the current stream."""
    import torch
    from . import abi
    from .integrators import device_scene, make_integrator
    rank, ws = world()
    dev = torch.cuda.current_device() if device is None else device
    integ = make_integrator(scene, integrator)
    ds = device_scene(scene, dev)
    H, W, _ = scene.film_shape
    # the raw block and the developed image are allocated once per device scene and reused frame after frame
    bufs = getattr(ds, "_dist_bufs", None)
    if bufs is None or bufs[0].shape[:2] != (H, W):
        bufs = (torch.empty((H, W, 4), dtype=torch.float32, device=f"cuda:{dev}"), torch.empty((H, W, 3), dtype=torch.float32, device=f"cuda:{dev}"))
        ds._dist_bufs = bufs
    film, out = bufs
    film.zero_()
    p = integ.params(scene, seed, spp, shard=(rank, ws), tile_size=tile_size)
    stream = torch.cuda.current_stream(dev).cuda_stream
    abi.check(ds.lib.b200pt_render_accumulate(ds.h, C.byref(p), C.c_void_p(film.data_ptr()), C.c_void_p(stream)), ds.lib)
    all_reduce_film(film)                      # the one NCCL collective of a frame
    abi.check(ds.lib.b200pt_develop(ds.h, C.c_void_p(film.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(stream)), ds.lib)
    return out


def render_backward_distributed(scene, grad_in, integrator=None, seed: int = 0, spp: int = 0, tile_size: int = 32, device=None):
    """PRB gradient step on all ranks: each rank back-propagates its pixel tiles, then ONE
    fused all-reduce of the flat gradient buffer (all differentiable parameters)."""
    import torch
    from . import abi
    from .integrators import PRBIntegrator, device_scene
    rank, ws = world()
    dev = torch.cuda.current_device() if device is None else device
    integ = integrator or PRBIntegrator.from_scene(scene)
    ds = device_scene(scene, dev)
    g = torch.as_tensor(grad_in, dtype=torch.float32, device=f"cuda:{dev}").contiguous()
    ds.grad_zero()
    p = integ.params(scene, seed, spp, shard=(rank, ws), tile_size=tile_size)
    p.prb = 1
    stream = torch.cuda.current_stream(dev).cuda_stream
    abi.check(ds.lib.b200pt_render_backward_device(ds.h, C.byref(p), C.c_void_p(g.data_ptr()), C.c_void_p(stream)), ds.lib)
    ptr, n = ds.grad_device_view()
    flat = _wrap_device_floats(ptr, n, dev)
    all_reduce_film(flat)
    torch.cuda.current_stream(dev).synchronize()
    return {k: ds.grad(i) for k, i in scene.parameters().items() if scene.textures[i].differentiable}


def _wrap_device_floats(ptr: int, n: int, dev: int):
    """Zero-copy torch view of library-owned device memory (for NCCL)."""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": "<f4", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device=f"cuda:{dev}")[:n]
