"""ctypes mirror of ``include/b200pt.h`` and the loader of ``libb200pt.so``.

The structures below must stay field-for-field identical to the C header; the
CPU test-suite checks their sizes against ``b200pt_abi_sizeof()`` exported by
the library, and that every symbol the header declares is exported.

There is no fallback: if the CUDA library has not been built, ``load()`` raises
``RuntimeError`` (the product path must fail loudly, never silently degrade).
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 2
MAX_SLOTS = 12

# enums ---------------------------------------------------------------------
TEX_CONST, TEX_BITMAP, TEX_CHECKERBOARD = 0, 1, 2
EMITTER_AREA, EMITTER_CONSTANT, EMITTER_ENVMAP = 0, 1, 2
WRAP_REPEAT, WRAP_MIRROR, WRAP_CLAMP = 0, 1, 2
FILTER_BILINEAR, FILTER_NEAREST = 0, 1
BSDF_DIFFUSE, BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_PRINCIPLED, BSDF_PLASTIC = 0, 1, 2, 3, 4
SAMPLING_NONE, SAMPLING_RECTANGLE, SAMPLING_MESH = 0, 1, 2
LAYOUT_NORMALS, LAYOUT_TANGENTS, LAYOUT_TEXCOORDS = 1, 2, 4
RFILTER_BOX, RFILTER_GAUSSIAN, RFILTER_GAUSSIAN_EXP2, RFILTER_GAUSSIAN_TABLE = 0, 1, 2, 3

# texture slots
SLOT_REFLECTANCE = 0
SLOT_ETA, SLOT_K, SLOT_SPEC_REFL, SLOT_ALPHA_U, SLOT_ALPHA_V = 0, 1, 2, 3, 4
SLOT_D_SPEC_REFL, SLOT_D_SPEC_TRANS, SLOT_D_ALPHA_U, SLOT_D_ALPHA_V = 0, 1, 2, 3
(SLOT_P_BASE_COLOR, SLOT_P_ROUGHNESS, SLOT_P_ANISOTROPIC, SLOT_P_METALLIC,
 SLOT_P_SPEC_TRANS, SLOT_P_SPECULAR, SLOT_P_SPEC_TINT, SLOT_P_SHEEN,
 SLOT_P_SHEEN_TINT, SLOT_P_FLATNESS, SLOT_P_CLEARCOAT, SLOT_P_CLEARCOAT_GLOSS) = range(12)

P_HAS_CLEARCOAT, P_HAS_SHEEN, P_HAS_SPEC_TRANS, P_HAS_METALLIC = 1, 2, 4, 8
P_HAS_SPEC_TINT, P_HAS_SHEEN_TINT, P_HAS_ANISOTROPIC, P_HAS_FLATNESS = 16, 32, 64, 128
P_ETA_SPECULAR = 256
M_ROUGH, M_GGX, M_NONLINEAR = 1 << 16, 1 << 17, 1 << 18     # roughconductor / roughdielectric / plastic (B200PT_M_*)
SLOT_PL_DIFFUSE, SLOT_PL_SPEC_REFL = 0, 1

STATUS = {0: "ok", 1: "invalid argument", 2: "CUDA error / no device",
          3: "unsupported", 4: "out of memory"}


class Texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("channels", C.c_int32), ("value", C.c_float * 3), ("value1", C.c_float * 3),
                ("width", C.c_int32), ("height", C.c_int32), ("data", C.POINTER(C.c_float)),
                ("wrap", C.c_int32), ("filter", C.c_int32), ("to_uv", C.c_float * 9),
                ("differentiable", C.c_int32)]


class Bsdf(C.Structure):
    _fields_ = [("type", C.c_int32), ("twosided", C.c_int32), ("tex", C.c_int32 * MAX_SLOTS),
                ("eta", C.c_float), ("spec_srate", C.c_float), ("clearcoat_srate", C.c_float),
                ("diff_refl_srate", C.c_float), ("flags", C.c_uint32),
                ("plastic_fdr_int", C.c_float), ("plastic_spec_weight", C.c_float)]


class Shape(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_faces", C.c_uint32),
                ("vertices", C.POINTER(C.c_float)), ("faces", C.POINTER(C.c_uint32)),
                ("layout", C.c_uint32), ("bsdf", C.c_int32), ("emitter", C.c_int32),
                ("sampling", C.c_int32), ("to_world", C.c_float * 16),
                ("frame_n", C.c_float * 3), ("inv_area", C.c_float)]


class Emitter(C.Structure):
    _fields_ = [("shape", C.c_int32), ("radiance_tex", C.c_int32), ("sampling_weight", C.c_float),
                ("type", C.c_int32), ("env_scale", C.c_float),
                ("env_mis_compensation", C.c_int32), ("to_world", C.c_float * 16),
                ("to_world_inv", C.c_float * 16)]


class Sensor(C.Structure):
    _fields_ = [("sample_to_camera", C.c_float * 16), ("to_world", C.c_float * 16),
                ("near_clip", C.c_float), ("far_clip", C.c_float),
                ("film_size", C.c_uint32 * 2), ("crop_size", C.c_uint32 * 2),
                ("crop_offset", C.c_uint32 * 2), ("rfilter", C.c_int32),
                ("rfilter_stddev", C.c_float), ("base_seed", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32),
                ("n_shapes", C.c_uint32), ("shapes", C.POINTER(Shape)),
                ("n_bsdfs", C.c_uint32), ("bsdfs", C.POINTER(Bsdf)),
                ("n_emitters", C.c_uint32), ("emitters", C.POINTER(Emitter)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
                ("sensor", Sensor)]


class RenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("spp", C.c_uint32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("hide_emitters", C.c_int32),
                ("shard_rank", C.c_uint32), ("shard_count", C.c_uint32),
                ("tile_size", C.c_uint32), ("chunk_lanes", C.c_uint32), ("prb", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("bounces", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_double),
                ("trace_ms", C.c_double), ("trace_launches", C.c_uint64),
                ("trace_rays", C.c_uint64)]


# Every symbol include/b200pt.h declares (checked by tests/test_abi.py).
SYMBOLS = [
    "b200pt_abi_version", "b200pt_last_error", "b200pt_device_count", "b200pt_set_devices",
    "b200pt_scene_create", "b200pt_scene_destroy", "b200pt_scene_update_texture", "b200pt_scene_update_vertices",
    "b200pt_render", "b200pt_render_accumulate", "b200pt_develop",
    "b200pt_render_backward", "b200pt_render_backward_device", "b200pt_grad_zero",
    "b200pt_tangent_zero", "b200pt_tangent_write", "b200pt_render_forward",
    "b200pt_grad_read", "b200pt_grad_device_view", "b200pt_grad_offset",
    "b200pt_ray_intersect", "b200pt_ray_test", "b200pt_bsdf_eval_pdf_sample", "b200pt_env_query",
    "b200pt_get_stats", "b200pt_abi_sizeof",
]

LIB_PATH = os.environ.get("B200PT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200pt.so")
_lib = None


class B200PTError(RuntimeError):
    pass


def adjoint_covers_slot(bsdf_type: int, flags: int, slot: int) -> bool:
    """BSDF slots whose parameter derivative the PRB adjoint implements (csrc/api.cu: adjoint_covers_slot): all texture
    slots of all models (delta lobes: an exact zero, prb.py:296) except principled `specular`."""
    if bsdf_type == BSDF_PRINCIPLED:
        return slot != SLOT_P_SPECULAR
    return BSDF_DIFFUSE <= bsdf_type <= BSDF_PLASTIC


def load() -> C.CDLL:
    """Load ``libb200pt.so`` (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "mitsuba3_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, f32p = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
    lib.b200pt_abi_version.restype = C.c_uint32
    lib.b200pt_last_error.restype = C.c_char_p
    lib.b200pt_device_count.restype = C.c_int
    lib.b200pt_set_devices.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.b200pt_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(vp)]
    lib.b200pt_scene_destroy.argtypes = [vp]
    lib.b200pt_scene_destroy.restype = None
    lib.b200pt_scene_update_texture.argtypes = [vp, u32, f32p, C.c_size_t]
    lib.b200pt_scene_update_vertices.argtypes = [vp, u32, f32p, u32]
    lib.b200pt_render.argtypes = [vp, C.POINTER(RenderParams), f32p]
    lib.b200pt_render_accumulate.argtypes = [vp, C.POINTER(RenderParams), vp, vp]
    lib.b200pt_develop.argtypes = [vp, vp, vp, vp]
    lib.b200pt_render_backward.argtypes = [vp, C.POINTER(RenderParams), f32p]
    lib.b200pt_render_backward_device.argtypes = [vp, C.POINTER(RenderParams), vp, vp]
    lib.b200pt_grad_zero.argtypes = [vp]
    lib.b200pt_grad_read.argtypes = [vp, u32, f32p, C.c_size_t]
    lib.b200pt_grad_device_view.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.b200pt_grad_offset.argtypes = [vp, u32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.b200pt_ray_intersect.argtypes = [vp, u32, f32p, f32p, f32p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    lib.b200pt_ray_test.argtypes = [vp, u32, f32p, C.POINTER(C.c_uint8)]
    lib.b200pt_bsdf_eval_pdf_sample.argtypes = [vp, u32, u32, f32p, f32p]
    lib.b200pt_tangent_zero.argtypes = [vp]
    lib.b200pt_tangent_write.argtypes = [vp, u32, f32p, C.c_size_t]
    lib.b200pt_render_forward.argtypes = [vp, C.POINTER(RenderParams), f32p]
    lib.b200pt_env_query.argtypes = [vp, u32, f32p, f32p]
    lib.b200pt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.b200pt_abi_sizeof.argtypes = [C.c_int]
    lib.b200pt_abi_sizeof.restype = C.c_size_t
    if lib.b200pt_abi_version() != ABI_VERSION:
        raise RuntimeError("libb200pt.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, lib=None) -> None:
    if status != 0:
        lib = lib or load()
        msg = lib.b200pt_last_error()
        raise B200PTError(f"b200pt: {STATUS.get(status, status)}: "
                          f"{msg.decode() if msg else ''}")
