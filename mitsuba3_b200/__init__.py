"""mitsuba3_b200 -- B200-native wavefront path tracer behind Mitsuba 3's
integrator interface (hot path only: `path` / `prb`).

The compute path is hand-written sm_100a CUDA in ``csrc/`` exposed through the
C ABI of ``include/b200pt.h``; this package is the thin host-side mirror of
the reference's Python interface (``load_dict`` / ``render`` / integrator
``render`` + ``render_backward`` / parameter map). There is no CPU fallback.
"""
from .scene import (Scene, load_dict, cornell_box, cornell_box_heightfield, heightfield_mesh,  # noqa: F401
                    matpreview_like, matpreview_scene, uv_sphere_mesh, synthetic_sky)
from .transform import Transform4f  # noqa: F401

ScalarTransform4f = Transform4f


def __getattr__(name):
    # integrators import the CUDA library lazily so that host-only logic
    # (scene parsing, sharding, descriptors) stays importable on CPU boxes.
    if name in ("PathIntegrator", "PRBIntegrator", "render", "DeviceScene", "render_torch", "update_params", "update_vertices"):
        from . import integrators
        return getattr(integrators, name)
    raise AttributeError(name)
