"""``ScalarTransform4f`` work-alike in fp32 (host-side scene construction).

Mirrors ``include/mitsuba/core/transform.h`` of the reference: a transform is
the pair (matrix, inverse_transpose), composition keeps both
(transform.h:357-396), points/vectors/normals are mapped with the same
fused-multiply-add chains (transform.h:288-347) so that the packed vertex
records produced here agree with the reference's to the last bit where the
libm ``sin``/``cos`` agree.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def fma(a, b, c):
    """fp32 fused multiply-add (exact product in fp64, one rounding)."""
    return f32(np.float64(f32(a)) * np.float64(f32(b)) + np.float64(f32(c)))


class Transform4f:
    """Affine 4x4 transform; ``T().translate(a).rotate(axis, deg).scale(s)``
    composes on the right exactly like the reference's chaining API."""

    def __init__(self, matrix=None, inverse_transpose=None):
        self.matrix = np.eye(4, dtype=f32) if matrix is None else np.array(matrix, dtype=f32).reshape(4, 4)
        if inverse_transpose is None:
            inverse_transpose = np.linalg.inv(self.matrix.astype(np.float64)).T.astype(f32)
        self.inverse_transpose = np.array(inverse_transpose, dtype=f32).reshape(4, 4)

    # -- composition (transform.h:357-396, affine branch) --------------------
    def __matmul__(self, o: "Transform4f") -> "Transform4f":
        m, it = np.zeros((4, 4), f32), np.zeros((4, 4), f32)
        m[3, 3] = it[3, 3] = 1
        for i in range(3):
            for j in range(3):
                s, sit = f32(0), f32(0)
                for k in range(3):
                    s = fma(self.matrix[i, k], o.matrix[k, j], s)
                for k in range(3):
                    sit = fma(self.inverse_transpose[i, k], o.inverse_transpose[k, j], sit)
                m[i, j], it[i, j] = s, sit
        for l in range(3):
            s, sit = self.matrix[l, 3], o.inverse_transpose[3, l]
            for k in range(3):
                s = fma(self.matrix[l, k], o.matrix[k, 3], s)
            for k in range(3):
                sit = fma(self.inverse_transpose[3, k], o.inverse_transpose[k, l], sit)
            m[l, 3], it[3, l] = s, sit
        return Transform4f(m, it)

    # -- factories (transform.h:130-196) -------------------------------------
    def translate(self, v) -> "Transform4f":
        v = np.asarray(v, f32)
        m = np.eye(4, dtype=f32); m[:3, 3] = v
        it = np.eye(4, dtype=f32); it[3, :3] = -v
        return self @ Transform4f(m, it)

    def scale(self, v) -> "Transform4f":
        v = np.broadcast_to(np.asarray(v, f32), (3,)).astype(f32)
        m = np.diag(np.concatenate([v, [f32(1)]])).astype(f32)
        it = np.diag(np.concatenate([f32(1) / v, [f32(1)]])).astype(f32)
        return self @ Transform4f(m, it)

    def rotate(self, axis, angle_deg) -> "Transform4f":
        # drjit/transform.h:46-69 with angle = deg_to_rad(angle) in fp32
        axis = np.asarray(axis, f32)
        angle = f32(angle_deg) * f32(np.pi / 180.0)
        s, c = f32(np.sin(np.float64(angle))), f32(np.cos(np.float64(angle)))   # correctly rounded fp32
        cm = f32(1) - c
        sh1, sh2 = axis[[1, 2, 0]], axis[[2, 0, 1]]
        tmp0 = np.array([fma(axis[i] * axis[i], cm, c) for i in range(3)], f32)
        tmp1 = np.array([fma(axis[i] * sh1[i], cm, sh2[i] * s) for i in range(3)], f32)
        tmp2 = np.array([fma(axis[i] * sh2[i], cm, -(sh1[i] * s)) for i in range(3)], f32)
        m = np.array([[tmp0[0], tmp2[1], tmp1[2], 0],
                      [tmp1[0], tmp0[1], tmp2[2], 0],
                      [tmp2[0], tmp1[1], tmp0[2], 0],
                      [0, 0, 0, 1]], f32)
        # (dr::Matrix's Vector4 constructor arguments are rows)
        return self @ Transform4f(m, m)

    def look_at(self, origin, target, up) -> "Transform4f":
        # transform.h:172-196
        origin, target, up = (np.asarray(x, f32) for x in (origin, target, up))
        d = _normalize(target - origin)
        left = _normalize(_cross(up, d))
        new_up = _cross(d, left)
        m = np.eye(4, dtype=f32)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
        inv = np.eye(4, dtype=f32)  # rows before the final transposes in the reference
        inv[:3, 0], inv[:3, 1], inv[:3, 2] = left, new_up, d
        # inverse[3] = transpose(inverse) * (-origin, 1): translation of the inverse
        r = inv[:3, :3].T
        tr = np.array([-_dot(r[i], origin) for i in range(3)], f32)
        it = inv.copy()
        it[3, :3] = tr
        return self @ Transform4f(m, it)

    # -- application ----------------------------------------------------------
    def point(self, p):
        p = np.asarray(p, f32)
        r = [self.matrix[i, 3] for i in range(3)]
        for j in range(3):
            for i in range(3):
                r[i] = fma(self.matrix[i, j], p[j], r[i])
        return np.array(r, f32)

    def vector(self, v):
        v = np.asarray(v, f32)
        r = [f32(self.matrix[i, 0] * v[0]) for i in range(3)]
        for j in range(1, 3):
            for i in range(3):
                r[i] = fma(self.matrix[i, j], v[j], r[i])
        return np.array(r, f32)

    def normal(self, n):
        n = np.asarray(n, f32)
        r = [f32(self.inverse_transpose[i, 0] * n[0]) for i in range(3)]
        for j in range(1, 3):
            for i in range(3):
                r[i] = fma(self.inverse_transpose[i, j], n[j], r[i])
        return np.array(r, f32)

    def det3(self) -> float:
        return float(np.linalg.det(self.matrix[:3, :3].astype(np.float64)))

    def __repr__(self):
        return f"Transform4f({self.matrix.tolist()})"


def _dot(a, b):
    return fma(a[2], b[2], fma(a[1], b[1], f32(a[0] * b[0])))


def _cross(a, b):
    return np.array([fma(a[1], b[2], -f32(a[2] * b[1])),
                     fma(a[2], b[0], -f32(a[0] * b[2])),
                     fma(a[0], b[1], -f32(a[1] * b[0]))], f32)


def _sqnorm(a):
    return fma(a[2], a[2], fma(a[1], a[1], f32(a[0] * a[0])))


def _normalize(a):
    a = np.asarray(a, f32)
    return (a * (f32(1) / np.sqrt(_sqnorm(a), dtype=f32))).astype(f32)


def _matmul4(a, b):
    """dr::Matrix product in fp32: result(i,j) = fmadd chain over k."""
    r = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            s = f32(a[i, 0] * b[0, j])
            for k in range(1, 4):
                s = fma(a[i, k], b[k, j], s)
            r[i, j] = s
    return r


def perspective_sample_to_camera(film_size, crop_size, crop_offset, fov_x, near, far):
    """``perspective_projection(...).inverse()`` (sensor.h:234-269) in fp32.

    The reference keeps (matrix, inverse_transpose) for every factor
    (transform.h:130-140,419-437) and multiplies the inverse transposes in the
    same order as the matrices; the inverse is the transpose of that product.
    """
    fw, fh = f32(film_size[0]), f32(film_size[1])
    rel_size = np.array([f32(crop_size[0]) / fw, f32(crop_size[1]) / fh], f32)
    rel_off = np.array([f32(crop_offset[0]) / fw, f32(crop_offset[1]) / fh], f32)
    aspect = f32(fw / fh)
    near, far = f32(near), f32(far)

    def it_scale(v):
        v = np.asarray(v, f32)
        return np.diag(np.concatenate([f32(1) / v, [f32(1)]])).astype(f32)

    def it_translate(v):
        m = np.eye(4, dtype=f32); m[3, :3] = -np.asarray(v, f32); return m

    ang = f32(f32(fov_x) * f32(0.5)) * f32(np.pi / 180.0)
    tan = f32(np.tan(np.float64(ang)))   # correctly rounded fp32 tangent (numpy's fp32 tan is 1 ulp off here)
    inv_p = np.diag(np.array([tan, tan, 0, f32(1) / near], f32)).astype(f32)
    inv_p[2, 3] = 1
    inv_p[3, 2] = f32(near - far) / f32(far * near)
    it = it_scale([f32(1) / rel_size[0], f32(1) / rel_size[1], 1])
    for nxt in (it_translate([-rel_off[0], -rel_off[1], 0]),
                it_scale([-0.5, f32(-0.5) * aspect, 1]),
                it_translate([-1, f32(-1) / aspect, 0]),
                inv_p.T.copy()):
        it = _matmul4(it, nxt)
    return it.T.copy() + f32(0)   # +0 normalises -0.0


def parse_fov(fov, fov_axis, aspect):
    """sensor.cpp:142-195 parse_fov -> horizontal field of view in degrees."""
    axis = fov_axis.lower()
    if axis == "smaller":
        axis = "y" if aspect > 1 else "x"
    elif axis == "larger":
        axis = "x" if aspect > 1 else "y"
    if axis == "x":
        return float(fov)
    if axis == "y":
        return float(np.rad2deg(2.0 * np.arctan(np.tan(0.5 * np.deg2rad(fov)) * aspect)))
    if axis == "diagonal":
        diag = 2.0 * np.tan(0.5 * np.deg2rad(fov))
        width = diag / np.sqrt(1.0 + 1.0 / (aspect * aspect))
        return float(np.rad2deg(2.0 * np.arctan(width * 0.5)))
    raise ValueError(f"unknown fov_axis {fov_axis!r}")
