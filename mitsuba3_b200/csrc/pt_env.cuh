// pt_env.cuh -- environment emitters on the device: `envmap`
// (src/emitters/envmap.cpp:276-396,425-590) with its Hierarchical2D<Float,0> sample warp
// (include/mitsuba/core/distr_2d.h:549-729, warp.h:447-521) and `constant`
// (src/emitters/constant.cpp:95-152). The elementary functions follow the single precision
// branches of drjit/math.h (sincos :74-186, acos :458-502, atan2 :520-597) operation by
// operation, so the sampled directions are the ones the reference's own polynomials give.
// Included by pt_device.cuh (needs DevEnv, DirectionSample).
#pragma once

namespace pt {

#define PT_DR_EPSILON 5.9604644775390625e-08f   // dr::Epsilon<float> = 2^-24

PT_DEV float dr_lerp(float a, float b, float t) { return __fmaf_rn(b, t, __fmaf_rn(-a, t, a)); }

// drjit/math.h:74-186 detail::sincos<true, true>, float
PT_DEV void dr_sincos(float x, float &s_out, float &c_out) {
    float xa = fabsf(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t sign_sin = (((uint32_t) j) << 29) ^ __float_as_uint(x);
    uint32_t sign_cos = ((uint32_t) ~(j - 2)) << 29;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y;
    if (xa == PT_INF) z = __int_as_float(0x7fc00000);
    float z2 = z * z;
    float s = __fmaf_rn(z2, -1.9515295891e-4f, __fmaf_rn(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = __fmaf_rn(z2, 2.443315711809948e-5f, __fmaf_rn(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = __fmaf_rn(s, y, y);
    c = __fmaf_rn(c, z, __fmaf_rn(z, -0.5f, 1.f));
    bool polymask = (j & 2) == 0;
    float rs = polymask ? s : c, rc = polymask ? c : s;
    s_out = (sign_sin >> 31) ? -rs : rs;
    c_out = (sign_cos >> 31) ? -rc : rc;
}

// drjit/math.h:458-502 acos, float (estrin with 5 coefficients)
PT_DEV float dr_acos(float x) {
    float xa = fabsf(x), x2 = x * x;
    bool big = xa > 0.5f;
    float x1 = 0.5f * (1.f - xa);
    float x3 = big ? x1 : x2, x4 = big ? __fsqrt_rn(x1) : xa;
    float p0 = __fmaf_rn(x3, 7.4953002686e-2f, 1.666675242e-1f), p1 = __fmaf_rn(x3, 2.4181311049e-2f, 4.5470025998e-2f), p2 = 4.2163199048e-2f;
    float y2 = x3 * x3;
    float q0 = __fmaf_rn(y2, p1, p0);
    float y4 = y2 * y2;
    float z1 = __fmaf_rn(y4, p2, q0);
    z1 = __fmaf_rn(z1, x3 * x4, x4);
    float z2 = z1 + z1;
    z2 = x < 0.f ? PT_PI - z2 : z2;
    float z3 = (PT_PI * .5f) - copysignf(z1, x);
    return big ? z2 : z3;
}
PT_DEV float dr_safe_acos(float x) { return dr_acos(fminf(fmaxf(x, -1.f), 1.f)); }

// drjit/math.h:520-597 atan2, float (estrin with 7 coefficients)
PT_DEV float dr_atan2(float y, float x) {
    float abs_x = fabsf(x), abs_y = fabsf(y);
    float min_val = fminf(abs_y, abs_x), max_val = fmaxf(abs_x, abs_y);
    float scaled_min = fdiv(min_val, max_val), z = scaled_min * scaled_min;
    float p0 = __fmaf_rn(z, -0.33326497518773606976f, 0.99999934166683966009f), p1 = __fmaf_rn(z, -0.13486708938456973185f, 0.19881342388439013552f),
          p2 = __fmaf_rn(z, -0.037006525670417265220f, 0.083863120428809689910f), p3 = 0.0078613793713198150252f;
    float z2 = z * z;
    float q0 = __fmaf_rn(z2, p1, p0), q1 = __fmaf_rn(z2, p3, p2);
    float z4 = z2 * z2;
    float t = __fmaf_rn(z4, q1, q0);
    t = t * scaled_min;
    t = abs_y > abs_x ? PT_PI * .5f - t : t;
    t = x < 0.f ? PT_PI - t : t;
    float r = y < 0.f ? -t : t;
    if (!(max_val != 0.f)) r = 0.f;
    return r;
}

// transform.h:288-299 vector transform with a row-major 3x3
PT_DEV float3 env_xform(const float *m, float3 v) {
    float q0 = m[0] * v.x, q1 = m[3] * v.x, q2 = m[6] * v.x;
    q0 = __fmaf_rn(m[1], v.y, q0); q1 = __fmaf_rn(m[4], v.y, q1); q2 = __fmaf_rn(m[7], v.y, q2);
    q0 = __fmaf_rn(m[2], v.z, q0); q1 = __fmaf_rn(m[5], v.z, q1); q2 = __fmaf_rn(m[8], v.z, q2);
    return V(q0, q1, q2);
}

// distr_2d.h:782-785 Level::index (2x2 blocks stored contiguously)
PT_DEV uint32_t h2d_index(uint32_t width, uint32_t x, uint32_t y) {
    return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1)) + ((y & ~1u) * width);
}

// warp.h:447-453
PT_DEV float interval_to_linear(float v0, float v1, float sample) {
    if (fabsf(v0 - v1) > 1e-4f * (v0 + v1))
        return fdiv(v0 - safe_sqrt(dr_lerp(v0 * v0, v1 * v1, sample)), v0 - v1);
    return sample;
}

// distr_2d.h:549-622 Hierarchical2D::sample: one 128-bit load per MIP level
PT_DEV void h2d_sample(const DevEnv &e, float sx, float sy, float &ox, float &oy, float &pdf) {
    sx = fminf(fmaxf(sx, 0.f), 1.f); sy = fminf(fmaxf(sy, 0.f), 1.f);
    uint32_t offx = 0, offy = 0;
    for (int l = (int) e.n_levels - 1; l > 0; --l) {
        offx <<= 1; offy <<= 1;
        uint32_t packet = (e.lvl_offset[l] + h2d_index(e.lvl_width[l], offx, offy)) >> 2;
        float4 v = __ldg((const float4 *) e.warp + packet);
        float v00 = v.x, v10 = v.y, v01 = v.z, v11 = v.w;
        sx = fminf(fmaxf(sx, 0.f), 1.f); sy = fminf(fmaxf(sy, 0.f), 1.f);
        float r0 = v00 + v10, r1 = v01 + v11;
        sy *= r0 + r1;
        bool y_mask = sy > r0;
        if (y_mask) { offy += 1u; sy -= r0; }
        float dy = y_mask ? r1 : r0;
        float c0 = y_mask ? v01 : v00, c1 = y_mask ? v11 : v10;
        sx *= dy;
        bool x_mask = sx > c0;
        if (x_mask) { sx -= c0; offx += 1u; }
        float dx = x_mask ? c1 : c0;
        float inv = rcp_(dy * dx);
        sy *= dx * inv;
        sx *= dy * inv;
    }
    uint32_t w0 = e.lvl_width[0], oi = e.lvl_offset[0] + offx + offy * w0;
    float v00 = __ldg(e.warp + oi), v10 = __ldg(e.warp + oi + 1), v01 = __ldg(e.warp + oi + w0), v11 = __ldg(e.warp + oi + w0 + 1);
    float r0 = v00 + v10, r1 = v01 + v11;                       // warp.h:480-494 square_to_bilinear
    sy = interval_to_linear(r0, r1, sy);
    float c0 = dr_lerp(v00, v01, sy), c1 = dr_lerp(v10, v11, sy);
    sx = interval_to_linear(c0, c1, sx);
    pdf = dr_lerp(c0, c1, sx);
    ox = ((float) (int32_t) offx + sx) * e.patch_size[0];
    oy = ((float) (int32_t) offy + sy) * e.patch_size[1];
}

// distr_2d.h:706-729 Hierarchical2D::eval
PT_DEV float h2d_eval(const DevEnv &e, float px, float py) {
    px = fminf(fmaxf(px, 0.f), 1.f); py = fminf(fmaxf(py, 0.f), 1.f);
    px *= e.inv_patch_size[0]; py *= e.inv_patch_size[1];
    uint32_t ox = min((uint32_t) (int32_t) px, e.max_patch_index[0]), oy = min((uint32_t) (int32_t) py, e.max_patch_index[1]);
    px -= (float) (int32_t) ox; py -= (float) (int32_t) oy;
    uint32_t w0 = e.lvl_width[0], oi = e.lvl_offset[0] + ox + oy * w0;
    float v00 = __ldg(e.warp + oi), v10 = __ldg(e.warp + oi + 1), v01 = __ldg(e.warp + oi + w0), v11 = __ldg(e.warp + oi + w0 + 1);
    return dr_lerp(dr_lerp(v00, v10, px), dr_lerp(v01, v11, px), py);      // warp.h:516-521
}

// envmap.cpp:531-590 eval_spectrum (rgb): dr::Texture, linear filter, clamp (texture_impl.h:150-205).
// Texels are padded to float4 -> four 128-bit loads.
PT_DEV float3 env_eval_spectrum(const DevEnv &e, float uvx, float uvy) {
    float rx = (float) e.W, ry = (float) e.H;
    float u = uvx - floorf(uvx), v = fminf(fmaxf(uvy, 0.f), 1.f);
    float posx = fdiv(__fmaf_rn(u, rx, 1.f), rx + 2.f), posy = fdiv(__fmaf_rn(v, ry - 1.f, 0.5f), ry);
    int32_t resx = (int32_t) e.W + 2, resy = (int32_t) e.H;
    float fx = __fmaf_rn(posx, (float) resx, -0.5f), fy = __fmaf_rn(posy, (float) resy, -0.5f);
    int32_t ix = (int32_t) floorf(fx), iy = (int32_t) floorf(fy);
    float wx1 = fx - (float) ix, wx0 = 1.f - wx1, wy1 = fy - (float) iy, wy0 = 1.f - wy1;
    int32_t x0 = min(max(ix, 0), resx - 1), x1 = min(max(ix + 1, 0), resx - 1);
    int32_t y0 = min(max(iy, 0), resy - 1), y1 = min(max(iy + 1, 0), resy - 1);
    float4 t00 = __ldg(e.tex + (size_t) y0 * resx + x0), t10 = __ldg(e.tex + (size_t) y0 * resx + x1);
    float4 t01 = __ldg(e.tex + (size_t) y1 * resx + x0), t11 = __ldg(e.tex + (size_t) y1 * resx + x1);
    float w00 = (1.f * wx0) * wy0, w10 = (1.f * wx1) * wy0, w01 = (1.f * wx0) * wy1, w11 = (1.f * wx1) * wy1;
    float3 out = V(0.f, 0.f, 0.f);
    out = V(__fmaf_rn(t00.x, w00, out.x), __fmaf_rn(t00.y, w00, out.y), __fmaf_rn(t00.z, w00, out.z));
    out = V(__fmaf_rn(t10.x, w10, out.x), __fmaf_rn(t10.y, w10, out.y), __fmaf_rn(t10.z, w10, out.z));
    out = V(__fmaf_rn(t01.x, w01, out.x), __fmaf_rn(t01.y, w01, out.y), __fmaf_rn(t01.z, w01, out.z));
    out = V(__fmaf_rn(t11.x, w11, out.x), __fmaf_rn(t11.y, w11, out.y), __fmaf_rn(t11.z, w11, out.z));
    return out * e.scale;
}

// envmap.cpp:449-453
PT_DEV float2 env_direction_to_uv(float3 d) {
    return make_float2(dr_atan2(d.x, -d.z) * (0.5f * PT_INV_PI), dr_safe_acos(d.y) * PT_INV_PI);
}

// Emitter::eval for a ray that left the scene, d = ray direction = -si.wi
// (envmap.cpp:276-285, constant.cpp:95-98). `crad` is the radiance of a `constant` emitter.
// The three entry points below are NOT inlined and take the descriptor by pointer (global
// memory): the register budget and stack of the shading kernels stay what they are for
// scenes without an environment.
static __device__ __noinline__ float3 env_eval(const DevEnv *ep, float3 crad, float3 d) {
    const DevEnv &e = *ep;
    if (e.type == B200PT_EMITTER_CONSTANT) return crad;
    float2 uv = env_direction_to_uv(env_xform(e.mi, d));
    return env_eval_spectrum(e, uv.x, uv.y);
}

// envmap.cpp:381-396 / constant.cpp:148-152 pdf_direction (without the scene's emitter pmf)
static __device__ __noinline__ float env_pdf_direction(const DevEnv *ep, float3 d_world) {
    const DevEnv &e = *ep;
    if (e.type == B200PT_EMITTER_CONSTANT) return 0.25f * PT_INV_PI;
    float3 d = env_xform(e.mi, d_world);
    float2 uv = env_direction_to_uv(d);
    float u = uv.x - fdiv(.5f, (float) e.W), v = uv.y;
    u -= floorf(u); v -= floorf(v);
    float q = fmaxf(d.x * d.x + d.z * d.z, PT_DR_EPSILON * PT_DR_EPSILON);
    float inv_sin_theta = rsqrt_(fmaxf(q, 0.f));
    return h2d_eval(e, u, v) * inv_sin_theta * fdiv(1.f, 2.f * (PT_PI * PT_PI));
}

// envmap.cpp:336-379 / constant.cpp:119-146 sample_direction; returns radiance / pdf
static __device__ __noinline__ float3 env_sample_direction(const DevEnv *ep, float3 crad, float3 ref_p, float sx, float sy, DirectionSample &ds) {
    const DevEnv &e = *ep;
    float3 rel = ref_p - V(e.center[0], e.center[1], e.center[2]);
    float radius = fmaxf(e.radius, __fsqrt_rn(vsqnorm(rel))), dist = 2.f * radius;
    if (e.type == B200PT_EMITTER_CONSTANT) {
        float z = __fmaf_rn(-2.f, sy, 1.f), r = safe_sqrt(__fmaf_rn(-z, z, 1.f)), sn, cs;   // warp.h:250-255
        dr_sincos(2.f * PT_PI * sx, sn, cs);
        float3 d = V(r * cs, r * sn, z);
        ds.p = vfmas(d, dist, ref_p); ds.n = -d; ds.uv = make_float2(sx, sy);
        ds.pdf = 0.25f * PT_INV_PI; ds.d = d; ds.dist = dist;
        return V(fdiv(crad.x, ds.pdf), fdiv(crad.y, ds.pdf), fdiv(crad.z, ds.pdf));
    }
    float ux, uy, pdf;
    h2d_sample(e, sx, sy, ux, uy, pdf);
    ux += fdiv(.5f, (float) e.W);
    bool active = pdf > 0.f;
    float st, ct, sp, cp;
    dr_sincos(uy * PT_PI, st, ct); dr_sincos(ux * (2.f * PT_PI), sp, cp);
    float inv_sin_theta = rcp_(fmaxf(st, PT_DR_EPSILON));
    float3 d = env_xform(e.m, V(sp * st, ct, -cp * st));
    ds.p = ref_p + d * dist; ds.n = -d; ds.uv = make_float2(ux, uy);
    ds.pdf = active ? pdf * inv_sin_theta * fdiv(1.f, 2.f * (PT_PI * PT_PI)) : 0.f;
    ds.d = d; ds.dist = dist;
    if (!active) return V(0.f, 0.f, 0.f);
    float3 rad = env_eval_spectrum(e, ux, uy);
    return V(fdiv(rad.x, ds.pdf), fdiv(rad.y, ds.pdf), fdiv(rad.z, ds.pdf));
}

} // namespace pt
