// pt_device.cuh -- device-side building blocks of the wavefront path tracer:
// fp32 vector helpers, TEA/PCG32, surface-interaction reconstruction, textures,
// warps, BSDFs and area-light sampling. Each block names the reference code
// whose arithmetic (operation order, fused multiply-adds) it follows; paths are
// relative to the reference checkout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/b200pt.h"

#define PT_DEV __device__ __forceinline__

namespace pt {

// ----------------------------------------------------------------------------
// constants (core/math.h:18-22 without Embree; drjit array_constants.h)
// ----------------------------------------------------------------------------
#define PT_PI          3.14159265358979323846f
#define PT_INV_PI      0.31830988618379067154f
#define PT_RAY_EPS     (1500.f * 5.9604644775390625e-08f)
#define PT_SHADOW_EPS  (PT_RAY_EPS * 10.f)
#define PT_LARGEST     3.402823466e+38f
#define PT_INF         __int_as_float(0x7f800000)

// BSDFFlags (bsdf.h:31-125)
#define F_NULL                0x00001u
#define F_DIFFUSE_REFLECTION  0x00002u
#define F_GLOSSY_REFLECTION   0x00008u
#define F_GLOSSY_TRANSMISSION 0x00010u
#define F_DELTA_REFLECTION    0x00020u
#define F_DELTA_TRANSMISSION  0x00040u
#define F_DELTA  (F_NULL | F_DELTA_REFLECTION | F_DELTA_TRANSMISSION)
#define F_SMOOTH (0x2u | 0x4u | 0x8u | 0x10u)

// ----------------------------------------------------------------------------
// float3 helpers. vdot/vcross/vsqnorm spell out the fma chains of
// drjit/array_base.h:671-690 and array_router.h:658-689.
// ----------------------------------------------------------------------------
PT_DEV float3 V(float x, float y, float z) { return make_float3(x, y, z); }
PT_DEV float3 operator+(float3 a, float3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_DEV float3 operator-(float3 a, float3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_DEV float3 operator*(float3 a, float3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_DEV float3 operator*(float3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
PT_DEV float3 operator-(float3 a) { return V(-a.x, -a.y, -a.z); }
PT_DEV float3 vfma(float3 a, float3 b, float3 c) { return V(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y), __fmaf_rn(a.z, b.z, c.z)); }
PT_DEV float3 vfmas(float3 a, float s, float3 c) { return V(__fmaf_rn(a.x, s, c.x), __fmaf_rn(a.y, s, c.y), __fmaf_rn(a.z, s, c.z)); }
PT_DEV float vdot(float3 a, float3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
PT_DEV float3 vcross(float3 a, float3 b) {
    return V(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
PT_DEV float vsqnorm(float3 a) { return __fmaf_rn(a.z, a.z, __fmaf_rn(a.y, a.y, __fmul_rn(a.x, a.x))); }
PT_DEV float rcp_(float x) { return __frcp_rn(x); }
PT_DEV float rsqrt_(float x) { return __frcp_rn(__fsqrt_rn(x)); }
PT_DEV float3 vnormalize(float3 a) { return a * rsqrt_(vsqnorm(a)); }
PT_DEV float vmaxc(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
PT_DEV float mulsign(float a, float b) { return __int_as_float(__float_as_int(a) ^ (__float_as_int(b) & 0x80000000)); }
PT_DEV float mulsign_neg(float a, float b) { return __int_as_float(__float_as_int(a) ^ (~__float_as_int(b) & 0x80000000)); }
PT_DEV float safe_sqrt(float x) { return __fsqrt_rn(fmaxf(x, 0.f)); }
PT_DEV float sqr(float x) { return __fmul_rn(x, x); }
PT_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// ----------------------------------------------------------------------------
// RNG (core/random.h:77-90, drjit/random.h:135-161,287-289, sampler.cpp:129-148)
// ----------------------------------------------------------------------------
PT_DEV void tea32(uint32_t v0, uint32_t v1, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0; o1 = v1;
}

struct Pcg32 {
    uint64_t state, inc;
    PT_DEV uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
    }
    PT_DEV float next_f32() { return __uint_as_float((next_u32() >> 9) | 0x3f800000u) - 1.f; }
    // per-lane stream of the JIT variants: (v0, v1) = TEA(seed, lane); seed(v0, v1)
    PT_DEV void seed_lane(uint32_t seed_value, uint32_t lane) {
        uint32_t v0, v1; tea32(seed_value, lane, v0, v1);
        state = 0; inc = (((uint64_t) v1) << 1) | 1u;
        next_u32(); state += (uint64_t) v0; next_u32();
    }
    // `inc` is not stored in the wavefront state: it is a pure function of the lane
    PT_DEV void restore(uint32_t seed_value, uint32_t lane, uint64_t st) {
        uint32_t v0, v1; tea32(seed_value, lane, v0, v1);
        inc = (((uint64_t) v1) << 1) | 1u; state = st;
    }
};

// ----------------------------------------------------------------------------
// Device scene
// ----------------------------------------------------------------------------
struct DevTexture {
    int32_t kind, channels, width, height, wrap, filter, differentiable;
    uint32_t grad_offset;      // into the flat gradient buffer (floats)
    float value[3];
    float value1[3];           // checkerboard color1
    float to_uv[9];
    const float *data;
};

struct DevBsdf {
    int32_t type, twosided;
    int32_t tex[B200PT_MAX_SLOTS];
    float eta, spec_srate, clearcoat_srate, diff_refl_srate;
    uint32_t flags;
    float plastic_fdr_int, plastic_spec_weight;
};
// internal: `plastic` shades in the conductor queue (type = B200PT_BSDF_CONDUCTOR + this flag)
#define PT_M_PLASTIC (1u << 24)

struct DevShape {
    uint32_t layout; int32_t bsdf, emitter, sampling;
    uint32_t first_prim, n_prims, first_vertex, pad;
    float to_world[16];
    float frame_n[3]; float inv_area;
    float area_sum, area_norm; const float *area_cdf, *area_pmf; // B200PT_SAMPLING_MESH (core/distr_1d.h)
};

struct DevEmitter { int32_t shape, radiance_tex; float sampling_weight; int32_t type; };

// The environment emitter of the scene (pt_env.cuh). `tex` holds the lat-long map with its
// periodic halo columns, one float4 per texel; `warp` is the Hierarchical2D buffer
// (level 0 row-major, levels >= 1 in 2x2 blocks, distr_2d.h:431-450).
constexpr int ENV_MAX_LEVELS = 20;
struct DevEnv {
    int32_t type;                 // -1: none, else B200PT_EMITTER_CONSTANT / _ENVMAP
    int32_t emitter_index, radiance_tex;
    uint32_t W, H;                // real resolution
    const float4 *tex;            // H x (W + 2)
    const float *warp;
    float scale;
    float m[9], mi[9];            // linear part of to_world / its inverse, row-major
    float center[3], radius;      // bounding sphere of the scene (envmap.cpp:260-274)
    uint32_t n_levels, lvl_width[ENV_MAX_LEVELS], lvl_offset[ENV_MAX_LEVELS];
    float patch_size[2], inv_patch_size[2];
    uint32_t max_patch_index[2];
};

struct DevScene {
    // acceleration structure (bvh.h)
    const float4 *nodes; uint32_t n_nodes; uint32_t n_tris;
    const float4 *tris;        // 3 float4 per triangle in leaf order
    // shading data in (shape, prim) order
    const uint4  *prim_verts;  // v0, v1, v2 (global vertex index), shape
    const uint32_t *uv_flipped; // one bit per primitive: FaceUVFlipped of its face record (mesh.cpp:634-644); nullptr when no mesh packs tangents
    const float4 *vertices;    // 2 float4 per vertex: (p.xyz, n.x) (n.y, n.z, u, v)
    const DevShape *shapes; const DevBsdf *bsdfs; const DevEmitter *emitters; const DevTexture *textures;
    uint32_t n_shapes, n_bsdfs, n_emitters, n_textures;
    float *grad;               // flat gradient buffer of all differentiable textures
    const float *tangent;      // forward mode: d(parameter) in the same layout (b200pt_tangent_write)
    // The four tables above live in ONE contiguous blob (shapes | bsdfs | emitters | textures);
    // kernels stage it into shared memory when it is small (kernels.cu: stage_tables).
    const unsigned char *tables; uint32_t tables_bytes, off_bsdfs, off_emitters, off_textures;
    uint32_t geom_bytes, n_vertices;   // bytes of prim_verts + vertices (staged together when small)
    // sensor + film
    float s2c[16]; float cam_to_world[16];
    float near_clip, far_clip;
    uint32_t film_w, film_h, crop_w, crop_h, crop_x, crop_y;
    int32_t rfilter; float gauss_radius, gauss_alpha, gauss_bias; float gauss_coeff[10];
    uint32_t base_seed;
    // environment emitter: the descriptor lives in global memory (the out-of-line functions of
    // pt_env.cuh take the pointer, so kernels never spill a copy of the scene for them)
    const DevEnv *env; int32_t env_type, env_emitter, env_radiance_tex;   // env_type -1: none; radiance_tex: constant rgb / envmap `data`
    float env_scale;
    // Scene::m_emitter_distr (scene.cpp:120-140): set only when some sampling_weight != 1
    const float *em_cdf, *em_pmf; float em_sum, em_norm;
};

struct Ray { float3 o, d; float maxt; };

struct SurfaceInteraction {
    float t; float3 p, n, sh_s, sh_t, sh_n, wi; float2 uv; uint32_t shape;
    PT_DEV float3 to_local(float3 v) const { return V(vdot(v, sh_s), vdot(v, sh_t), vdot(v, sh_n)); }
    // frame.h:39-41
    PT_DEV float3 to_world(float3 v) const { return vfmas(sh_n, v.z, vfmas(sh_t, v.y, sh_s * v.x)); }
};

// vector.h:118-138 coordinate_system
PT_DEV void coordinate_system(float3 n, float3 &s, float3 &t) {
    float sign = copysignf(1.f, n.z), a = -rcp_(sign + n.z), b = n.x * n.y * a;
    s = V(mulsign(sqr(n.x) * a, n.z) + 1.f, mulsign(b, n.z), mulsign_neg(n.x, n.z));
    t = V(b, __fmaf_rn(n.y, n.y * a, sign), -n.y);
}

// mesh.h:1132-1153 moeller_trumbore with precomputed edges
PT_DEV bool moeller_trumbore(float3 o, float3 d, float maxt, float3 p0, float3 e1, float3 e2, float &t, float &u, float &v) {
    float3 pvec = vcross(d, e2);
    float inv_det = rcp_(vdot(e1, pvec));
    float3 tvec = o - p0;
    u = vdot(tvec, pvec) * inv_det;
    bool active = u >= 0.f && u <= 1.f;
    float3 qvec = vcross(tvec, e1);
    v = vdot(d, qvec) * inv_det;
    active = active && v >= 0.f && u + v <= 1.f;
    t = vdot(e2, qvec) * inv_det;
    return active && t >= 0.f && t <= maxt;
}

// mesh_utils.h:89-118 frame_decode: the three floats of a packed tangent frame (modified Rodrigues parameters of the frame's
// quaternion) -> unit normal (third column of the rotation) and tangent (first column)
PT_DEV void frame_decode(float3 p, float3 &n, float3 &s) {
    float a = rcp_(1.f + vsqnorm(p)), A = 8.f * sqr(a), B = __fmaf_rn(-4.f, a, A);
    float X = A * p.x, Y = A * p.y, Z = A * p.z, yy = p.y * Y, xy = p.x * Y, xz = p.x * Z, yz = p.y * Z, u = 1.f - yy;
    n = V(__fmaf_rn(B, p.y, xz), __fmaf_rn(-B, p.x, yz), __fmaf_rn(-p.x, X, u));
    s = V(__fmaf_rn(-p.z, Z, u), __fmaf_rn(B, p.z, xy), __fmaf_rn(-B, p.y, xz));
}

// interaction.h:804-830 + mesh.cpp:2254-2437 + interaction.h:558-603
PT_DEV SurfaceInteraction compute_si(const DevScene &sc, float t, float b1, float b2, uint32_t prim, float3 ray_d) {
    SurfaceInteraction si;
    uint4 pv = sc.prim_verts[prim];
    const DevShape &sh = sc.shapes[pv.w];
    float4 a0 = sc.vertices[2 * pv.x], a1 = sc.vertices[2 * pv.x + 1];
    float4 c0 = sc.vertices[2 * pv.y], c1 = sc.vertices[2 * pv.y + 1];
    float4 g0 = sc.vertices[2 * pv.z], g1 = sc.vertices[2 * pv.z + 1];
    float3 p0 = V(a0.x, a0.y, a0.z), p1 = V(c0.x, c0.y, c0.z), p2 = V(g0.x, g0.y, g0.z);
    float b0 = 1.f - b1 - b2;
    float3 e1 = p1 - p0, e2 = p2 - p0;
    si.p = vfmas(p0, b0, vfmas(p1, b1, p2 * b2));   // mesh.cpp:2300
    si.n = vnormalize(vcross(e1, e2));
    si.t = t;
    const bool tangents = (sh.layout & B200PT_LAYOUT_TANGENTS) != 0;     // mesh.cpp:2274 need_tangents (Shading is always requested)
    float3 sh_s = V(0.f, 0.f, 0.f);
    if (sh.layout & B200PT_LAYOUT_NORMALS) {
        float3 n0 = V(a0.w, a1.x, a1.y), n1 = V(c0.w, c1.x, c1.y), n2 = V(g0.w, g1.x, g1.y);
        if (tangents) {
            // packed tangent frame (mesh.cpp:2339-2351, 2417-2424): one decode per vertex yields the normal and the tangent
            float3 t0, t1, t2;
            frame_decode(n0, n0, t0); frame_decode(n1, n1, t1); frame_decode(n2, n2, t2);
            sh_s = vfmas(t1 - t0, b1, vfmas(t2 - t0, b2, t0));
        }
        float3 dn1 = n1 - n0, dn2 = n2 - n0;
        float3 n = vfmas(dn1, b1, vfmas(dn2, b2, n0));
        si.sh_n = n * rsqrt_(vsqnorm(n));
    } else {
        si.sh_n = si.n;
    }
    if (sh.layout & B200PT_LAYOUT_TEXCOORDS) {
        float u0 = a1.z, v0 = a1.w;
        float du0 = c1.z - u0, dv0 = c1.w - v0, du1 = g1.z - u0, dv1 = g1.w - v0;
        si.uv = make_float2(__fmaf_rn(du0, b1, __fmaf_rn(du1, b2, u0)), __fmaf_rn(dv0, b1, __fmaf_rn(dv1, b2, v0)));
    } else {
        si.uv = make_float2(b1, b2);
    }
    // finalize_surface_interaction (interaction.h:571-597): orthogonalise the shape's tangent against the shading normal; a zero
    // tangent (no packed frames) falls back to coordinate_system; the bitangent follows the orientation of the parameterisation
    float3 s_o = vfmas(si.sh_n, -vdot(si.sh_n, sh_s), sh_s);
    float sqr_norm = vsqnorm(s_o);
    if (tangents && sqr_norm > 0.f) {
        si.sh_s = s_o * rsqrt_(sqr_norm);
        float3 t2 = vcross(si.sh_n, si.sh_s);
        const bool flipped = (sc.uv_flipped[prim >> 5] >> (prim & 31u)) & 1u;
        si.sh_t = flipped ? V(-t2.x, -t2.y, -t2.z) : t2;
    } else {
        coordinate_system(si.sh_n, si.sh_s, si.sh_t);
    }
    si.wi = si.to_local(-ray_d);
    si.shape = pv.w;
    return si;
}

// interaction.h:161-190
PT_DEV float3 offset_p(float3 p, float3 n, float3 d) {
    float mag = (1.f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * PT_RAY_EPS;
    mag = mulsign(mag, vdot(n, d));
    return vfmas(n, mag, p);
}
PT_DEV Ray spawn_ray(float3 p, float3 n, float3 d) { Ray r; r.o = offset_p(p, n, d); r.d = d; r.maxt = PT_LARGEST; return r; }
PT_DEV Ray spawn_ray_to(float3 p, float3 n, float3 t) {
    Ray r; r.o = offset_p(p, n, t - p);
    float3 d = t - r.o; float dist = __fsqrt_rn(vsqnorm(d));
    r.d = V(fdiv(d.x, dist), fdiv(d.y, dist), fdiv(d.z, dist));
    r.maxt = dist * (1.f - PT_SHADOW_EPS);
    return r;
}

// ----------------------------------------------------------------------------
// Textures (bitmap.cpp:496-519, drjit/texture_impl.h:87-205)
// ----------------------------------------------------------------------------
PT_DEV int32_t tex_wrap(int32_t pos, int32_t shape, int mode) {
    if (mode == B200PT_WRAP_CLAMP) return min(max(pos, 0), shape - 1);
    int32_t div = (pos < 0 ? pos + 1 : pos) / shape;
    if (pos < 0) div -= 1;
    int32_t mod = pos - div * shape;
    if (mode == B200PT_WRAP_MIRROR && (div & 1)) mod = shape - 1 - mod;
    return mod;
}
struct TexTaps { int32_t idx[4]; float w[4]; int n; };
// internal wrap mode of the envmap's `data` texture: taps of EnvironmentMapEmitter::eval_spectrum
// (envmap.cpp:531-590) in the halo storage, mapped back to the REAL column of the parameter
// (halo col 0 <- last real column, col W+1 <- first; envmap.cpp:228-246 routes gradients the same way)
#define PT_WRAP_ENVMAP 100
PT_DEV void tex_lookup(const DevTexture &t, float2 uv, TexTaps &tp) {
    if (t.wrap == PT_WRAP_ENVMAP) {
        int32_t W = t.width, H = t.height, resx = W + 2;
        float rx = (float) W, ry = (float) H;
        float u = uv.x - floorf(uv.x), v = fminf(fmaxf(uv.y, 0.f), 1.f);
        float posx = fdiv(__fmaf_rn(u, rx, 1.f), rx + 2.f), posy = fdiv(__fmaf_rn(v, ry - 1.f, 0.5f), ry);
        float fx = __fmaf_rn(posx, (float) resx, -0.5f), fy = __fmaf_rn(posy, (float) H, -0.5f);
        int32_t ix = (int32_t) floorf(fx), iy = (int32_t) floorf(fy);
        float wx1 = fx - (float) ix, wx0 = 1.f - wx1, wy1 = fy - (float) iy, wy0 = 1.f - wy1;
        int32_t xs0 = min(max(ix, 0), resx - 1), xs1 = min(max(ix + 1, 0), resx - 1);
        int32_t y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
        int32_t x0 = xs0 == 0 ? W - 1 : (xs0 == W + 1 ? 0 : xs0 - 1), x1 = xs1 == 0 ? W - 1 : (xs1 == W + 1 ? 0 : xs1 - 1);
        tp.n = 4;
        tp.idx[0] = y0 * W + x0; tp.w[0] = (1.f * wx0) * wy0; tp.idx[1] = y0 * W + x1; tp.w[1] = (1.f * wx1) * wy0;
        tp.idx[2] = y1 * W + x0; tp.w[2] = (1.f * wx0) * wy1; tp.idx[3] = y1 * W + x1; tp.w[3] = (1.f * wx1) * wy1;
        return;
    }
    float u = __fmaf_rn(t.to_uv[1], uv.y, __fmaf_rn(t.to_uv[0], uv.x, t.to_uv[2]));
    float v = __fmaf_rn(t.to_uv[4], uv.y, __fmaf_rn(t.to_uv[3], uv.x, t.to_uv[5]));
    int32_t W = t.width, H = t.height;
    if (t.filter == B200PT_FILTER_NEAREST) {
        int32_t px = tex_wrap((int32_t) floorf(u * (float) W), W, t.wrap), py = tex_wrap((int32_t) floorf(v * (float) H), H, t.wrap);
        tp.n = 1; tp.idx[0] = py * W + px; tp.w[0] = 1.f; return;
    }
    float fx = __fmaf_rn(u, (float) W, -.5f), fy = __fmaf_rn(v, (float) H, -.5f);
    float flx = floorf(fx), fly = floorf(fy);
    int32_t ix = (int32_t) flx, iy = (int32_t) fly;
    float w1x = fx - flx, w1y = fy - fly, w0x = 1.f - w1x, w0y = 1.f - w1y;
    int32_t x0 = tex_wrap(ix, W, t.wrap), x1 = tex_wrap(ix + 1, W, t.wrap);
    int32_t y0 = tex_wrap(iy, H, t.wrap), y1 = tex_wrap(iy + 1, H, t.wrap);
    tp.n = 4;
    tp.idx[0] = y0 * W + x0; tp.w[0] = w0x * w0y; tp.idx[1] = y0 * W + x1; tp.w[1] = w1x * w0y;
    tp.idx[2] = y1 * W + x0; tp.w[2] = w0x * w1y; tp.idx[3] = y1 * W + x1; tp.w[3] = w1x * w1y;
}
// checkerboard.cpp:70-110: which of the two colours is seen at uv. eval() picks color0 where the two
// half-cell masks are EQUAL, eval_1() where they DIFFER (the reference's own asymmetry, kept).
PT_DEV bool checker_masks_equal(const DevTexture &t, float2 uv) {
    float u = __fmaf_rn(t.to_uv[1], uv.y, __fmaf_rn(t.to_uv[0], uv.x, t.to_uv[2]));
    float v = __fmaf_rn(t.to_uv[4], uv.y, __fmaf_rn(t.to_uv[3], uv.x, t.to_uv[5]));
    bool mx = u - floorf(u) > .5f, my = v - floorf(v) > .5f;
    return mx == my;
}
PT_DEV float3 tex_eval3(const DevScene &sc, int32_t tex, float2 uv) {
    if (tex < 0) return V(0.f, 0.f, 0.f);
    const DevTexture &t = sc.textures[tex];
    if (t.kind == B200PT_TEX_CONST) return t.channels == 1 ? V(t.value[0], t.value[0], t.value[0]) : V(t.value[0], t.value[1], t.value[2]);
    if (t.kind == B200PT_TEX_CHECKERBOARD) {
        const float *c = checker_masks_equal(t, uv) ? t.value : t.value1;
        return t.channels == 1 ? V(c[0], c[0], c[0]) : V(c[0], c[1], c[2]);
    }
    TexTaps tp; tex_lookup(t, uv, tp);
    float out[3]; int C = t.channels;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int cc = C == 1 ? 0 : c;
        if (tp.n == 1) { out[c] = __ldg(&t.data[(size_t) tp.idx[0] * C + cc]); continue; }
        float v00 = __ldg(&t.data[(size_t) tp.idx[0] * C + cc]), v01 = __ldg(&t.data[(size_t) tp.idx[1] * C + cc]);
        float v10 = __ldg(&t.data[(size_t) tp.idx[2] * C + cc]), v11 = __ldg(&t.data[(size_t) tp.idx[3] * C + cc]);
        out[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tp.w[0], v00), __fmul_rn(tp.w[1], v01)), __fmul_rn(tp.w[2], v10)), __fmul_rn(tp.w[3], v11));
    }
    return V(out[0], out[1], out[2]);
}
PT_DEV float tex_eval1(const DevScene &sc, int32_t tex, float2 uv) {
    if (tex >= 0 && sc.textures[tex].kind == B200PT_TEX_CHECKERBOARD) {
        const DevTexture &t = sc.textures[tex];
        return checker_masks_equal(t, uv) ? t.value1[0] : t.value[0];      // eval_1: color0 where the masks differ
    }
    return tex_eval3(sc, tex, uv).x;
}

// fp32 atomicAdd scatter of a texture-parameter gradient (adjoint of tex_eval3)
PT_DEV void tex_scatter3(const DevScene &sc, int32_t tex, float2 uv, float3 g) {
    if (tex < 0) return;
    const DevTexture &t = sc.textures[tex];
    if (!t.differentiable) return;
    float *grad = sc.grad + t.grad_offset;
    float gv[3] = { g.x, g.y, g.z };
    if (t.kind == B200PT_TEX_CONST) {
        if (t.channels == 1) atomicAdd(grad, gv[0] + gv[1] + gv[2]);
        else { atomicAdd(grad + 0, gv[0]); atomicAdd(grad + 1, gv[1]); atomicAdd(grad + 2, gv[2]); }
        return;
    }
    TexTaps tp; tex_lookup(t, uv, tp);
    int C = t.channels;
    for (int k = 0; k < tp.n; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(grad + (size_t) tp.idx[k] * C + (C == 1 ? 0 : c), tp.w[k] * gv[c]);
}

// ----------------------------------------------------------------------------
// Warps (warp.h:54-89, 412-420)
// ----------------------------------------------------------------------------
PT_DEV float3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = __fmaf_rn(2.f, sx, -1.f), y = __fmaf_rn(2.f, sy, -1.f);
    bool is_zero = (x == 0.f) && (y == 0.f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = fdiv(__fmul_rn(__fmul_rn(0.25f, PT_PI), rp), r);
    if (q13) phi = __fmul_rn(0.5f, PT_PI) - phi;
    if (is_zero) phi = 0.f;
    float s, c; sincosf(phi, &s, &c);
    float px = r * c, py = r * s;
    float z = safe_sqrt(1.f - __fmaf_rn(py, py, __fmul_rn(px, px)));
    return V(px, py, z);
}
PT_DEV float2 square_to_uniform_triangle(float sx, float sy) {
    float t = safe_sqrt(1.f - sx);
    return make_float2(1.f - t, t * sy);
}

// ----------------------------------------------------------------------------
// Fresnel (fresnel.h:35-117)
// ----------------------------------------------------------------------------
PT_DEV void fresnel(float cos_theta_i, float eta, float &r, float &cos_theta_t, float &eta_it, float &eta_ti) {
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = rcp_(eta);
    eta_it = outside ? eta : rcp_eta; eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = __fmaf_rn(-__fmaf_rn(-cos_theta_i, cos_theta_i, 1.f), __fmul_rn(eta_ti, eta_ti), 1.f);
    float cti = fabsf(cos_theta_i), ctt = safe_sqrt(cos_theta_t_sqr);
    bool index_matched = eta == 1.f, special = index_matched || cti == 0.f;
    float a_s = fdiv(__fmaf_rn(-eta_it, ctt, cti), __fmaf_rn(eta_it, ctt, cti));
    float a_p = fdiv(__fmaf_rn(-eta_it, cti, ctt), __fmaf_rn(eta_it, cti, ctt));
    r = __fmul_rn(0.5f, __fadd_rn(sqr(a_s), sqr(a_p)));
    if (special) r = index_matched ? 0.f : 1.f;
    cos_theta_t = mulsign_neg(ctt, cos_theta_i);
}
PT_DEV float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    float c2 = __fmul_rn(cos_theta_i, cos_theta_i), s2 = __fsub_rn(1.f, c2), s4 = __fmul_rn(s2, s2);
    float temp_1 = __fsub_rn(__fsub_rn(__fmul_rn(eta_r, eta_r), __fmul_rn(eta_i, eta_i)), s2);
    float a_2_pb_2 = safe_sqrt(__fadd_rn(__fmul_rn(temp_1, temp_1), __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(4.f, eta_i), eta_i), eta_r), eta_r)));
    float a = safe_sqrt(__fmul_rn(.5f, __fadd_rn(a_2_pb_2, temp_1)));
    float term_1 = __fadd_rn(a_2_pb_2, c2), term_2 = __fmul_rn(__fmul_rn(2.f, cos_theta_i), a);
    float r_s = fdiv(__fsub_rn(term_1, term_2), __fadd_rn(term_1, term_2));
    float term_3 = __fadd_rn(__fmul_rn(a_2_pb_2, c2), s4), term_4 = __fmul_rn(term_2, s2);
    float r_p = fdiv(__fmul_rn(r_s, __fsub_rn(term_3, term_4)), __fadd_rn(term_3, term_4));
    return __fmul_rn(0.5f, __fadd_rn(r_s, r_p));
}

// ----------------------------------------------------------------------------
// BSDFs. One branch-flattened implementation per material type; the shading
// kernels are instantiated per type (template argument), so that a warp of a
// material queue executes a single model.
// ----------------------------------------------------------------------------
struct BsdfSample { float3 wo; float pdf, eta; uint32_t sampled_type, sampled_component; };
struct BsdfResult { float3 value; float pdf; BsdfSample bs; float3 weight; };

PT_DEV uint32_t bsdf_flags(const DevBsdf &b) {
    switch (b.type) {
        case B200PT_BSDF_DIFFUSE: return F_DIFFUSE_REFLECTION;
        case B200PT_BSDF_CONDUCTOR: return (b.flags & PT_M_PLASTIC) ? (F_DELTA_REFLECTION | F_DIFFUSE_REFLECTION) : (b.flags & B200PT_M_ROUGH) ? F_GLOSSY_REFLECTION : F_DELTA_REFLECTION;
        case B200PT_BSDF_DIELECTRIC: return (b.flags & B200PT_M_ROUGH) ? (F_GLOSSY_REFLECTION | F_GLOSSY_TRANSMISSION) : (F_DELTA_REFLECTION | F_DELTA_TRANSMISSION);
        default: return F_DIFFUSE_REFLECTION | F_GLOSSY_REFLECTION | ((b.flags & B200PT_P_HAS_SPEC_TRANS) ? F_GLOSSY_TRANSMISSION : 0u);
    }
}

} // namespace pt

#include "pt_principled.cuh"
#include "pt_rough.cuh"
#include "pt_bsdf_grad.cuh"

namespace pt {

template <int TYPE>
PT_DEV void bsdf_eval_pdf_inner(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float3 &value, float &pdf) {
    value = V(0.f, 0.f, 0.f); pdf = 0.f;
    if (TYPE == B200PT_BSDF_DIFFUSE) {
        // diffuse.cpp:160-179
        if (wi.z > 0.f && wo.z > 0.f) {
            float3 refl = tex_eval3(sc, b.tex[B200PT_SLOT_REFLECTANCE], uv);
            value = (refl * PT_INV_PI) * wo.z;
            pdf = PT_INV_PI * wo.z;
        }
    } else if (TYPE == B200PT_BSDF_PRINCIPLED) {
        principled_eval_pdf(sc, b, uv, wi, wo, value, pdf);
    } else if (TYPE == B200PT_BSDF_CONDUCTOR) {
        if (b.flags & PT_M_PLASTIC) plastic_eval_pdf(sc, b, uv, wi, wo, value, pdf);
        else if (b.flags & B200PT_M_ROUGH) roughconductor_eval_pdf(sc, b, uv, wi, wo, value, pdf); // smooth: delta lobe, 0
    } else if (TYPE == B200PT_BSDF_DIELECTRIC) {
        if (b.flags & B200PT_M_ROUGH) roughdielectric_eval_pdf(sc, b, uv, wi, wo, value, pdf);
    }
}

template <int TYPE>
PT_DEV void bsdf_sample_inner(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float s1, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    bs.wo = V(0.f, 0.f, 0.f); bs.pdf = 0.f; bs.eta = 0.f; bs.sampled_type = 0; bs.sampled_component = 0;
    weight = V(0.f, 0.f, 0.f);
    if (TYPE == B200PT_BSDF_DIFFUSE) {
        // diffuse.cpp:100-123
        if (!(wi.z > 0.f)) return;
        bs.wo = square_to_cosine_hemisphere(s2x, s2y);
        bs.pdf = PT_INV_PI * bs.wo.z;
        bs.eta = 1.f; bs.sampled_type = F_DIFFUSE_REFLECTION;
        if (bs.pdf > 0.f) weight = tex_eval3(sc, b.tex[B200PT_SLOT_REFLECTANCE], uv);
    } else if (TYPE == B200PT_BSDF_CONDUCTOR) {
        if (b.flags & PT_M_PLASTIC) { plastic_sample(sc, b, uv, wi, s1, s2x, s2y, bs, weight); return; }
        if (b.flags & B200PT_M_ROUGH) { roughconductor_sample(sc, b, uv, wi, s2x, s2y, bs, weight); return; }
        // conductor.cpp:247-307
        if (!(wi.z > 0.f)) return;
        bs.sampled_type = F_DELTA_REFLECTION; bs.wo = V(-wi.x, -wi.y, wi.z); bs.eta = 1.f; bs.pdf = 1.f;
        float3 eta = tex_eval3(sc, b.tex[B200PT_SLOT_ETA], uv), k = tex_eval3(sc, b.tex[B200PT_SLOT_K], uv);
        float3 refl = b.tex[B200PT_SLOT_SPEC_REFL] >= 0 ? tex_eval3(sc, b.tex[B200PT_SLOT_SPEC_REFL], uv) : V(1.f, 1.f, 1.f);
        weight = V(refl.x * fresnel_conductor(wi.z, eta.x, k.x), refl.y * fresnel_conductor(wi.z, eta.y, k.y), refl.z * fresnel_conductor(wi.z, eta.z, k.z));
    } else if (TYPE == B200PT_BSDF_DIELECTRIC) {
        if (b.flags & B200PT_M_ROUGH) { roughdielectric_sample(sc, b, uv, wi, s1, s2x, s2y, bs, weight); return; }
        // dielectric.cpp:245-370
        float r_i, ctt, eta_it, eta_ti;
        fresnel(wi.z, b.eta, r_i, ctt, eta_it, eta_ti);
        float t_i = 1.f - r_i;
        bool sel_r = s1 <= r_i;
        bs.pdf = sel_r ? r_i : t_i;
        bs.sampled_component = sel_r ? 0 : 1;
        bs.sampled_type = sel_r ? F_DELTA_REFLECTION : F_DELTA_TRANSMISSION;
        bs.wo = sel_r ? V(-wi.x, -wi.y, wi.z) : V(-eta_ti * wi.x, -eta_ti * wi.y, ctt);
        bs.eta = sel_r ? 1.f : eta_it;
        float3 w = V(1.f, 1.f, 1.f);
        if (sel_r) { if (b.tex[B200PT_SLOT_D_SPEC_REFL] >= 0) w = w * tex_eval3(sc, b.tex[B200PT_SLOT_D_SPEC_REFL], uv); }
        else {
            if (b.tex[B200PT_SLOT_D_SPEC_TRANS] >= 0) w = w * tex_eval3(sc, b.tex[B200PT_SLOT_D_SPEC_TRANS], uv);
            w = w * sqr(eta_ti);
        }
        weight = w;
    } else {
        principled_sample(sc, b, uv, wi, s1, s2x, s2y, bs, weight);
    }
}

// BSDF::eval_pdf_sample (bsdf.cpp:21-31) including the `twosided` adapter
// (twosided.cpp:60-230, same BSDF on both sides)
template <int TYPE>
PT_DEV BsdfResult bsdf_eval_pdf_sample(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float s1, float s2x, float s2y) {
    BsdfResult r;
    if (b.twosided) {
        bool back = wi.z < 0.f;
        if (wi.z == 0.f) {
            r.value = V(0.f, 0.f, 0.f); r.pdf = 0.f; r.weight = V(0.f, 0.f, 0.f);
            r.bs.wo = V(0.f, 0.f, 0.f); r.bs.pdf = 0.f; r.bs.eta = 0.f; r.bs.sampled_type = 0; r.bs.sampled_component = 0;
            return r;
        }
        if (back) { wi.z = -wi.z; wo.z = -wo.z; }
        bsdf_eval_pdf_inner<TYPE>(sc, b, uv, wi, wo, r.value, r.pdf);
        bsdf_sample_inner<TYPE>(sc, b, uv, wi, s1, s2x, s2y, r.bs, r.weight);
        if (back) r.bs.wo.z = -r.bs.wo.z;
        return r;
    }
    bsdf_eval_pdf_inner<TYPE>(sc, b, uv, wi, wo, r.value, r.pdf);
    bsdf_sample_inner<TYPE>(sc, b, uv, wi, s1, s2x, s2y, r.bs, r.weight);
    return r;
}

// ----------------------------------------------------------------------------
// Area lights
// ----------------------------------------------------------------------------
struct DirectionSample { float3 p, n, d; float2 uv; float pdf, dist; int32_t emitter; };

// transform.h:324-335
PT_DEV float3 xform_point_affine(const float *m, float3 p) {
    float r0 = m[3], r1 = m[7], r2 = m[11];
    r0 = __fmaf_rn(m[0], p.x, r0); r1 = __fmaf_rn(m[4], p.x, r1); r2 = __fmaf_rn(m[8], p.x, r2);
    r0 = __fmaf_rn(m[1], p.y, r0); r1 = __fmaf_rn(m[5], p.y, r1); r2 = __fmaf_rn(m[9], p.y, r2);
    r0 = __fmaf_rn(m[2], p.z, r0); r1 = __fmaf_rn(m[6], p.z, r1); r2 = __fmaf_rn(m[10], p.z, r2);
    return V(r0, r1, r2);
}

// Shape::sample_position (rectangle.cpp:159-172 / mesh.cpp:1662-1712)
PT_DEV void shape_sample_position(const DevScene &sc, const DevShape &sh, float sx, float sy, float3 &p, float3 &n, float &pdf, float2 &uv) {
    if (sh.sampling == B200PT_SAMPLING_RECTANGLE) {
        p = xform_point_affine(sh.to_world, V(__fmaf_rn(sx, 2.f, -1.f), __fmaf_rn(sy, 2.f, -1.f), 0.f));
        n = V(sh.frame_n[0], sh.frame_n[1], sh.frame_n[2]); pdf = sh.inv_area; uv = make_float2(sx, sy);
        return;
    }
    // DiscreteDistribution::sample_reuse (core/distr_1d.h:137-183)
    float value = sy * sh.area_sum;
    uint32_t lo = 0, hi = sh.n_prims - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) / 2; float c = __ldg(&sh.area_cdf[mid]);
        if (((c < value) || c == 0.f) && c != sh.area_sum) lo = mid + 1; else hi = mid;
    }
    uint32_t face = lo;
    float pmf_n = __ldg(&sh.area_pmf[face]) * sh.area_norm, cdf_n = face ? __ldg(&sh.area_cdf[face - 1]) * sh.area_norm : 0.f;
    float sy_re = fdiv(sy - cdf_n, pmf_n);
    uint4 pv = sc.prim_verts[sh.first_prim + face];
    float4 a0 = sc.vertices[2 * pv.x], a1 = sc.vertices[2 * pv.x + 1];
    float4 c0 = sc.vertices[2 * pv.y], c1 = sc.vertices[2 * pv.y + 1];
    float4 g0 = sc.vertices[2 * pv.z], g1 = sc.vertices[2 * pv.z + 1];
    float3 p0 = V(a0.x, a0.y, a0.z), e0 = V(c0.x, c0.y, c0.z) - p0, e1 = V(g0.x, g0.y, g0.z) - p0;
    float2 b = square_to_uniform_triangle(sx, sy_re);
    p = vfmas(e0, b.x, vfmas(e1, b.y, p0));
    pdf = sh.area_norm;
    float b0 = 1.f - b.x - b.y;
    if (sh.layout & B200PT_LAYOUT_TEXCOORDS)
        uv = make_float2(__fmaf_rn(a1.z, b0, __fmaf_rn(c1.z, b.x, g1.z * b.y)), __fmaf_rn(a1.w, b0, __fmaf_rn(c1.w, b.x, g1.w * b.y)));
    else uv = b;
    if (sh.layout & B200PT_LAYOUT_NORMALS)
        n = vnormalize(vfmas(V(a0.w, a1.x, a1.y), b0, vfmas(V(c0.w, c1.x, c1.y), b.x, V(g0.w, g1.x, g1.y) * b.y)));
    else n = vnormalize(vcross(e0, e1));
}

} // namespace pt
#include "pt_env.cuh"
namespace pt {

// m_emitter_distr->sample_reuse_pmf (scene.cpp:257-260, core/distr_1d.h:137-216): (index, pmf, reused
// sample). Out of line: scenes with uniform emitter selection keep their register budget.
static __device__ __noinline__ float3 sample_emitter_weighted(const float *cdf, const float *pmf_w, float sum, float norm, uint32_t n, float sx) {
    if (n < 2) return V(__uint_as_float(0u), __ldg(&pmf_w[0]) * norm, sx);
    float value = sx * sum;
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) { uint32_t mid = (lo + hi) / 2; float c = __ldg(&cdf[mid]); if (((c < value) || c == 0.f) && c != sum) lo = mid + 1; else hi = mid; }
    float cdf_n = lo ? __ldg(&cdf[lo - 1]) * norm : 0.f;
    float pmf = __ldg(&pmf_w[lo]) * norm;
    return V(__uint_as_float(lo), pmf, fdiv(sx - cdf_n, pmf));
}

// Scene::sample_emitter_direction (scene.cpp:316-366) -> AreaLight::sample_direction
// (area.cpp:118-168) -> Shape::sample_direction (shape.cpp:94-111); visibility is
// resolved by the trace kernel. Returns em_weight.
// EXT = false: scenes with area lights only and uniform emitter selection (the common case, e.g. the
// Cornell box): the environment / weighted-selection code is compiled out so that the shading kernels
// keep their register budget (the mere presence of the out-of-line calls costs ~8 % of k_shade).
template <bool EXT>
PT_DEV float3 sample_emitter_direction(const DevScene &sc, float3 ref_p, float sx, float sy, DirectionSample &ds) {
    uint32_t n = sc.n_emitters;
    ds.pdf = 0.f; ds.emitter = -1;
    if (n == 0) return V(0.f, 0.f, 0.f);
    float nf = (float) n, scaled = sx * nf;
    uint32_t index = min((uint32_t) scaled, n - 1);
    float sx_re = n < 2 ? sx : scaled - (float) index;
    float emitter_weight = n < 2 ? 1.f : nf;
    float pmf = fdiv(1.f, nf);
    if (EXT && sc.em_cdf) {
        float3 r = sample_emitter_weighted(sc.em_cdf, sc.em_pmf, sc.em_sum, sc.em_norm, n, sx);
        index = __float_as_uint(r.x); pmf = r.y; sx_re = r.z; emitter_weight = n >= 2 ? rcp_(pmf) : 1.f;
    }
    const DevEmitter &em = sc.emitters[index];
    if (EXT && em.type != B200PT_EMITTER_AREA) {
        float3 crad = em.type == B200PT_EMITTER_CONSTANT ? tex_eval3(sc, em.radiance_tex, make_float2(0.f, 0.f)) : V(0.f, 0.f, 0.f);
        float3 spec_env = env_sample_direction(sc.env, crad, ref_p, sx_re, sy, ds);
        ds.emitter = (int32_t) index;
        ds.pdf *= pmf;
        return spec_env * emitter_weight;
    }
    const DevShape &sh = sc.shapes[em.shape];
    shape_sample_position(sc, sh, sx_re, sy, ds.p, ds.n, ds.pdf, ds.uv);
    ds.d = ds.p - ref_p;
    float dist2 = vsqnorm(ds.d);
    ds.dist = __fsqrt_rn(dist2);
    ds.d = V(fdiv(ds.d.x, ds.dist), fdiv(ds.d.y, ds.dist), fdiv(ds.d.z, ds.dist));
    float dp = fabsf(vdot(ds.d, ds.n));
    float x = fdiv(dist2, dp);
    ds.pdf *= isfinite(x) ? x : 0.f;
    bool active = vdot(ds.d, ds.n) < 0.f && ds.pdf != 0.f;
    float3 radiance = tex_eval3(sc, em.radiance_tex, ds.uv);
    float3 spec = active ? V(fdiv(radiance.x, ds.pdf), fdiv(radiance.y, ds.pdf), fdiv(radiance.z, ds.pdf)) : V(0.f, 0.f, 0.f);
    ds.emitter = (int32_t) index;
    ds.pdf *= pmf;
    return spec * emitter_weight;
}

// scene.cpp:378-389: m_emitter_pmf, or sampling_weight * normalization of the emitter distribution
PT_DEV float emitter_pmf(const DevScene &sc, int32_t emitter) {
    return sc.em_cdf ? sc.emitters[emitter].sampling_weight * sc.em_norm : fdiv(1.f, (float) sc.n_emitters);
}

// Scene::pdf_emitter_direction (scene.cpp:378-389) -> area.cpp:170-197 -> shape.cpp:113-124
PT_DEV float pdf_emitter_direction(const DevScene &sc, int32_t emitter, float3 d, float3 n, float dist) {
    float dp = vdot(d, n);
    if (!(dp < 0.f)) return 0.f;
    const DevShape &sh = sc.shapes[sc.emitters[emitter].shape];
    float pdf = sh.sampling == B200PT_SAMPLING_RECTANGLE ? sh.inv_area : sh.area_norm;
    float adp = fabsf(dp);
    pdf *= adp != 0.f ? fdiv(dist * dist, adp) : 0.f;
    return pdf * emitter_pmf(sc, emitter);
}

// path.cpp:359-364
PT_DEV float mis_weight(float a, float b) { a *= a; b *= b; float w = fdiv(a, a + b); return isfinite(w) ? w : 0.f; }

// perspective.cpp:239-279 (primary ray; differentials are not consumed by the hot-path BSDFs)
PT_DEV Ray sample_camera_ray(const DevScene &sc, float px, float py) {
    const float *m = sc.s2c;
    float r0 = m[3], r1 = m[7], r2 = m[11], r3 = m[15];
    r0 = __fmaf_rn(m[0], px, r0); r1 = __fmaf_rn(m[4], px, r1); r2 = __fmaf_rn(m[8], px, r2); r3 = __fmaf_rn(m[12], px, r3);
    r0 = __fmaf_rn(m[1], py, r0); r1 = __fmaf_rn(m[5], py, r1); r2 = __fmaf_rn(m[9], py, r2); r3 = __fmaf_rn(m[13], py, r3);
    r0 = __fmaf_rn(m[2], 0.f, r0); r1 = __fmaf_rn(m[6], 0.f, r1); r2 = __fmaf_rn(m[10], 0.f, r2); r3 = __fmaf_rn(m[14], 0.f, r3);
    float3 near_p = V(fdiv(r0, r3), fdiv(r1, r3), fdiv(r2, r3));
    float3 d = vnormalize(near_p);
    const float *w = sc.cam_to_world;
    Ray ray;
    ray.o = V(w[3], w[7], w[11]);
    float q0 = w[0] * d.x, q1 = w[4] * d.x, q2 = w[8] * d.x;
    q0 = __fmaf_rn(w[1], d.y, q0); q1 = __fmaf_rn(w[5], d.y, q1); q2 = __fmaf_rn(w[9], d.y, q2);
    q0 = __fmaf_rn(w[2], d.z, q0); q1 = __fmaf_rn(w[6], d.z, q1); q2 = __fmaf_rn(w[10], d.z, q2);
    ray.d = V(q0, q1, q2);
    float inv_z = rcp_(d.z), near_t = sc.near_clip * inv_z, far_t = sc.far_clip * inv_z;
    ray.o = ray.o + ray.d * near_t;
    ray.maxt = far_t - near_t;
    return ray;
}

// gaussian.cpp:57-102 / drjit estrin (degree 9)
PT_DEV float estrin9(float x, const float *c) {
    float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    float p01 = __fmaf_rn(c[1], x, c[0]), p23 = __fmaf_rn(c[3], x, c[2]), p45 = __fmaf_rn(c[5], x, c[4]);
    float p67 = __fmaf_rn(c[7], x, c[6]), p89 = __fmaf_rn(c[9], x, c[8]);
    float p0123 = __fmaf_rn(p23, x2, p01), p4567 = __fmaf_rn(p67, x2, p45);
    return __fmaf_rn(p89, x8, __fmaf_rn(p4567, x4, p0123));
}
PT_DEV float rfilter_eval(const DevScene &sc, float x) {
    if (sc.rfilter == B200PT_RFILTER_GAUSSIAN_EXP2)
        return fmaxf(0.f, exp2f((1.4426950408889634074f * sc.gauss_alpha) * (x * x)) - sc.gauss_bias);
    return fmaxf(estrin9(x * x, sc.gauss_coeff), 0.f);
}

} // namespace pt
