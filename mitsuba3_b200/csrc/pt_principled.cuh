// pt_principled.cuh -- Disney "principled" BSDF on the device.
//
// Follows src/bsdfs/principled.cpp:332-837 (sample / eval / pdf),
// src/bsdfs/principledhelpers.h (GTR1, Schlick terms, principled_fresnel,
// mac_mic_compatibility, calc_dist_params) and the GGX branch of
// include/mitsuba/render/microfacet.h:185-421 with sample_visible = true.
// The feature mask (b.flags, = the reference's m_has_* booleans) is uniform per
// BSDF, so all lanes of a material-queue warp take the same side of every
// `has_*` test: the lobes are branch-flattened per material, not per lane.
#pragma once
namespace pt {

struct PrParams {
    float anisotropic, roughness, flatness, spec_trans, metallic, clearcoat, sheen, spec_tint, sheen_tint, clearcoat_gloss;
    float3 base_color;
    bool has_anisotropic, has_spec_trans, has_sheen, has_sheen_tint, has_flatness, has_spec_tint, has_metallic, has_clearcoat;
    float eta, spec_srate, clearcoat_srate, diff_refl_srate;
};

PT_DEV PrParams pr_load(const DevScene &sc, const DevBsdf &b, float2 uv) {
    PrParams p; uint32_t f = b.flags;
    p.has_anisotropic = f & B200PT_P_HAS_ANISOTROPIC; p.has_spec_trans = f & B200PT_P_HAS_SPEC_TRANS;
    p.has_sheen = f & B200PT_P_HAS_SHEEN; p.has_sheen_tint = f & B200PT_P_HAS_SHEEN_TINT;
    p.has_flatness = f & B200PT_P_HAS_FLATNESS; p.has_spec_tint = f & B200PT_P_HAS_SPEC_TINT;
    p.has_metallic = f & B200PT_P_HAS_METALLIC; p.has_clearcoat = f & B200PT_P_HAS_CLEARCOAT;
    p.anisotropic = p.has_anisotropic ? tex_eval1(sc, b.tex[B200PT_SLOT_P_ANISOTROPIC], uv) : 0.f;
    p.roughness = tex_eval1(sc, b.tex[B200PT_SLOT_P_ROUGHNESS], uv);
    p.flatness = p.has_flatness ? tex_eval1(sc, b.tex[B200PT_SLOT_P_FLATNESS], uv) : 0.f;
    p.spec_trans = p.has_spec_trans ? tex_eval1(sc, b.tex[B200PT_SLOT_P_SPEC_TRANS], uv) : 0.f;
    p.metallic = p.has_metallic ? tex_eval1(sc, b.tex[B200PT_SLOT_P_METALLIC], uv) : 0.f;
    p.clearcoat = p.has_clearcoat ? tex_eval1(sc, b.tex[B200PT_SLOT_P_CLEARCOAT], uv) : 0.f;
    p.sheen = p.has_sheen ? tex_eval1(sc, b.tex[B200PT_SLOT_P_SHEEN], uv) : 0.f;
    p.spec_tint = p.has_spec_tint ? tex_eval1(sc, b.tex[B200PT_SLOT_P_SPEC_TINT], uv) : 0.f;
    p.sheen_tint = p.has_sheen_tint ? tex_eval1(sc, b.tex[B200PT_SLOT_P_SHEEN_TINT], uv) : 0.f;
    p.clearcoat_gloss = p.has_clearcoat ? tex_eval1(sc, b.tex[B200PT_SLOT_P_CLEARCOAT_GLOSS], uv) : 0.f;
    p.base_color = tex_eval3(sc, b.tex[B200PT_SLOT_P_BASE_COLOR], uv);
    p.eta = b.eta; p.spec_srate = b.spec_srate; p.clearcoat_srate = b.clearcoat_srate; p.diff_refl_srate = b.diff_refl_srate;
    return p;
}

PT_DEV float lerpf(float a, float b, float t) { return __fmaf_rn(b, t, __fmaf_rn(-a, t, a)); }   // drjit lerp
PT_DEV float clipf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
PT_DEV float3 vmulsign(float3 v, float s) { return V(mulsign(v.x, s), mulsign(v.y, s), mulsign(v.z, s)); }
PT_DEV float3 vmulsign_neg(float3 v, float s) { return V(mulsign_neg(v.x, s), mulsign_neg(v.y, s), mulsign_neg(v.z, s)); }

PT_DEV float schlick_weight(float cos_i) { float m = clipf(1.f - cos_i, 0.f, 1.f); return sqr(sqr(m)) * m; }
PT_DEV float calc_schlick(float R0, float cos_theta_i, float eta) {
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = rcp_(eta), eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = __fmaf_rn(-__fmaf_rn(-cos_theta_i, cos_theta_i, 1.f), sqr(eta_ti), 1.f);
    float cos_theta_t = safe_sqrt(cos_theta_t_sqr);
    return eta_it > 1.f ? lerpf(schlick_weight(fabsf(cos_theta_i)), 1.f, R0) : lerpf(schlick_weight(cos_theta_t), 1.f, R0);
}
PT_DEV float schlick_R0_eta(float eta) { return sqr(fdiv(eta - 1.f, eta + 1.f)); }

PT_DEV bool mac_mic(float3 m, float3 wi, float3 wo, float cos_theta_i, bool reflection) {
    float3 ms = vmulsign(m, cos_theta_i);
    if (reflection) return vdot(wi, ms) > 0.f && vdot(wo, ms) > 0.f;
    return vdot(wi, ms) > 0.f && vdot(wo, vmulsign_neg(m, cos_theta_i)) > 0.f;
}

PT_DEV float3 principled_fresnel(float F_dielectric, float metallic, float spec_tint, float3 base_color, float lum, float cos_theta_i,
                                 bool front_side, float bsdf, float eta, bool has_metallic, bool has_spec_tint) {
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = rcp_(eta), eta_it = outside ? eta : rcp_eta;
    float3 F_schlick = V(0.f, 0.f, 0.f);
    if (has_metallic)
        F_schlick = F_schlick + V(calc_schlick(base_color.x, cos_theta_i, eta), calc_schlick(base_color.y, cos_theta_i, eta), calc_schlick(base_color.z, cos_theta_i, eta)) * metallic;
    if (has_spec_tint) {
        float3 c_tint = lum > 0.f ? V(fdiv(base_color.x, lum), fdiv(base_color.y, lum), fdiv(base_color.z, lum)) : V(1.f, 1.f, 1.f);
        float3 F0 = c_tint * schlick_R0_eta(eta_it);
        float k = (1.f - metallic) * spec_tint;
        F_schlick = F_schlick + V(calc_schlick(F0.x, cos_theta_i, eta), calc_schlick(F0.y, cos_theta_i, eta), calc_schlick(F0.z, cos_theta_i, eta)) * k;
    }
    float fd = (1.f - metallic) * (1.f - spec_tint) * F_dielectric;
    float fb = bsdf * F_dielectric;
    return front_side ? V(fd + F_schlick.x, fd + F_schlick.y, fd + F_schlick.z) : V(fb, fb, fb);
}

PT_DEV void calc_dist_params(float anisotropic, float roughness, bool has_anisotropic, float &ax, float &ay) {
    float r2 = sqr(roughness);
    if (!has_anisotropic) { ax = ay = fmaxf(0.001f, r2); return; }
    float aspect = __fsqrt_rn(1.f - 0.9f * anisotropic);
    ax = fmaxf(0.001f, fdiv(r2, aspect)); ay = fmaxf(0.001f, r2 * aspect);
}

// ---- GGX (microfacet.h) -----------------------------------------------------------
struct Ggx { float au, av; };
PT_DEV Ggx ggx_make(float au, float av) { Ggx g; g.au = fmaxf(au, 1e-4f); g.av = fmaxf(av, 1e-4f); return g; }
PT_DEV float ggx_eval(const Ggx &g, float3 m) {
    float alpha_uv = g.au * g.av;
    float result = rcp_(PT_PI * alpha_uv * sqr(sqr(fdiv(m.x, g.au)) + sqr(fdiv(m.y, g.av)) + sqr(m.z)));
    return result * m.z > 1e-20f ? result : 0.f;
}
PT_DEV float ggx_smith_g1(const Ggx &g, float3 v, float3 m) {
    float xy_alpha_2 = sqr(g.au * v.x) + sqr(g.av * v.y), tan_theta_alpha_2 = fdiv(xy_alpha_2, sqr(v.z));
    float result = fdiv(2.f, 1.f + __fsqrt_rn(1.f + tan_theta_alpha_2));
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (vdot(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
PT_DEV float ggx_pdf(const Ggx &g, float3 wi, float3 m) { return fdiv(ggx_eval(g, m) * ggx_smith_g1(g, wi, m) * fabsf(vdot(wi, m)), wi.z); }

PT_DEV float2 square_to_uniform_disk_concentric(float sx, float sy) {
    float x = __fmaf_rn(2.f, sx, -1.f), y = __fmaf_rn(2.f, sy, -1.f);
    bool is_zero = (x == 0.f) && (y == 0.f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = fdiv(__fmul_rn(__fmul_rn(0.25f, PT_PI), rp), r);
    if (q13) phi = __fmul_rn(0.5f, PT_PI) - phi;
    if (is_zero) phi = 0.f;
    float s, c; sincosf(phi, &s, &c);
    return make_float2(r * c, r * s);
}
PT_DEV float2 ggx_sample_visible_11(float cos_theta_i, float sx, float sy) {
    float2 p = square_to_uniform_disk_concentric(sx, sy);
    float s = 0.5f * (1.f + cos_theta_i);
    p.y = lerpf(safe_sqrt(1.f - sqr(p.x)), p.y, s);
    float x = p.x, y = p.y, z = safe_sqrt(1.f - __fmaf_rn(p.y, p.y, __fmul_rn(p.x, p.x)));
    float sin_theta_i = safe_sqrt(1.f - sqr(cos_theta_i));
    float norm = rcp_(__fmaf_rn(sin_theta_i, y, cos_theta_i * z));
    return make_float2(__fmaf_rn(cos_theta_i, y, -(sin_theta_i * z)) * norm, x * norm);
}
PT_DEV float3 ggx_sample(const Ggx &g, float3 wi, float sx, float sy) {
    float3 wi_p = vnormalize(V(g.au * wi.x, g.av * wi.y, wi.z));
    float sin_theta_2 = __fmaf_rn(wi_p.x, wi_p.x, sqr(wi_p.y)), inv_sin_theta = rsqrt_(sin_theta_2);   // frame.h:111-122
    float cos_phi = wi_p.x * inv_sin_theta, sin_phi = wi_p.y * inv_sin_theta;
    if (fabsf(sin_theta_2) <= 4.f * 5.9604644775390625e-08f) { cos_phi = 1.f; sin_phi = 0.f; }
    else { cos_phi = clipf(cos_phi, -1.f, 1.f); sin_phi = clipf(sin_phi, -1.f, 1.f); }
    float2 sl = ggx_sample_visible_11(wi_p.z, sx, sy);
    float s0 = __fmaf_rn(cos_phi, sl.x, -(sin_phi * sl.y)) * g.au, s1 = __fmaf_rn(sin_phi, sl.x, cos_phi * sl.y) * g.av;
    return vnormalize(V(-s0, -s1, 1.f));
}

// ---- GTR1 / clearcoat (principledhelpers.h:21-111) ------------------------------------
PT_DEV float gtr1_eval(float alpha, float3 m) {
    float cos_theta2 = sqr(m.z), alpha2 = sqr(alpha);
    float result = fdiv(alpha2 - 1.f, PT_PI * logf(alpha2) * (1.f + (alpha2 - 1.f) * cos_theta2));
    return result * m.z > 1e-20f ? result : 0.f;
}
PT_DEV float gtr1_pdf(float alpha, float3 m) { return m.z < 0.f ? 0.f : m.z * gtr1_eval(alpha, m); }
PT_DEV float3 gtr1_sample(float alpha, float sx, float sy) {
    float sin_phi, cos_phi; sincosf((2.f * PT_PI) * sx, &sin_phi, &cos_phi);
    float alpha2 = sqr(alpha);
    float cos_theta2 = fdiv(1.f - powf(alpha2, 1.f - sy), 1.f - alpha2);
    float sin_theta = __fsqrt_rn(fmaxf(0.f, 1.f - cos_theta2)), cos_theta = __fsqrt_rn(fmaxf(0.f, cos_theta2));
    return V(cos_phi * sin_theta, sin_phi * sin_theta, cos_theta);
}
PT_DEV float smith_ggx1(float3 v, float3 wh, float alpha) {
    float alpha_2 = sqr(alpha), cos_theta = fabsf(v.z), cos_theta_2 = sqr(cos_theta), tan_theta_2 = fdiv(1.f - cos_theta_2, cos_theta_2);
    float result = 2.f * rcp_(1.f + __fsqrt_rn(1.f + alpha_2 * tan_theta_2));
    if (v.z == 1.f) result = 1.f;
    if (vdot(v, wh) * v.z <= 0.f) result = 0.f;
    return result;
}
PT_DEV float luminance(float3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }

// principled.cpp:489-703
PT_DEV float3 pr_eval(const PrParams &p, float3 wi, float3 wo, bool active) {
    float cos_theta_i = wi.z;
    if (!active || cos_theta_i == 0.f) return V(0.f, 0.f, 0.f);
    float brdf = (1.f - p.metallic) * (1.f - p.spec_trans), bsdf = (1.f - p.metallic) * p.spec_trans;
    float cos_theta_o = wo.z;
    bool reflect = cos_theta_i * cos_theta_o > 0.f, refract = cos_theta_i * cos_theta_o < 0.f, front_side = cos_theta_i > 0.f;
    float inv_eta = rcp_(p.eta), eta_path = front_side ? p.eta : inv_eta, inv_eta_path = front_side ? inv_eta : p.eta;
    float ax, ay; calc_dist_params(p.anisotropic, p.roughness, p.has_anisotropic, ax, ay);
    Ggx dist = ggx_make(ax, ay);
    float3 wh = vnormalize(wi + wo * (reflect ? 1.f : eta_path));
    wh = vmulsign(wh, wh.z);
    float F, cos_theta_t, eta_it, eta_ti;
    fresnel(vdot(wi, wh), p.eta, F, cos_theta_t, eta_it, eta_ti);
    bool refl_c = mac_mic(wh, wi, wo, cos_theta_i, true), refr_c = mac_mic(wh, wi, wo, cos_theta_i, false);
    bool spec_reflect_active = reflect && refl_c && F > 0.f;
    bool clearcoat_active = p.has_clearcoat && p.clearcoat > 0.f && reflect && refl_c && front_side;
    bool spec_trans_active = p.has_spec_trans && bsdf > 0.f && refract && refr_c && F < 1.f;
    bool diffuse_active = brdf > 0.f && reflect && front_side;
    bool sheen_active = p.has_sheen && p.sheen > 0.f && reflect && (1.f - p.metallic > 0.f) && front_side;
    float D = ggx_eval(dist, wh), G = ggx_smith_g1(dist, wi, wh) * ggx_smith_g1(dist, wo, wh);
    float3 value = V(0.f, 0.f, 0.f);
    if (spec_reflect_active) {
        float lum = p.has_spec_tint ? luminance(p.base_color) : 1.f;
        float3 Fp = principled_fresnel(F, p.metallic, p.spec_tint, p.base_color, lum, vdot(wi, wh), front_side, bsdf, p.eta, p.has_metallic, p.has_spec_tint);
        float den = 4.f * fabsf(cos_theta_i);
        value = value + V(fdiv(Fp.x * D * G, den), fdiv(Fp.y * D * G, den), fdiv(Fp.z * D * G, den));
    }
    if (spec_trans_active) {
        float scale = sqr(inv_eta_path);
        float wih = vdot(wi, wh), woh = vdot(wo, wh);
        float t = fabsf(fdiv(scale * (1.f - F) * D * G * eta_path * eta_path * wih * woh, cos_theta_i * sqr(wih + eta_path * woh)));
        value = value + (V(__fsqrt_rn(p.base_color.x), __fsqrt_rn(p.base_color.y), __fsqrt_rn(p.base_color.z)) * bsdf) * t;
    }
    if (clearcoat_active) {
        float Fcc = calc_schlick(0.04f, vdot(wi, wh), p.eta);
        float Dcc = gtr1_eval(lerpf(0.1f, 0.001f, p.clearcoat_gloss), wh);
        float Gcc = smith_ggx1(wi, wh, 0.25f) * smith_ggx1(wo, wh, 0.25f);
        float c = (p.clearcoat * 0.25f) * Fcc * Dcc * Gcc * fabsf(cos_theta_o);
        value = value + V(c, c, c);
    }
    if (diffuse_active) {
        float Fo = schlick_weight(fabsf(cos_theta_o)), Fi = schlick_weight(fabsf(cos_theta_i));
        float f_diff = (1.f - 0.5f * Fi) * (1.f - 0.5f * Fo);
        float cos_theta_d = vdot(wh, wo);
        float Rr = 2.f * p.roughness * sqr(cos_theta_d);
        float f_retro = Rr * (Fo + Fi + Fo * Fi * (Rr - 1.f));
        float k;
        if (p.has_flatness) {
            float Fss90 = fdiv(Rr, 2.f);
            float Fss = lerpf(1.f, Fss90, Fo) * lerpf(1.f, Fss90, Fi);
            float f_ss = 1.25f * (Fss * (fdiv(1.f, fabsf(cos_theta_o) + fabsf(cos_theta_i)) - 0.5f) + 0.5f);
            k = lerpf(f_diff + f_retro, f_ss, p.flatness);
        } else k = f_diff + f_retro;
        value = value + ((p.base_color * (brdf * fabsf(cos_theta_o))) * PT_INV_PI) * k;
        if (sheen_active) {
            float Fd = schlick_weight(fabsf(cos_theta_d));
            float sh = p.sheen * (1.f - p.metallic) * Fd;
            if (p.has_sheen_tint) {
                float lum = luminance(p.base_color);
                float3 c_tint = lum > 0.f ? V(fdiv(p.base_color.x, lum), fdiv(p.base_color.y, lum), fdiv(p.base_color.z, lum)) : V(1.f, 1.f, 1.f);
                float3 c_sheen = V(lerpf(1.f, c_tint.x, p.sheen_tint), lerpf(1.f, c_tint.y, p.sheen_tint), lerpf(1.f, c_tint.z, p.sheen_tint));
                value = value + (c_sheen * sh) * fabsf(cos_theta_o);
            } else {
                float c = sh * fabsf(cos_theta_o);
                value = value + V(c, c, c);
            }
        }
    }
    return value;
}

// principled.cpp:705-822
PT_DEV float pr_pdf(const PrParams &p, float3 wi, float3 wo, bool active) {
    float cos_theta_i = wi.z;
    if (!active || cos_theta_i == 0.f) return 0.f;
    float brdf = (1.f - p.metallic) * (1.f - p.spec_trans), bsdf = (1.f - p.metallic) * p.spec_trans;
    bool front_side = cos_theta_i > 0.f;
    float eta_path = front_side ? p.eta : rcp_(p.eta);
    float cos_theta_o = wo.z;
    bool reflect = cos_theta_i * cos_theta_o > 0.f, refract = cos_theta_i * cos_theta_o < 0.f;
    float3 wh = vnormalize(wi + wo * (reflect ? 1.f : eta_path));
    wh = vmulsign(wh, wh.z);
    float ax, ay; calc_dist_params(p.anisotropic, p.roughness, p.has_anisotropic, ax, ay);
    Ggx dist = ggx_make(ax, ay);
    float F, cos_theta_t, eta_it, eta_ti;
    fresnel(vdot(wi, wh), p.eta, F, cos_theta_t, eta_it, eta_ti);
    float prob_spec_reflect = front_side ? p.spec_srate * (1.f - bsdf * (1.f - F)) : F;
    float prob_spec_trans = p.has_spec_trans ? (front_side ? p.spec_srate * bsdf * (1.f - F) : (1.f - F)) : 0.f;
    float prob_clearcoat = p.has_clearcoat ? (front_side ? 0.25f * p.clearcoat * p.clearcoat_srate : 0.f) : 0.f;
    float prob_diffuse = front_side ? brdf * p.diff_refl_srate : 0.f;
    float rcp_tot = rcp_(prob_spec_reflect + prob_spec_trans + prob_clearcoat + prob_diffuse);
    prob_spec_reflect *= rcp_tot; prob_spec_trans *= rcp_tot; prob_clearcoat *= rcp_tot; prob_diffuse *= rcp_tot;
    float dwh_dwo_abs;
    if (p.has_spec_trans) {
        float wih = vdot(wi, wh), woh = vdot(wo, wh);
        dwh_dwo_abs = fabsf(reflect ? rcp_(4.f * woh) : fdiv(sqr(eta_path) * woh, sqr(wih + eta_path * woh)));
    } else dwh_dwo_abs = fabsf(rcp_(4.f * vdot(wo, wh)));
    float pdf = 0.f;
    bool mf_reflect = mac_mic(wh, wi, wo, cos_theta_i, true) && reflect;
    if (mf_reflect) pdf += prob_spec_reflect * ggx_pdf(dist, vmulsign(wi, cos_theta_i), wh) * dwh_dwo_abs;
    if (reflect) pdf += prob_diffuse * (PT_INV_PI * wo.z);
    if (p.has_spec_trans) {
        bool mf_trans = mac_mic(wh, wi, wo, cos_theta_i, false) && refract;
        if (mf_trans) pdf += prob_spec_trans * ggx_pdf(dist, vmulsign(wi, cos_theta_i), wh) * dwh_dwo_abs;
    }
    if (p.has_clearcoat && mf_reflect) pdf += prob_clearcoat * gtr1_pdf(lerpf(0.1f, 0.001f, p.clearcoat_gloss), wh) * dwh_dwo_abs;
    return pdf;
}

// principled.cpp:332-487
PT_DEV void pr_sample(const PrParams &p, float3 wi, float sample1, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    bs.wo = V(0.f, 0.f, 0.f); bs.pdf = 0.f; bs.eta = 0.f; bs.sampled_type = 0; bs.sampled_component = 0;
    weight = V(0.f, 0.f, 0.f);
    float cos_theta_i = wi.z;
    bool active = cos_theta_i != 0.f;
    if (!active) return;
    float brdf = (1.f - p.metallic) * (1.f - p.spec_trans);
    float bsdf = p.has_spec_trans ? (1.f - p.metallic) * p.spec_trans : 0.f;
    bool front_side = cos_theta_i > 0.f;
    float ax, ay; calc_dist_params(p.anisotropic, p.roughness, p.has_anisotropic, ax, ay);
    Ggx dist = ggx_make(ax, ay);
    float3 m_spec = ggx_sample(dist, vmulsign(wi, cos_theta_i), s2x, s2y);
    float F, cos_theta_t, eta_it, eta_ti;
    fresnel(vdot(wi, m_spec), p.eta, F, cos_theta_t, eta_it, eta_ti);
    active = active && (front_side || bsdf > 0.f);
    float prob_spec_reflect = front_side ? p.spec_srate * (1.f - bsdf * (1.f - F)) : F;
    float prob_spec_trans = p.has_spec_trans ? (front_side ? p.spec_srate * bsdf * (1.f - F) : (1.f - F)) : 0.f;
    float prob_clearcoat = p.has_clearcoat ? (front_side ? 0.25f * p.clearcoat * p.clearcoat_srate : 0.f) : 0.f;
    float prob_diffuse = front_side ? brdf * p.diff_refl_srate : 0.f;
    float rcp_tot = rcp_(prob_spec_reflect + prob_spec_trans + prob_clearcoat + prob_diffuse);
    prob_spec_trans *= rcp_tot; prob_clearcoat *= rcp_tot; prob_diffuse *= rcp_tot;
    float curr = 0.f;
    bool sample_diffuse = active && sample1 < prob_diffuse;
    curr += prob_diffuse;
    bool sample_clearcoat = p.has_clearcoat && active && sample1 >= curr && sample1 < curr + prob_clearcoat;
    curr += prob_clearcoat;
    bool sample_spec_trans = p.has_spec_trans && active && sample1 >= curr && sample1 < curr + prob_spec_trans;
    curr += prob_spec_trans;
    bool sample_spec_reflect = active && sample1 >= curr;
    bs.eta = 1.f;
    if (sample_spec_reflect) {
        float k = 2.f * vdot(wi, m_spec);
        float3 wo = V(__fmaf_rn(m_spec.x, k, -wi.x), __fmaf_rn(m_spec.y, k, -wi.y), __fmaf_rn(m_spec.z, k, -wi.z));
        bs.wo = wo; bs.sampled_component = 3; bs.sampled_type = F_GLOSSY_REFLECTION;
        active = active && mac_mic(m_spec, wi, wo, cos_theta_i, true) && cos_theta_i * wo.z > 0.f;
    }
    if (sample_spec_trans) {
        float k = __fmaf_rn(vdot(wi, m_spec), eta_ti, cos_theta_t);
        float3 wo = V(__fmaf_rn(m_spec.x, k, -(wi.x * eta_ti)), __fmaf_rn(m_spec.y, k, -(wi.y * eta_ti)), __fmaf_rn(m_spec.z, k, -(wi.z * eta_ti)));
        bs.wo = wo; bs.sampled_component = 2; bs.sampled_type = F_GLOSSY_TRANSMISSION; bs.eta = eta_it;
        active = active && mac_mic(m_spec, wi, wo, cos_theta_i, false) && cos_theta_i * wo.z < 0.f;
    }
    if (sample_clearcoat) {
        float3 m_cc = gtr1_sample(lerpf(0.1f, 0.001f, p.clearcoat_gloss), s2x, s2y);
        float k = 2.f * vdot(wi, m_cc);
        float3 wo = V(__fmaf_rn(m_cc.x, k, -wi.x), __fmaf_rn(m_cc.y, k, -wi.y), __fmaf_rn(m_cc.z, k, -wi.z));
        bs.wo = wo; bs.sampled_component = 1; bs.sampled_type = F_GLOSSY_REFLECTION;
        active = active && mac_mic(m_cc, wi, wo, cos_theta_i, true) && cos_theta_i * wo.z > 0.f;
    }
    if (sample_diffuse) {
        float3 wo = square_to_cosine_hemisphere(s2x, s2y);
        bs.wo = wo; bs.sampled_component = 0; bs.sampled_type = F_DIFFUSE_REFLECTION;
        active = active && cos_theta_i * wo.z > 0.f;
    }
    bs.pdf = pr_pdf(p, wi, bs.wo, true);   // pdf() does not mask its result
    active = active && bs.pdf > 0.f;
    if (active) {
        float3 r = pr_eval(p, wi, bs.wo, true);
        weight = V(fdiv(r.x, bs.pdf), fdiv(r.y, bs.pdf), fdiv(r.z, bs.pdf));
    }
}

PT_DEV void principled_eval_pdf(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float3 &value, float &pdf) {
    PrParams p = pr_load(sc, b, uv);
    value = pr_eval(p, wi, wo, true);
    pdf = pr_pdf(p, wi, wo, true);
}
PT_DEV void principled_sample(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float s1, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    PrParams p = pr_load(sc, b, uv);
    pr_sample(p, wi, s1, s2x, s2y, bs, weight);
}

} // namespace pt
