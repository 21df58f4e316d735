// pt_rough.cuh -- rough conductor / rough dielectric on the device:
//   src/bsdfs/roughconductor.cpp:222-520, src/bsdfs/roughdielectric.cpp:247-610 (unpolarized,
//   both lobes, Radiance mode), include/mitsuba/render/microfacet.h:185-421 (Beckmann + GGX,
//   visible-normal sampling), fresnel.h:276-314, drjit/math.h:1557-1578 (erfinv).
// They are the B200PT_M_ROUGH variants of B200PT_BSDF_CONDUCTOR / _DIELECTRIC, so they shade
// in the same material queues / kernels as their smooth siblings (one uniform branch per BSDF).
// Included by pt_device.cuh after pt_principled.cuh (reuses its GGX helpers).
#pragma once

namespace pt {

// drjit/math.h:27-43 estrin_impl for 9 coefficients
PT_DEV float estrin9c(float x, const float *c) {
    float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    float p0 = __fmaf_rn(x, c[1], c[0]), p1 = __fmaf_rn(x, c[3], c[2]), p2 = __fmaf_rn(x, c[5], c[4]), p3 = __fmaf_rn(x, c[7], c[6]), p4 = c[8];
    float q0 = __fmaf_rn(x2, p1, p0), q1 = __fmaf_rn(x2, p3, p2), q2 = p4;
    float r0 = __fmaf_rn(x4, q1, q0), r1 = q2;
    return __fmaf_rn(x8, r1, r0);
}

PT_DEV float dr_erfinv(float x) {
    float w = -logf((1.f - x) * (1.f + x));
    float w1 = w - 2.5f, w2 = __fsqrt_rn(w) - 3.f;
    const float c1[9] = { 1.50140941f, 0.246640727f, -0.00417768164f, -0.00125372503f, 0.00021858087f,
                          -4.39150654e-06f, -3.5233877e-06f, 3.43273939e-07f, 2.81022636e-08f };
    const float c2[9] = { 2.83297682f, 1.00167406f, 0.00943887047f, -0.0076224613f, 0.00573950773f,
                          -0.00367342844f, 0.00134934322f, 0.000100950558f, -0.000200214257f };
    return (w < 5.f ? estrin9c(w1, c1) : estrin9c(w2, c2)) * x;
}

// ---- MicrofacetDistribution, sample_visible = true --------------------------------------
struct Mfd { bool is_ggx; float au, av; };
PT_DEV Mfd mfd_make(bool is_ggx, float au, float av) { Mfd d; d.is_ggx = is_ggx; d.au = fmaxf(au, 1e-4f); d.av = fmaxf(av, 1e-4f); return d; }

PT_DEV float mfd_eval(const Mfd &d, float3 m) {                                 // microfacet.h:185-208
    float alpha_uv = d.au * d.av, cos_theta_2 = sqr(m.z), result;
    float e = sqr(fdiv(m.x, d.au)) + sqr(fdiv(m.y, d.av));
    if (!d.is_ggx) result = fdiv(expf(fdiv(-e, cos_theta_2)), PT_PI * alpha_uv * sqr(cos_theta_2));
    else result = rcp_(PT_PI * alpha_uv * sqr(e + sqr(m.z)));
    return result * m.z > 1e-20f ? result : 0.f;
}

PT_DEV float mfd_smith_g1(const Mfd &d, float3 v, float3 m) {                   // microfacet.h:341-365
    float xy_alpha_2 = sqr(d.au * v.x) + sqr(d.av * v.y), tan_theta_alpha_2 = fdiv(xy_alpha_2, sqr(v.z)), result;
    if (!d.is_ggx) {
        float a = rsqrt_(tan_theta_alpha_2), a_sqr = sqr(a);
        result = a >= 1.6f ? 1.f : fdiv(3.535f * a + 2.181f * a_sqr, 1.f + 2.276f * a + 2.577f * a_sqr);
    } else result = fdiv(2.f, 1.f + __fsqrt_rn(1.f + tan_theta_alpha_2));
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (vdot(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
PT_DEV float mfd_pdf(const Mfd &d, float3 wi, float3 m) { return fdiv(mfd_eval(d, m) * mfd_smith_g1(d, wi, m) * fabsf(vdot(wi, m)), wi.z); }

PT_DEV float2 mfd_sample_visible_11(const Mfd &d, float cos_theta_i, float sx, float sy) {   // microfacet.h:368-421
    if (d.is_ggx) return ggx_sample_visible_11(cos_theta_i, sx, sy);
    float tan_theta_i = fdiv(safe_sqrt(__fmaf_rn(-cos_theta_i, cos_theta_i, 1.f)), cos_theta_i);
    float cot_theta_i = rcp_(tan_theta_i);
    float maxval = erff(cot_theta_i);
    sx = fmaxf(fminf(sx, 1.f - 1e-6f), 1e-6f); sy = fmaxf(fminf(sy, 1.f - 1e-6f), 1e-6f);
    float x = maxval - (maxval + 1.f) * erff(__fsqrt_rn(-logf(sx)));
    const float inv_sqrt_pi = 0.56418958354775628695f;
    sx *= 1.f + maxval + inv_sqrt_pi * tan_theta_i * expf(-sqr(cot_theta_i));
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        float slope = dr_erfinv(x);
        float value = 1.f + x + inv_sqrt_pi * tan_theta_i * expf(-sqr(slope)) - sx;
        float derivative = 1.f - slope * tan_theta_i;
        x -= fdiv(value, derivative);
    }
    return make_float2(dr_erfinv(x), dr_erfinv(__fmaf_rn(2.f, sy, -1.f)));
}

PT_DEV float3 mfd_sample(const Mfd &d, float3 wi, float sx, float sy, float &pdf) {          // microfacet.h:296-325
    float3 wi_p = vnormalize(V(d.au * wi.x, d.av * wi.y, wi.z));
    float sin_theta_2 = __fmaf_rn(wi_p.x, wi_p.x, sqr(wi_p.y)), inv_sin_theta = rsqrt_(sin_theta_2);   // frame.h:111-122
    float cos_phi = wi_p.x * inv_sin_theta, sin_phi = wi_p.y * inv_sin_theta;
    if (fabsf(sin_theta_2) <= 4.f * 5.9604644775390625e-08f) { cos_phi = 1.f; sin_phi = 0.f; }
    else { cos_phi = clipf(cos_phi, -1.f, 1.f); sin_phi = clipf(sin_phi, -1.f, 1.f); }
    float2 sl = mfd_sample_visible_11(d, wi_p.z, sx, sy);
    float s0 = __fmaf_rn(cos_phi, sl.x, -(sin_phi * sl.y)) * d.au, s1 = __fmaf_rn(sin_phi, sl.x, cos_phi * sl.y) * d.av;
    float3 m = vnormalize(V(-s0, -s1, 1.f));
    pdf = fdiv(mfd_eval(d, m) * mfd_smith_g1(d, wi, m) * fabsf(vdot(wi, m)), wi.z);
    return m;
}

PT_DEV float3 reflect_m(float3 wi, float3 m) { float k = 2.f * vdot(wi, m); return V(__fmaf_rn(m.x, k, -wi.x), __fmaf_rn(m.y, k, -wi.y), __fmaf_rn(m.z, k, -wi.z)); }
PT_DEV float3 refract_m(float3 wi, float3 m, float cos_theta_t, float eta_ti) {
    float k = __fmaf_rn(vdot(wi, m), eta_ti, cos_theta_t);
    return V(__fmaf_rn(m.x, k, -(wi.x * eta_ti)), __fmaf_rn(m.y, k, -(wi.y * eta_ti)), __fmaf_rn(m.z, k, -(wi.z * eta_ti)));
}

PT_DEV Mfd rough_distr(const DevScene &sc, const DevBsdf &b, float2 uv, int slot_u, int slot_v) {
    return mfd_make((b.flags & B200PT_M_GGX) != 0, tex_eval1(sc, b.tex[slot_u], uv), tex_eval1(sc, b.tex[slot_v], uv));
}

// ---- roughconductor.cpp ---------------------------------------------------------------
PT_DEV float3 rc_fresnel(const DevScene &sc, const DevBsdf &b, float2 uv, float cos_i) {
    float3 eta = tex_eval3(sc, b.tex[B200PT_SLOT_ETA], uv), k = tex_eval3(sc, b.tex[B200PT_SLOT_K], uv);
    return V(fresnel_conductor(cos_i, eta.x, k.x), fresnel_conductor(cos_i, eta.y, k.y), fresnel_conductor(cos_i, eta.z, k.z));
}

PT_DEV void roughconductor_eval_pdf(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float3 &value, float &pdf) {
    float cti = wi.z, cto = wo.z;
    float3 H = vnormalize(wo + wi);
    if (!(cti > 0.f && cto > 0.f && vdot(wi, H) > 0.f && vdot(wo, H) > 0.f)) return;
    Mfd d = rough_distr(sc, b, uv, B200PT_SLOT_ALPHA_U, B200PT_SLOT_ALPHA_V);
    float D = mfd_eval(d, H);
    if (!(D != 0.f)) return;
    float g1_wi = mfd_smith_g1(d, wi, H), G = g1_wi * mfd_smith_g1(d, wo, H);
    float val = fdiv(D * G, 4.f * cti);
    float3 v = V(val, val, val);
    float3 F = rc_fresnel(sc, b, uv, vdot(wi, H));
    if (b.tex[B200PT_SLOT_SPEC_REFL] >= 0) v = v * tex_eval3(sc, b.tex[B200PT_SLOT_SPEC_REFL], uv);
    value = F * v;
    pdf = fdiv(D * g1_wi, 4.f * cti);
}

PT_DEV void roughconductor_sample(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    float cti = wi.z;
    if (!(cti > 0.f)) return;
    Mfd d = rough_distr(sc, b, uv, B200PT_SLOT_ALPHA_U, B200PT_SLOT_ALPHA_V);
    float pdf; float3 m = mfd_sample(d, wi, s2x, s2y, pdf);
    bs.wo = reflect_m(wi, m); bs.eta = 1.f; bs.sampled_component = 0; bs.sampled_type = F_GLOSSY_REFLECTION;
    bool active = pdf != 0.f && bs.wo.z > 0.f;
    float w = mfd_smith_g1(d, bs.wo, m);
    bs.pdf = fdiv(pdf, 4.f * vdot(bs.wo, m));
    float3 F = rc_fresnel(sc, b, uv, vdot(wi, m));
    float3 wv = V(w, w, w);
    if (b.tex[B200PT_SLOT_SPEC_REFL] >= 0) wv = wv * tex_eval3(sc, b.tex[B200PT_SLOT_SPEC_REFL], uv);
    if (active) weight = F * wv;
}

// ---- roughdielectric.cpp --------------------------------------------------------------
PT_DEV void roughdielectric_eval_pdf(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float3 &value, float &pdf) {
    float cti = wi.z, cto = wo.z;
    if (!(cti != 0.f)) return;
    bool reflect = cti * cto > 0.f;
    float m_eta = b.eta, m_inv_eta = rcp_(b.eta);
    float eta = cti > 0.f ? m_eta : m_inv_eta, inv_eta = cti > 0.f ? m_inv_eta : m_eta;
    float3 m = vnormalize(wi + wo * (reflect ? 1.f : eta));
    m = vmulsign(m, m.z);
    float dot_wi_m = vdot(wi, m), dot_wo_m = vdot(wo, m);
    if (!(dot_wi_m * cti > 0.f && dot_wo_m * cto > 0.f)) return;
    Mfd d = rough_distr(sc, b, uv, B200PT_SLOT_D_ALPHA_U, B200PT_SLOT_D_ALPHA_V);
    float D = mfd_eval(d, m);
    float F, ctt, eta_it, eta_ti; fresnel(dot_wi_m, m_eta, F, ctt, eta_it, eta_ti);
    float G = mfd_smith_g1(d, wi, m) * mfd_smith_g1(d, wo, m);
    float dwh_dwo, val;
    if (reflect) {
        val = fdiv(F * D * G, 4.f * fabsf(cti));
        dwh_dwo = rcp_(4.f * dot_wo_m);
    } else {
        float scale = sqr(inv_eta);
        val = fabsf(fdiv(scale * (1.f - F) * D * G * eta * eta * dot_wi_m * dot_wo_m, cti * sqr(dot_wi_m + eta * dot_wo_m)));
        dwh_dwo = fdiv(eta * eta * dot_wo_m, sqr(dot_wi_m + eta * dot_wo_m));
    }
    float3 v = V(val, val, val);
    int slot = reflect ? B200PT_SLOT_D_SPEC_REFL : B200PT_SLOT_D_SPEC_TRANS;
    if (b.tex[slot] >= 0) v = v * tex_eval3(sc, b.tex[slot], uv);
    value = v;
    float prob = mfd_pdf(d, vmulsign(wi, cti), m);
    prob *= reflect ? F : 1.f - F;
    pdf = prob * fabsf(dwh_dwo);
}

PT_DEV void roughdielectric_sample(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float s1, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    float cti = wi.z;
    bool active = cti != 0.f;
    Mfd d = rough_distr(sc, b, uv, B200PT_SLOT_D_ALPHA_U, B200PT_SLOT_D_ALPHA_V);
    float pdf; float3 m = mfd_sample(d, vmulsign(wi, cti), s2x, s2y, pdf);
    active = active && pdf != 0.f;
    float F, ctt, eta_it, eta_ti; fresnel(vdot(wi, m), b.eta, F, ctt, eta_it, eta_ti);
    bool sel_r = s1 <= F && active, sel_t = !sel_r && active;
    bs.pdf = pdf * (sel_r ? F : 1.f - F);
    bs.eta = sel_r ? 1.f : eta_it;
    bs.sampled_component = sel_r ? 0 : 1;
    bs.sampled_type = sel_r ? F_GLOSSY_REFLECTION : F_GLOSSY_TRANSMISSION;
    float3 w = V(1.f, 1.f, 1.f); float dwh_dwo = 0.f;
    if (sel_r) {
        bs.wo = reflect_m(wi, m);
        if (b.tex[B200PT_SLOT_D_SPEC_REFL] >= 0) w = w * tex_eval3(sc, b.tex[B200PT_SLOT_D_SPEC_REFL], uv);
        dwh_dwo = rcp_(4.f * vdot(bs.wo, m));
    }
    if (sel_t) {
        bs.wo = refract_m(wi, m, ctt, eta_ti);
        float3 factor = V(sqr(eta_ti), sqr(eta_ti), sqr(eta_ti));
        if (b.tex[B200PT_SLOT_D_SPEC_TRANS] >= 0) factor = factor * tex_eval3(sc, b.tex[B200PT_SLOT_D_SPEC_TRANS], uv);
        w = w * factor;
        dwh_dwo = fdiv(sqr(bs.eta) * vdot(bs.wo, m), sqr(vdot(wi, m) + bs.eta * vdot(bs.wo, m)));
    }
    w = w * mfd_smith_g1(d, bs.wo, m);
    bs.pdf *= fabsf(dwh_dwo);
    if (active) weight = w;
}

} // namespace pt

namespace pt {

// ---- plastic.cpp:209-380 (smooth plastic, both components enabled). Shares the conductor
// queue / kernel: DevBsdf.type = B200PT_BSDF_CONDUCTOR with PT_M_PLASTIC set (api.cu). ------
PT_DEV float3 plastic_diffuse(const DevScene &sc, const DevBsdf &b, float2 uv) {
    float3 diff = tex_eval3(sc, b.tex[B200PT_SLOT_PL_DIFFUSE], uv);
    float f = b.plastic_fdr_int;
    if (b.flags & B200PT_M_NONLINEAR) return V(fdiv(diff.x, 1.f - diff.x * f), fdiv(diff.y, 1.f - diff.y * f), fdiv(diff.z, 1.f - diff.z * f));
    return V(fdiv(diff.x, 1.f - f), fdiv(diff.y, 1.f - f), fdiv(diff.z, 1.f - f));
}

PT_DEV void plastic_eval_pdf(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, float3 &value, float &pdf) {
    float cti = wi.z, cto = wo.z;
    if (!(cti > 0.f && cto > 0.f)) return;
    float f_i, f_o, t0, t1, t2;
    fresnel(cti, b.eta, f_i, t0, t1, t2); fresnel(cto, b.eta, f_o, t0, t1, t2);
    float3 diff = plastic_diffuse(sc, b, uv);
    float hemi_pdf = PT_INV_PI * cto, inv_eta_2 = fdiv(1.f, b.eta * b.eta);
    value = diff * (hemi_pdf * inv_eta_2 * (1.f - f_i) * (1.f - f_o));
    float w = b.plastic_spec_weight, prob_specular = f_i * w, prob_diffuse = (1.f - f_i) * (1.f - w);
    prob_diffuse = fdiv(prob_diffuse, prob_specular + prob_diffuse);
    pdf = hemi_pdf * prob_diffuse;
}

PT_DEV void plastic_sample(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float s1, float s2x, float s2y, BsdfSample &bs, float3 &weight) {
    float cti = wi.z;
    if (!(cti > 0.f)) return;
    float f_i, t0, t1, t2; fresnel(cti, b.eta, f_i, t0, t1, t2);
    float w = b.plastic_spec_weight, prob_specular = f_i * w, prob_diffuse = (1.f - f_i) * (1.f - w);
    prob_specular = fdiv(prob_specular, prob_specular + prob_diffuse);
    prob_diffuse = 1.f - prob_specular;
    bs.eta = 1.f;
    if (s1 < prob_specular) {
        bs.wo = V(-wi.x, -wi.y, wi.z); bs.pdf = prob_specular; bs.sampled_component = 0; bs.sampled_type = F_DELTA_REFLECTION;
        float v = fdiv(f_i, bs.pdf); float3 val = V(v, v, v);
        if (b.tex[B200PT_SLOT_PL_SPEC_REFL] >= 0) val = val * tex_eval3(sc, b.tex[B200PT_SLOT_PL_SPEC_REFL], uv);
        weight = val;
    } else {
        bs.wo = square_to_cosine_hemisphere(s2x, s2y);
        bs.pdf = prob_diffuse * (PT_INV_PI * bs.wo.z); bs.sampled_component = 1; bs.sampled_type = F_DIFFUSE_REFLECTION;
        float f_o; fresnel(bs.wo.z, b.eta, f_o, t0, t1, t2);
        float3 val = plastic_diffuse(sc, b, uv);
        float inv_eta_2 = fdiv(1.f, b.eta * b.eta);
        weight = val * fdiv(inv_eta_2 * (1.f - f_i) * (1.f - f_o), prob_diffuse);
    }
}

} // namespace pt
