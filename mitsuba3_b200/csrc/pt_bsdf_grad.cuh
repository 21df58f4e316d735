// pt_bsdf_grad.cuh -- parameter derivatives of the non-diffuse BSDF models for the PRB adjoint.
//
// The reference differentiates `bsdf.eval(ctx, si, wo)` with Dr.Jit's reverse-mode AD (prb.py:214-215 for the emitter
// direction, :263-299 for the sampled direction; sampling densities and MIS weights are detached). Here the same
// derivative is obtained with forward-mode dual numbers: `bsdf_eval_dual<TYPE>` evaluates f(wo; theta) once more with
// every model parameter carried as (value, d/d theta_k) where theta_k is ONE channel of ONE texture of the BSDF (all
// slots that reference that texture are seeded, e.g. alpha_u and alpha_v of an isotropic `alpha`). The shading kernel
// loops over the differentiable channels of the material at hand (k_shade<TYPE, true>), so the cost is paid only for
// parameters the caller actually optimises.
//
// The dual evaluations restate the value formulas of pt_principled.cuh / pt_rough.cuh (principled.cpp:489-703,
// roughconductor.cpp:367-425, roughdielectric.cpp:404-480, plastic.cpp:285-330, microfacet.h, fresnel.h:276-314). They do
// not need the primal code's bit-exact operation order: a derivative is compared at 2e-3 against the reference's AD.
// Delta lobes contribute nothing (eval of a delta lobe is 0), as in the reference.
#pragma once
namespace pt {

// the per-model dual evaluations are real calls: four inlined copies of the principled model per shading kernel cost
// 4 KB of register spills and bought nothing
#define PT_DEV_CALL static __device__ __noinline__

struct Dual { float v, d; };
PT_DEV Dual D_(float v, float d = 0.f) { Dual r; r.v = v; r.d = d; return r; }
PT_DEV Dual operator+(Dual a, Dual b) { return D_(a.v + b.v, a.d + b.d); }
PT_DEV Dual operator+(Dual a, float b) { return D_(a.v + b, a.d); }
PT_DEV Dual operator+(float a, Dual b) { return D_(a + b.v, b.d); }
PT_DEV Dual operator-(Dual a, Dual b) { return D_(a.v - b.v, a.d - b.d); }
PT_DEV Dual operator-(Dual a, float b) { return D_(a.v - b, a.d); }
PT_DEV Dual operator-(float a, Dual b) { return D_(a - b.v, -b.d); }
PT_DEV Dual operator-(Dual a) { return D_(-a.v, -a.d); }
PT_DEV Dual operator*(Dual a, Dual b) { return D_(a.v * b.v, a.d * b.v + a.v * b.d); }
PT_DEV Dual operator*(Dual a, float b) { return D_(a.v * b, a.d * b); }
PT_DEV Dual operator*(float a, Dual b) { return D_(a * b.v, a * b.d); }
PT_DEV Dual operator/(Dual a, Dual b) { float r = 1.f / b.v, q = a.v * r; return D_(q, (a.d - q * b.d) * r); }
PT_DEV Dual operator/(Dual a, float b) { float r = 1.f / b; return D_(a.v * r, a.d * r); }
PT_DEV Dual operator/(float a, Dual b) { float r = 1.f / b.v, q = a * r; return D_(q, -q * b.d * r); }
PT_DEV Dual dsqr(Dual a) { return D_(a.v * a.v, 2.f * a.v * a.d); }
PT_DEV Dual dsqrt(Dual a) { float s = sqrtf(fmaxf(a.v, 0.f)); return D_(s, s > 0.f ? 0.5f * a.d / s : 0.f); }      // safe_sqrt
PT_DEV Dual dmax(Dual a, float b) { return a.v >= b ? a : D_(b); }
PT_DEV Dual dabs(Dual a) { return a.v >= 0.f ? a : -a; }
PT_DEV Dual dlog(Dual a) { return D_(logf(a.v), a.d / a.v); }
PT_DEV Dual dexp(Dual a) { float e = expf(a.v); return D_(e, e * a.d); }
PT_DEV Dual dlerp(Dual a, Dual b, Dual t) { return a + (b - a) * t; }
PT_DEV Dual dlerp(float a, Dual b, float t) { return a + (b - a) * t; }
PT_DEV Dual dlerp(float a, float b, Dual t) { return a + (b - a) * t; }
struct Dual3 { Dual x, y, z; };
PT_DEV Dual3 D3(Dual x, Dual y, Dual z) { Dual3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_DEV Dual3 D3(float3 v) { return D3(D_(v.x), D_(v.y), D_(v.z)); }
PT_DEV Dual3 operator+(Dual3 a, Dual3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_DEV Dual3 operator*(Dual3 a, Dual b) { return D3(a.x * b, a.y * b, a.z * b); }
PT_DEV Dual3 operator*(Dual3 a, float b) { return D3(a.x * b, a.y * b, a.z * b); }
PT_DEV Dual3 operator*(Dual3 a, Dual3 b) { return D3(a.x * b.x, a.y * b.y, a.z * b.z); }

// texture value with the seed: d = 1 on channel `seed_ch` if this slot's texture is the seeded one
PT_DEV Dual tex_dual1(const DevScene &sc, int32_t tex, float2 uv, int32_t seed_tex) {
    return D_(tex_eval1(sc, tex, uv), (tex >= 0 && tex == seed_tex) ? 1.f : 0.f);
}
PT_DEV Dual3 tex_dual3(const DevScene &sc, int32_t tex, float2 uv, int32_t seed_tex, int seed_ch) {
    float3 v = tex_eval3(sc, tex, uv);
    bool s = tex >= 0 && tex == seed_tex;
    // a one-channel texture in a colour slot feeds all three channels (tex_eval3 broadcasts it)
    bool mono = s && sc.textures[tex].channels == 1;
    return D3(D_(v.x, s && (mono || seed_ch == 0) ? 1.f : 0.f), D_(v.y, s && (mono || seed_ch == 1) ? 1.f : 0.f), D_(v.z, s && (mono || seed_ch == 2) ? 1.f : 0.f));
}

// ---- microfacet distribution (microfacet.h:185-208, 341-365), alpha as duals -------------------------------------
struct MfdD { bool is_ggx; Dual au, av; };
PT_DEV MfdD mfdD_make(bool is_ggx, Dual au, Dual av) { MfdD d; d.is_ggx = is_ggx; d.au = dmax(au, 1e-4f); d.av = dmax(av, 1e-4f); return d; }
PT_DEV Dual mfdD_eval(const MfdD &d, float3 m) {
    Dual alpha_uv = d.au * d.av; float cos_theta_2 = m.z * m.z;
    Dual e = dsqr(m.x / d.au) + dsqr(m.y / d.av), result;
    if (!d.is_ggx) result = dexp(-e / cos_theta_2) / (PT_PI * alpha_uv * (cos_theta_2 * cos_theta_2));
    else result = 1.f / (PT_PI * alpha_uv * dsqr(e + cos_theta_2));
    return result.v * m.z > 1e-20f ? result : D_(0.f);
}
PT_DEV Dual mfdD_smith_g1(const MfdD &d, float3 v, float3 m) {
    Dual xy_alpha_2 = dsqr(d.au * v.x) + dsqr(d.av * v.y), tan_theta_alpha_2 = xy_alpha_2 / (v.z * v.z), result;
    if (!d.is_ggx) {
        Dual a = 1.f / dsqrt(tan_theta_alpha_2), a_sqr = dsqr(a);
        result = a.v >= 1.6f ? D_(1.f) : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
    } else result = 2.f / (1.f + dsqrt(1.f + tan_theta_alpha_2));
    if (xy_alpha_2.v == 0.f) result = D_(1.f);
    if (vdot(v, m) * v.z <= 0.f) result = D_(0.f);
    return result;
}

// ---- principled (principled.cpp:489-703) -------------------------------------------------------------------------
PT_DEV Dual dschlick(Dual R0, float cos_theta_i, float eta) {        // calc_schlick with a dual R0
    bool outside = cos_theta_i >= 0.f;
    float rcp_eta = 1.f / eta, eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t = safe_sqrt(1.f - (1.f - cos_theta_i * cos_theta_i) * eta_ti * eta_ti);
    float sw = eta_it > 1.f ? schlick_weight(fabsf(cos_theta_i)) : schlick_weight(cos_theta_t);
    return sw + (1.f - sw) * R0;                                   // lerp(sw, 1, R0)
}
PT_DEV Dual dluminance(Dual3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }

PT_DEV_CALL Dual3 principled_eval_dual(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, int32_t st, int sc_) {
    const Dual3 zero = D3(V(0.f, 0.f, 0.f));
    float cos_theta_i = wi.z;
    if (cos_theta_i == 0.f) return zero;
    const uint32_t f = b.flags;
    const bool has_anisotropic = f & B200PT_P_HAS_ANISOTROPIC, has_spec_trans = f & B200PT_P_HAS_SPEC_TRANS, has_sheen = f & B200PT_P_HAS_SHEEN,
               has_sheen_tint = f & B200PT_P_HAS_SHEEN_TINT, has_flatness = f & B200PT_P_HAS_FLATNESS, has_spec_tint = f & B200PT_P_HAS_SPEC_TINT,
               has_metallic = f & B200PT_P_HAS_METALLIC, has_clearcoat = f & B200PT_P_HAS_CLEARCOAT;
    auto P1 = [&](bool has, int slot) { return has ? tex_dual1(sc, b.tex[slot], uv, st) : D_(0.f); };
    Dual anisotropic = P1(has_anisotropic, B200PT_SLOT_P_ANISOTROPIC), roughness = P1(true, B200PT_SLOT_P_ROUGHNESS),
         flatness = P1(has_flatness, B200PT_SLOT_P_FLATNESS), spec_trans = P1(has_spec_trans, B200PT_SLOT_P_SPEC_TRANS),
         metallic = P1(has_metallic, B200PT_SLOT_P_METALLIC), clearcoat = P1(has_clearcoat, B200PT_SLOT_P_CLEARCOAT),
         sheen = P1(has_sheen, B200PT_SLOT_P_SHEEN), spec_tint = P1(has_spec_tint, B200PT_SLOT_P_SPEC_TINT),
         sheen_tint = P1(has_sheen_tint, B200PT_SLOT_P_SHEEN_TINT), clearcoat_gloss = P1(has_clearcoat, B200PT_SLOT_P_CLEARCOAT_GLOSS);
    Dual3 base_color = tex_dual3(sc, b.tex[B200PT_SLOT_P_BASE_COLOR], uv, st, sc_);
    const float eta = b.eta;

    Dual brdf = (1.f - metallic) * (1.f - spec_trans), bsdf = (1.f - metallic) * spec_trans;
    float cos_theta_o = wo.z;
    bool reflect = cos_theta_i * cos_theta_o > 0.f, refract = cos_theta_i * cos_theta_o < 0.f, front_side = cos_theta_i > 0.f;
    float inv_eta = 1.f / eta, eta_path = front_side ? eta : inv_eta, inv_eta_path = front_side ? inv_eta : eta;
    Dual r2 = dsqr(roughness), ax, ay;
    if (!has_anisotropic) ax = ay = dmax(r2, 0.001f);
    else { Dual aspect = dsqrt(1.f - 0.9f * anisotropic); ax = dmax(r2 / aspect, 0.001f); ay = dmax(r2 * aspect, 0.001f); }
    MfdD dist = mfdD_make(true, ax, ay);
    float3 wh = vnormalize(wi + wo * (reflect ? 1.f : eta_path));
    wh = vmulsign(wh, wh.z);
    float F, cos_theta_t, eta_it, eta_ti;
    fresnel(vdot(wi, wh), eta, F, cos_theta_t, eta_it, eta_ti);
    bool refl_c = mac_mic(wh, wi, wo, cos_theta_i, true), refr_c = mac_mic(wh, wi, wo, cos_theta_i, false);
    bool spec_reflect_active = reflect && refl_c && F > 0.f;
    bool clearcoat_active = has_clearcoat && clearcoat.v > 0.f && reflect && refl_c && front_side;
    bool spec_trans_active = has_spec_trans && bsdf.v > 0.f && refract && refr_c && F < 1.f;
    bool diffuse_active = brdf.v > 0.f && reflect && front_side;
    bool sheen_active = has_sheen && sheen.v > 0.f && reflect && (1.f - metallic.v > 0.f) && front_side;
    Dual Dm = mfdD_eval(dist, wh), G = mfdD_smith_g1(dist, wi, wh) * mfdD_smith_g1(dist, wo, wh);
    Dual3 value = zero;
    if (spec_reflect_active) {
        Dual lum = has_spec_tint ? dluminance(base_color) : D_(1.f);
        float cwh = vdot(wi, wh);
        Dual3 Fp;
        if (front_side) {
            Dual3 F_schlick = zero;
            if (has_metallic) F_schlick = F_schlick + D3(dschlick(base_color.x, cwh, eta), dschlick(base_color.y, cwh, eta), dschlick(base_color.z, cwh, eta)) * metallic;
            if (has_spec_tint) {
                bool outside = cwh >= 0.f; float eta_it2 = outside ? eta : 1.f / eta, r0 = schlick_R0_eta(eta_it2);
                Dual3 c_tint = lum.v > 0.f ? D3(base_color.x / lum, base_color.y / lum, base_color.z / lum) : D3(V(1.f, 1.f, 1.f));
                Dual k = (1.f - metallic) * spec_tint;
                F_schlick = F_schlick + D3(dschlick(c_tint.x * r0, cwh, eta), dschlick(c_tint.y * r0, cwh, eta), dschlick(c_tint.z * r0, cwh, eta)) * k;
            }
            Dual fd = (1.f - metallic) * (1.f - spec_tint) * F;
            Fp = D3(fd + F_schlick.x, fd + F_schlick.y, fd + F_schlick.z);
        } else { Dual fb = bsdf * F; Fp = D3(fb, fb, fb); }
        value = value + Fp * (Dm * G / (4.f * fabsf(cos_theta_i)));
    }
    if (spec_trans_active) {
        float scale = inv_eta_path * inv_eta_path, wih = vdot(wi, wh), woh = vdot(wo, wh);
        float geo = scale * (1.f - F) * eta_path * eta_path * wih * woh / (cos_theta_i * sqr(wih + eta_path * woh));
        Dual t = dabs(Dm * G * geo);
        value = value + D3(dsqrt(base_color.x), dsqrt(base_color.y), dsqrt(base_color.z)) * (bsdf * t);
    }
    if (clearcoat_active) {
        float Fcc = calc_schlick(0.04f, vdot(wi, wh), eta);
        Dual alpha = dlerp(0.1f, 0.001f, clearcoat_gloss), alpha2 = dsqr(alpha);
        Dual Dcc = (alpha2 - 1.f) / (PT_PI * dlog(alpha2) * (1.f + (alpha2 - 1.f) * (wh.z * wh.z)));
        if (!(Dcc.v * wh.z > 1e-20f)) Dcc = D_(0.f);
        float Gcc = smith_ggx1(wi, wh, 0.25f) * smith_ggx1(wo, wh, 0.25f);
        Dual c = clearcoat * Dcc * (0.25f * Fcc * Gcc * fabsf(cos_theta_o));
        value = value + D3(c, c, c);
    }
    if (diffuse_active) {
        float Fo = schlick_weight(fabsf(cos_theta_o)), Fi = schlick_weight(fabsf(cos_theta_i));
        float f_diff = (1.f - 0.5f * Fi) * (1.f - 0.5f * Fo);
        float cos_theta_d = vdot(wh, wo);
        Dual Rr = 2.f * roughness * (cos_theta_d * cos_theta_d);
        Dual f_retro = Rr * (Fo + Fi + Fo * Fi * (Rr - 1.f));
        Dual k;
        if (has_flatness) {
            Dual Fss90 = Rr / 2.f;
            Dual Fss = dlerp(1.f, Fss90, Fo) * dlerp(1.f, Fss90, Fi);
            Dual f_ss = 1.25f * (Fss * (1.f / (fabsf(cos_theta_o) + fabsf(cos_theta_i)) - 0.5f) + 0.5f);
            k = dlerp(f_diff + f_retro, f_ss, flatness);
        } else k = f_diff + f_retro;
        value = value + base_color * (brdf * k * (fabsf(cos_theta_o) * PT_INV_PI));
        if (sheen_active) {
            float Fd = schlick_weight(fabsf(cos_theta_d));
            Dual sh = sheen * (1.f - metallic) * Fd;
            if (has_sheen_tint) {
                Dual lum = dluminance(base_color);
                Dual3 c_tint = lum.v > 0.f ? D3(base_color.x / lum, base_color.y / lum, base_color.z / lum) : D3(V(1.f, 1.f, 1.f));
                Dual one = D_(1.f);
                Dual3 c_sheen = D3(dlerp(one, c_tint.x, sheen_tint), dlerp(one, c_tint.y, sheen_tint), dlerp(one, c_tint.z, sheen_tint));
                value = value + c_sheen * (sh * fabsf(cos_theta_o));
            } else { Dual c = sh * fabsf(cos_theta_o); value = value + D3(c, c, c); }
        }
    }
    return value;
}

// ---- rough conductor (roughconductor.cpp:367-425), fresnel_conductor (fresnel.h:276-314) ---------------------------
PT_DEV Dual fresnel_conductor_dual(float cos_theta_i, Dual eta_r, Dual eta_i) {
    float c2 = cos_theta_i * cos_theta_i, s2 = 1.f - c2, s4 = s2 * s2;
    Dual temp_1 = dsqr(eta_r) - dsqr(eta_i) - s2;
    Dual a_2_pb_2 = dsqrt(dsqr(temp_1) + 4.f * dsqr(eta_i) * dsqr(eta_r));
    Dual a = dsqrt(0.5f * (a_2_pb_2 + temp_1));
    Dual term_1 = a_2_pb_2 + c2, term_2 = 2.f * cos_theta_i * a;
    Dual r_s = (term_1 - term_2) / (term_1 + term_2);
    Dual term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
    Dual r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return 0.5f * (r_s + r_p);
}

PT_DEV_CALL Dual3 roughconductor_eval_dual(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, int32_t st, int sc_) {
    const Dual3 zero = D3(V(0.f, 0.f, 0.f));
    float cti = wi.z, cto = wo.z;
    float3 H = vnormalize(wo + wi);
    if (!(cti > 0.f && cto > 0.f && vdot(wi, H) > 0.f && vdot(wo, H) > 0.f)) return zero;
    MfdD d = mfdD_make((b.flags & B200PT_M_GGX) != 0, tex_dual1(sc, b.tex[B200PT_SLOT_ALPHA_U], uv, st), tex_dual1(sc, b.tex[B200PT_SLOT_ALPHA_V], uv, st));
    Dual Dm = mfdD_eval(d, H);
    if (!(Dm.v != 0.f)) return zero;
    Dual G = mfdD_smith_g1(d, wi, H) * mfdD_smith_g1(d, wo, H);
    Dual val = Dm * G / (4.f * cti);
    Dual3 eta = tex_dual3(sc, b.tex[B200PT_SLOT_ETA], uv, st, sc_), k = tex_dual3(sc, b.tex[B200PT_SLOT_K], uv, st, sc_);
    float ci = vdot(wi, H);
    Dual3 F = D3(fresnel_conductor_dual(ci, eta.x, k.x), fresnel_conductor_dual(ci, eta.y, k.y), fresnel_conductor_dual(ci, eta.z, k.z));
    Dual3 v = D3(val, val, val);
    if (b.tex[B200PT_SLOT_SPEC_REFL] >= 0) v = v * tex_dual3(sc, b.tex[B200PT_SLOT_SPEC_REFL], uv, st, sc_);
    return F * v;
}

// ---- rough dielectric (roughdielectric.cpp:404-480) ---------------------------------------------------------------
PT_DEV_CALL Dual3 roughdielectric_eval_dual(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, int32_t st, int sc_) {
    const Dual3 zero = D3(V(0.f, 0.f, 0.f));
    float cti = wi.z, cto = wo.z;
    if (!(cti != 0.f)) return zero;
    bool reflect = cti * cto > 0.f;
    float m_eta = b.eta, m_inv_eta = 1.f / b.eta;
    float eta = cti > 0.f ? m_eta : m_inv_eta, inv_eta = cti > 0.f ? m_inv_eta : m_eta;
    float3 m = vnormalize(wi + wo * (reflect ? 1.f : eta));
    m = vmulsign(m, m.z);
    float dot_wi_m = vdot(wi, m), dot_wo_m = vdot(wo, m);
    if (!(dot_wi_m * cti > 0.f && dot_wo_m * cto > 0.f)) return zero;
    MfdD d = mfdD_make((b.flags & B200PT_M_GGX) != 0, tex_dual1(sc, b.tex[B200PT_SLOT_D_ALPHA_U], uv, st), tex_dual1(sc, b.tex[B200PT_SLOT_D_ALPHA_V], uv, st));
    Dual Dm = mfdD_eval(d, m);
    float F, ctt, eta_it, eta_ti; fresnel(dot_wi_m, m_eta, F, ctt, eta_it, eta_ti);
    Dual G = mfdD_smith_g1(d, wi, m) * mfdD_smith_g1(d, wo, m), val;
    if (reflect) val = Dm * G * (F / (4.f * fabsf(cti)));
    else val = dabs(Dm * G * ((inv_eta * inv_eta) * (1.f - F) * eta * eta * dot_wi_m * dot_wo_m / (cti * sqr(dot_wi_m + eta * dot_wo_m))));
    Dual3 v = D3(val, val, val);
    int slot = reflect ? B200PT_SLOT_D_SPEC_REFL : B200PT_SLOT_D_SPEC_TRANS;
    if (b.tex[slot] >= 0) v = v * tex_dual3(sc, b.tex[slot], uv, st, sc_);
    return v;
}

// ---- smooth plastic (plastic.cpp:285-330): the diffuse component; the specular one is a delta lobe ------------------
PT_DEV_CALL Dual3 plastic_eval_dual(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, int32_t st, int sc_) {
    float cti = wi.z, cto = wo.z;
    if (!(cti > 0.f && cto > 0.f)) return D3(V(0.f, 0.f, 0.f));
    float f_i, f_o, t0, t1, t2;
    fresnel(cti, b.eta, f_i, t0, t1, t2); fresnel(cto, b.eta, f_o, t0, t1, t2);
    Dual3 diff = tex_dual3(sc, b.tex[B200PT_SLOT_PL_DIFFUSE], uv, st, sc_);
    const float fdr = b.plastic_fdr_int;
    if (b.flags & B200PT_M_NONLINEAR) diff = D3(diff.x / (1.f - diff.x * fdr), diff.y / (1.f - diff.y * fdr), diff.z / (1.f - diff.z * fdr));
    else diff = diff * (1.f / (1.f - fdr));
    return diff * (PT_INV_PI * cto / (b.eta * b.eta) * (1.f - f_i) * (1.f - f_o));
}

// f(wo) and d f(wo) / d(texture `seed_tex`, channel `seed_ch`) of the model in queue TYPE (twosided handled by the caller)
template <int TYPE>
PT_DEV Dual3 bsdf_eval_dual(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo, int32_t seed_tex, int seed_ch) {
    if (TYPE == B200PT_BSDF_PRINCIPLED) return principled_eval_dual(sc, b, uv, wi, wo, seed_tex, seed_ch);
    if (TYPE == B200PT_BSDF_CONDUCTOR) {
        if (b.flags & PT_M_PLASTIC) return plastic_eval_dual(sc, b, uv, wi, wo, seed_tex, seed_ch);
        if (b.flags & B200PT_M_ROUGH) return roughconductor_eval_dual(sc, b, uv, wi, wo, seed_tex, seed_ch);
    }
    if (TYPE == B200PT_BSDF_DIELECTRIC && (b.flags & B200PT_M_ROUGH)) return roughdielectric_eval_dual(sc, b, uv, wi, wo, seed_tex, seed_ch);
    return D3(V(0.f, 0.f, 0.f));          // delta lobes: eval = 0 (prb.py: relative_grad of 0)
}

// The differentiable texture of slot `k` of this BSDF, or -1: not set, not differentiable, or already visited through
// an earlier slot (an isotropic `alpha` fills both alpha slots with one texture).
PT_DEV int32_t bsdf_grad_slot(const DevScene &sc, const DevBsdf &b, int k) {
    int32_t t = b.tex[k];
    if (t < 0 || !sc.textures[t].differentiable) return -1;
    for (int j = 0; j < k; ++j) if (b.tex[j] == t) return -1;
    return t;
}

// Coefficient of one parameter channel at one path vertex (prb.py:214-215, 263-299):
//   d/dtheta [ a_dir . f(wo_em) + a_ind . f(wo_s) / f(wo_s) ]  per colour channel of the radiance
template <int TYPE>
PT_DEV float3 bsdf_param_coeff(const DevScene &sc, const DevBsdf &b, float2 uv, float3 wi, float3 wo_em, float3 wo_s, float3 a_dir, float3 a_ind,
                               bool has_dir, bool has_ind, int32_t tex, int ch) {
    float3 c = V(0.f, 0.f, 0.f);
    if (has_dir) {
        Dual3 f = bsdf_eval_dual<TYPE>(sc, b, uv, wi, wo_em, tex, ch);
        c = V(a_dir.x * f.x.d, a_dir.y * f.y.d, a_dir.z * f.z.d);
    }
    if (has_ind) {
        Dual3 f = bsdf_eval_dual<TYPE>(sc, b, uv, wi, wo_s, tex, ch);
        c.x += f.x.v != 0.f ? a_ind.x * f.x.d / f.x.v : 0.f;
        c.y += f.y.v != 0.f ? a_ind.y * f.y.d / f.y.v : 0.f;
        c.z += f.z.v != 0.f ? a_ind.z * f.z.d / f.z.v : 0.f;
    }
    return c;
}

} // namespace pt
