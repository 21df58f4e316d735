// env_host.cpp -- see env_host.h. Plain fp32 host arithmetic in the reference's order
// (double accumulators exactly where the reference uses them).
#include "env_host.h"
#include <cmath>
#include <cstring>
#include <limits>

namespace pt {

static const float kPi = 3.14159265358979323846f;
static const float kRayEps = 1500.f * 5.9604644775390625e-08f;     // core/math.h:18-22 (no Embree)

// drjit/math.h:74-186 sin, float
static float dr_sin(float x) {
    float xa = std::fabs(x);
    int32_t j = (int32_t) (xa * 1.2732395447351626862f);
    j = (j + 1) & ~1;
    float y = (float) j;
    uint32_t xb; std::memcpy(&xb, &x, 4);
    uint32_t sign_sin = (((uint32_t) j) << 29) ^ xb;
    y = xa - y * 0.78515625f - y * 2.4187564849853515625e-4f - y * 3.77489497744594108e-8f;
    float z = y * y;
    if (xa == std::numeric_limits<float>::infinity()) z = std::numeric_limits<float>::quiet_NaN();
    float z2 = z * z;
    float s = std::fmaf(z2, -1.9515295891e-4f, std::fmaf(z, 8.3321608736e-3f, -1.6666654611e-1f)) * z;
    float c = std::fmaf(z2, 2.443315711809948e-5f, std::fmaf(z, -1.388731625493765e-3f, 4.166664568298827e-2f)) * z;
    s = std::fmaf(s, y, y);
    c = std::fmaf(c, z, std::fmaf(z, -0.5f, 1.f));
    float r = (j & 2) == 0 ? s : c;
    return (sign_sin >> 31) ? -r : r;
}

static inline uint32_t h2d_index(uint32_t width, uint32_t x, uint32_t y) {      // distr_2d.h:782-785
    return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1)) + ((y & ~1u) * width);
}

static uint32_t log2i_ceil(uint32_t v) {                                         // core/math.h:197-201
    uint32_t r = 31u - (uint32_t) __builtin_clz(v);
    if ((v & (v - 1u)) != 0u) r += 1u;
    return r;
}

// Hierarchical2D(data, size) with normalize = true, enable_sampling = true (distr_2d.h:403-540)
static void build_hierarchy(const float *data, uint32_t sx, uint32_t sy, EnvHost &h) {
    uint32_t npx = sx - 1, npy = sy - 1;
    uint32_t max_level = log2i_ceil(npx > npy ? npx : npy);
    h.max_patch_index[0] = npx - 1; h.max_patch_index[1] = npy - 1;
    h.patch_size[0] = 1.f / (float) npx; h.patch_size[1] = 1.f / (float) npy;
    h.inv_patch_size[0] = (float) npx; h.inv_patch_size[1] = (float) npy;
    h.lvl_width.assign(1, sx); h.lvl_size.assign(1, sx * sy);
    uint32_t lx = npx, ly = npy;
    for (uint32_t l = 0; l < max_level; ++l) {
        lx += lx & 1u; ly += ly & 1u;
        h.lvl_width.push_back(lx); h.lvl_size.push_back(lx * ly);
        lx >>= 1; ly >>= 1;
    }
    size_t n_levels = h.lvl_width.size();
    h.lvl_offset.resize(n_levels);
    uint32_t total = 0;
    for (size_t l = 0; l < n_levels; ++l) { total = (total + 3u) & ~3u; h.lvl_offset[l] = total; total += h.lvl_size[l]; }
    total = (total + 3u) & ~3u;
    h.warp.assign(total ? total : 4, 0.f);
    float *out = h.warp.data();
    bool has_mip = n_levels > 1;
    double sum = 0.0;
    const float *p = data;
    for (uint32_t y = 0; y < npy; ++y) {
        for (uint32_t x = 0; x < npx; ++x) {
            float avg = .25f * (p[0] + p[1] + p[sx] + p[sx + 1]);
            sum += (double) avg;
            if (has_mip) out[h.lvl_offset[1] + h2d_index(h.lvl_width[1], x, y)] = avg;
            ++p;
        }
        ++p;
    }
    float scale = (float) ((double) (npx * npy) / sum);
    for (uint32_t i = 0; i < h.lvl_size[0]; ++i) out[h.lvl_offset[0] + i] = data[i] * scale;
    if (has_mip) for (uint32_t i = 0; i < h.lvl_size[1]; ++i) out[h.lvl_offset[1] + i] *= scale;
    lx = npx; ly = npy;
    for (size_t l = 2; l < n_levels; ++l) {
        lx = (lx + 1u) >> 1; ly = (ly + 1u) >> 1;
        for (uint32_t y = 0; y < ly; ++y)
            for (uint32_t x = 0; x < lx; ++x) {
                const float *d0 = out + h.lvl_offset[l - 1] + h2d_index(h.lvl_width[l - 1], x * 2, y * 2);
                out[h.lvl_offset[l] + h2d_index(h.lvl_width[l], x, y)] = d0[0] + d0[1] + d0[2] + d0[3];
            }
    }
}

bool build_envmap(const float *env_data, uint32_t W, uint32_t H, bool mis_compensation, EnvHost &out) {
    if (W < 2 || H < 3 || !env_data || (uint64_t) (W + 2) * H >= (1ull << 30)) return false;
    uint32_t sw = W + 2;
    out.tex.assign((size_t) sw * H * 4, 0.f);
    for (uint32_t y = 0; y < H; ++y) {
        float *row = out.tex.data() + (size_t) y * sw * 4;
        for (uint32_t x = 0; x < W; ++x) std::memcpy(row + (size_t) (x + 1) * 4, env_data + ((size_t) y * W + x) * 3, 3 * sizeof(float));
        std::memcpy(row, row + (size_t) W * 4, 4 * sizeof(float));                    // col 0   <- last real column
        std::memcpy(row + (size_t) (W + 1) * 4, row + 4, 4 * sizeof(float));          // col W+1 <- first real column
    }
    // rebuild_distribution (envmap.cpp:474-529): (W + 1) x H luminance grid over storage columns 1..W+1
    uint32_t rx = W + 1, ry = H;
    std::vector<float> lum((size_t) rx * ry);
    for (uint32_t y = 0; y < ry; ++y)
        for (uint32_t x = 0; x < rx; ++x) {
            const float *c = out.tex.data() + ((size_t) y * sw + (x + 1)) * 4;
            lum[(size_t) y * rx + x] = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f;
        }
    float offset = 0.f;
    if (mis_compensation) {
        float min_lum = std::numeric_limits<float>::infinity(); double acc = 0.0;
        for (uint32_t y = 0; y < ry; ++y)
            for (uint32_t x = 0; x < rx - 1u; ++x) { float l = lum[(size_t) y * rx + x]; min_lum = std::fmin(min_lum, l); acc += (double) l; }
        offset = (float) (acc / (double) ((size_t) (rx - 1u) * (size_t) ry));
        if (offset - min_lum <= 0.01f * offset) offset = 0.f;
    }
    float theta_scale = 1.f / (float) (ry - 1) * kPi;
    for (uint32_t y = 0; y < ry; ++y) {
        float sin_theta = dr_sin((float) y * theta_scale);
        for (uint32_t x = 0; x < rx; ++x) { float &l = lum[(size_t) y * rx + x]; l = std::fmax(l - offset, 0.f) * sin_theta; }
    }
    build_hierarchy(lum.data(), rx, ry, out);
    return true;
}

void scene_bounding_sphere(const float *verts8, size_t n_verts, float center[3], float &radius) {
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (size_t v = 0; v < n_verts; ++v)
        for (int k = 0; k < 3; ++k) { float p = verts8[v * 8 + k]; lo[k] = std::fmin(lo[k], p); hi[k] = std::fmax(hi[k], p); }
    if (n_verts && lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) {
        float d[3];
        for (int k = 0; k < 3; ++k) { center[k] = (hi[k] + lo[k]) * .5f; d[k] = center[k] - hi[k]; }
        float r = std::sqrt(std::fmaf(d[2], d[2], std::fmaf(d[1], d[1], d[0] * d[0])));      // bbox.h:343-346
        radius = std::fmax(kRayEps, r * (1.f + kRayEps));
    } else {
        center[0] = center[1] = center[2] = 0.f; radius = kRayEps;
    }
}

} // namespace pt
