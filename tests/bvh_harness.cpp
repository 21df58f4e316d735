// bvh_harness.cpp -- CPU test harness around the product's host-side BVH builder
// (mitsuba3_b200/csrc/bvh.cpp). Built by tests/test_bvh_host.py with g++; not part of the library.
//
// It restates, independently of the CUDA kernels, what the node layout documented in bvh.h promises:
//   * validate(): structural invariants (permutation, leaf sizes, breadth-first numbering, every triangle
//     in exactly one leaf, child boxes enclose the triangles below them)
//   * the same two checks for the 4-wide tree of pt::collapse_bvh4 (validate4 / trace4)
//   * trace(): a plain stack walk over the 64-byte nodes against a brute-force loop over all triangles,
//     same triangle test, same tie-break (smaller t, then smaller primitive id) as pt::traverse.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bvh.h"

namespace {

struct Handle { pt::Bvh bvh; pt::Bvh4 wide; std::vector<float> tri; };

bool tri_test(const float *p, const float *o, const float *d, float maxt, float &t) {
    float e1[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] }, e2[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
    float pv[3] = { d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0] };
    float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (det == 0.f) return false;
    float inv = 1.f / det, tv[3] = { o[0] - p[0], o[1] - p[1], o[2] - p[2] };
    float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
    float q[3] = { tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0] };
    float v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
    t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    return u >= 0.f && v >= 0.f && u + v <= 1.f && t >= 0.f && t <= maxt;
}

bool slab(const float *lo, const float *hi, const float *o, const float *inv, float tmax) {
    float tmin = 0.f, tmx = tmax;
    for (int a = 0; a < 3; ++a) {
        float t0 = (lo[a] - o[a]) * inv[a], t1 = (hi[a] - o[a]) * inv[a];
        tmin = std::fmax(tmin, std::fmin(t0, t1)); tmx = std::fmin(tmx, std::fmax(t0, t1));
    }
    return tmin <= tmx * 1.0000004f;
}

// boxes of a node: child k in {0, 1} -> lo[3], hi[3] (layout of bvh.h)
void child_box(const pt::BvhNode &n, int k, float lo[3], float hi[3]) {
    const float *f = n.f + 6 * k;
    lo[0] = f[0]; lo[1] = f[1]; lo[2] = f[2]; hi[0] = f[3]; hi[1] = f[4]; hi[2] = f[5];
}

} // namespace

extern "C" {

void *bvh_h_build(const float *tri, uint32_t n) {
    Handle *h = new Handle;
    h->tri.assign(tri, tri + 9 * (size_t) n);
    h->bvh = pt::build_bvh(tri, n);
    h->wide = pt::collapse_bvh4(h->bvh);
    return h;
}
void bvh_h_free(void *p) { delete (Handle *) p; }
uint32_t bvh_h_n_nodes(void *p) { return (uint32_t) ((Handle *) p)->bvh.nodes.size(); }
uint32_t bvh_h_depth(void *p) { return ((Handle *) p)->bvh.depth; }
const void *bvh_h_nodes(void *p) { return ((Handle *) p)->bvh.nodes.data(); }
const uint32_t *bvh_h_order(void *p) { return ((Handle *) p)->bvh.order.data(); }

// 0 = ok, otherwise the number of the violated invariant
int bvh_h_validate(void *p, uint32_t max_leaf) {
    Handle *h = (Handle *) p; const pt::Bvh &b = h->bvh;
    const uint32_t n = (uint32_t) (h->tri.size() / 9);
    if (b.order.size() != n) return 1;
    std::vector<uint8_t> seen(n, 0);
    for (uint32_t t : b.order) { if (t >= n || seen[t]) return 2; seen[t] = 1; }
    if (b.nodes.empty()) return 3;
    std::vector<uint32_t> covered(n, 0);           // leaf-order position -> times referenced
    std::vector<uint8_t> referenced(b.nodes.size(), 0);
    // subtree check with an explicit stack: (node, enclosing lo/hi)
    struct Item { int32_t child; float lo[3], hi[3]; };
    std::vector<Item> st;
    auto push_children = [&](const pt::BvhNode &nd, uint32_t self) -> int {
        const int32_t ch[2] = { nd.left, nd.right };
        for (int k = 0; k < 2; ++k) {
            if (ch[k] == pt::BVH_EMPTY) continue;
            if (ch[k] >= 0 && ((uint32_t) ch[k] <= self || (uint32_t) ch[k] >= b.nodes.size())) return 4;   // breadth-first: children after the parent
            Item it; it.child = ch[k]; child_box(nd, k, it.lo, it.hi); st.push_back(it);
        }
        return 0;
    };
    if (int e = push_children(b.nodes[0], 0)) return e;
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        if (it.child < 0) {
            uint32_t enc = (uint32_t) ~it.child, first = enc >> 3, count = (enc & 7u) + 1u;
            if (first + count > n) return 5;
            if (count > max_leaf && count > 8) return 6;
            for (uint32_t i = first; i < first + count; ++i) {
                covered[i]++;
                const float *v = &h->tri[9 * (size_t) b.order[i]];
                for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a)
                    if (!(v[3 * k + a] >= it.lo[a] && v[3 * k + a] <= it.hi[a])) return 7;    // box encloses its triangles
            }
        } else {
            if (referenced[it.child]) return 8;      // a tree, not a DAG
            referenced[it.child] = 1;
            const pt::BvhNode &nd = b.nodes[it.child];
            // child boxes lie inside the (inflated) parent box, up to the inflation of the children themselves
            for (int k = 0; k < 2; ++k) {
                int32_t c = k ? nd.right : nd.left; if (c == pt::BVH_EMPTY) continue;
                float lo[3], hi[3]; child_box(nd, k, lo, hi);
                for (int a = 0; a < 3; ++a) {
                    float pad = 2e-5f * std::fmax(std::fabs(it.lo[a]), std::fabs(it.hi[a])) + 1e-6f;
                    if (lo[a] < it.lo[a] - pad || hi[a] > it.hi[a] + pad) return 9;
                }
            }
            if (int e = push_children(nd, (uint32_t) it.child)) return e;
        }
    }
    for (uint32_t i = 0; i < n; ++i) if (covered[i] != 1) return 10;   // every triangle in exactly one leaf
    for (size_t i = 1; i < b.nodes.size(); ++i) if (!referenced[i]) return 11;
    return 0;
}

// rays: n x 7 (o, d, maxt). brute != 0: loop over all triangles instead of the tree.
void bvh_h_trace(void *p, uint32_t n_rays, const float *rays, int brute, float *t_out, uint32_t *prim_out, uint64_t *n_tests, uint64_t *n_steps) {
    Handle *h = (Handle *) p; const pt::Bvh &b = h->bvh;
    const uint32_t n = (uint32_t) (h->tri.size() / 9);
    uint64_t tests = 0, steps = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const float *o = rays + 7 * (size_t) r, *d = o + 3; float maxt = o[6];
        float best = INFINITY; uint32_t prim = 0xffffffffu;
        auto test = [&](uint32_t global) {
            float t; tests++;
            if (tri_test(&h->tri[9 * (size_t) global], o, d, maxt, t) && (t < best || (t == best && global < prim))) { best = t; prim = global; }
        };
        if (brute) { for (uint32_t i = 0; i < n; ++i) test(i); }
        else {
            float inv[3]; for (int a = 0; a < 3; ++a) inv[a] = std::fabs(d[a]) > 1e-30f ? 1.f / d[a] : std::copysign(1e30f, d[a]);
            std::vector<int32_t> st; st.push_back(0);
            while (!st.empty()) {
                int32_t c = st.back(); st.pop_back();
                if (c < 0) { uint32_t enc = (uint32_t) ~c, first = enc >> 3, count = (enc & 7u) + 1u; for (uint32_t i = first; i < first + count; ++i) test(b.order[i]); continue; }
                const pt::BvhNode &nd = b.nodes[c]; steps++;
                for (int k = 0; k < 2; ++k) {
                    int32_t ch = k ? nd.right : nd.left; if (ch == pt::BVH_EMPTY) continue;
                    float lo[3], hi[3]; child_box(nd, k, lo, hi);
                    if (slab(lo, hi, o, inv, std::fmin(maxt, best))) st.push_back(ch);
                }
            }
        }
        t_out[r] = best; prim_out[r] = prim;
    }
    if (n_tests) *n_tests = tests;
    if (n_steps) *n_steps = steps;
}

// ---- 4-wide tree (pt::collapse_bvh4) -------------------------------------------------------------

uint32_t bvh_h_n_nodes4(void *p) { return (uint32_t) ((Handle *) p)->wide.nodes.size(); }
uint32_t bvh_h_depth4(void *p) { return ((Handle *) p)->wide.depth; }
const void *bvh_h_nodes4(void *p) { return ((Handle *) p)->wide.nodes.data(); }

int bvh_h_validate4(void *p) {
    Handle *h = (Handle *) p; const pt::Bvh &b = h->bvh; const pt::Bvh4 &w = h->wide;
    const uint32_t n = (uint32_t) (h->tri.size() / 9);
    if (w.nodes.empty()) return 3;
    std::vector<uint32_t> covered(n, 0);
    std::vector<uint8_t> referenced(w.nodes.size(), 0);
    uint32_t next_inner = 1;                         // breadth-first: inner children are numbered in order of appearance
    for (size_t i = 0; i < w.nodes.size(); ++i) {
        const pt::Bvh4Node &nd = w.nodes[i];
        bool seen_empty = false; int occupied = 0;
        for (int c = 0; c < 4; ++c) {
            if (nd.pad[c] != 0) return 12;
            int32_t ch = nd.child[c];
            if (ch == pt::BVH_EMPTY) {
                seen_empty = true;
                for (int a = 0; a < 3; ++a) if (!(nd.lo[a][c] > nd.hi[a][c])) return 13;      // inverted box, never entered
                continue;
            }
            if (seen_empty) return 14;               // occupied children first
            occupied++;
            if (ch >= 0) {
                if ((uint32_t) ch != next_inner || (uint32_t) ch >= w.nodes.size()) return 4;
                next_inner++; referenced[ch] = 1;
            } else {
                uint32_t enc = (uint32_t) ~ch, first = enc >> 3, count = (enc & 7u) + 1u;
                if (first + count > n) return 5;
                for (uint32_t k = first; k < first + count; ++k) {
                    covered[k]++;
                    const float *v = &h->tri[9 * (size_t) b.order[k]];
                    for (int j = 0; j < 3; ++j) for (int a = 0; a < 3; ++a)
                        if (!(v[3 * j + a] >= nd.lo[a][c] && v[3 * j + a] <= nd.hi[a][c])) return 7;
                }
            }
        }
        // a node with an inner child that could still have been opened must be full
        if (i == 0 && n == 0) { if (occupied != 0) return 15; }
    }
    if (next_inner != w.nodes.size()) return 11;
    for (uint32_t i = 0; i < n; ++i) if (covered[i] != 1) return 10;
    // child boxes of an inner child lie inside the box its parent stores for it
    for (size_t i = 0; i < w.nodes.size(); ++i) {
        const pt::Bvh4Node &nd = w.nodes[i];
        for (int c = 0; c < 4; ++c) {
            int32_t ch = nd.child[c]; if (ch < 0 || ch == pt::BVH_EMPTY) continue;
            const pt::Bvh4Node &sub = w.nodes[ch];
            for (int g = 0; g < 4; ++g) {
                if (sub.child[g] == pt::BVH_EMPTY) continue;
                for (int a = 0; a < 3; ++a) if (sub.lo[a][g] < nd.lo[a][c] || sub.hi[a][g] > nd.hi[a][c]) return 9;
            }
        }
    }
    return 0;
}

void bvh_h_trace4(void *p, uint32_t n_rays, const float *rays, float *t_out, uint32_t *prim_out, uint64_t *n_tests, uint64_t *n_steps) {
    Handle *h = (Handle *) p; const pt::Bvh &b = h->bvh; const pt::Bvh4 &w = h->wide;
    uint64_t tests = 0, steps = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const float *o = rays + 7 * (size_t) r, *d = o + 3; float maxt = o[6];
        float best = INFINITY; uint32_t prim = 0xffffffffu;
        float inv[3]; for (int a = 0; a < 3; ++a) inv[a] = std::fabs(d[a]) > 1e-30f ? 1.f / d[a] : std::copysign(1e30f, d[a]);
        std::vector<int32_t> st; st.push_back(0);
        while (!st.empty()) {
            int32_t c = st.back(); st.pop_back();
            if (c < 0) {
                uint32_t enc = (uint32_t) ~c, first = enc >> 3, count = (enc & 7u) + 1u;
                for (uint32_t i = first; i < first + count; ++i) {
                    float t; uint32_t g = b.order[i]; tests++;
                    if (tri_test(&h->tri[9 * (size_t) g], o, d, maxt, t) && (t < best || (t == best && g < prim))) { best = t; prim = g; }
                }
                continue;
            }
            const pt::Bvh4Node &nd = w.nodes[c]; steps++;
            for (int k = 0; k < 4; ++k) {
                if (nd.child[k] == pt::BVH_EMPTY) continue;
                float lo[3] = { nd.lo[0][k], nd.lo[1][k], nd.lo[2][k] }, hi[3] = { nd.hi[0][k], nd.hi[1][k], nd.hi[2][k] };
                if (slab(lo, hi, o, inv, std::fmin(maxt, best))) st.push_back(nd.child[k]);
            }
        }
        t_out[r] = best; prim_out[r] = prim;
    }
    if (n_tests) *n_tests = tests;
    if (n_steps) *n_steps = steps;
}

// The node step of the wide walk exactly as the traversal kernel does it (kernels.cu, k_trace_dyn<.., WIDE>):
// four slab tests with the near distance, misses sorted last by the 5-exchange network, far children pushed
// farthest first, descent into the nearest; a fixed-size stack whose high-water mark is reported.
static bool slab_near(const float *lo, const float *hi, const float *o, const float *inv, float tmax, float &tnear) {
    float t0x = (lo[0] - o[0]) * inv[0], t1x = (hi[0] - o[0]) * inv[0];
    float t0y = (lo[1] - o[1]) * inv[1], t1y = (hi[1] - o[1]) * inv[1];
    float t0z = (lo[2] - o[2]) * inv[2], t1z = (hi[2] - o[2]) * inv[2];
    float tmin = std::fmax(std::fmax(std::fmin(t0x, t1x), std::fmin(t0y, t1y)), std::fmax(std::fmin(t0z, t1z), 0.f));
    float tmx = std::fmin(std::fmin(std::fmax(t0x, t1x), std::fmax(t0y, t1y)), std::fmin(std::fmax(t0z, t1z), tmax));
    tnear = tmin;
    return tmin <= tmx * 1.0000004f;
}

void bvh_h_trace4_ordered(void *p, uint32_t n_rays, const float *rays, int mode, float *t_out, uint32_t *prim_out, uint32_t *max_sp_out, uint64_t *n_steps) {
    const int any_hit = mode & 1; const bool nearest_only = (mode & 2) != 0;   // nearest_only: descend into the nearest child, push the others unsorted
    uint64_t steps = 0;
    Handle *h = (Handle *) p; const pt::Bvh &b = h->bvh; const pt::Bvh4 &w = h->wide;
    const int32_t SENT = 0x76543210;
    int max_sp = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const float *o = rays + 7 * (size_t) r, *d = o + 3; float maxt = o[6];
        float best = INFINITY; uint32_t prim = 0xffffffffu; bool occluded = false;
        float inv[3]; for (int a = 0; a < 3; ++a) inv[a] = std::fabs(d[a]) > 1e-30f ? 1.f / d[a] : std::copysign(1e30f, d[a]);
        int32_t stack[128]; int sp = 0; stack[0] = SENT; int32_t node = 0;
        while (node != SENT && !occluded) {
            if (node < 0) {
                uint32_t enc = (uint32_t) ~node, first = enc >> 3, count = (enc & 7u) + 1u;
                for (uint32_t i = first; i < first + count; ++i) {
                    float t; uint32_t g = b.order[i];
                    if (tri_test(&h->tri[9 * (size_t) g], o, d, maxt, t)) {
                        if (any_hit) occluded = true;
                        else if (t < best || (t == best && g < prim)) { best = t; prim = g; maxt = t; }
                    }
                }
                node = stack[sp--];
                continue;
            }
            const pt::Bvh4Node &nd = w.nodes[node]; steps++;
            float t[4]; int32_t c[4];
            for (int k = 0; k < 4; ++k) {
                float lo[3] = { nd.lo[0][k], nd.lo[1][k], nd.lo[2][k] }, hi[3] = { nd.hi[0][k], nd.hi[1][k], nd.hi[2][k] };
                c[k] = nd.child[k];
                bool hit = slab_near(lo, hi, o, inv, maxt, t[k]) & (c[k] != pt::BVH_EMPTY);
                if (c[k] == pt::BVH_EMPTY) {
                    // the kernel does not look at the child id: the inverted box of an empty child must miss by itself
                    float nr[3], fr[3];
                    for (int a = 0; a < 3; ++a) { bool neg = inv[a] < 0.f; nr[a] = neg ? hi[a] : lo[a]; fr[a] = neg ? lo[a] : hi[a]; }
                    float tn = std::fmax(std::fmax((nr[0] - o[0]) * inv[0], (nr[1] - o[1]) * inv[1]), std::fmax((nr[2] - o[2]) * inv[2], 0.f));
                    float tf = std::fmin(std::fmin((fr[0] - o[0]) * inv[0], (fr[1] - o[1]) * inv[1]), std::fmin((fr[2] - o[2]) * inv[2], maxt));
                    if (tn <= tf * 1.0000004f) { t_out[r] = -3.f; prim_out[r] = 0xfffffffcu; goto next_ray; }
                }
                if (c[k] != pt::BVH_EMPTY) {
                    // the kernel picks the near / far plane by the sign of the direction instead of min/max per axis
                    // (box_hit_nf): must be the same numbers
                    float nr[3], fr[3];
                    for (int a = 0; a < 3; ++a) { bool neg = inv[a] < 0.f; nr[a] = neg ? hi[a] : lo[a]; fr[a] = neg ? lo[a] : hi[a]; }
                    float tn = std::fmax(std::fmax((nr[0] - o[0]) * inv[0], (nr[1] - o[1]) * inv[1]), std::fmax((nr[2] - o[2]) * inv[2], 0.f));
                    float tf = std::fmin(std::fmin((fr[0] - o[0]) * inv[0], (fr[1] - o[1]) * inv[1]), std::fmin((fr[2] - o[2]) * inv[2], maxt));
                    bool hit_nf = tn <= tf * 1.0000004f;
                    if (hit_nf != hit || (hit && tn != t[k])) { t_out[r] = -2.f; prim_out[r] = 0xfffffffdu; goto next_ray; }
                }
                t[k] = hit ? t[k] : INFINITY;
            }
#define CSWAP(i, j) { bool sw = t[j] < t[i]; float tlo = sw ? t[j] : t[i], thi = sw ? t[i] : t[j]; int32_t clo = sw ? c[j] : c[i], chi = sw ? c[i] : c[j]; t[i] = tlo; t[j] = thi; c[i] = clo; c[j] = chi; }
            if (nearest_only) { CSWAP(0, 1) CSWAP(0, 2) CSWAP(0, 3) }      // minimum to the front, the rest as they come
            else {
                CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
                if (!(t[0] <= t[1] && t[1] <= t[2] && t[2] <= t[3])) { t_out[r] = -1.f; prim_out[r] = 0xfffffffeu; goto next_ray; }   // network failed to sort
            }
#undef CSWAP
            if (t[3] < INFINITY) stack[++sp] = c[3];
            if (t[2] < INFINITY) stack[++sp] = c[2];
            if (t[1] < INFINITY) stack[++sp] = c[1];
            if (sp > max_sp) max_sp = sp;
            node = t[0] < INFINITY ? c[0] : stack[sp--];
        }
        t_out[r] = any_hit ? (occluded ? 1.f : 0.f) : best; prim_out[r] = prim;
    next_ray:;
    }
    if (max_sp_out) *max_sp_out = (uint32_t) max_sp;
    if (n_steps) *n_steps = steps;
}

} // extern "C"
