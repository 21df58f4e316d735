// bvh_harness.cpp -- CPU test harness around the product's host-side BVH builder
// (mitsuba3_b200/csrc/bvh.cpp). Built by tests/test_bvh_host.py with g++; not part of the library.
//
// It restates, independently of the CUDA kernels, what the node layout documented in bvh.h promises:
//   * validate(): structural invariants (permutation, leaf sizes, breadth-first numbering, every triangle
//     in exactly one leaf, child boxes enclose the triangles below them)
//   * trace(): a plain stack walk over the 64-byte nodes against a brute-force loop over all triangles,
//     same triangle test, same tie-break (smaller t, then smaller primitive id) as pt::traverse.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bvh.h"

namespace {

struct Handle { pt::Bvh bvh; std::vector<float> tri; };

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

} // extern "C"
