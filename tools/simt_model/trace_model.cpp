// trace_model.cpp -- SIMT issue model of the traversal kernel (design tooling, NOT product code and
// not a measurement). It replays the control flow of k_trace_dyn / k_trace (kernels.cu) warp by warp
// on the CPU for a path-traced scene and counts, per code region, the warp instructions issued and the
// threads active in them. Region costs are instruction counts read off the SASS of the shipped kernel
// (cuobjdump -sass, see profiles/r01_simt_model.md). The kernel is issue-bound (ncu: SM throughput 82 %,
// 13 of 32 threads active), so issued warp instructions are the quantity that sets its run time.
//
//   g++ -O2 -std=c++17 -I../../mitsuba3_b200/csrc trace_model.cpp ../../mitsuba3_b200/csrc/bvh.cpp -o /tmp/trace_model
//   /tmp/trace_model /tmp/cornell.bin [--res 256] [--spp 16] [--idle 8] [--static] [--wide 4] ...
//
// The BVH is built by the product's own builder (bvh.cpp). The rays come from a simplified path tracer
// (diffuse surfaces, one area light, NEE shadow rays, Russian roulette from depth 5, 8 bounces) -- the
// wave structure of the real integrator, not its radiometry.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

#include "bvh.h"

using pt::BvhNode;

// ------------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------------
struct Costs {
    // read off the SASS of k_trace_dyn<false, true> (16 B per instruction)
    int node = 93;        // one iteration of the inner-node loop (two slab tests, ordering, stack)    0x2540..0x2b40
    int tri = 68;         // one Moeller-Trumbore iteration incl. the IEEE reciprocal fast path         0x2c80..0x3160
    int leaf_ovh = 28;    // per leaf: decode, loop set-up, pop of the next postponed leaf            0x2ba0..0x2c80 + 0x3160..0x3240
    int outer = 19;       // per pass of the outer while: head + refill-break test                     0x3260..0x3390
    int head = 12;        // ballot / any / branch at the top of the job loop
    int refill_common = 25, refill_start = 85;   // counter fetch; per started ray (loads, 3 IEEE reciprocals)
    int ret_shadow_add = 18, ret_restart = 85, ret_finish = 15, ret_hit = 30, ret_miss = 15;
    int bucket_empty = 6, bucket_hit = 35;       // per queue of the bucket pass
    int n_queues = 5;
    // wide-node variants: cost of one node iteration with W children
    int wide_override = 0;
    int node_wide(int w) const { return wide_override ? wide_override : 40 + 28 * w; }   // loads + W slab tests + ordered push
    int node_coop = 62;   // cooperative node step: own child box (2 loads, 1 slab test), rank by shuffles, ordered push
    int leaf_coop = 92;   // cooperative leaf: one triangle per lane + shuffle reduction of (t, prim), incl. leaf overhead
};

struct Options {
    std::string scene;
    int res = 256, spp = 16, max_depth = 8, rr_depth = 5;
    int idle = 8;
    bool dynamic = true;
    int wide = 2;
    int warps = 148 * 5 * 8;
    int tri_reject = 0;       // > 0: cost of a triangle iteration that the sign prefilter rejects
    int coop = 0;             // > 0: `coop` lanes share one ray of a `coop`-wide tree: one child box / one leaf triangle per lane
    bool leaf_once = false;   // one round of triangle tests per pass of the outer loop (lanes holding a second leaf keep it for the next round)
    bool split_inplace = false; // same launch, work counter over [0, n) = shadow jobs, [n, 2n) = path jobs (slots without one idle)
    bool split_phases = false; // shadow rays of a wave in their own launch (no shadow/path mix inside a warp)
    int sort_bits = 0;        // > 0: sort the slots of each wave by a Morton key of (origin cell, direction octant)
    int window = 0;           // > 0: after ordering, shuffle the slots inside consecutive windows of this many entries
                              //      (the shading kernels compact survivors in atomic order: order is only kept at the
                              //      granularity of one grid-stride iteration)
    bool verbose = false;
};

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
static inline V3 normalize(V3 a) { return a * (1.f / std::sqrt(dot(a, a))); }

struct Tri { V3 p0, e1, e2, n; float albedo_max; bool emitter; };
struct Scene {
    std::vector<Tri> tris;          // in BVH leaf order
    std::vector<uint32_t> lights;   // indices into tris
    float cam[7];
    // BVH as generic wide nodes: children boxes + child ids (>= 0 inner, < 0 leaf ~((first << 3) | (count - 1)))
    struct Node { int n; float lo[8][3], hi[8][3]; int32_t child[8]; };
    std::vector<Node> nodes;
};

static const int32_t SENT = 0x76543210;
static int WS = 32;   // rays per warp
static int TH = 1;    // threads cooperating on one ray (--coop: TH = tree width, WS = 32 / TH)

static void load_scene(const Options &opt, Scene &sc) {
    FILE *f = fopen(opt.scene.c_str(), "rb");
    if (!f) { perror("scene"); exit(1); }
    uint32_t hdr[2];
    if (fread(hdr, 4, 2, f) != 2 || fread(sc.cam, 4, 7, f) != 7) exit(1);
    std::vector<float> raw((size_t) hdr[0] * 13);
    if (fread(raw.data(), 4, raw.size(), f) != raw.size()) exit(1);
    fclose(f);
    std::vector<float> tri9((size_t) hdr[0] * 9);
    for (uint32_t i = 0; i < hdr[0]; ++i) memcpy(&tri9[9 * i], &raw[13 * i], 36);
    pt::Bvh bvh = pt::build_bvh(tri9.data(), hdr[0]);
    sc.tris.resize(hdr[0]);
    for (uint32_t k = 0; k < hdr[0]; ++k) {
        const float *r = &raw[13 * (size_t) bvh.order[k]];
        Tri t; t.p0 = { r[0], r[1], r[2] };
        t.e1 = V3{ r[3], r[4], r[5] } - t.p0; t.e2 = V3{ r[6], r[7], r[8] } - t.p0;
        t.n = normalize(cross(t.e1, t.e2));
        t.albedo_max = std::max(r[9], std::max(r[10], r[11])); t.emitter = r[12] > 0.f;
        sc.tris[k] = t;
        if (t.emitter) sc.lights.push_back(k);
    }
    // binary nodes -> generic nodes
    std::vector<Scene::Node> bin(bvh.nodes.size());
    for (size_t i = 0; i < bvh.nodes.size(); ++i) {
        const BvhNode &b = bvh.nodes[i]; Scene::Node n; n.n = 0;
        const float L[6] = { b.f[0], b.f[1], b.f[2], b.f[3], b.f[4], b.f[5] }, R[6] = { b.f[6], b.f[7], b.f[8], b.f[9], b.f[10], b.f[11] };
        if (b.left != pt::BVH_EMPTY) { for (int a = 0; a < 3; ++a) { n.lo[n.n][a] = L[a]; n.hi[n.n][a] = L[3 + a]; } n.child[n.n++] = b.left; }
        if (b.right != pt::BVH_EMPTY) { for (int a = 0; a < 3; ++a) { n.lo[n.n][a] = R[a]; n.hi[n.n][a] = R[3 + a]; } n.child[n.n++] = b.right; }
        bin[i] = n;
    }
    if (opt.wide <= 2) { sc.nodes = bin; return; }
    // collapse to width W: repeatedly replace the inner child with the largest surface area by its children
    std::vector<int32_t> remap(bin.size(), -1);
    std::vector<Scene::Node> out; out.reserve(bin.size());
    std::vector<std::pair<int32_t, int32_t>> todo;   // (binary node, wide node)
    out.push_back(Scene::Node()); todo.push_back({ 0, 0 });
    while (!todo.empty()) {
        auto [bi, wi] = todo.back(); todo.pop_back();
        Scene::Node w = bin[bi];
        while (true) {
            int best = -1; float best_area = -1.f;
            for (int c = 0; c < w.n; ++c) {
                if (w.child[c] < 0) continue;
                if (w.n - 1 + bin[w.child[c]].n > opt.wide) continue;
                float dx = w.hi[c][0] - w.lo[c][0], dy = w.hi[c][1] - w.lo[c][1], dz = w.hi[c][2] - w.lo[c][2];
                float area = dx * dy + dy * dz + dz * dx;
                if (area > best_area) { best_area = area; best = c; }
            }
            if (best < 0) break;
            Scene::Node ch = bin[w.child[best]];
            // replace slot `best` by the first grandchild, append the others
            for (int a = 0; a < 3; ++a) { w.lo[best][a] = ch.lo[0][a]; w.hi[best][a] = ch.hi[0][a]; }
            w.child[best] = ch.child[0];
            for (int g = 1; g < ch.n; ++g) {
                for (int a = 0; a < 3; ++a) { w.lo[w.n][a] = ch.lo[g][a]; w.hi[w.n][a] = ch.hi[g][a]; }
                w.child[w.n++] = ch.child[g];
            }
        }
        for (int c = 0; c < w.n; ++c)
            if (w.child[c] >= 0) { int32_t id = (int32_t) out.size(); out.push_back(Scene::Node()); todo.push_back({ w.child[c], id }); w.child[c] = id; }
        out[wi] = w;
    }
    sc.nodes = out;
}

// ------------------------------------------------------------------------------------------------
// rays, waves
// ------------------------------------------------------------------------------------------------
struct Ray { V3 o, d; float maxt; };
struct Slot { bool has_shadow, alive; Ray sh, path; int depth; float thr; };
struct HitRec { float t; uint32_t prim; };

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    float next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float) ((s >> 40) * (1.0 / 16777216.0)); }
};

static bool tri_hit(const Tri &t, const Ray &r, float maxt, float &tt) {
    V3 pvec = cross(r.d, t.e2); float det = dot(t.e1, pvec);
    if (det == 0.f) return false;
    float inv = 1.f / det; V3 tv = r.o - t.p0; float u = dot(tv, pvec) * inv;
    if (u < 0.f || u > 1.f) return false;
    V3 q = cross(tv, t.e1); float v = dot(r.d, q) * inv;
    if (v < 0.f || u + v > 1.f) return false;
    tt = dot(t.e2, q) * inv;
    return tt >= 0.f && tt <= maxt;
}
// would the sign-only prefilter (no reciprocal) reject this triangle?
static bool tri_prefilter_rejects(const Tri &t, const Ray &r, float maxt) {
    V3 pvec = cross(r.d, t.e2); float det = dot(t.e1, pvec);
    if (det == 0.f) return true;
    V3 tv = r.o - t.p0; float un = dot(tv, pvec); float s = det > 0 ? 1.f : -1.f; float ad = std::fabs(det);
    if (un * s < 0.f || un * s > ad) return true;
    V3 q = cross(tv, t.e1); float vn = dot(r.d, q);
    if (vn * s < 0.f || (un + vn) * s > ad) return true;
    float tn = dot(t.e2, q);
    return tn * s < 0.f || tn * s > maxt * ad;
}

static inline bool box_hit(const float lo[3], const float hi[3], const V3 &o, const V3 &inv, float tmax, float &tnear) {
    float t0x = (lo[0] - o.x) * inv.x, t1x = (hi[0] - o.x) * inv.x;
    float t0y = (lo[1] - o.y) * inv.y, t1y = (hi[1] - o.y) * inv.y;
    float t0z = (lo[2] - o.z) * inv.z, t1z = (hi[2] - o.z) * inv.z;
    float tmin = std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::max(std::min(t0z, t1z), 0.f));
    float tmx = std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::min(std::max(t0z, t1z), tmax));
    tnear = tmin;
    return tmin <= tmx * 1.0000004f;
}
static inline float safe_inv(float d) { return std::fabs(d) > 1e-30f ? 1.f / d : std::copysign(1e30f, d); }

// ------------------------------------------------------------------------------------------------
// the warp model
// ------------------------------------------------------------------------------------------------
enum Region { R_NODE, R_TRI, R_LEAF, R_OUTER, R_HEAD, R_REFILL, R_RETIRE, R_BUCKET, R_COUNT };
static const char *region_name[R_COUNT] = { "node loop", "triangle tests", "leaf overhead", "outer loop", "job-loop head", "refill", "retire", "bucket pass" };

struct Counters {
    double warp_instr[R_COUNT] = {}, thread_instr[R_COUNT] = {};
    double rays = 0, node_steps = 0, tri_tests = 0;
    void add(Region r, int cost, int active) { warp_instr[r] += cost; thread_instr[r] += (double) cost * active * TH; }
    double total_warp() const { double s = 0; for (double v : warp_instr) s += v; return s; }
    double total_thread() const { double s = 0; for (double v : thread_instr) s += v; return s; }
};

struct Lane {
    int kind = 0; uint32_t slot = 0; Ray r; V3 inv; float maxt; float hit_t; uint32_t hit_prim;
    int32_t stack[64]; int sp; int32_t node, leaf; bool occluded; bool searching;
};

struct Warp { Lane ln[32]; bool exhausted = false; double clock = 0; uint32_t static_base = 0; bool static_loaded = false; };

struct Wave {
    const Scene *sc; const Options *opt; const Costs *cost;
    std::vector<Slot> *slots; std::vector<HitRec> *hits; std::vector<uint32_t> *retire_order;
    uint32_t work_counter = 0; bool first;
    Counters cnt;
};

static void start_ray(Lane &l, const Ray &r) {
    l.r = r; l.maxt = r.maxt; l.inv = { safe_inv(r.d.x), safe_inv(r.d.y), safe_inv(r.d.z) };
    l.hit_t = INFINITY; l.hit_prim = 0xffffffffu; l.stack[0] = SENT; l.sp = 0; l.node = 0; l.leaf = 0; l.occluded = false;
}

static void node_step(Wave &w, Lane &l) {
    const Scene::Node &n = w.sc->nodes[l.node];
    // children hit, sorted near to far
    int32_t ids[8]; float ts[8]; int k = 0;
    for (int c = 0; c < n.n; ++c) {
        float t;
        if (box_hit(n.lo[c], n.hi[c], l.r.o, l.inv, l.maxt, t)) {
            int j = k++;
            while (j > 0 && ts[j - 1] > t) { ts[j] = ts[j - 1]; ids[j] = ids[j - 1]; --j; }
            ts[j] = t; ids[j] = n.child[c];
        }
    }
    if (k == 0) l.node = l.stack[l.sp--];
    else {
        for (int j = k - 1; j >= 1; --j) l.stack[++l.sp] = ids[j];
        l.node = ids[0];
    }
    if (l.node < 0 && l.leaf >= 0) { l.searching = false; l.leaf = l.node; l.node = l.stack[l.sp--]; }
    w.cnt.node_steps++;
}

// traversal section of one pass of the job loop; `dyn_break` = leave when too few lanes still walk
static void traverse_section(Wave &w, Warp &wp, bool dyn_break) {
    const Costs &C = *w.cost; const Options &O = *w.opt;
    const int node_cost = O.coop ? C.node_coop : O.wide <= 2 ? C.node : C.node_wide(O.wide);
    bool in_outer[32]; int n_outer = 0;
    for (int i = 0; i < WS; ++i) { in_outer[i] = wp.ln[i].kind != 0 && (wp.ln[i].node != SENT || wp.ln[i].leaf < 0); n_outer += in_outer[i]; }
    while (n_outer > 0) {
        // ---- inner-node loop ----
        bool in_node[32]; int n_node = 0;
        for (int i = 0; i < WS; ++i) {
            Lane &l = wp.ln[i];
            if (in_outer[i]) l.searching = O.leaf_once ? l.leaf >= 0 : true;
            in_node[i] = in_outer[i] && l.node >= 0 && l.node != SENT; n_node += in_node[i];
        }
        while (n_node > 0) {
            w.cnt.add(R_NODE, node_cost, n_node);
            bool any_searching = false;
            for (int i = 0; i < WS; ++i) if (in_node[i]) { node_step(w, wp.ln[i]); any_searching |= wp.ln[i].searching; }
            if (!any_searching) break;
            n_node = 0;
            for (int i = 0; i < WS; ++i) { in_node[i] = in_node[i] && wp.ln[i].node >= 0 && wp.ln[i].node != SENT; n_node += in_node[i]; }
        }
        // ---- leaf loop ----
        while (true) {
            int n_leaf = 0, maxcount = 0; int cnts[32];
            for (int i = 0; i < WS; ++i) {
                cnts[i] = 0;
                if (in_outer[i] && wp.ln[i].leaf < 0) { uint32_t enc = (uint32_t) ~wp.ln[i].leaf; cnts[i] = (int) (enc & 7u) + 1; n_leaf++; maxcount = std::max(maxcount, cnts[i]); }
            }
            if (!n_leaf) break;
            for (int k = 0; k < maxcount; ++k) {
                const bool charge = !O.coop || k == 0;     // cooperative leaf: all its triangles in one iteration
                int act = 0, act_full = 0;
                for (int i = 0; i < WS; ++i) {
                    if (cnts[i] <= k) continue;
                    Lane &l = wp.ln[i]; act++;
                    uint32_t first = ((uint32_t) ~l.leaf) >> 3, ti = first + (uint32_t) k;
                    const Tri &t = w.sc->tris[ti]; float tt;
                    bool rejected = O.tri_reject > 0 && tri_prefilter_rejects(t, l.r, l.maxt);
                    if (!rejected) act_full++;
                    if (tri_hit(t, l.r, l.maxt, tt)) {
                        if (l.kind == 1) l.occluded = true;
                        else if (tt < l.hit_t || (tt == l.hit_t && ti < l.hit_prim)) { l.hit_t = tt; l.hit_prim = ti; l.maxt = tt; }
                    }
                    w.cnt.tri_tests++;
                }
                if (!charge) continue;
                if (O.coop) { w.cnt.add(R_TRI, C.leaf_coop, n_leaf); continue; }
                if (O.tri_reject > 0) {
                    // prefilter executed by all, the exact tail only by the lanes that pass it
                    w.cnt.add(R_TRI, O.tri_reject, act);
                    if (act_full) w.cnt.add(R_TRI, C.tri - O.tri_reject + 8, act_full);
                } else w.cnt.add(R_TRI, C.tri, act);
            }
            if (!O.coop) w.cnt.add(R_LEAF, C.leaf_ovh, n_leaf);
            for (int i = 0; i < WS; ++i) {
                if (!cnts[i]) continue;
                Lane &l = wp.ln[i];
                l.leaf = l.node;
                if (l.node < 0) l.node = l.stack[l.sp--];
                if (l.occluded) { l.node = SENT; l.leaf = 0; }
            }
            if (O.leaf_once) break;
        }
        w.cnt.add(R_OUTER, C.outer, n_outer);
        if (dyn_break && !wp.exhausted && n_outer < WS - std::max(1, O.idle * WS / 32)) break;
        n_outer = 0;
        for (int i = 0; i < WS; ++i) { in_outer[i] = in_outer[i] && (wp.ln[i].node != SENT || wp.ln[i].leaf < 0); n_outer += in_outer[i]; }
    }
}

static void begin_job(Wave &w, Lane &l, uint32_t slot, int &n_sh, int &n_pa) {
    if (w.opt->split_inplace && !w.first) {
        const uint32_t n = (uint32_t) w.slots->size();
        bool shadow = slot < n; if (!shadow) slot -= n;
        const Slot &s = (*w.slots)[slot];
        l.slot = slot;
        if (shadow) { if (!s.has_shadow) return; l.kind = 1; start_ray(l, s.sh); n_sh++; }
        else { if (!s.alive) return; l.kind = 2; start_ray(l, s.path); n_pa++; }
        w.cnt.rays++;
        return;
    }
    const Slot &s = (*w.slots)[slot];
    l.slot = slot;
    if (!w.first && s.has_shadow) { l.kind = 1; start_ray(l, s.sh); n_sh++; }
    else { l.kind = 2; start_ray(l, s.path); n_pa++; }
    w.cnt.rays++;
}

// retire finished rays of the warp; returns per-queue hit flags for the bucket pass
static void retire_section(Wave &w, Warp &wp) {
    const Costs &C = *w.cost;
    int n_add = 0, n_restart = 0, n_finish = 0, n_hit = 0, n_miss = 0; bool any_bucket = false;
    for (int i = 0; i < WS; ++i) {
        Lane &l = wp.ln[i];
        if (l.kind == 0 || l.node != SENT || l.leaf < 0) continue;
        const Slot &s = (*w.slots)[l.slot];
        if (l.kind == 1) {
            if (!l.occluded) n_add++;
            if (s.alive && w.opt->split_inplace) l.kind = 0;
            else if (s.alive) { n_restart++; l.kind = 2; start_ray(l, s.path); w.cnt.rays++; }
            else { n_finish++; l.kind = 0; (*w.hits)[l.slot] = { INFINITY, 0xffffffffu }; }
        } else {
            (*w.hits)[l.slot] = { l.hit_t, l.hit_prim };
            if (l.hit_prim != 0xffffffffu) { n_hit++; any_bucket = true; w.retire_order->push_back(l.slot); } else n_miss++;
            l.kind = 0;
        }
    }
    if (n_add) w.cnt.add(R_RETIRE, C.ret_shadow_add, n_add);
    if (n_restart) w.cnt.add(R_RETIRE, C.ret_restart, n_restart);
    if (n_finish) w.cnt.add(R_RETIRE, C.ret_finish, n_finish);
    if (n_hit) w.cnt.add(R_RETIRE, C.ret_hit, n_hit);
    if (n_miss) w.cnt.add(R_RETIRE, C.ret_miss, n_miss);
    w.cnt.add(R_BUCKET, C.bucket_empty * C.n_queues + (any_bucket ? C.bucket_hit : 0), WS);
}

// one pass of the job loop of k_trace_dyn; false when the warp is done
static bool dyn_pass(Wave &w, Warp &wp) {
    const Costs &C = *w.cost; const Options &O = *w.opt;
    const uint32_t n = (uint32_t) w.slots->size() * ((w.opt->split_inplace && !w.first) ? 2u : 1u);
    double before = w.cnt.total_warp();
    int idle = 0; for (int i = 0; i < WS; ++i) idle += wp.ln[i].kind == 0;
    w.cnt.add(R_HEAD, C.head, WS);
    if (!wp.exhausted && idle >= std::max(1, O.idle * WS / 32)) {
        uint32_t base = w.work_counter; w.work_counter += (uint32_t) idle;
        if (base + idle >= n) wp.exhausted = true;
        w.cnt.add(R_REFILL, C.refill_common, WS);
        int n_sh = 0, n_pa = 0; uint32_t k = 0;
        for (int i = 0; i < WS; ++i) if (wp.ln[i].kind == 0) { uint32_t s = base + k++; if (s < n) begin_job(w, wp.ln[i], s, n_sh, n_pa); }
        if (n_sh) w.cnt.add(R_REFILL, C.refill_start, n_sh);
        if (n_pa) w.cnt.add(R_REFILL, C.refill_start, n_pa);
    }
    bool any = false; for (int i = 0; i < WS; ++i) any |= wp.ln[i].kind != 0;
    if (!any) { wp.clock += w.cnt.total_warp() - before; return !wp.exhausted; }   // empty jobs (split ranges): fetch again
    traverse_section(w, wp, true);
    retire_section(w, wp);
    wp.clock += w.cnt.total_warp() - before;
    return true;
}

// k_trace (static assignment): 32 consecutive slots, shadow ray then path ray, bucket pass per batch
static bool static_pass(Wave &w, Warp &wp, uint32_t n_warps) {
    const Costs &C = *w.cost;
    const uint32_t n = (uint32_t) w.slots->size();
    double before = w.cnt.total_warp();
    if (!wp.static_loaded) { wp.static_loaded = true; } else wp.static_base += n_warps * (uint32_t) WS;
    if (wp.static_base >= n) return false;
    int n_sh = 0, n_pa = 0;
    wp.exhausted = true;
    // phase 1: shadow rays
    for (int i = 0; i < WS; ++i) {
        Lane &l = wp.ln[i]; l.kind = 0; uint32_t s = wp.static_base + i;
        if (s < n && !w.first && (*w.slots)[s].has_shadow) { l.slot = s; l.kind = 1; start_ray(l, (*w.slots)[s].sh); n_sh++; w.cnt.rays++; }
    }
    if (n_sh) { w.cnt.add(R_REFILL, C.refill_start, n_sh); traverse_section(w, wp, false); }
    int n_add = 0; for (int i = 0; i < WS; ++i) if (wp.ln[i].kind == 1 && !wp.ln[i].occluded) n_add++;
    if (n_add) w.cnt.add(R_RETIRE, C.ret_shadow_add, n_add);
    // phase 2: path rays
    for (int i = 0; i < WS; ++i) {
        Lane &l = wp.ln[i]; l.kind = 0; uint32_t s = wp.static_base + i;
        if (s < n) { (*w.hits)[s] = { INFINITY, 0xffffffffu }; if (w.first || (*w.slots)[s].alive) { l.slot = s; l.kind = 2; start_ray(l, (*w.slots)[s].path); n_pa++; w.cnt.rays++; } }
    }
    if (n_pa) { w.cnt.add(R_REFILL, C.refill_start, n_pa); traverse_section(w, wp, false); }
    retire_section(w, wp);
    w.cnt.add(R_HEAD, C.head, WS);
    wp.clock += w.cnt.total_warp() - before;
    return true;
}

static Counters run_wave(const Scene &sc, const Options &opt, const Costs &cost, std::vector<Slot> &slots, std::vector<HitRec> &hits,
                         std::vector<uint32_t> &retire_order, bool first) {
    Wave w; w.sc = &sc; w.opt = &opt; w.cost = &cost; w.slots = &slots; w.hits = &hits; w.retire_order = &retire_order; w.first = first;
    hits.assign(slots.size(), { INFINITY, 0xffffffffu }); retire_order.clear();
    uint32_t n_warps = (uint32_t) std::min<size_t>((size_t) opt.warps, (slots.size() * TH + 255) / 256 * 8);
    std::vector<Warp> warps(n_warps);
    for (uint32_t i = 0; i < n_warps; ++i) warps[i].static_base = i * (uint32_t) WS;
    using QE = std::pair<double, uint32_t>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (uint32_t i = 0; i < n_warps; ++i) pq.push({ 0.0, i });
    while (!pq.empty()) {
        uint32_t wi = pq.top().second; pq.pop();
        bool more = opt.dynamic ? dyn_pass(w, warps[wi]) : static_pass(w, warps[wi], n_warps);
        if (more) pq.push({ warps[wi].clock, wi });
    }
    return w.cnt;
}

// ------------------------------------------------------------------------------------------------
// simplified path tracer producing the waves
// ------------------------------------------------------------------------------------------------
static V3 cosine_dir(V3 n, Rng &rng) {
    float u1 = rng.next(), u2 = rng.next(); float r = std::sqrt(u1), phi = 6.2831853f * u2;
    V3 a = std::fabs(n.x) > 0.5f ? V3{ 0, 1, 0 } : V3{ 1, 0, 0 }; V3 s = normalize(cross(n, a)), t = cross(n, s);
    return normalize(s * (r * std::cos(phi)) + t * (r * std::sin(phi)) + n * std::sqrt(std::max(0.f, 1.f - u1)));
}

static uint32_t morton_key(const Ray &r, int bits) {
    auto q = [&](float v) { int x = (int) ((v * 0.5f + 0.5f) * (1 << bits)); return (uint32_t) std::min(std::max(x, 0), (1 << bits) - 1); };
    uint32_t oct = (r.d.x < 0) | ((r.d.y < 0) << 1) | ((r.d.z < 0) << 2);
    if (bits >= 100) return oct;                       // direction octant only
    bool cell_only = bits >= 10; if (cell_only) bits -= 10;   // 11, 12, 13: origin cell only
    uint32_t x = q(r.o.x), y = q(r.o.y), z = q(r.o.z), key = 0;
    for (int b = bits - 1; b >= 0; --b) key = (key << 3) | (((x >> b) & 1) << 2) | (((y >> b) & 1) << 1) | ((z >> b) & 1);
    return cell_only ? key : (key << 3) | oct;
}

int main(int argc, char **argv) {
    Options opt; Costs cost;
    if (argc < 2) { fprintf(stderr, "usage: trace_model scene.bin [options]\n"); return 1; }
    opt.scene = argv[1];
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() { return atoi(argv[++i]); };
        if (a == "--res") opt.res = val(); else if (a == "--spp") opt.spp = val(); else if (a == "--idle") opt.idle = val();
        else if (a == "--static") opt.dynamic = false; else if (a == "--wide") opt.wide = val(); else if (a == "--warps") opt.warps = val();
        else if (a == "--tri-reject") opt.tri_reject = val(); else if (a == "--split") opt.split_phases = true; else if (a == "--split2") opt.split_inplace = true; else if (a == "--leaf-once") opt.leaf_once = true; else if (a == "--window") opt.window = val(); else if (a == "--coop") { opt.coop = val(); opt.wide = opt.coop; TH = opt.coop; WS = 32 / TH; }
        else if (a == "--node-coop") cost.node_coop = val(); else if (a == "--wide-cost") cost.wide_override = val();
        else if (a == "--sort") opt.sort_bits = val(); else if (a == "--node-cost") cost.node = val(); else if (a == "--tri-cost") cost.tri = val();
        else if (a == "--retire-scale") { int p = val(); cost.ret_restart = cost.ret_restart * p / 100; cost.ret_hit = cost.ret_hit * p / 100; cost.refill_start = cost.refill_start * p / 100; }
        else if (a == "-v") opt.verbose = true;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    Scene sc; load_scene(opt, sc);
    printf("scene %s: %zu triangles, %zu nodes (width %d); %dx%d, %d spp, idle %d, %s fetch, %d resident warps\n", opt.scene.c_str(), sc.tris.size(),
           sc.nodes.size(), opt.wide, opt.res, opt.res, opt.spp, opt.idle, opt.dynamic ? "dynamic" : "static", opt.warps);

    // wave 0: camera rays, lane = pixel * spp + s
    V3 eye = { sc.cam[0], sc.cam[1], sc.cam[2] }, fwd = normalize(V3{ sc.cam[3], sc.cam[4], sc.cam[5] } - eye);
    V3 right = normalize(cross(fwd, V3{ 0, 1, 0 })), up = cross(right, fwd);
    float tanh_ = std::tan(sc.cam[6] * 0.5f * 3.14159265f / 180.f);
    std::vector<Slot> slots((size_t) opt.res * opt.res * opt.spp);
    Rng rng(7);
    for (size_t i = 0; i < slots.size(); ++i) {
        size_t pix = i / opt.spp; int px = (int) (pix % opt.res), py = (int) (pix / opt.res);
        float sx = ((px + rng.next()) / opt.res * 2.f - 1.f) * tanh_, sy = (1.f - (py + rng.next()) / opt.res * 2.f) * tanh_;
        Slot s; s.has_shadow = false; s.alive = true; s.depth = 0; s.thr = 1.f;
        s.path = { eye, normalize(fwd + right * sx + up * sy), INFINITY };
        slots[i] = s;
    }
    Counters total; double total_rays = 0;
    std::vector<HitRec> hits; std::vector<uint32_t> order;
    printf("%-5s %10s %10s %9s %9s %9s %9s\n", "wave", "slots", "rays", "thr/warp", "slot-i/ray", "thr-i/ray", "Mwarp-i");
    for (int wave = 0; wave < opt.max_depth && !slots.empty(); ++wave) {
        Counters c;
        if (opt.split_phases && wave > 0) {
            // shadow rays in their own launch, then the path rays
            std::vector<Slot> sh, pa; std::vector<uint32_t> pa_src;
            for (size_t i = 0; i < slots.size(); ++i) {
                if (slots[i].has_shadow) { Slot s = slots[i]; s.alive = false; sh.push_back(s); }
                if (slots[i].alive) { Slot s = slots[i]; s.has_shadow = false; pa.push_back(s); pa_src.push_back((uint32_t) i); }
            }
            std::vector<HitRec> h2; std::vector<uint32_t> o2;
            Counters c1 = run_wave(sc, opt, cost, sh, h2, o2, false);
            Counters c2 = run_wave(sc, opt, cost, pa, hits, order, false);
            for (int r = 0; r < R_COUNT; ++r) { c.warp_instr[r] = c1.warp_instr[r] + c2.warp_instr[r]; c.thread_instr[r] = c1.thread_instr[r] + c2.thread_instr[r]; }
            c.rays = c1.rays + c2.rays; c.node_steps = c1.node_steps + c2.node_steps; c.tri_tests = c1.tri_tests + c2.tri_tests;
            // map back
            std::vector<HitRec> hfull(slots.size(), { INFINITY, 0xffffffffu });
            for (size_t i = 0; i < pa.size(); ++i) hfull[pa_src[i]] = hits[i];
            for (auto &o : order) o = pa_src[o];
            hits = hfull;
        } else c = run_wave(sc, opt, cost, slots, hits, order, wave == 0);
        double W = c.total_warp(), T = c.total_thread();
        printf("%-5d %10zu %10.0f %9.2f %9.0f %9.0f %9.1f\n", wave, slots.size(), c.rays, T / W, W * 32.0 / c.rays, T / c.rays, W * 1e-6);
        if (opt.verbose) {
            for (int r = 0; r < R_COUNT; ++r)
                printf("      %-16s %5.1f %% of issue, %5.2f threads/warp\n", region_name[r], 100.0 * c.warp_instr[r] / W, c.warp_instr[r] ? c.thread_instr[r] / c.warp_instr[r] : 0.0);
            printf("      node steps/ray %.2f, triangle tests/ray %.2f\n", c.node_steps / c.rays, c.tri_tests / c.rays);
        }
        for (int r = 0; r < R_COUNT; ++r) { total.warp_instr[r] += c.warp_instr[r]; total.thread_instr[r] += c.thread_instr[r]; }
        total.node_steps += c.node_steps; total.tri_tests += c.tri_tests; total_rays += c.rays;
        // shade in queue (retire) order -> next wave
        std::vector<Slot> next; next.reserve(order.size());
        for (uint32_t si : order) {
            const Slot &s = slots[si]; const HitRec &h = hits[si]; const Tri &t = sc.tris[h.prim];
            int depth = s.depth + 1;
            if (depth >= opt.max_depth) continue;                 // active_next = depth + 1 < max_depth
            V3 p = s.path.o + s.path.d * h.t; V3 n = dot(t.n, s.path.d) < 0 ? t.n : t.n * -1.f;
            Slot ns; ns.depth = depth; ns.has_shadow = false; ns.alive = true; ns.thr = s.thr * t.albedo_max;
            // NEE towards the light
            if (!sc.lights.empty()) {
                const Tri &lt = sc.tris[sc.lights[(size_t) (rng.next() * sc.lights.size()) % sc.lights.size()]];
                float a = rng.next(), b = rng.next(); if (a + b > 1.f) { a = 1.f - a; b = 1.f - b; }
                V3 lp = lt.p0 + lt.e1 * a + lt.e2 * b, dv = lp - p; float dist = std::sqrt(dot(dv, dv)); V3 wo = dv * (1.f / dist);
                if (dot(n, wo) > 0.f && dot(lt.n, wo) < 0.f && !t.emitter) { ns.has_shadow = true; ns.sh = { p + n * 1e-4f, wo, dist * (1.f - 1e-3f) }; }
            }
            ns.path = { p + n * 1e-4f, cosine_dir(n, rng), INFINITY };
            if (depth >= opt.rr_depth) { float q = std::min(ns.thr, 0.95f); if (rng.next() >= q) ns.alive = false; else ns.thr /= q; }
            if (depth + 1 > opt.max_depth) ns.alive = false;
            if (ns.alive || ns.has_shadow) next.push_back(ns);
        }
        if (opt.sort_bits > 0) std::stable_sort(next.begin(), next.end(), [&](const Slot &a, const Slot &b) { return morton_key(a.path, opt.sort_bits) < morton_key(b.path, opt.sort_bits); });
        if (opt.window > 1)
            for (size_t b = 0; b < next.size(); b += (size_t) opt.window) {
                size_t e = std::min(next.size(), b + (size_t) opt.window);
                for (size_t i = e - 1; i > b; --i) { size_t j = b + (size_t) (rng.next() * (float) (i - b + 1)); if (j > i) j = i; std::swap(next[i], next[j]); }
            }
        slots.swap(next);
    }
    double W = total.total_warp(), T = total.total_thread();
    printf("frame: %.0f rays, %.2f threads/warp, %.0f slot-instr/ray, %.0f thread-instr/ray, %.1f M warp-instr  (%.2f node steps, %.2f triangle tests per ray)\n",
           total_rays, T / W, W * 32.0 / total_rays, T / total_rays, W * 1e-6, total.node_steps / total_rays, total.tri_tests / total_rays);
    for (int r = 0; r < R_COUNT; ++r)
        printf("  %-16s %5.1f %% of issue, %5.2f threads/warp\n", region_name[r], 100.0 * total.warp_instr[r] / W, total.warp_instr[r] ? total.thread_instr[r] / total.warp_instr[r] : 0.0);
    return 0;
}
