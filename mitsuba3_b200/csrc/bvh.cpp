// bvh.cpp -- binned-SAH BVH2 builder, see bvh.h for the layout.
#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <queue>
#include <cstdlib>

namespace pt {
namespace {

struct Box {
    float lo[3], hi[3];
    Box() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        float d[3] = { hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] };
        if (d[0] < 0) return 0.f;
        return 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    }
};

struct TmpNode { Box box; int32_t left = -1, right = -1; uint32_t first = 0, count = 0; };

struct Builder {
    const float *tri; uint32_t n;
    uint32_t max_leaf = BVH_MAX_LEAF;
    std::vector<Box> tbox; std::vector<float> cent;   // per triangle
    std::vector<uint32_t> order;
    std::vector<TmpNode> tmp;
    uint32_t max_depth = 0;

    int build(uint32_t first, uint32_t count, uint32_t depth) {
        max_depth = std::max(max_depth, depth);
        int id = (int) tmp.size();
        tmp.emplace_back();
        Box box, cbox;
        for (uint32_t i = first; i < first + count; ++i) { box.grow(tbox[order[i]]); cbox.grow(&cent[3 * order[i]]); }
        tmp[id].box = box; tmp[id].first = first; tmp[id].count = count;
        if (count <= max_leaf || depth > 60)
            if (count <= 8) return id;
        // binned SAH over the 3 axes
        constexpr int NB = 16;
        float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_split = -1;
        for (int a = 0; a < 3; ++a) {
            float lo = cbox.lo[a], ext = cbox.hi[a] - lo;
            if (!(ext > 0.f)) continue;
            Box bb[NB]; uint32_t bc[NB] = { 0 };
            float scale = NB / ext;
            for (uint32_t i = first; i < first + count; ++i) {
                int b = std::min(NB - 1, std::max(0, (int) ((cent[3 * order[i] + a] - lo) * scale)));
                bb[b].grow(tbox[order[i]]); bc[b]++;
            }
            float la[NB], ra[NB]; uint32_t lc[NB], rc[NB];
            Box acc; uint32_t c = 0;
            for (int b = 0; b < NB; ++b) { acc.grow(bb[b]); c += bc[b]; la[b] = acc.area(); lc[b] = c; }
            acc = Box(); c = 0;
            for (int b = NB - 1; b >= 0; --b) { acc.grow(bb[b]); c += bc[b]; ra[b] = acc.area(); rc[b] = c; }
            for (int b = 0; b < NB - 1; ++b) {
                if (lc[b] == 0 || rc[b + 1] == 0) continue;
                float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
            }
        }
        uint32_t mid;
        float leaf_cost = box.area() * count;
        if (best_axis < 0 || (count <= max_leaf && best_cost >= leaf_cost)) {
            if (count <= max_leaf) return id;
            // degenerate centroids: median split in index order
            mid = first + count / 2;
        } else {
            float lo = cbox.lo[best_axis], scale = NB / (cbox.hi[best_axis] - lo);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                int b = std::min(NB - 1, std::max(0, (int) ((cent[3 * t + best_axis] - lo) * scale)));
                return b <= best_split;
            });
            mid = (uint32_t) (it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        int l = build(first, mid - first, depth + 1);
        int r = build(mid, first + count - mid, depth + 1);
        tmp[id].left = l; tmp[id].right = r; tmp[id].count = 0;
        return id;
    }
};

inline float down(float x) { return std::nextafter(x, -std::numeric_limits<float>::infinity()); }
inline float up(float x) { return std::nextafter(x, std::numeric_limits<float>::infinity()); }

void inflate(Box &b) {
    for (int a = 0; a < 3; ++a) {
        float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
        float pad = m * 4e-6f + 1e-7f;
        b.lo[a] = down(b.lo[a] - pad); b.hi[a] = up(b.hi[a] + pad);
    }
}

} // namespace

Bvh build_bvh(const float *tri, uint32_t n) {
    Bvh out;
    if (n == 0) {
        BvhNode root; std::memset(&root, 0, sizeof(root));
        for (int k = 0; k < 2; ++k) for (int a = 0; a < 3; ++a) { root.f[6 * k + a] = std::numeric_limits<float>::infinity(); root.f[6 * k + 3 + a] = -std::numeric_limits<float>::infinity(); }
        root.left = root.right = BVH_EMPTY;
        out.nodes.push_back(root);
        out.level_start = { 0u, 1u };
        return out;
    }
    Builder b; b.tri = tri; b.n = n;
    if (const char *e = getenv("B200PT_BVH_LEAF")) b.max_leaf = (uint32_t) std::min(8, std::max(1, atoi(e)));
    b.tbox.resize(n); b.cent.resize(3 * (size_t) n); b.order.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        b.order[i] = i;
        for (int k = 0; k < 3; ++k) b.tbox[i].grow(tri + 9 * (size_t) i + 3 * k);
        for (int a = 0; a < 3; ++a) b.cent[3 * (size_t) i + a] = 0.5f * (b.tbox[i].lo[a] + b.tbox[i].hi[a]);
    }
    b.tmp.reserve(2 * (size_t) n);
    int root = b.build(0, n, 0);
    out.depth = b.max_depth;
    out.order = b.order;

    auto encode_child = [&](int id, std::vector<int> &bfs_index) -> int32_t {
        const TmpNode &t = b.tmp[id];
        if (t.left < 0) return ~(int32_t) ((t.first << 3) | (t.count - 1));
        return bfs_index[id];
    };
    // breadth-first numbering of the inner nodes
    std::vector<int> bfs_index(b.tmp.size(), -1);
    std::vector<int> bfs;
    if (b.tmp[root].left < 0) {
        // single leaf: synthesise a root whose left child is the leaf
        BvhNode rn; std::memset(&rn, 0, sizeof(rn));
        Box lb = b.tmp[root].box; inflate(lb);
        rn.f[0] = lb.lo[0]; rn.f[1] = lb.lo[1]; rn.f[2] = lb.lo[2]; rn.f[3] = lb.hi[0]; rn.f[4] = lb.hi[1]; rn.f[5] = lb.hi[2];
        for (int a = 0; a < 3; ++a) { rn.f[6 + a] = std::numeric_limits<float>::infinity(); rn.f[9 + a] = -std::numeric_limits<float>::infinity(); }
        rn.left = ~(int32_t) ((b.tmp[root].first << 3) | (b.tmp[root].count - 1)); rn.right = BVH_EMPTY;
        out.nodes.push_back(rn);
        out.level_start = { 0u, 1u };
        return out;
    }
    std::queue<int> q; q.push(root);
    size_t level_left = 1, next_level = 0;
    out.level_start.push_back(0u);
    while (!q.empty()) {
        int id = q.front(); q.pop();
        bfs_index[id] = (int) bfs.size(); bfs.push_back(id);
        const TmpNode &t = b.tmp[id];
        if (b.tmp[t.left].left >= 0) { q.push(t.left); ++next_level; }
        if (b.tmp[t.right].left >= 0) { q.push(t.right); ++next_level; }
        if (--level_left == 0) { out.level_start.push_back((uint32_t) bfs.size()); level_left = next_level; next_level = 0; }
    }
    out.nodes.resize(bfs.size());
    for (size_t i = 0; i < bfs.size(); ++i) {
        const TmpNode &t = b.tmp[bfs[i]];
        BvhNode &nd = out.nodes[i]; std::memset(&nd, 0, sizeof(nd));
        Box lb = b.tmp[t.left].box, rb = b.tmp[t.right].box; inflate(lb); inflate(rb);
        nd.f[0] = lb.lo[0]; nd.f[1] = lb.lo[1]; nd.f[2] = lb.lo[2]; nd.f[3] = lb.hi[0]; nd.f[4] = lb.hi[1]; nd.f[5] = lb.hi[2];
        nd.f[6] = rb.lo[0]; nd.f[7] = rb.lo[1]; nd.f[8] = rb.lo[2]; nd.f[9] = rb.hi[0]; nd.f[10] = rb.hi[1]; nd.f[11] = rb.hi[2];
        nd.left = encode_child(t.left, bfs_index); nd.right = encode_child(t.right, bfs_index);
    }
    return out;
}

} // namespace pt
