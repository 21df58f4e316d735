// bvh.h -- host-side BVH2 builder (binned SAH) for the triangle soup of a scene.
//
// Replaces, for triangle meshes, the acceleration-structure build the reference
// delegates to Embree / its kd-tree / OptiX (scene.cpp:93, scene_embree.inl,
// kdtree.h, scene_optix.inl:446). The layout is designed for the GPU traversal
// kernels in kernels.cu:
//
//   node (64 B, 4 x float4), Aila/Laine style -- a node stores the boxes of
//   BOTH children so that one 64 B fetch decides where to descend:
//     n0 = (L.lo.x, L.lo.y, L.lo.z, L.hi.x)
//     n1 = (L.hi.y, L.hi.z, R.lo.x, R.lo.y)
//     n2 = (R.lo.z, R.hi.x, R.hi.y, R.hi.z)
//     n3 = (left, right, 0, 0) as int bits
//   child >= 0            : inner node index
//   child <  0            : leaf, ~child = (first << 3) | (count - 1), count <= 8
//   child == BVH_EMPTY    : no child (box is inverted, never entered)
//
//   Nodes are emitted in breadth-first order, so the first K nodes are the top
//   of the tree: the traversal kernels stage them in shared memory with one
//   bulk (TMA) copy.
//
//   triangle (48 B, 3 x float4) in leaf order:
//     t0 = (p0.xyz, as_float(global prim id)), t1 = (e1.xyz, 0), t2 = (e2.xyz, 0)
//   with e1 = p1 - p0, e2 = p2 - p0 evaluated in fp32 exactly like
//   Mesh::moeller_trumbore does at run time (mesh.h:1139).
//
// Boxes are inflated by a few ulps: the ray/triangle test is the fp32
// Moeller-Trumbore of the reference, whose accepted hits can lie marginally
// outside the exact triangle; the boxes must never cull such a hit.
#pragma once
#include <cstdint>
#include <vector>

namespace pt {

constexpr int32_t BVH_EMPTY = 0x7fffffff;
constexpr uint32_t BVH_MAX_LEAF = 2;   // measured: 2 beats 1 and 4 on B200 (profiles/r01_tuning.md)

struct BvhNode { float f[12]; int32_t left, right, pad0, pad1; };
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> order;   // leaf order -> input triangle index
    uint32_t depth = 0;
    // breadth-first order is level order: level l = nodes [level_start[l], level_start[l + 1]). The device refit
    // (kernels.cu: k_refit_level) recomputes the boxes level by level from the deepest one up.
    std::vector<uint32_t> level_start;
};

// tri: n x 9 floats (p0, p1, p2)
Bvh build_bvh(const float *tri, uint32_t n);

} // namespace pt
