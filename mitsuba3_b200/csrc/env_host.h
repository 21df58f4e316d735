// env_host.h -- host-side construction of the environment emitter's device data:
// the halo-extended lat-long texture (envmap.cpp:155-192), the luminance x sin(theta)
// Hierarchical2D sample warp (envmap.cpp:474-529, core/distr_2d.h:403-540) and the
// bounding sphere of the scene (envmap.cpp:260-274, constant.cpp:76-93).
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/b200pt.h"

namespace pt {

struct EnvHost {
    std::vector<float> tex;      // H x (W + 2) x 4 (rgb + pad), with the periodic halo columns
    std::vector<float> warp;     // Hierarchical2D::m_data
    std::vector<uint32_t> lvl_width, lvl_size, lvl_offset;
    float patch_size[2], inv_patch_size[2];
    uint32_t max_patch_index[2];
};

// Builds texture + warp of an `envmap` emitter from its lat-long map `data` (height x width x 3,
// real columns). Returns false on an invalid size.
bool build_envmap(const float *data, uint32_t width, uint32_t height, bool mis_compensation, EnvHost &out);

// Bounding sphere of all vertices (8 floats per vertex, position first), inflated as the
// environment emitters do; center[3], radius.
void scene_bounding_sphere(const float *verts8, size_t n_verts, float center[3], float &radius);

} // namespace pt
