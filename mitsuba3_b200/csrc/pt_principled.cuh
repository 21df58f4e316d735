// pt_principled.cuh -- Disney principled BSDF (principled.cpp:332-837,
// principledhelpers.h, microfacet.h:185-421). Filled in by a later milestone.
#pragma once
namespace pt {
PT_DEV void principled_eval_pdf(const DevScene &, const DevBsdf &, float2, float3, float3, float3 &value, float &pdf) {
    value = V(0.f, 0.f, 0.f); pdf = 0.f;
}
PT_DEV void principled_sample(const DevScene &, const DevBsdf &, float2, float3, float, float, float, BsdfSample &bs, float3 &weight) {
    bs.wo = V(0.f, 0.f, 0.f); bs.pdf = 0.f; bs.eta = 0.f; bs.sampled_type = 0; bs.sampled_component = 0;
    weight = V(0.f, 0.f, 0.f);
}
} // namespace pt
