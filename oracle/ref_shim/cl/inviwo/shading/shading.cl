// shading/shading.cl (Inviwo, un-vendored) -- stand-in: the two phase functions the oracle restates (isotropic,
// Henyey-Greenstein); arithmetic = oracle/orc_tracer.c samplePhase.  material = AdvancedMaterialProperty::
// getCombinedMaterialParameters() = (anisotropy g, roughness, IOR, 0).
#ifndef SHADING_CL
#define SHADING_CL
#include "samplers.cl"
#include "shading/shadingmath.cl"
typedef int ShadingType;
#define HENYEY_GREENSTEIN 1
CLC_INLINE float isotropicPhaseFunction() { return CPM_INV_4PI_F; }
CLC_INLINE float3 clc_samplePhase(ShadingType phase, float4 material, float3 wi, float u1, float u2) {
    if (phase != HENYEY_GREENSTEIN) return uniformSampleSphere(make_float2(u1, u2));
    float g = material.x, ct;
    if (fabsf(g) < 1e-3f) {
        ct = fmaf(-2.0f, u1, 1.0f);
    } else {
        float q = (1.0f - g * g) / fmaf(2.0f * g, u1, 1.0f - g);
        ct = (1.0f + g * g - q * q) / (2.0f * g);
    }
    ct = cpm_clamp(ct, -1.0f, 1.0f);
    float st = sqrtf(cpm_fmax(0.0f, fmaf(-ct, ct, 1.0f)));
    float sp, cp;
    cpm_sincosf(CPM_2PI_F * u2, &sp, &cp);
    float3 v2;
    if (fabsf(wi.x) > fabsf(wi.y)) {
        float inv = 1.0f / sqrtf(fmaf(wi.x, wi.x, wi.z * wi.z));
        v2 = make_float3(-wi.z * inv, 0.0f, wi.x * inv);
    } else {
        float inv = 1.0f / sqrtf(fmaf(wi.y, wi.y, wi.z * wi.z));
        v2 = make_float3(0.0f, wi.z * inv, -wi.y * inv);
    }
    float3 v3 = cross(wi, v2);
    float a = st * cp, b = st * sp;
    return make_float3(fmaf(a, v2.x, fmaf(b, v3.x, ct * wi.x)), fmaf(a, v2.y, fmaf(b, v3.y, ct * wi.y)),
                       fmaf(a, v2.z, fmaf(b, v3.z, ct * wi.z)));
}
CLC_INLINE void sampleShadingFunction(image3d_t volumeTex, __constant VolumeParameters* volumeParams, float volumeSample,
                                      float4 material, float3 sample, float3* direction, float2 rnd, ShadingType shadingType) {
    *direction = clc_samplePhase(shadingType, material, *direction, rnd.x, rnd.y);
}
CLC_INLINE void sampleShadingFunctionPdf(image3d_t volumeTex, __constant VolumeParameters* volumeParams, float volumeSample,
                                         float4 material, float3 sample, float3* direction, float* pdf, float2 rnd,
                                         ShadingType shadingType) {
    *direction = clc_samplePhase(shadingType, material, *direction, rnd.x, rnd.y);
    *pdf = CPM_INV_4PI_F;
}
#endif
