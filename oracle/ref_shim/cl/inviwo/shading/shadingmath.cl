// shading/shadingmath.cl (Inviwo, un-vendored) -- stand-in; arithmetic = oracle/orc_common.h uniformSampleSphere
#ifndef SHADINGMATH_CL
#define SHADINGMATH_CL
CLC_INLINE float3 uniformSampleSphere(float2 uv) {
    float z = fmaf(-2.0f, uv.x, 1.0f);
    float r = sqrtf(cpm_fmax(0.0f, fmaf(-z, z, 1.0f)));
    float s, c;
    cpm_sincosf(CPM_2PI_F * uv.y, &s, &c);
    return make_float3(r * c, r * s, z);
}
CLC_INLINE float uniformSpherePdf() { return CPM_INV_4PI_F; }
// cone around +z with half-angle acos(cosThetaMax) (cone lights; not on the measured path)
CLC_INLINE float3 uniformSampleCone(float2 uv, float cosThetaMax) {
    float ct = fmaf(uv.x, cosThetaMax - 1.0f, 1.0f);
    float st = sqrtf(cpm_fmax(0.0f, fmaf(-ct, ct, 1.0f)));
    float s, c;
    cpm_sincosf(CPM_2PI_F * uv.y, &s, &c);
    return make_float3(st * c, st * s, ct);
}
CLC_INLINE float uniformConePdf(float cosThetaMax) { return 1.0f / (CPM_2PI_F * (1.0f - cosThetaMax)); }
#endif
