// transformations.cl (Inviwo, un-vendored) -- stand-in; arithmetic = oracle/orc_common.h, orc_grid.c.
// float16 matrices are column-major (glm); direction codec: host twin ppm/photondata.cpp:100-117.
#ifndef TRANSFORMATIONS_CL
#define TRANSFORMATIONS_CL
CLC_INLINE float3 transformPoint(float16 m, float3 p) {
    return make_float3(fmaf(m.s[8], p.z, fmaf(m.s[4], p.y, fmaf(m.s[0], p.x, m.s[12]))),
                       fmaf(m.s[9], p.z, fmaf(m.s[5], p.y, fmaf(m.s[1], p.x, m.s[13]))),
                       fmaf(m.s[10], p.z, fmaf(m.s[6], p.y, fmaf(m.s[2], p.x, m.s[14]))));
}
CLC_INLINE float3 transformVector(float16 m, float3 p) {
    return make_float3(fmaf(m.s[8], p.z, fmaf(m.s[4], p.y, m.s[0] * p.x)), fmaf(m.s[9], p.z, fmaf(m.s[5], p.y, m.s[1] * p.x)),
                       fmaf(m.s[10], p.z, fmaf(m.s[6], p.y, m.s[2] * p.x)));
}
// (theta, phi) = (acos(z), atan2(y, x))
CLC_INLINE float2 encodeDirection(float3 d) {
    return make_float2(cpm_acosf(cpm_clamp(d.z, -1.0f, 1.0f)), cpm_atan2f(d.y, d.x));
}
CLC_INLINE float3 decodeDirection(float2 a) {
    float st, ct, sp, cp;
    cpm_sincosf(a.x, &st, &ct);
    cpm_sincosf(a.y, &sp, &cp);
    return make_float3(st * cp, st * sp, ct);
}
#endif
