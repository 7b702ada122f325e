// intersection/rayboxintersection.cl (Inviwo, un-vendored) -- stand-in: slab test that narrows [t0, t1];
// arithmetic = oracle/orc_common.h rayBoxIntersection
#ifndef RAYBOXINTERSECTION_CL
#define RAYBOXINTERSECTION_CL
#include "datastructures/bbox.cl"
CLC_INLINE bool rayBoxIntersection(BBox box, float3 o, float3 d, float* __restrict t0, float* __restrict t1) {
    float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    float ax = (box.pMin.x - o.x) * ix, bx = (box.pMax.x - o.x) * ix;
    float ay = (box.pMin.y - o.y) * iy, by = (box.pMax.y - o.y) * iy;
    float az = (box.pMin.z - o.z) * iz, bz = (box.pMax.z - o.z) * iz;
    float n = cpm_fmax(cpm_fmax(cpm_fmin(ax, bx), cpm_fmin(ay, by)), cpm_fmin(az, bz));
    float f = cpm_fmin(cpm_fmin(cpm_fmax(ax, bx), cpm_fmax(ay, by)), cpm_fmax(az, bz));
    *t0 = cpm_fmax(*t0, n);
    *t1 = cpm_fmin(*t1, f);
    return *t0 < *t1;
}
#endif
