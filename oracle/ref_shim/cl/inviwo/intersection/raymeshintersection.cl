// intersection/raymeshintersection.cl (Inviwo, un-vendored) -- stand-in: Moeller-Trumbore over the index triples,
// nearest / farthest hit narrow [t0, t1]; arithmetic = oracle/orc_emission.c orc_light_mesh_intersect
#ifndef RAYMESHINTERSECTION_CL
#define RAYMESHINTERSECTION_CL
CLC_INLINE bool rayMeshIntersection(__global float const* __restrict vertices, __global int const* __restrict indices,
                                    int nIndices, float3 o, float3 d, float* t0, float* t1) {
    int nTri = nIndices / 3;
    float tn = FLT_MAX, tf = -FLT_MAX;
    bool hit = false;
    for (int t = 0; t < nTri; ++t) {
        __global const float* p0 = vertices + 3 * (size_t)indices[3 * t];
        __global const float* p1 = vertices + 3 * (size_t)indices[3 * t + 1];
        __global const float* p2 = vertices + 3 * (size_t)indices[3 * t + 2];
        float3 v0 = make_float3(p0[0], p0[1], p0[2]);
        float3 e1 = make_float3(p1[0] - v0.x, p1[1] - v0.y, p1[2] - v0.z);
        float3 e2 = make_float3(p2[0] - v0.x, p2[1] - v0.y, p2[2] - v0.z);
        float3 p = cross(d, e2);
        float det = dot(e1, p);
        if (fabsf(det) < 1e-12f) continue;
        float inv = 1.0f / det;
        float3 tv = o - v0;
        float u = dot(tv, p) * inv;
        if (u < 0.0f || u > 1.0f) continue;
        float3 q = cross(tv, e1);
        float v = dot(d, q) * inv;
        if (v < 0.0f || u + v > 1.0f) continue;
        float tt = dot(e2, q) * inv;
        tn = cpm_fmin(tn, tt);
        tf = cpm_fmax(tf, tt);
        hit = true;
    }
    if (hit) {
        *t0 = cpm_fmax(*t0, tn);
        *t1 = cpm_fmin(*t1, tf);
        hit = *t0 < *t1;
    }
    return hit;
}
#endif
