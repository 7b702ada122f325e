// samplers.cl (Inviwo, un-vendored) -- stand-in.  Volume / transfer-function sampling as OpenCL 1.2 section 8.2 states
// it for CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_CLAMP_TO_EDGE | CLK_FILTER_LINEAR, arithmetic = oracle/orc_tracer.c.
#ifndef SAMPLERS_CL
#define SAMPLERS_CL

typedef struct VolumeParameters {
    float16 textureToIndex;
    float16 indexToTexture;
    float16 textureToWorld;
    float formatScaling;   // (v + formatOffset) * formatScaling: "scaling for 12-bit data"
    float formatOffset;
} VolumeParameters;

__constant sampler_t smpNormClampEdgeLinear = 1;
__constant sampler_t smpUNormNoClampNearest = 2;

CLC_INLINE float clc_texel(image3d_t img, int i, int j, int k) {
    size_t idx = ((size_t)k * img->dims[1] + (size_t)j) * img->dims[0] + (size_t)i;
    switch (img->format) {
        case 0: return (float)((const uchar*)img->data)[idx] / 255.0f;     // CL_UNORM_INT8
        case 1: return (float)((const ushort*)img->data)[idx] / 65535.0f;  // CL_UNORM_INT16
        default: return ((const float*)img->data)[idx];
    }
}
CLC_INLINE int clc_clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
CLC_INLINE float clc_lerp(float p, float q, float a) { return fmaf(a, q - p, p); }

// un-normalised value at normalised position pos.xyz
CLC_INLINE float getVoxel(image3d_t V, float4 pos) {
    float fx = (float)V->dims[0], fy = (float)V->dims[1], fz = (float)V->dims[2];
    float u = fmaf(pos.x, fx, -0.5f), v = fmaf(pos.y, fy, -0.5f), w = fmaf(pos.z, fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    float a = u - fu, b = v - fv, c = w - fw;
    int i0 = (int)cpm_clamp(fu, -1.0f, fx - 1.0f);
    int j0 = (int)cpm_clamp(fv, -1.0f, fy - 1.0f);
    int k0 = (int)cpm_clamp(fw, -1.0f, fz - 1.0f);
    int i1 = clc_clampi(i0 + 1, 0, V->dims[0] - 1), j1 = clc_clampi(j0 + 1, 0, V->dims[1] - 1),
        k1 = clc_clampi(k0 + 1, 0, V->dims[2] - 1);
    i0 = clc_clampi(i0, 0, V->dims[0] - 1);
    j0 = clc_clampi(j0, 0, V->dims[1] - 1);
    k0 = clc_clampi(k0, 0, V->dims[2] - 1);
    float x00 = clc_lerp(clc_texel(V, i0, j0, k0), clc_texel(V, i1, j0, k0), a);
    float x10 = clc_lerp(clc_texel(V, i0, j1, k0), clc_texel(V, i1, j1, k0), a);
    float x01 = clc_lerp(clc_texel(V, i0, j0, k1), clc_texel(V, i1, j0, k1), a);
    float x11 = clc_lerp(clc_texel(V, i0, j1, k1), clc_texel(V, i1, j1, k1), a);
    float y0 = clc_lerp(x00, x10, b), y1 = clc_lerp(x01, x11, b);
    return clc_lerp(y0, y1, c);
}
CLC_INLINE float4 getNormalizedVoxel(image3d_t V, __constant VolumeParameters* p, float4 pos) {
    float v = (getVoxel(V, pos) + p->formatOffset) * p->formatScaling;
    return make_float4(v, v, v, v);
}
// integer voxel coordinate, no filtering
CLC_INLINE float4 getNormalizedVoxelUnorm(image3d_t V, __constant VolumeParameters* p, int4 c) {
    float v = (clc_texel(V, c.x, c.y, c.z) + p->formatOffset) * p->formatScaling;
    return make_float4(v, v, v, v);
}

// RGBA float image (transfer function: height 1; entry / exit images: nearest with integer coordinates)
CLC_INLINE float4 read_imagef(image2d_t img, sampler_t smp, float2 coord) {
    const float* t = (const float*)img->data;
    int width = img->dims[0];
    float fw = (float)width;
    float u = fmaf(coord.x, fw, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fw - 1.0f);
    int i1 = clc_clampi(i0 + 1, 0, width - 1);
    i0 = clc_clampi(i0, 0, width - 1);
    return make_float4(clc_lerp(t[4 * i0], t[4 * i1], a), clc_lerp(t[4 * i0 + 1], t[4 * i1 + 1], a),
                       clc_lerp(t[4 * i0 + 2], t[4 * i1 + 2], a), clc_lerp(t[4 * i0 + 3], t[4 * i1 + 3], a));
}
CLC_INLINE float4 read_imagef(image2d_t img, sampler_t smp, int2 c) {
    const float* t = (const float*)img->data + 4 * ((size_t)c.y * img->dims[0] + c.x);
    return make_float4(t[0], t[1], t[2], t[3]);
}
#endif
