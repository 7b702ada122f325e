// image3d_write.cl (Inviwo, un-vendored) -- stand-in: buffer-backed 3-D "image" writes, out[x + y dx + z dx dy];
// ushort conversion assumed convert_ushort_sat_rte(v * 65535) (parity unpinned for the rounding mode)
#ifndef IMAGE3D_WRITE_CL
#define IMAGE3D_WRITE_CL
typedef __global ushort2* image_3d_write_vec2_uint16_t;
typedef __global float* image_3d_write_float32_t;
typedef __global float4* image_3d_write_vec4_float32_t;
CLC_INLINE void writeImageVec2UInt16f(image_3d_write_vec2_uint16_t out, int4 c, int4 dim, float2 v) {
    out[c.x + c.y * dim.x + c.z * dim.x * dim.y] =
        make_ushort2(convert_ushort_sat_rte(v.x * 65535.0f), convert_ushort_sat_rte(v.y * 65535.0f));
}
CLC_INLINE void writeImageFloat32f(image_3d_write_float32_t out, int4 c, int4 dim, float v) {
    out[c.x + c.y * dim.x + c.z * dim.x * dim.y] = v;
}
CLC_INLINE void writeImageVec4Float32f(image_3d_write_vec4_float32_t out, int4 c, int4 dim, float4 v) {
    out[c.x + c.y * dim.x + c.z * dim.x * dim.y] = v;
}
#endif
