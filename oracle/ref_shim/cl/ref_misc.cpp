// ref_misc.cpp -- ppm/cl/threshold.cl, ppm/cl/indextobuffer.cl, ppm/cl/hashlightsample.cl and ugc/cl/buffermixer.cl of the
// reference, compiled where they lie and run one work-item at a time.  TEST INFRASTRUCTURE (oracle/_ref/libcl_ref.so).
#include "ref_common.h"
namespace {
#include "threshold.cl"
#include "indextobuffer.cl"
#include "hashlightsample.cl"
namespace mixf {
#define MIX_T float
#include "buffermixer.cl"
#undef MIX_T
#undef BUFFER_MIX_CL
}  // namespace mixf
namespace mixu8 {
// as BufferMixerCL builds it for integer buffers (ugc/buffermixercl.cpp): mix in float, convert back (round toward zero)
inline uchar convert_uchar(float v) { return (uchar)clc_f2i(v); }
#define MIX_T uchar
#define CONVERT_T_TO_FLOAT convert_float
#define CONVERT_FLOAT_TO_T convert_uchar
#include "buffermixer.cl"
#undef MIX_T
}  // namespace mixu8
}  // namespace

REF_API void ref_threshold(const uint* data, uint threshold, int n, uint* out) {
    REF_FOR_EACH_WORK_ITEM(n, thresholdKernel(data, threshold, n, out));
}
REF_API void ref_index_to_buffer(uint* indices, int n) { REF_FOR_EACH_WORK_ITEM(n, indexToBufferKernel(indices, n)); }
REF_API void ref_hash_light_samples(const float* ls, const float* isect, int n_src, const uint* ids, int n_ids,
                                    const float cell_size[3], const int n_blocks[3], uint* which_bucket, int out_offset) {
    REF_FOR_EACH_WORK_ITEM(n_ids, hashLightSampleKernel((const float8*)ls, (const float2*)isect, n_src, ids, n_ids,
                                                        make_float3(cell_size[0], cell_size[1], cell_size[2]),
                                                        make_int3(n_blocks[0], n_blocks[1], n_blocks[2]), which_bucket,
                                                        out_offset));
}
REF_API void ref_mix_f32(const float* x, const float* y, float a, uint len, float* out) {
    REF_FOR_EACH_WORK_ITEM(len, mixf::mixKernel(x, y, a, len, out));
}
REF_API void ref_mix_u8(const uchar* x, const uchar* y, float a, uint len, uchar* out) {
    REF_FOR_EACH_WORK_ITEM(len, mixu8::mixKernel(x, y, a, len, out));
}
