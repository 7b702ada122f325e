// ref_common.h -- shared by the drivers that run the reference's kernels on the host (TEST INFRASTRUCTURE).
#pragma once
#include "clc.h"
#include "cpm_oracle.h"   // orc_volume / orc_trace_params: the drivers take the oracle's argument structs

#define REF_API extern "C" __attribute__((visibility("default")))

static inline clc_image ref_image3d(const orc_volume* v) {
    clc_image im;
    im.data = v->data;
    im.dims[0] = v->dims[0]; im.dims[1] = v->dims[1]; im.dims[2] = v->dims[2];
    im.format = v->format;
    return im;
}
static inline clc_image2d ref_image_tf(const float* rgba, int width) {
    clc_image2d im;
    im.data = rgba;
    im.dims[0] = width; im.dims[1] = 1; im.dims[2] = 1;
    im.format = 3;
    return im;
}
static inline float16 ref_mat(const float m[16]) {
    float16 r;
    for (int i = 0; i < 16; ++i) r.s[i] = m[i];
    return r;
}
// run `body` once per work-item of a 1-D range
#define REF_FOR_EACH_WORK_ITEM(n, body)                         \
    do {                                                        \
        clc::wi().gsize[0] = (size_t)(n);                       \
        clc::wi().gsize[1] = clc::wi().gsize[2] = 1;            \
        clc::wi().gid[1] = clc::wi().gid[2] = 0;                \
        for (size_t gid_ = 0; gid_ < (size_t)(n); ++gid_) {     \
            clc::wi().gid[0] = gid_;                            \
            body;                                               \
        }                                                       \
    } while (0)
