// common.cuh -- context object, error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "cpm_b200.h"
#include "cpm_detmath.h"

struct cpm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    uint64_t launches = 0;
    // small persistent device scratch (counters, histograms), grown on demand
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned host word for synchronous scalar read-backs
    uint32_t* pinned = nullptr;
    void* comm = nullptr;  // ncclComm_t when multi-GPU is initialised
    cudaStream_t xfer_stream = nullptr;  // transfer stream of cpm_mem_prefetch_h2d (lazy)
    cudaEvent_t xfer_fence = nullptr;
    unsigned* trace_cursor = nullptr;    // work cursor of walk_kernel (tracer.cu)
    void* walk_buf = nullptr;            // walk entries + power of the wavefront tracer
    size_t walk_bytes = 0;
    cudaStream_t d2h_stream = nullptr;   // read-back stream of cpm_mem_readback_d2h (lazy)
    cudaEvent_t d2h_fence = nullptr;
    cudaEvent_t select_done = nullptr;   // cpm_select_below_begin / _end
    bool select_pending = false;
};

struct cpm_event {
    cudaEvent_t ev;
};

struct cpm_volume {
    int dims[3];
    int format;
    float scale, offset;
    int layout;
    const void* linear;      // caller-owned (LINEAR) or null
    cudaArray_t array;       // TEXTURE layout: 2-D layered array, one layer per z slice
    cudaTextureObject_t tex;
};

// the opacity-bound grid as a point-sampled 3-D texture (bound.cu: cpm_bound_tex_*)
struct cpm_bound_tex {
    int dims[3];
    cudaArray_t array;
    cudaTextureObject_t tex;
};

int cpm_fail(cpm_ctx* ctx, int code, const char* fmt, ...);
int cpm_scratch(cpm_ctx* ctx, size_t bytes, void** out);

#define CPM_CUDA(ctx, call)                                                                 \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return cpm_fail((ctx), e_ == cudaErrorMemoryAllocation ? CPM_E_NOMEM : CPM_E_CUDA, \
                            "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define CPM_REQUIRE(ctx, cond, msg)                                              \
    do {                                                                         \
        if (!(cond)) return cpm_fail((ctx), CPM_E_INVALID, "%s: %s", __func__, (msg)); \
    } while (0)

// every kernel launch goes through this so that ctx->launches is an honest count
#define CPM_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                         \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
        (ctx)->launches++;                                                       \
        CPM_CUDA((ctx), cudaGetLastError());                                     \
    } while (0)

static inline unsigned cpm_div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- MWC64X core (rng/cl/random.cl:44-95), device side --------------------------------
#define CPM_MWC64X_A 4294883355u
#define CPM_MWC64X_M 18446383549859758079ull

struct cpm_rng {
    uint32_t x, c;
};
__device__ __forceinline__ uint32_t cpm_rng_next(cpm_rng& s) {
    uint32_t res = s.x ^ s.c;
    // A*X + C: low word = Xn, high = Cn -- one mad.wide (the C++ form leaves the compiler adding the zero high word of C
    // in a separate instruction)
    asm("{\n\t.reg .u64 t, cc;\n\tcvt.u64.u32 cc, %1;\n\tmad.wide.u32 t, %0, %2, cc;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "+r"(s.x), "+r"(s.c)
        : "r"(CPM_MWC64X_A));
    return res;
}
// random_01 = (float)u / 4294967295.0f ; the divisor rounds to 2^32 in fp32 so the
// result is (float)u * 2^-32 exactly (can be 1.0f and 0.0f).
__device__ __forceinline__ float cpm_u01(uint32_t k) { return __uint2float_rn(k) * 2.3283064365386963e-10f; }
__device__ __forceinline__ float cpm_rng_01(cpm_rng& s) { return cpm_u01(cpm_rng_next(s)); }
