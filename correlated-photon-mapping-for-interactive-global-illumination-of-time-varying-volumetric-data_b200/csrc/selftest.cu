// selftest.cu -- device-side evaluation of the deterministic math layer, so that tests can
// prove bit-equality between the GPU and the host compilation of include/cpm_detmath.h.
#include "sampling.cuh"

namespace {
__device__ const float g_nlog[64] = {CPM_NLOG_TABLE};
__global__ void math_kernel(int fn, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                            size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = x[i], b = y ? y[i] : 0.0f, r, s, c;
    switch (fn) {
        case 0: r = cpm_logf(a); break;
        case 1: cpm_sincosf(a, &s, &c); r = s; break;
        case 2: cpm_sincosf(a, &s, &c); r = c; break;
        case 3: r = cpm_acosf(a); break;
        case 4: r = cpm_atan2f(a, b); break;
        case 5: r = unorm8(a); break;
        case 6: r = unorm16(a); break;
        case 7: r = cpm_powf(a, b); break;
        case 8: r = cpm_cbrtf(a); break;
        case 9: r = cpm_expf_sym(a); break;
        case 10: r = cpm_native_logf_tab(a, g_nlog); break;
        default: r = 0.0f;
    }
    out[i] = r;
}
}  // namespace

extern "C" int cpm_selftest_math(cpm_ctx* ctx, int fn, const float* x, const float* y, float* out, size_t n) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, fn >= 0 && fn <= 10, "unknown function id");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, x && out, "null buffer");
    CPM_LAUNCH(ctx, math_kernel, cpm_div_up(n, 256), 256, 0, fn, x, y, out, n);
    return CPM_OK;
}
