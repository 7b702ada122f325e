// importance.cu -- view / light importance image and importance-driven 2-D sample generation
// (north-star subsystem 2: "importance-map-driven photon emission").
//
// cpm_view_importance replaces uniformGridImportanceKernel + uniformGridImportance + stepToNextCell2
// (isc/cl/minmaxuniformgrid3dimportance.cl:336-378, :86-133, :42-68) with the traversal set-up of
// setupUniformGridTraversal (ugc/cl/uniformgrid/uniformgrid.cl:38-69); host side
// MinMaxUniformGrid3DImportanceCL::computeImportance (isc/minmaxuniformgrid3dimportancecl.cpp:99-132).
// The reference compiles this class but no processor instantiates it.
//
// cpm_sample_importance2d is new: the reference only has the SampleGenerator2DCL interface
// (lcl/samplegenerator2dcl.h:53-88) with a uniform implementation.  It warps the uniform generator's
// stratified samples through the inverse CDF of a piecewise-constant 2-D density (importance + floor) and
// writes the density into the pdf slot that directionalLightSamplerKernel divides by
// (lcl/cl/directionallightsampler.cl:55-62: power = radiance / (pdf / area)).  Parity: oracle restatement.
//
// Both are tiny next to the tracer (one thread per pixel / per image row / per sample).
#include <string.h>

#include "sampling.cuh"

namespace {

struct M4 {
    float m[16];
};
__device__ __forceinline__ float3_ xf(const M4& M, float3_ p) {
    return {fmaf(M.m[8], p.z, fmaf(M.m[4], p.y, fmaf(M.m[0], p.x, M.m[12]))),
            fmaf(M.m[9], p.z, fmaf(M.m[5], p.y, fmaf(M.m[1], p.x, M.m[13]))),
            fmaf(M.m[10], p.z, fmaf(M.m[6], p.y, fmaf(M.m[2], p.x, M.m[14])))};
}

struct ViewArgs {
    const ushort2* grid;
    int dims[3];
    float cell[3];
    M4 tex2idx, idx2tex;
    const float4* entry;
    const float4* exit;
    int width, height;
    float tf_min, tf_max;
    float* out;
};

__global__ void __launch_bounds__(128) view_importance_kernel(const ViewArgs A) {
    int gx = blockIdx.x * 16 + (threadIdx.x & 15), gy = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (gx >= A.width || gy >= A.height) return;
    int id = gx + gy * A.width;
    float4 e = A.entry[id], x = A.exit[id];
    float3_ x1 = xf(A.tex2idx, {e.x, e.y, e.z}), x2 = xf(A.tex2idx, {x.x, x.y, x.z});
    x1 = {x1.x + 0.5f, x1.y + 0.5f, x1.z + 0.5f};
    x2 = {x2.x + 0.5f, x2.y + 0.5f, x2.z + 0.5f};
    if (x1.x == x2.x && x1.y == x2.y && x1.z == x2.z) {
        A.out[id] = 0.0f;
        return;
    }
    float a1[3] = {x1.x, x1.y, x1.z}, a2[3] = {x2.x, x2.y, x2.z};
    float dt[3], deltatx[3];
    int cell[3], cell_end[3], di[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float mx = (float)(A.dims[k] - 1);
        float cf = cpm_clamp(floorf(a1[k] / A.cell[k]), 0.0f, mx);
        cell[k] = (int)cf;
        cell_end[k] = (int)cpm_clamp(truncf(a2[k] / A.cell[k]), 0.0f, mx);
        di[k] = (a1[k] < a2[k]) ? 1 : ((a1[k] > a2[k]) ? -1 : 0);
        float inv_abs = 1.0f / fabsf(a2[k] - a1[k]);
        float minx = A.cell[k] * cf;
        float maxx = minx + A.cell[k];
        dt[k] = ((a1[k] > a2[k]) ? (a1[k] - minx) : (maxx - a1[k])) * inv_abs;
        deltatx[k] = A.cell[k] * inv_abs;
    }
    float3_ t1 = xf(A.idx2tex, x1), t2 = xf(A.idx2tex, x2);
    float lx = t2.x - t1.x, ly = t2.y - t1.y, lz = t2.z - t1.z;
    float len = sqrtf(fmaf(lz, lz, fmaf(ly, ly, lx * lx)));
    const int sy = A.dims[0], sz = A.dims[0] * A.dims[1];
    float importance = 0.0f, dt1 = 0.0f;
    bool go = true;
    while (go) {
        ushort2 mm = A.grid[cell[0] + cell[1] * sy + cell[2] * sz];
        float lo = (1.0f / 65535.0f) * (float)mm.x, hi = (1.0f / 65535.0f) * (float)mm.y;
        float dt0 = dt1;
        // stepToNextCell2: x wins ties, then y
        if (dt[0] <= dt[1] && dt[0] <= dt[2]) {
            dt1 = dt[0];
            if (cell[0] == cell_end[0]) go = false; else { dt[0] += deltatx[0]; cell[0] += di[0]; }
        } else if (dt[1] <= dt[0] && dt[1] <= dt[2]) {
            dt1 = dt[1];
            if (cell[1] == cell_end[1]) go = false; else { dt[1] += deltatx[1]; cell[1] += di[1]; }
        } else {
            dt1 = dt[2];
            if (cell[2] == cell_end[2]) go = false; else { dt[2] += deltatx[2]; cell[2] += di[2]; }
        }
        if (!(hi < A.tf_min || lo > A.tf_max)) importance += cpm_fmin(1.0f, dt1) - dt0;
    }
    A.out[id] = importance * len;
}

// cdf layout: rows [h][w+1] then marginal [h+1]
__global__ void __launch_bounds__(128) cdf_rows_kernel(const float* __restrict__ imp, int w, int h, float floor_value,
                                                       float* __restrict__ cdf) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    float* row = cdf + (size_t)y * (w + 1);
    float acc = 0.0f;
    row[0] = 0.0f;
    for (int x = 0; x < w; ++x) {
        acc += cpm_fmax(imp[(size_t)y * w + x], 0.0f) + floor_value;
        row[x + 1] = acc;
    }
}
__global__ void cdf_marginal_kernel(int w, int h, float* __restrict__ cdf) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float* marg = cdf + (size_t)h * (w + 1);
    float acc = 0.0f;
    marg[0] = 0.0f;
    for (int y = 0; y < h; ++y) {
        acc += cdf[(size_t)y * (w + 1) + w];
        marg[y + 1] = acc;
    }
}

// largest i in [0, n-1] with c[i] <= t   (c ascending, c[0] = 0)
__device__ __forceinline__ int upper_cell(const float* __restrict__ c, int n, float t) {
    int lo = 0, hi = n;   // invariant: c[lo] <= t, answer in [lo, hi)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (c[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) warp_samples_kernel(const float* __restrict__ cdf, int w, int h,
                                                           const float4* __restrict__ uni, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = uni[i];
    const float* marg = cdf + (size_t)h * (w + 1);
    const float total = marg[h];
    const float one_m = 0.99999994f;
    float tv = cpm_clamp(s.y, 0.0f, one_m) * total;
    int y = upper_cell(marg, h, tv);
    float rowsum = marg[y + 1] - marg[y];
    float dv = rowsum > 0.0f ? cpm_clamp((tv - marg[y]) / rowsum, 0.0f, one_m) : 0.5f;
    const float* row = cdf + (size_t)y * (w + 1);
    float tu = cpm_clamp(s.x, 0.0f, one_m) * row[w];
    int x = upper_cell(row, w, tu);
    float f = row[x + 1] - row[x];
    float du = f > 0.0f ? cpm_clamp((tu - row[x]) / f, 0.0f, one_m) : 0.5f;
    float pdf = total > 0.0f ? f * ((float)w * (float)h) / total : 1.0f;
    out[i] = make_float4(((float)x + du) / (float)w, ((float)y + dv) / (float)h, s.z, pdf * s.w);
}

}  // namespace

extern "C" {

int cpm_view_importance(cpm_ctx* ctx, const uint16_t* minmax, const int grid_dims[3], const float cell_size[3],
                        const float texture_to_index[16], const float index_to_texture[16], const float* entry,
                        const float* exit, int width, int height, float tf_min, float tf_max, float* importance) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, minmax && grid_dims && cell_size && texture_to_index && index_to_texture && entry && exit && importance,
                "null argument");
    CPM_REQUIRE(ctx, width > 0 && height > 0, "bad image size");
    CPM_REQUIRE(ctx, grid_dims[0] > 0 && grid_dims[1] > 0 && grid_dims[2] > 0, "grid dims must be positive");
    ViewArgs a;
    a.grid = (const ushort2*)minmax;
    for (int k = 0; k < 3; ++k) {
        a.dims[k] = grid_dims[k];
        a.cell[k] = cell_size[k];
    }
    for (int k = 0; k < 16; ++k) {
        a.tex2idx.m[k] = texture_to_index[k];
        a.idx2tex.m[k] = index_to_texture[k];
    }
    a.entry = (const float4*)entry;
    a.exit = (const float4*)exit;
    a.width = width;
    a.height = height;
    a.tf_min = tf_min;
    a.tf_max = tf_max;
    a.out = importance;
    dim3 grid((width + 15) / 16, (height + 7) / 8);
    CPM_LAUNCH(ctx, view_importance_kernel, grid, 128, 0, a);
    return CPM_OK;
}

size_t cpm_sample_importance2d_scratch_floats(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    return (size_t)height * ((size_t)width + 1) + (size_t)height + 1;
}

int cpm_sample_importance2d(cpm_ctx* ctx, const float* importance, int width, int height, float floor_value,
                            const float* uniform_samples, int n, float* cdf_scratch, float* samples_out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    CPM_REQUIRE(ctx, importance && cdf_scratch && width > 0 && height > 0, "null argument / bad image size");
    CPM_REQUIRE(ctx, floor_value >= 0.0f, "floor must be >= 0");
    CPM_LAUNCH(ctx, cdf_rows_kernel, cpm_div_up(height, 128), 128, 0, importance, width, height, floor_value, cdf_scratch);
    CPM_LAUNCH(ctx, cdf_marginal_kernel, 1, 32, 0, width, height, cdf_scratch);
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, uniform_samples && samples_out, "null sample buffer");
    CPM_LAUNCH(ctx, warp_samples_kernel, cpm_div_up(n, 128), 128, 0, cdf_scratch, width, height, (const float4*)uniform_samples,
               n, (float4*)samples_out);
    return CPM_OK;
}

}  // extern "C"
