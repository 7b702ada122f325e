// emission.cu -- sample generation and light-source sampling (north-star subsystem 2).
//
// Replaces uniformSampleGenerator2DKernel (isc/cl/uniformsamplegenerator2d.cl:35-52),
// directionalLightSamplerKernel (lcl/cl/directionallightsampler.cl:38-63) and
// lightSampleMeshIntersectionKernel (lcl/cl/intersection/lightsamplemeshintersection.cl:37-59).
// All three are pure streaming kernels (16-32 B per sample); 16-byte vector stores.
#include "sampling.cuh"

namespace {

__global__ void __launch_bounds__(256) uniform2d_kernel(float nx, float ny, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float fi = (float)i;
    // coord = (fmod(id, dims.x), id / dims.x): the y coordinate is deliberately not floored
    float cx = fmodf(fi, nx);
    float cy = fi / nx;
    out[i] = make_float4((0.5f + cx) / nx, (0.5f + cy) / ny, 0.0f, 1.0f);
}

struct DirLight {
    float radiance[3], dir[3], origin[3], u[3], v[3], area;
};

__global__ void __launch_bounds__(256) directional_kernel(const float4* __restrict__ samples, DirLight L, int n,
                                                          float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = samples[i];  // u, v, w, pdf
    // planeOrigin + planeTangentU * s.x + planeTangentV * s.y as written (lcl/cl/directionallightsampler.cl:56): rounded
    // products and sums, left to right
    float ox = __fadd_rn(__fadd_rn(L.origin[0], __fmul_rn(L.u[0], s.x)), __fmul_rn(L.v[0], s.y));
    float oy = __fadd_rn(__fadd_rn(L.origin[1], __fmul_rn(L.u[1], s.x)), __fmul_rn(L.v[1], s.y));
    float oz = __fadd_rn(__fadd_rn(L.origin[2], __fmul_rn(L.u[2], s.x)), __fmul_rn(L.v[2], s.y));
    float pdf = s.w / L.area;
    float2 ang = encode_direction({L.dir[0], L.dir[1], L.dir[2]});
    out[2 * (size_t)i] = make_float4(ox, oy, oz, L.radiance[0] / pdf);
    out[2 * (size_t)i + 1] = make_float4(L.radiance[1] / pdf, L.radiance[2] / pdf, ang.x, ang.y);
}

struct PointLight {
    float radiance[3], pos[3];
};

__global__ void __launch_bounds__(256) point_kernel(const float4* __restrict__ samples, PointLight L, int n,
                                                    float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = samples[i];
    float3_ d = uniform_sample_sphere(s.x, s.y);
    d = {-d.x, -d.y, -d.z};
    float2 ang = encode_direction(d);
    const float pdf = CPM_INV_4PI_F;  // uniformSpherePdf()
    out[2 * (size_t)i] = make_float4(L.pos[0], L.pos[1], L.pos[2], L.radiance[0] / pdf);
    out[2 * (size_t)i + 1] = make_float4(L.radiance[1] / pdf, L.radiance[2] / pdf, ang.x, ang.y);
}

__device__ __forceinline__ float dot3(float3_ a, float3_ b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3_ cross3(float3_ a, float3_ b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}

// Moeller-Trumbore over the index triples; (t0,t1) = (max(0,nearest hit), min(FLT_MAX, farthest hit)).
// The mesh (scene proxy: 8 vertices, 36 indices) is staged in shared memory.
__global__ void __launch_bounds__(128) mesh_intersect_kernel(const float* __restrict__ vertices,
                                                             const int* __restrict__ indices, int n_indices,
                                                             const float4* __restrict__ light_samples, int n,
                                                             float2* __restrict__ out) {
    extern __shared__ float s_tri[];  // 9 floats per triangle
    int n_tri = n_indices / 3;
    for (int k = threadIdx.x; k < n_tri * 9; k += blockDim.x) {
        int tri = k / 9, r = k % 9;
        s_tri[k] = vertices[3 * (size_t)indices[3 * tri + r / 3] + r % 3];
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 l0 = light_samples[2 * (size_t)i], l1 = light_samples[2 * (size_t)i + 1];
    float3_ o = {l0.x, l0.y, l0.z};
    float3_ d = decode_direction(l1.z, l1.w);
    float tn = 3.402823466e+38f, tf = -3.402823466e+38f;
    bool hit = false;
    for (int t = 0; t < n_tri; ++t) {
        const float* T = s_tri + 9 * t;
        float3_ v0 = {T[0], T[1], T[2]};
        float3_ e1 = {T[3] - v0.x, T[4] - v0.y, T[5] - v0.z};
        float3_ e2 = {T[6] - v0.x, T[7] - v0.y, T[8] - v0.z};
        float3_ p = cross3(d, e2);
        float det = dot3(e1, p);
        if (fabsf(det) < 1e-12f) continue;
        float inv = 1.0f / det;
        float3_ tv = {o.x - v0.x, o.y - v0.y, o.z - v0.z};
        float u = dot3(tv, p) * inv;
        if (u < 0.0f || u > 1.0f) continue;
        float3_ q = cross3(tv, e1);
        float v = dot3(d, q) * inv;
        if (v < 0.0f || u + v > 1.0f) continue;
        float tt = dot3(e2, q) * inv;
        tn = cpm_fmin(tn, tt);
        tf = cpm_fmax(tf, tt);
        hit = true;
    }
    float t0 = 0.0f, t1 = 3.402823466e+38f;
    if (hit) {
        t0 = cpm_fmax(t0, tn);
        t1 = cpm_fmin(t1, tf);
        hit = t0 < t1;
    }
    if (!hit) {
        t0 = 0.0f;
        t1 = -1.0f;
    }
    out[i] = make_float2(t0, t1);
}

}  // namespace

extern "C" {

int cpm_sample_uniform2d(cpm_ctx* ctx, float nx, float ny, int n_elements, float* out) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n_elements >= 0, "negative n_elements");
    if (n_elements == 0) return CPM_OK;
    CPM_REQUIRE(ctx, out != nullptr, "out is NULL");
    CPM_REQUIRE(ctx, nx > 0.0f && ny > 0.0f, "dimensions must be positive");
    CPM_LAUNCH(ctx, uniform2d_kernel, cpm_div_up(n_elements, 256), 256, 0, nx, ny, n_elements, (float4*)out);
    return CPM_OK;
}

int cpm_light_sample_directional(cpm_ctx* ctx, const float* samples, const float radiance[3], const float direction[3],
                                 const float plane_origin[3], const float plane_u[3], const float plane_v[3],
                                 float plane_area, int n, float* light_samples) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, samples && radiance && direction && plane_origin && plane_u && plane_v && light_samples,
                "null argument");
    DirLight L;
    for (int k = 0; k < 3; ++k) {
        L.radiance[k] = radiance[k];
        L.dir[k] = direction[k];
        L.origin[k] = plane_origin[k];
        L.u[k] = plane_u[k];
        L.v[k] = plane_v[k];
    }
    L.area = plane_area;
    CPM_LAUNCH(ctx, directional_kernel, cpm_div_up(n, 256), 256, 0, (const float4*)samples, L, n, (float4*)light_samples);
    return CPM_OK;
}

int cpm_light_sample_point(cpm_ctx* ctx, const float* samples, const float radiance[3], const float position[3], int n,
                           float* light_samples) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, samples && radiance && position && light_samples, "null argument");
    PointLight L;
    for (int k = 0; k < 3; ++k) {
        L.radiance[k] = radiance[k];
        L.pos[k] = position[k];
    }
    CPM_LAUNCH(ctx, point_kernel, cpm_div_up(n, 256), 256, 0, (const float4*)samples, L, n, (float4*)light_samples);
    return CPM_OK;
}

int cpm_light_mesh_intersect(cpm_ctx* ctx, const float* vertices, const int32_t* indices, int n_indices,
                             const float* light_samples, int n, float* intersections) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, vertices && indices && light_samples && intersections, "null argument");
    CPM_REQUIRE(ctx, n_indices >= 3 && n_indices % 3 == 0, "n_indices must be a positive multiple of 3");
    CPM_REQUIRE(ctx, n_indices <= 3 * 1200, "mesh too large for the shared-memory proxy path (max 1200 triangles)");
    size_t smem = (size_t)(n_indices / 3) * 9 * sizeof(float);
    CPM_LAUNCH(ctx, mesh_intersect_kernel, cpm_div_up(n, 128), 128, smem, vertices, indices, n_indices,
               (const float4*)light_samples, n, (float2*)intersections);
    return CPM_OK;
}

}  // extern "C"
