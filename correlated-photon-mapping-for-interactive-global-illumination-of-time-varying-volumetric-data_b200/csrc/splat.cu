// splat.cu -- photon density estimation by splatting into the light volume (the reference's real
// form of north-star subsystem 7).
//
// Replaces splatPhotonsToLightVolumeKernel and splatSelectedPhotonsToLightVolumeKernel
// (ppm/cl/photonstolightvolume.cl:31-79, 139-166, 168-202) with the Epanechnikov kernel of
// ppm/cl/densityestimationkernel.cl:56-60.
//
// B200 notes: the reference adds with a compare-and-swap loop per voxel; here every add is one
// fire-and-forget RED.E.ADD.F32 resolved in L2 (no return value, no retry loop).  Summation order
// is as undefined as in the reference; tests compare against a double-precision accumulation.
#include "sampling.cuh"

namespace {

struct Mat4s {
    float m[16];
};
__device__ __forceinline__ float3_ xform(const Mat4s& M, float x, float y, float z) {
    return {fmaf(M.m[8], z, fmaf(M.m[4], y, fmaf(M.m[0], x, M.m[12]))),
            fmaf(M.m[9], z, fmaf(M.m[5], y, fmaf(M.m[1], x, M.m[13]))),
            fmaf(M.m[10], z, fmaf(M.m[6], y, fmaf(M.m[2], x, M.m[14])))};
}

struct SplatArgs {
    float* vol;
    Mat4s tex2idx, idx2tex;
    int dim[3];
    const float4* photons;
    const uint32_t* indices;
    int n;
    int per_interaction;
    int n_interactions;
    float radius, scale, multiplier;
};

template <int CH>
__device__ __forceinline__ void splat_one(const SplatArgs& A, float4 p0, float pr, float pg, float pb) {
    if (p0.x == CPM_FLT_MAX_ || p0.y == CPM_FLT_MAX_ || p0.z == CPM_FLT_MAX_) return;
    const float r = A.radius;
    float3_ lo = xform(A.tex2idx, p0.x - r, p0.y - r, p0.z - r);
    float3_ hi = xform(A.tex2idx, p0.x + r, p0.y + r, p0.z + r);
    int sx = (int)cpm_clamp(truncf(lo.x), 0.f, 2147483520.f), sy = (int)cpm_clamp(truncf(lo.y), 0.f, 2147483520.f),
        sz = (int)cpm_clamp(truncf(lo.z), 0.f, 2147483520.f);
    int ex = (int)cpm_clamp(truncf(hi.x + 1.f), -2147483520.f, (float)A.dim[0]),
        ey = (int)cpm_clamp(truncf(hi.y + 1.f), -2147483520.f, (float)A.dim[1]),
        ez = (int)cpm_clamp(truncf(hi.z + 1.f), -2147483520.f, (float)A.dim[2]);
    for (int z = sz; z < ez; ++z)
        for (int y = sy; y < ey; ++y)
            for (int x = sx; x < ex; ++x) {
                size_t vi = (size_t)x + (size_t)y * A.dim[0] + (size_t)z * A.dim[0] * A.dim[1];
                float3_ c = xform(A.idx2tex, (float)x, (float)y, (float)z);
                float dx = c.x - p0.x, dy = c.y - p0.y, dz = c.z - p0.z;
                float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                float xk = dist / r;
                float w = xk <= 1.0f ? 0.75f * (1.0f - xk * xk) : 0.0f;
                float fr = pr * w;
                if (CH == 1) {
                    if (fr != 0.0f) atomicAdd(A.vol + vi, fr);
                } else {
                    float fg = pg * w, fb = pb * w;
                    if (fr != 0.0f) atomicAdd(A.vol + 4 * vi, fr);
                    if (fg != 0.0f) atomicAdd(A.vol + 4 * vi + 1, fg);
                    if (fb != 0.0f) atomicAdd(A.vol + 4 * vi + 2, fb);
                }
            }
}

template <int CH>
__global__ void __launch_bounds__(128) splat_kernel(const SplatArgs A) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.n) return;
    const float s = CPM_INV_4PI_F * A.scale;
    if (!A.indices) {
        float4 p0 = A.photons[2 * (size_t)g], p1 = A.photons[2 * (size_t)g + 1];
        splat_one<CH>(A, p0, p0.w * s, p1.x * s, p1.y * s);
    } else {
        uint32_t id = A.indices[g];
        for (int k = 0; k < A.n_interactions; ++k) {
            size_t pid = (size_t)k * A.per_interaction + id;
            float4 p0 = A.photons[2 * pid], p1 = A.photons[2 * pid + 1];
            splat_one<CH>(A, p0, p0.w * s * A.multiplier, p1.x * s * A.multiplier, p1.y * s * A.multiplier);
        }
    }
}

}  // namespace

extern "C" int cpm_splat_photons(cpm_ctx* ctx, float* light_volume, int channels, const float texture_to_index[16],
                                 const float index_to_texture[16], const int out_dims[3], const float* photons,
                                 const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
                                 float radius, float relative_irradiance_scale, float multiplier) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0, "negative n");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, light_volume && texture_to_index && index_to_texture && out_dims && photons, "null argument");
    CPM_REQUIRE(ctx, channels == 1 || channels == 4, "channels must be 1 or 4");
    CPM_REQUIRE(ctx, radius > 0.0f, "radius must be positive");
    SplatArgs a;
    a.vol = light_volume;
    for (int k = 0; k < 16; ++k) {
        a.tex2idx.m[k] = texture_to_index[k];
        a.idx2tex.m[k] = index_to_texture[k];
    }
    for (int k = 0; k < 3; ++k) a.dim[k] = out_dims[k];
    a.photons = (const float4*)photons;
    a.indices = indices;
    a.n = n;
    a.per_interaction = photons_per_interaction;
    a.n_interactions = n_interactions;
    a.radius = radius;
    a.scale = relative_irradiance_scale;
    a.multiplier = multiplier;
    if (channels == 1)
        CPM_LAUNCH(ctx, splat_kernel<1>, cpm_div_up(n, 128), 128, 0, a);
    else
        CPM_LAUNCH(ctx, splat_kernel<4>, cpm_div_up(n, 128), 128, 0, a);
    return CPM_OK;
}
