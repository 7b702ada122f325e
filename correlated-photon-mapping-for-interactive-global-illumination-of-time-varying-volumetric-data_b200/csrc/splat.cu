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
    const float4* old_photons;  // cpm_splat_photons_update: records to remove (same ids), or null
    float4* old_sync;           // cpm_splat_photons_update_sync: == old_photons, overwritten with the new records
    const uint32_t* indices;
    int n;
    int per_interaction;
    int n_interactions;
    float radius, scale, multiplier;
};

// One photon record into the (2 r s + 2)^3 voxels around it.  The weight 0.75 (1 - (d/r)^2) for d <= r is
// evaluated from the squared distance (no square root, no division) and the voxel centre is built per loop level
// (3 fma per voxel instead of a full matrix product): a few ulp from the reference's operation order, far
// inside the order-dependence of the float adds themselves (tests: relative RMSE <= 1e-5 vs a double sum).
template <int CH, bool DIAG>
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
    const float* M = A.idx2tex.m;
    const float r2 = r * r, k = 0.75f / r2;
    // voxel centre minus photon position, split by loop level
    const float bx = M[12] - p0.x, by = M[13] - p0.y, bz = M[14] - p0.z;
    for (int z = sz; z < ez; ++z) {
        const float fz = (float)z;
        const float zx = fmaf(M[8], fz, bx), zy = fmaf(M[9], fz, by), zz = fmaf(M[10], fz, bz);
        for (int y = sy; y < ey; ++y) {
            const float fy = (float)y;
            const float yx = fmaf(M[4], fy, zx), yy = fmaf(M[5], fy, zy), yz = fmaf(M[6], fy, zz);
            float* row = A.vol + ((size_t)y + (size_t)z * A.dim[1]) * A.dim[0] * CH;
            // axis-aligned volumes (the index-to-texture matrix is diagonal): dy and dz do not change along the row
            const float s_yz = fmaf(yz, yz, yy * yy);
            for (int x = sx; x < ex; ++x) {
                const float fx = (float)x;
                float d2;
                if (DIAG) {
                    const float dx = fmaf(M[0], fx, yx);
                    d2 = fmaf(dx, dx, s_yz);
                } else {
                    float dx = fmaf(M[0], fx, yx), dy = fmaf(M[1], fx, yy), dz = fmaf(M[2], fx, yz);
                    d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                }
                float w = d2 <= r2 ? fmaf(-k, d2, 0.75f) : 0.0f;
                float fr = pr * w;
                if (CH == 1) {
                    if (fr != 0.0f) atomicAdd(row + x, fr);
                } else {
                    float fg = pg * w, fb = pb * w;
                    if (fr != 0.0f) atomicAdd(row + 4 * x, fr);
                    if (fg != 0.0f) atomicAdd(row + 4 * x + 1, fg);
                    if (fb != 0.0f) atomicAdd(row + 4 * x + 2, fb);
                }
            }
        }
    }
}

template <int CH, bool DIAG>
__global__ void __launch_bounds__(128) splat_kernel(const SplatArgs A) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.n) return;
    const float s = CPM_INV_4PI_F * A.scale;
    if (!A.indices) {
        float4 p0 = A.photons[2 * (size_t)g], p1 = A.photons[2 * (size_t)g + 1];
        splat_one<CH, DIAG>(A, p0, p0.w * s, p1.x * s, p1.y * s);
    } else {
        uint32_t id = A.indices[g];
        for (int k = 0; k < A.n_interactions; ++k) {
            size_t pid = (size_t)k * A.per_interaction + id;
            float4 p0 = A.photons[2 * pid], p1 = A.photons[2 * pid + 1];
            if (A.old_photons) {
                // incremental update: -old +new in one pass; a record the re-trace reproduced bit for bit
                // contributes nothing and is skipped (the two passes would cancel up to rounding)
                float4 q0 = A.old_photons[2 * pid], q1 = A.old_photons[2 * pid + 1];
                if (A.old_sync) {   // leave the copy equal to the new records: no whole-buffer copy after the update
                    A.old_sync[2 * pid] = p0;
                    A.old_sync[2 * pid + 1] = p1;
                }
                if (q0.x == p0.x && q0.y == p0.y && q0.z == p0.z && q0.w == p0.w && q1.x == p1.x && q1.y == p1.y) continue;
                splat_one<CH, DIAG>(A, q0, -(q0.w * s * A.multiplier), -(q1.x * s * A.multiplier), -(q1.y * s * A.multiplier));
            }
            splat_one<CH, DIAG>(A, p0, p0.w * s * A.multiplier, p1.x * s * A.multiplier, p1.y * s * A.multiplier);
        }
    }
}

// copyIndexPhotonsKernel (ppm/cl/photonstolightvolume.cl:225-247): the records of the listed ids, every interaction,
// power times `multiplier`, packed as out[out_offset + g + k * n] -- the "mem-aligned changed photons" of
// PhotonToLightVolumeProcessorCL (alignChangedPhotons).  One thread per (id, interaction): 2 x LDG.128 gathered,
// 2 x STG.128 coalesced.
__global__ void __launch_bounds__(256) copy_index_photons_kernel(const float4* __restrict__ photons, const uint32_t* __restrict__ indices,
                                                                 int n, float multiplier, int per_interaction, int n_interactions,
                                                                 float4* __restrict__ out, size_t out_offset) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)n * n_interactions) return;
    const int k = (int)(g / (size_t)n);
    const size_t i = g - (size_t)k * n;
    const size_t pid = (size_t)k * per_interaction + indices[i];
    float4 p0 = photons[2 * pid], p1 = photons[2 * pid + 1];
    p0.w *= multiplier;
    p1.x *= multiplier;
    p1.y *= multiplier;
    out[2 * (out_offset + g)] = p0;
    out[2 * (out_offset + g) + 1] = p1;
}

}  // namespace

extern "C" int cpm_copy_index_photons(cpm_ctx* ctx, const float* photons, const uint32_t* indices, int n, float multiplier,
                                      int photons_per_interaction, int n_interactions, float* aligned_photons, size_t out_offset) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n >= 0 && n_interactions >= 1 && photons_per_interaction >= 0, "negative size");
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, photons && indices && aligned_photons, "null argument");
    const size_t total = (size_t)n * n_interactions;
    CPM_LAUNCH(ctx, copy_index_photons_kernel, cpm_div_up(total, 256), 256, 0, (const float4*)photons, indices, n, multiplier,
               photons_per_interaction, n_interactions, (float4*)aligned_photons, out_offset);
    return CPM_OK;
}

static int splat_common(cpm_ctx* ctx, float* light_volume, int channels, const float texture_to_index[16],
                        const float index_to_texture[16], const int out_dims[3], const float* photons,
                        const float* old_photons, const uint32_t* indices, int n, int photons_per_interaction,
                        int n_interactions, float radius, float relative_irradiance_scale, float multiplier,
                        bool sync_old = false) {
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
    a.old_photons = (const float4*)old_photons;
    a.old_sync = sync_old ? (float4*)const_cast<float*>(old_photons) : nullptr;
    a.indices = indices;
    a.n = n;
    a.per_interaction = photons_per_interaction;
    a.n_interactions = n_interactions;
    a.radius = radius;
    a.scale = relative_irradiance_scale;
    a.multiplier = multiplier;
    // diagonal index-to-texture matrix (an axis-aligned light volume, the usual case): the kernel's row loop is shorter
    const float* m = index_to_texture;
    const bool diag = m[1] == 0.f && m[2] == 0.f && m[4] == 0.f && m[6] == 0.f && m[8] == 0.f && m[9] == 0.f;
    if (channels == 1) {
        if (diag) CPM_LAUNCH(ctx, (splat_kernel<1, true>), cpm_div_up(n, 128), 128, 0, a);
        else CPM_LAUNCH(ctx, (splat_kernel<1, false>), cpm_div_up(n, 128), 128, 0, a);
    } else {
        if (diag) CPM_LAUNCH(ctx, (splat_kernel<4, true>), cpm_div_up(n, 128), 128, 0, a);
        else CPM_LAUNCH(ctx, (splat_kernel<4, false>), cpm_div_up(n, 128), 128, 0, a);
    }
    return CPM_OK;
}

extern "C" int cpm_splat_photons(cpm_ctx* ctx, float* light_volume, int channels, const float texture_to_index[16],
                                 const float index_to_texture[16], const int out_dims[3], const float* photons,
                                 const uint32_t* indices, int n, int photons_per_interaction, int n_interactions,
                                 float radius, float relative_irradiance_scale, float multiplier) {
    if (!ctx) return CPM_E_INVALID;
    return splat_common(ctx, light_volume, channels, texture_to_index, index_to_texture, out_dims, photons, nullptr, indices, n,
                        photons_per_interaction, n_interactions, radius, relative_irradiance_scale, multiplier);
}

extern "C" int cpm_splat_photons_update(cpm_ctx* ctx, float* light_volume, int channels, const float texture_to_index[16],
                                        const float index_to_texture[16], const int out_dims[3], const float* old_photons,
                                        const float* new_photons, const uint32_t* indices, int n,
                                        int photons_per_interaction, int n_interactions, float radius,
                                        float relative_irradiance_scale) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, old_photons && indices, "null argument");
    return splat_common(ctx, light_volume, channels, texture_to_index, index_to_texture, out_dims, new_photons, old_photons,
                        indices, n, photons_per_interaction, n_interactions, radius, relative_irradiance_scale, 1.0f);
}

extern "C" int cpm_splat_photons_update_sync(cpm_ctx* ctx, float* light_volume, int channels, const float texture_to_index[16],
                                             const float index_to_texture[16], const int out_dims[3], float* old_photons,
                                             const float* new_photons, const uint32_t* indices, int n,
                                             int photons_per_interaction, int n_interactions, float radius,
                                             float relative_irradiance_scale) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, old_photons && indices, "null argument");
    CPM_REQUIRE(ctx, old_photons != new_photons, "old and new records must be different buffers");
    return splat_common(ctx, light_volume, channels, texture_to_index, index_to_texture, out_dims, new_photons, old_photons,
                        indices, n, photons_per_interaction, n_interactions, radius, relative_irradiance_scale, 1.0f, true);
}
