// raycast.cu -- final image from the light volume: front-to-back emission-absorption ray march of
// volume x transfer function x light volume.
//
// The step the workspace network performs with Inviwo's LightingRaycaster after PhotonToLightVolumeProcessorCL
// (workspaces/CorrelatedPhotonMappingSingleVolume.inv:1178-1271).  LightingRaycaster is an Inviwo core (OpenGL)
// processor and not part of the reference tree: PARITY UNPINNED, the checker is oracle/orc_gather.c's
// orc_raycast_light_volume.  Camera, sample positions and compositing are those of cpm_gather_raymarch, so the
// two images differ only in how the in-scattered radiance of a sample is obtained (light-volume fetch vs photon
// gather).
//
// B200 notes: one ray per lane, warps take 8 x 4 pixel tiles from a global counter; all-transparent cells of the
// opacity-bound grid are left in one jump; visible samples fetch 8 light-volume taps from the linear f32 buffer
// (256^3: 64 MB, L2 resident after the first rays).
#include <string.h>

#include <algorithm>

#include "raymarch.cuh"

namespace {

struct RaycastArgs {
    cpm_gather_params p;
    VolumeView vol;
    const float4* tf;
    int tf_width;
    const float* lv;
    int lx, ly, lz;
    float4* image;
    BoundGrid bound;
};

// trilinear light-volume sample (normalised coordinates, clamp-to-edge), channel ch of NCH interleaved
template <int NCH>
__device__ __forceinline__ void sample_light(const RaycastArgs& A, float px, float py, float pz, float e[3]) {
    const float fx = (float)A.lx, fy = (float)A.ly, fz = (float)A.lz;
    float u = fmaf(px, fx, -0.5f), v = fmaf(py, fy, -0.5f), w = fmaf(pz, fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    float a = u - fu, b = v - fv, c = w - fw;
    int i0 = (int)cpm_clamp(fu, -1.0f, fx - 1.0f), j0 = (int)cpm_clamp(fv, -1.0f, fy - 1.0f), k0 = (int)cpm_clamp(fw, -1.0f, fz - 1.0f);
    int i1 = min(i0 + 1, A.lx - 1), j1 = min(j0 + 1, A.ly - 1), k1 = min(k0 + 1, A.lz - 1);
    i0 = max(i0, 0); j0 = max(j0, 0); k0 = max(k0, 0);
    const size_t sy = (size_t)A.lx, sz = (size_t)A.lx * A.ly;
    const size_t b00 = (size_t)k0 * sz + (size_t)j0 * sy, b10 = (size_t)k0 * sz + (size_t)j1 * sy;
    const size_t b01 = (size_t)k1 * sz + (size_t)j0 * sy, b11 = (size_t)k1 * sz + (size_t)j1 * sy;
    if (NCH == 1) {
        const float* L = A.lv;
        float x00 = lerpf(__ldg(L + b00 + i0), __ldg(L + b00 + i1), a), x10 = lerpf(__ldg(L + b10 + i0), __ldg(L + b10 + i1), a);
        float x01 = lerpf(__ldg(L + b01 + i0), __ldg(L + b01 + i1), a), x11 = lerpf(__ldg(L + b11 + i0), __ldg(L + b11 + i1), a);
        e[0] = e[1] = e[2] = lerpf(lerpf(x00, x10, b), lerpf(x01, x11, b), c);
    } else {
        const float4* L = (const float4*)A.lv;
        float4 t000 = __ldg(L + b00 + i0), t100 = __ldg(L + b00 + i1), t010 = __ldg(L + b10 + i0), t110 = __ldg(L + b10 + i1);
        float4 t001 = __ldg(L + b01 + i0), t101 = __ldg(L + b01 + i1), t011 = __ldg(L + b11 + i0), t111 = __ldg(L + b11 + i1);
#define CH(m, q)                                                                                               \
    e[q] = lerpf(lerpf(lerpf(t000.m, t100.m, a), lerpf(t010.m, t110.m, a), b),                                 \
                 lerpf(lerpf(t001.m, t101.m, a), lerpf(t011.m, t111.m, a), b), c);
        CH(x, 0) CH(y, 1) CH(z, 2)
#undef CH
    }
}

// the same sample for the one-channel light volume in two steps, so that the taps of several samples can be in flight
struct LightTaps {
    float t[8];   // t000 t100 t010 t110 t001 t101 t011 t111
    float a, b, c;
};
__device__ __forceinline__ LightTaps fetch_light1(const RaycastArgs& A, float px, float py, float pz) {
    LightTaps T;
    const float fx = (float)A.lx, fy = (float)A.ly, fz = (float)A.lz;
    float u = fmaf(px, fx, -0.5f), v = fmaf(py, fy, -0.5f), w = fmaf(pz, fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    T.a = u - fu; T.b = v - fv; T.c = w - fw;
    int i0 = (int)cpm_clamp(fu, -1.0f, fx - 1.0f), j0 = (int)cpm_clamp(fv, -1.0f, fy - 1.0f), k0 = (int)cpm_clamp(fw, -1.0f, fz - 1.0f);
    int i1 = min(i0 + 1, A.lx - 1), j1 = min(j0 + 1, A.ly - 1), k1 = min(k0 + 1, A.lz - 1);
    i0 = max(i0, 0); j0 = max(j0, 0); k0 = max(k0, 0);
    const size_t sy = (size_t)A.lx, sz = (size_t)A.lx * A.ly;
    const size_t b00 = (size_t)k0 * sz + (size_t)j0 * sy, b10 = (size_t)k0 * sz + (size_t)j1 * sy;
    const size_t b01 = (size_t)k1 * sz + (size_t)j0 * sy, b11 = (size_t)k1 * sz + (size_t)j1 * sy;
    const float* L = A.lv;
    T.t[0] = __ldg(L + b00 + i0); T.t[1] = __ldg(L + b00 + i1); T.t[2] = __ldg(L + b10 + i0); T.t[3] = __ldg(L + b10 + i1);
    T.t[4] = __ldg(L + b01 + i0); T.t[5] = __ldg(L + b01 + i1); T.t[6] = __ldg(L + b11 + i0); T.t[7] = __ldg(L + b11 + i1);
    return T;
}
__device__ __forceinline__ float blend_light1(const LightTaps& T) {   // the arithmetic of sample_light<1>, same order
    float x00 = lerpf(T.t[0], T.t[1], T.a), x10 = lerpf(T.t[2], T.t[3], T.a);
    float x01 = lerpf(T.t[4], T.t[5], T.a), x11 = lerpf(T.t[6], T.t[7], T.a);
    return lerpf(lerpf(x00, x10, T.b), lerpf(x01, x11, T.b), T.c);
}

// Samples are taken RB at a time: the voxel taps of the whole batch are requested together, then the light-volume taps
// of its visible samples, then the batch is composited in order.  A sample in a transparent cell that a one-at-a-time
// march would have stepped over has opacity exactly 0 and contributes nothing, and samples behind the one that ends the
// ray are dropped: the image is that of the plain loop, bit for bit.
template <int FMT, int LAYOUT, int NCH, int RB>
__global__ void __launch_bounds__(128) raycast_kernel(const RaycastArgs A, unsigned* __restrict__ tile_counter) {
    extern __shared__ float4 s_tf[];
    for (int i = threadIdx.x; i < A.tf_width; i += blockDim.x) s_tf[i] = A.tf[i];
    __syncthreads();
    const cpm_gather_params& P = A.p;
    const int tiles_x = (P.width + 7) / 8, tiles_y = (P.height + 3) / 4;
    const unsigned n_tiles = (unsigned)tiles_x * (unsigned)tiles_y;
    const int lane = threadIdx.x & 31;
    const float ftfw = (float)A.tf_width;
    const float inv_step = 1.0f / P.step;
    while (true) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int px = (int)(tile % tiles_x) * 8 + (lane & 7);
        const int py = (int)(tile / tiles_x) * 4 + (lane >> 3);
        if (px >= P.width || py >= P.height) continue;
        float3_ d = camera_ray(P, px, camera_row(P, py));
        float3_ o = {P.cam_origin[0], P.cam_origin[1], P.cam_origin[2]};
        float t0 = 0.0f, t1 = CPM_FLT_MAX_;
        float lr = 0.f, lg = 0.f, lb = 0.f, T = 1.0f;
        if (ray_box(P.aabb_min, P.aabb_max, o, d, t0, t1)) {
            CellRay R = {0, 0, 0, 0, 0, 0};
            float ix = 0.f, iy = 0.f, iz = 0.f;
            if (A.bound.g) {
                R = cell_ray(A.bound, o, d);
                ix = 1.0f / R.dx; iy = 1.0f / R.dy; iz = 1.0f / R.dz;
            }
            int k = 0;
            bool done = false;
            while (!done) {
                float t = fmaf((float)k + 0.5f, P.step, t0);
                if (A.bound.g) k = skip_transparent(A.bound, R, ix, iy, iz, t0, t1, P.step, inv_step, k, t);
                if (!(t < t1)) break;
                float tj[RB];
                Taps tp[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    tj[j] = fmaf((float)(k + j) + 0.5f, P.step, t0);
                    tp[j] = fetch_taps<FMT, LAYOUT>(A.vol, fmaf(tj[j], d.x, o.x), fmaf(tj[j], d.y, o.y), fmaf(tj[j], d.z, o.z));
                }
                float4 c[RB];
                bool vis[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    c[j] = sample_tf_rgba(s_tf, A.tf_width, ftfw, blend_taps<FMT>(A.vol, tp[j]));
                    vis[j] = tj[j] < t1 && c[j].w > 0.0f;
                }
                LightTaps lt[NCH == 1 ? RB : 1];
                if (NCH == 1) {
#pragma unroll
                    for (int j = 0; j < RB; ++j)
                        if (vis[j]) lt[j] = fetch_light1(A, fmaf(tj[j], d.x, o.x), fmaf(tj[j], d.y, o.y), fmaf(tj[j], d.z, o.z));
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    if (done) continue;
                    if (!(tj[j] < t1)) {
                        done = true;
                        continue;
                    }
                    if (vis[j]) {
                        float e[3];
                        if (NCH == 1)
                            e[0] = e[1] = e[2] = blend_light1(lt[j]);
                        else
                            sample_light<NCH>(A, fmaf(tj[j], d.x, o.x), fmaf(tj[j], d.y, o.y), fmaf(tj[j], d.z, o.z), e);
                        float Ts = cpm_expf(-(c[j].w * P.sigma_scale) * P.step);
                        float wgt = T * (1.0f - Ts);
                        lr = fmaf(wgt * c[j].x, e[0], lr);
                        lg = fmaf(wgt * c[j].y, e[1], lg);
                        lb = fmaf(wgt * c[j].z, e[2], lb);
                        T *= Ts;
                        if (T < 1e-4f) done = true;
                    }
                }
                k += RB;
            }
        }
        A.image[(size_t)py * P.width + px] = make_float4(lr, lg, lb, 1.0f - T);
    }
}

template <int FMT, int LAYOUT, int NCH, int RB>
int launch_raycast_rb(cpm_ctx* ctx, const RaycastArgs& a) {
    size_t smem = (size_t)a.tf_width * sizeof(float4);
    if (smem > 48 * 1024)
        CPM_CUDA(ctx, cudaFuncSetAttribute(raycast_kernel<FMT, LAYOUT, NCH, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* scratch;
    int rc = cpm_scratch(ctx, 4096, &scratch);
    if (rc != CPM_OK) return rc;
    unsigned* counter = (unsigned*)((char*)scratch + 3072);
    CPM_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    int tiles = ((a.p.width + 7) / 8) * ((a.p.height + 3) / 4);
    int per_sm = 0;
    CPM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raycast_kernel<FMT, LAYOUT, NCH, RB>, 128, smem));
    unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * std::max(per_sm, 1), (long long)cpm_div_up(tiles, 4));
    CPM_LAUNCH(ctx, (raycast_kernel<FMT, LAYOUT, NCH, RB>), grid, 128, smem, a, counter);
    return CPM_OK;
}

}  // namespace

#ifndef CPM_RAYCAST_BATCH
#define CPM_RAYCAST_BATCH 4   // samples whose taps are in flight together
#endif
template <int FMT, int LAYOUT, int NCH>
int launch_raycast(cpm_ctx* ctx, const RaycastArgs& a) {
    static const int rb = getenv("CPM_RAYCAST_BATCH") ? atoi(getenv("CPM_RAYCAST_BATCH")) : CPM_RAYCAST_BATCH;   // tuning sweeps only
    if (rb >= 4) return launch_raycast_rb<FMT, LAYOUT, NCH, 4>(ctx, a);
    if (rb >= 2) return launch_raycast_rb<FMT, LAYOUT, NCH, 2>(ctx, a);
    return launch_raycast_rb<FMT, LAYOUT, NCH, 1>(ctx, a);
}

extern "C" int cpm_raycast_light_volume(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                                        const cpm_gather_params* params, const float* light_volume, const int lv_dims[3],
                                        int channels, float* image) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && tf_rgba && params && light_volume && lv_dims && image, "null argument");
    CPM_REQUIRE(ctx, tf_width >= 1 && tf_width <= 8192, "tf_width out of range");
    CPM_REQUIRE(ctx, channels == 1 || channels == 4, "channels must be 1 or 4");
    CPM_REQUIRE(ctx, lv_dims[0] > 0 && lv_dims[1] > 0 && lv_dims[2] > 0, "light volume dims must be positive");
    CPM_REQUIRE(ctx, params->width > 0 && params->height > 0 && params->step > 0.0f, "bad image size / step");
    CPM_REQUIRE(ctx, params->strip_first >= 0 && params->strip_stride >= 0, "negative strip_first / strip_stride");
    RaycastArgs a;
    a.p = *params;
    a.vol = make_view(vol);
    a.tf = (const float4*)tf_rgba;
    a.tf_width = tf_width;
    a.lv = light_volume;
    a.lx = lv_dims[0];
    a.ly = lv_dims[1];
    a.lz = lv_dims[2];
    a.image = (float4*)image;
    CPM_REQUIRE(ctx, make_bound_grid(a.bound, params->opacity_bound, vol->dims, params->bound_cell_log2),
                "bound_cell_log2 must be in 0..8 and the bound grid smaller than 2^31 cells");
#define CPM_RC(F, L)                                                                  \
    return channels == 1 ? launch_raycast<F, L, 1>(ctx, a) : launch_raycast<F, L, 4>(ctx, a);
#define CPM_DISPATCH(F)                                        \
    if (vol->layout == CPM_VOLUME_TEXTURE) {                   \
        CPM_RC(F, CPM_VOLUME_TEXTURE)                          \
    } else {                                                   \
        CPM_RC(F, CPM_VOLUME_LINEAR)                           \
    }
    switch (vol->format) {
        case CPM_FMT_U8: CPM_DISPATCH(CPM_FMT_U8)
        case CPM_FMT_U16: CPM_DISPATCH(CPM_FMT_U16)
        default: CPM_DISPATCH(CPM_FMT_F32)
    }
#undef CPM_DISPATCH
#undef CPM_RC
}
