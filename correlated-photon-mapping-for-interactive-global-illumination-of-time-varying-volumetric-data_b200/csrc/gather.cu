// gather.cu -- photon map for gathering: photon cell keys, cell-sorted photon records and the view-ray-march
// gathering / density-estimation kernel (north-star subsystems 5, 6 and 7).
//
// None of this exists in the reference as a launched kernel: its density
// estimation is the splat of splat.cu, and its only per-point gather, photonsToLightVolumeKernel
// (ppm/cl/photonstolightvolume.cl:81-134), is commented out of the host code.  What is kept from the
// reference is the estimator itself -- Epanechnikov kernel 0.75 (1 - (d/r)^2) for d <= r
// (ppm/cl/densityestimationkernel.cl:56-60), power * isotropic phase (1/4pi) * relativeIrradianceScale
// (ppm/cl/photonstolightvolume.cl:160-165) -- so that a gather at a light-volume voxel centre equals the
// splat's value there (tests check exactly that).  Parity is against the oracle's restatement
// (oracle/orc_gather.c): "parity unpinned".
//
// Pipeline: cpm_photon_cell_keys (cell id of every photon record, sentinel records -> n_cells)
//        -> cpm_radix_sort_u32 (keys, ids)  -> cpm_build_cell_ranges  -> cpm_reorder_photons (records in cell
//        order, so a cell's photons are one contiguous 32 B-strided run)  -> cpm_gather_raymarch.
#include <string.h>

#include <algorithm>

#include "raymarch.cuh"

namespace {

__device__ __forceinline__ int cell_coord(float p, float g, int n) {
    return (int)cpm_clamp(truncf(p * g), 0.0f, (float)(n - 1));
}

__global__ void __launch_bounds__(256) photon_cell_keys_kernel(const float4* __restrict__ photons, size_t n, int gx, int gy,
                                                               int gz, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = photons[2 * i];
    uint32_t key;
    if (p.x == CPM_FLT_MAX_ || p.y == CPM_FLT_MAX_ || p.z == CPM_FLT_MAX_) {
        key = (uint32_t)gx * (uint32_t)gy * (uint32_t)gz;   // empty slot: after every real cell
    } else {
        int cx = cell_coord(p.x, (float)gx, gx), cy = cell_coord(p.y, (float)gy, gy), cz = cell_coord(p.z, (float)gz, gz);
        key = (uint32_t)cx + (uint32_t)gx * ((uint32_t)cy + (uint32_t)gy * (uint32_t)cz);
    }
    keys[i] = key;
    if (ids) ids[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) reorder_kernel(const float4* __restrict__ photons, const uint32_t* __restrict__ ids,
                                                      size_t n, float4* __restrict__ out) {
    // two threads per record: each moves one 16 B half
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = g >> 1;
    if (i >= n) return;
    out[g] = photons[2 * (size_t)ids[i] + (g & 1)];
}
// the same records, planar: out[j] = first half of record ids[j], out[n + j] = second half (the gather tests every
// candidate against its first half only: packed halves put two candidates in a sector and halve the cache footprint
// of a map that outgrows L2)
__global__ void __launch_bounds__(256) reorder_planar_kernel(const float4* __restrict__ photons, const uint32_t* __restrict__ ids,
                                                             size_t n, float4* __restrict__ out) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 2 * n) return;
    const size_t j = g >> 1, h = g & 1;
    out[h * n + j] = photons[2 * (size_t)ids[j] + h];
}

struct GatherArgs {
    cpm_gather_params p;
    VolumeView vol;
    const float4* tf;
    int tf_width;
    const float4* photons;   // cell-sorted records: (pos, power.r) of record i at [ps * i], (power.g, power.b, ..) ps * i + p1
    size_t ps, p1;           // interleaved 32-byte records: 2, 1; planar: 1, number of records
    const uint32_t* cell_start;
    const uint32_t* cell_end;
    float4* image;
    BoundGrid bound;   // optional per-cell opacity bound of (volume, tf): zero cells are skipped
};

// irradiance estimate at x: sum over photons within `radius` of power * Epanechnikov weight
__device__ __forceinline__ void gather_point(const GatherArgs& A, float x, float y, float z, float& er, float& eg, float& eb) {
    const cpm_gather_params& P = A.p;
    const float r = P.radius;
    const int gx = P.grid_dims[0], gy = P.grid_dims[1], gz = P.grid_dims[2];
    int x0 = cell_coord(x - r, (float)gx, gx), x1 = cell_coord(x + r, (float)gx, gx);
    int y0 = cell_coord(y - r, (float)gy, gy), y1 = cell_coord(y + r, (float)gy, gy);
    int z0 = cell_coord(z - r, (float)gz, gz), z1 = cell_coord(z + r, (float)gz, gz);
    for (int cz = z0; cz <= z1; ++cz)
        for (int cy = y0; cy <= y1; ++cy)
            for (int cx = x0; cx <= x1; ++cx) {
                uint32_t c = (uint32_t)cx + (uint32_t)gx * ((uint32_t)cy + (uint32_t)gy * (uint32_t)cz);
                uint32_t b = __ldg(A.cell_start + c), e = __ldg(A.cell_end + c);
                for (uint32_t j = b; j < e; ++j) {
                    float4 p0 = __ldg(A.photons + A.ps * (size_t)j);
                    float dx = p0.x - x, dy = p0.y - y, dz = p0.z - z;
                    float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                    float xk = dist / r;
                    if (xk <= 1.0f) {
                        float w = 0.75f * (1.0f - xk * xk);
                        float4 p1 = __ldg(A.photons + A.ps * (size_t)j + A.p1);
                        er = fmaf(p0.w, w, er);
                        eg = fmaf(p1.x, w, eg);
                        eb = fmaf(p1.y, w, eb);
                    }
                }
            }
}

// View-ray-march gather.  Same samples t_k = t0 + (k + 1/2) step, same estimator and compositing as the
// per-sample formulation (oracle/orc_gather.c), organised for the machine:
//  * warps take 8 x 4 pixel tiles from a global counter (rays differ in cost by orders of magnitude: a static
//    tile assignment leaves two thirds of the warp slots idle behind the few expensive tiles);
//  * samples in cells whose opacity bound is exactly zero are never fetched: the march jumps to the cell's exit
//    (the bound grid pads cells by one voxel, i.e. two samples, which covers the rounding of the exit point);
//  * the remaining samples are taken GATHER_S at a time: their taps are requested together, and if any of them is
//    visible the photons of the cells around the whole segment are read ONCE -- per photon the distance to
//    the ray (|q|^2 - (q.d)^2) rejects most candidates, and the survivors update the batch's estimates through
//    d_k^2 = d_perp^2 + (t_k - q.d)^2 -- instead of once per sample through 8 cells.
constexpr int GATHER_S = 12;
#ifndef CPM_GATHER_MIN_CTAS
#define CPM_GATHER_MIN_CTAS 4   // __launch_bounds__(128, .): register cap 128 (5 -> 102 and 6 -> 85 spill and measure slower)
#endif

template <int FMT, int LAYOUT>
__global__ void __launch_bounds__(128, CPM_GATHER_MIN_CTAS) gather_kernel(const GatherArgs A, unsigned* __restrict__ tile_counter) {
    extern __shared__ float4 s_tf[];
    for (int i = threadIdx.x; i < A.tf_width; i += blockDim.x) s_tf[i] = A.tf[i];
    __syncthreads();
    const cpm_gather_params& P = A.p;
    const int tiles_x = (P.width + 7) / 8, tiles_y = (P.height + 3) / 4;
    const unsigned n_tiles = (unsigned)tiles_x * (unsigned)tiles_y;
    const int lane = threadIdx.x & 31;
    const float ftfw = (float)A.tf_width;
    const float s = CPM_INV_4PI_F * P.scale;
    const float r = P.radius, r2 = r * r, kw = 0.75f / r2;
    const int gx = P.grid_dims[0], gy = P.grid_dims[1], gz = P.grid_dims[2];
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
            bool live = true;
            while (live) {
                float t = fmaf((float)k + 0.5f, P.step, t0);
                if (A.bound.g) k = skip_transparent(A.bound, R, ix, iy, iz, t0, t1, P.step, inv_step, k, t);
                if (!(t < t1)) break;
                // ---- a batch of GATHER_S samples: request all taps, then classify
                float v[GATHER_S];
                bool any = false;
                int nb = 0;   // samples of the batch that lie before t1
#pragma unroll
                for (int h = 0; h < GATHER_S; h += 4) {   // four samples' taps in flight at a time
                    Taps tp[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float tj = fmaf((float)(k + h + j) + 0.5f, P.step, t0);
                        tp[j] = fetch_taps<FMT, LAYOUT>(A.vol, fmaf(tj, d.x, o.x), fmaf(tj, d.y, o.y), fmaf(tj, d.z, o.z));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float tj = fmaf((float)(k + h + j) + 0.5f, P.step, t0);
                        v[h + j] = blend_taps<FMT>(A.vol, tp[j]);
                        if (tj < t1) {
                            nb = h + j + 1;
                            any = any || sample_tf_alpha4(s_tf, A.tf_width, ftfw, v[h + j]) > 0.0f;
                        }
                    }
                }
                if (any) {
                    float er[GATHER_S], eg[GATHER_S], eb[GATHER_S];
#pragma unroll
                    for (int j = 0; j < GATHER_S; ++j) er[j] = eg[j] = eb[j] = 0.0f;
                    // cells touched by the segment [x_first, x_last] grown by r
                    const float ta = fmaf((float)k + 0.5f, P.step, t0), tb = fmaf((float)(k + nb - 1) + 0.5f, P.step, t0);
                    const float ax = fmaf(ta, d.x, o.x), ay = fmaf(ta, d.y, o.y), az = fmaf(ta, d.z, o.z);
                    const float bx = fmaf(tb, d.x, o.x), by = fmaf(tb, d.y, o.y), bz = fmaf(tb, d.z, o.z);
                    const int x0 = cell_coord(fminf(ax, bx) - r, (float)gx, gx), x1 = cell_coord(fmaxf(ax, bx) + r, (float)gx, gx);
                    const int y0 = cell_coord(fminf(ay, by) - r, (float)gy, gy), y1 = cell_coord(fmaxf(ay, by) + r, (float)gy, gy);
                    const int z0 = cell_coord(fminf(az, bz) - r, (float)gz, gz), z1 = cell_coord(fmaxf(az, bz) + r, (float)gz, gz);
                    const float tlen = tb - ta;
                    // one photon against the batch: the distance to the ray rejects most, survivors update all estimates
                    auto visit = [&](const float4& p0, const float4& p1) {
                        float qx = p0.x - ax, qy = p0.y - ay, qz = p0.z - az;
                        float tc = fmaf(qz, d.z, fmaf(qy, d.y, qx * d.x));       // parameter of the closest approach
                        float q2 = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
                        float dp2 = fmaf(-tc, tc, q2);                              // squared distance to the ray
                        if (dp2 <= r2 && tc >= -r && tc <= tlen + r) {
#pragma unroll
                            for (int j = 0; j < GATHER_S; ++j) {
                                float dt = fmaf((float)j, P.step, -tc);
                                float d2 = fmaf(dt, dt, fmaxf(dp2, 0.0f));
                                float w = d2 <= r2 ? fmaf(-kw, d2, 0.75f) : 0.0f;
                                er[j] = fmaf(p0.w, w, er[j]);
                                eg[j] = fmaf(p1.x, w, eg[j]);
                                eb[j] = fmaf(p1.y, w, eb[j]);
                            }
                        }
                    };
                    // Cells x0..x1 of a row (cy, cz) are contiguous in the cell-sorted record array: one run per row.
                    // The loads are dependent (range -> records) and the lanes of a warp walk different runs, so the
                    // latency is hidden inside the lane: the next row's range is requested while this row's records
                    // are tested, and two whole records are requested per iteration -- both halves, so that a
                    // survivor never makes its warp wait for a dependent second load (for 32-byte records the
                    // second half is the same sector).  Same visiting order as a plain loop.
                    const int nyr = y1 - y0 + 1, nruns = nyr * (z1 - z0 + 1);
                    const uint32_t xspan = (uint32_t)(x1 - x0);
                    int ry = 0, rz = 0;   // row cursor of the next range request
                    auto next_range = [&](uint32_t& b_, uint32_t& e_) {
                        const uint32_t c0 = (uint32_t)x0 + (uint32_t)gx * ((uint32_t)(y0 + ry) + (uint32_t)gy * (uint32_t)(z0 + rz));
                        b_ = __ldg(A.cell_start + c0);
                        e_ = __ldg(A.cell_end + c0 + xspan);
                        if (++ry == nyr) ry = 0, ++rz;
                    };
                    uint32_t nb_, ne_;
                    next_range(nb_, ne_);
                    for (int rr = 0; rr < nruns; ++rr) {
                        const uint32_t b = nb_, e = ne_;
                        if (rr + 1 < nruns) next_range(nb_, ne_);
                        uint32_t i = b;
                        for (; i + 2 <= e; i += 2) {
                            const float4* rec = A.photons + A.ps * (size_t)i;
                            const float4 a0 = __ldg(rec), a1 = __ldg(rec + A.p1), c0 = __ldg(rec + A.ps), c1 = __ldg(rec + A.ps + A.p1);
                            visit(a0, a1);
                            visit(c0, c1);
                        }
                        if (i < e) {
                            const float4* rec = A.photons + A.ps * (size_t)i;
                            visit(__ldg(rec), __ldg(rec + A.p1));
                        }
                    }
                    // ---- composite front to back
#pragma unroll
                    for (int j = 0; j < GATHER_S; ++j) {
                        if (j < nb && live) {
                            float4 c = sample_tf_rgba(s_tf, A.tf_width, ftfw, v[j]);
                            if (c.w > 0.0f) {
                                float Ts = cpm_expf(-(c.w * P.sigma_scale) * P.step);
                                float wgt = T * (1.0f - Ts);
                                lr = fmaf(wgt * c.x, er[j] * s, lr);
                                lg = fmaf(wgt * c.y, eg[j] * s, lg);
                                lb = fmaf(wgt * c.z, eb[j] * s, lb);
                                T *= Ts;
                                if (T < 1e-4f) live = false;
                            }
                        }
                    }
                }
                k += GATHER_S;
            }
        }
        A.image[(size_t)py * P.width + px] = make_float4(lr, lg, lb, 1.0f - T);
    }
}

// irradiance at arbitrary points (the per-voxel formulation of the reference's disabled
// photonsToLightVolumeKernel, with the Epanechnikov kernel): out[i] = s * sum, one thread per point
__global__ void __launch_bounds__(128) gather_points_kernel(const GatherArgs A, const float* __restrict__ pts, int n,
                                                            float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float er = 0.f, eg = 0.f, eb = 0.f;
    gather_point(A, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], er, eg, eb);
    const float s = CPM_INV_4PI_F * A.p.scale;
    out[3 * i] = er * s;
    out[3 * i + 1] = eg * s;
    out[3 * i + 2] = eb * s;
}

}  // namespace

static int fill_gather_args(cpm_ctx* ctx, GatherArgs& a, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                            const cpm_gather_params* params, const float* sorted_photons, const uint32_t* cell_start,
                            const uint32_t* cell_end) {
    CPM_REQUIRE(ctx, params && sorted_photons && cell_start && cell_end, "null argument");
    CPM_REQUIRE(ctx, params->radius > 0.0f, "radius must be positive");
    CPM_REQUIRE(ctx, params->grid_dims[0] > 0 && params->grid_dims[1] > 0 && params->grid_dims[2] > 0, "grid dims must be positive");
    a.p = *params;
    if (vol) a.vol = make_view(vol);
    a.tf = (const float4*)tf_rgba;
    a.tf_width = tf_width;
    a.photons = (const float4*)sorted_photons;
    CPM_REQUIRE(ctx, params->planar_records >= 0, "negative planar_records");
    a.ps = params->planar_records > 0 ? 1 : 2;
    a.p1 = params->planar_records > 0 ? (size_t)params->planar_records : 1;
    a.cell_start = cell_start;
    a.cell_end = cell_end;
    a.image = nullptr;
    memset(&a.bound, 0, sizeof(a.bound));
    return CPM_OK;
}

template <int FMT, int LAYOUT>
static int launch_gather(cpm_ctx* ctx, const GatherArgs& a) {
    size_t smem = (size_t)a.tf_width * sizeof(float4);
    if (smem > 48 * 1024)
        CPM_CUDA(ctx, cudaFuncSetAttribute(gather_kernel<FMT, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent warps: tiles are handed out through a counter in the context scratch
    void* scratch;
    int rc = cpm_scratch(ctx, 4096, &scratch);
    if (rc != CPM_OK) return rc;
    unsigned* counter = (unsigned*)((char*)scratch + 3072);
    CPM_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    int tiles = ((a.p.width + 7) / 8) * ((a.p.height + 3) / 4);
    int per_sm = 0;
    CPM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_kernel<FMT, LAYOUT>, 128, smem));
    unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * std::max(per_sm, 1), (long long)cpm_div_up(tiles, 4));
    CPM_LAUNCH(ctx, (gather_kernel<FMT, LAYOUT>), grid, 128, smem, a, counter);
    return CPM_OK;
}

extern "C" {

int cpm_photon_cell_keys(cpm_ctx* ctx, const float* photons, size_t n_records, const int grid_dims[3], uint32_t* keys,
                         uint32_t* ids) {
    if (!ctx) return CPM_E_INVALID;
    if (n_records == 0) return CPM_OK;
    CPM_REQUIRE(ctx, photons && grid_dims && keys, "null argument");
    CPM_REQUIRE(ctx, grid_dims[0] > 0 && grid_dims[1] > 0 && grid_dims[2] > 0, "grid dims must be positive");
    CPM_REQUIRE(ctx, (double)grid_dims[0] * grid_dims[1] * grid_dims[2] < 4294967295.0, "too many cells for 32-bit keys");
    CPM_LAUNCH(ctx, photon_cell_keys_kernel, cpm_div_up(n_records, 256), 256, 0, (const float4*)photons, n_records,
               grid_dims[0], grid_dims[1], grid_dims[2], keys, ids);
    return CPM_OK;
}

int cpm_reorder_photons(cpm_ctx* ctx, const float* photons, const uint32_t* ids, size_t n, float* out) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, photons && ids && out, "null argument");
    CPM_REQUIRE(ctx, photons != out, "in-place reorder is not supported");
    CPM_LAUNCH(ctx, reorder_kernel, cpm_div_up(2 * n, 256), 256, 0, (const float4*)photons, ids, n, (float4*)out);
    return CPM_OK;
}

int cpm_reorder_photons_planar(cpm_ctx* ctx, const float* photons, const uint32_t* ids, size_t n, float* out) {
    if (!ctx) return CPM_E_INVALID;
    if (n == 0) return CPM_OK;
    CPM_REQUIRE(ctx, photons && ids && out, "null argument");
    CPM_REQUIRE(ctx, photons != out, "in-place reorder is not supported");
    CPM_REQUIRE(ctx, n <= 2147483647u, "too many records for cpm_gather_params::planar_records");
    CPM_LAUNCH(ctx, reorder_planar_kernel, cpm_div_up(2 * n, 256), 256, 0, (const float4*)photons, ids, n, (float4*)out);
    return CPM_OK;
}

int cpm_gather_raymarch(cpm_ctx* ctx, const cpm_volume* vol, const float* tf_rgba, int tf_width,
                        const cpm_gather_params* params, const float* sorted_photons, const uint32_t* cell_start,
                        const uint32_t* cell_end, float* image) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, vol && tf_rgba && image, "null argument");
    CPM_REQUIRE(ctx, tf_width >= 1 && tf_width <= 8192, "tf_width out of range");
    GatherArgs a;
    int rc = fill_gather_args(ctx, a, vol, tf_rgba, tf_width, params, sorted_photons, cell_start, cell_end);
    if (rc != CPM_OK) return rc;
    CPM_REQUIRE(ctx, params->width > 0 && params->height > 0 && params->step > 0.0f, "bad image size / step");
    CPM_REQUIRE(ctx, params->strip_first >= 0 && params->strip_stride >= 0, "negative strip_first / strip_stride");
    a.image = (float4*)image;
    CPM_REQUIRE(ctx, make_bound_grid(a.bound, params->opacity_bound, vol->dims, params->bound_cell_log2),
                "bound_cell_log2 must be in 0..8 and the bound grid smaller than 2^31 cells");
#define CPM_DISPATCH(F)                                                                       \
    return vol->layout == CPM_VOLUME_TEXTURE ? launch_gather<F, CPM_VOLUME_TEXTURE>(ctx, a)  \
                                             : launch_gather<F, CPM_VOLUME_LINEAR>(ctx, a);
    switch (vol->format) {
        case CPM_FMT_U8: CPM_DISPATCH(CPM_FMT_U8)
        case CPM_FMT_U16: CPM_DISPATCH(CPM_FMT_U16)
        default: CPM_DISPATCH(CPM_FMT_F32)
    }
#undef CPM_DISPATCH
}

int cpm_gather_points(cpm_ctx* ctx, const cpm_gather_params* params, const float* sorted_photons,
                      const uint32_t* cell_start, const uint32_t* cell_end, const float* points, int n_points,
                      float* irradiance) {
    if (!ctx) return CPM_E_INVALID;
    if (n_points == 0) return CPM_OK;
    CPM_REQUIRE(ctx, points && irradiance && n_points > 0, "null argument");
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    int rc = fill_gather_args(ctx, a, nullptr, nullptr, 0, params, sorted_photons, cell_start, cell_end);
    if (rc != CPM_OK) return rc;
    CPM_LAUNCH(ctx, gather_points_kernel, cpm_div_up(n_points, 128), 128, 0, a, points, n_points, irradiance);
    return CPM_OK;
}

}  // extern "C"
