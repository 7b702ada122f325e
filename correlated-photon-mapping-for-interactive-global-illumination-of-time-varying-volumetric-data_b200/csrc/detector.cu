// detector.cu -- temporal-correlation pass: which stored photon paths cross changed bricks
// (north-star subsystem 4, detection half).
//
// Replaces photonRecomputationDetectorKernel / ...EqualImportanceKernel
// (ppm/cl/photonrecomputationdetector.cl:92-157, :160-194) and the DDA library they use
// (ugc/cl/uniformgrid/uniformgrid.cl:38-69, :147-197, OPTIMIZE_STEP_FOR_SIMD branch).
//
// B200 notes: one thread per photon path; the streams (light sample 32 B, intersection 8 B,
// photon 32 B per interaction, key 4 B RMW) are read with 16 B vector loads; the importance grid
// (1-8 MB) is read through the read-only path and stays L2 resident.  The DDA loop is
// divergent by nature (3*dim/8 cells at most).
#include <string.h>

#include "sampling.cuh"

namespace {

#define CPM_FLT_MAX 3.402823466e+38f
#ifndef CPM_DETECT_PTX_LOOP
#define CPM_DETECT_PTX_LOOP 1  // 0: the reference-shaped loop for every segment (A/B builds)
#endif

struct Mat4 {
    float m[16];
};

__device__ __forceinline__ float3_ transform_point(const Mat4& M, float3_ p) {
    return {fmaf(M.m[8], p.z, fmaf(M.m[4], p.y, fmaf(M.m[0], p.x, M.m[12]))),
            fmaf(M.m[9], p.z, fmaf(M.m[5], p.y, fmaf(M.m[1], p.x, M.m[13]))),
            fmaf(M.m[10], p.z, fmaf(M.m[6], p.y, fmaf(M.m[2], p.x, M.m[14])))};
}

struct GridArgs {
    const float* grid;
    int dims[3];
    float cell[3];
};

__device__ float uniform_grid_importance(const GridArgs& G, float3_ x1, float3_ x2) {
    float a1[3] = {x1.x, x1.y, x1.z}, a2[3] = {x2.x, x2.y, x2.z};
    float dt[3], deltatx[3];
    int cell[3], cell_end[3], di[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float mx = (float)(G.dims[k] - 1);
        float cf = cpm_clamp(floorf(a1[k] / G.cell[k]), 0.0f, mx);
        cell[k] = (int)cf;
        cell_end[k] = (int)cpm_clamp(truncf(a2[k] / G.cell[k]), 0.0f, mx);
        di[k] = (a1[k] < a2[k]) ? 1 : ((a1[k] > a2[k]) ? -1 : 0);
        float inv_abs = 1.0f / fabsf(a2[k] - a1[k]);
        float minx = G.cell[k] * cf;
        float maxx = minx + G.cell[k];
        dt[k] = ((a1[k] > a2[k]) ? (a1[k] - minx) : (maxx - a1[k])) * inv_abs;
        deltatx[k] = G.cell[k] * inv_abs;
    }
    // The walk keeps one linear cell index and, per axis, the number of cells left to the end cell.
    // "selected axis already in its end cell" (uniformgrid.cl:147-197) is rem == 0: a cell only ever moves
    // toward its end cell (di = sign(x2 - x1), end cell on that side after clamping), and an axis with
    // di == 0 starts in its end cell.
    int rem0 = abs(cell_end[0] - cell[0]), rem1 = abs(cell_end[1] - cell[1]), rem2 = abs(cell_end[2] - cell[2]);
    const int sy = G.dims[0], sz = G.dims[0] * G.dims[1];
    const int s0 = di[0], s1 = di[1] * sy, s2 = di[2] * sz;
    int idx = cell[0] + cell[1] * sy + cell[2] * sz;
    float d0 = dt[0], d1 = dt[1], d2 = dt[2];
    float importance = 0.0f;
    const float del0 = deltatx[0], del1 = deltatx[1], del2 = deltatx[2];
    const float* __restrict__ grid = G.grid;
    if (CPM_DETECT_PTX_LOOP && d0 == d0 && d1 == d1 && d2 == d2) {
        // No NaN among the three parameters (they arise only as 0 * inf at set-up: a segment that lies in a cell face and
        // does not move along that axis): the reference's comparisons (uniformgrid.cl:163-180) then select the smallest
        // parameter with ties resolved x before y before z, i.e. dt1 = min(d0, d1, d2), x iff d0 == dt1, else y iff
        // d1 == dt1.  The loop is branch-free apart from its exit: lanes of a warp step along different axes in the same
        // iteration, and with a branch per axis every iteration would run all three paths.
        // Written in PTX: from the C++ form of this loop the compiler builds 34 instructions per step (predicate
        // shuffles, copies, two-instruction conditional decrements); this is 24 per step (two steps per trip, so that the previous parameter needs no copy).
        asm volatile(
            "{\n\t"
            ".reg .pred ax, ay, axy, go;\n\t"
            ".reg .f32 dta, dtb, c, w, val;\n\t"
            ".reg .s32 r, st;\n\t"
            ".reg .u64 addr;\n\t"
            "mov.f32 dta, 0f00000000;\n"
            "DDA_STEP:\n\t"
            "mad.wide.s32 addr, %0, 4, %8;\n\t"
            "ld.global.nc.f32 val, [addr];\n\t"
            "min.f32 dtb, %1, %2;\n\t"
            "min.f32 dtb, dtb, %3;\n\t"
            "setp.eq.f32 ax, %1, dtb;\n\t"
            "setp.eq.and.f32 ay, %2, dtb, !ax;\n\t"
            "selp.s32 r, %5, %6, ay;\n\t"
            "selp.s32 r, %4, r, ax;\n\t"
            "setp.ne.s32 go, r, 0;\n\t"
            "min.f32 c, dtb, 0f3F800000;\n\t"
            "sub.rn.f32 w, c, dta;\n\t"
            "mul.rn.f32 w, val, w;\n\t"
            "add.rn.f32 %7, %7, w;\n\t"
            "@!go bra DDA_DONE;\n\t"
            "selp.s32 st, %13, %14, ay;\n\t"
            "selp.s32 st, %12, st, ax;\n\t"
            "add.s32 %0, %0, st;\n\t"
            "or.pred axy, ax, ay;\n\t"
            "@ax add.rn.f32 %1, %1, %9;\n\t"
            "@ay add.rn.f32 %2, %2, %10;\n\t"
            "@!axy add.rn.f32 %3, %3, %11;\n\t"
            "@ax sub.s32 %4, %4, 1;\n\t"
            "@ay sub.s32 %5, %5, 1;\n\t"
            "@!axy sub.s32 %6, %6, 1;\n\t"
            "mad.wide.s32 addr, %0, 4, %8;\n\t"
            "ld.global.nc.f32 val, [addr];\n\t"
            "min.f32 dta, %1, %2;\n\t"
            "min.f32 dta, dta, %3;\n\t"
            "setp.eq.f32 ax, %1, dta;\n\t"
            "setp.eq.and.f32 ay, %2, dta, !ax;\n\t"
            "selp.s32 r, %5, %6, ay;\n\t"
            "selp.s32 r, %4, r, ax;\n\t"
            "setp.ne.s32 go, r, 0;\n\t"
            "min.f32 c, dta, 0f3F800000;\n\t"
            "sub.rn.f32 w, c, dtb;\n\t"
            "mul.rn.f32 w, val, w;\n\t"
            "add.rn.f32 %7, %7, w;\n\t"
            "@!go bra DDA_DONE;\n\t"
            "selp.s32 st, %13, %14, ay;\n\t"
            "selp.s32 st, %12, st, ax;\n\t"
            "add.s32 %0, %0, st;\n\t"
            "or.pred axy, ax, ay;\n\t"
            "@ax add.rn.f32 %1, %1, %9;\n\t"
            "@ay add.rn.f32 %2, %2, %10;\n\t"
            "@!axy add.rn.f32 %3, %3, %11;\n\t"
            "@ax sub.s32 %4, %4, 1;\n\t"
            "@ay sub.s32 %5, %5, 1;\n\t"
            "@!axy sub.s32 %6, %6, 1;\n\t"
            "bra DDA_STEP;\n"
            "DDA_DONE:\n\t"
            "}"
            : "+r"(idx), "+f"(d0), "+f"(d1), "+f"(d2), "+r"(rem0), "+r"(rem1), "+r"(rem2), "+f"(importance)
            : "l"(grid), "f"(del0), "f"(del1), "f"(del2), "r"(s0), "r"(s1), "r"(s2));
    } else {
        bool go = true;
        float dt1 = 0.0f;
        while (go) {
            float val = __ldg(grid + idx);
            float dt0 = dt1;
            bool ax = (d0 <= d1 && d0 <= d2);
            bool ay = !ax && (d0 > d1 && d1 <= d2);
            dt1 = ax ? d0 : (ay ? d1 : d2);
            int rsel = ax ? rem0 : (ay ? rem1 : rem2);
            if (rsel == 0) {
                go = false;
            } else if (ax) {
                d0 += del0; idx += s0; --rem0;
            } else if (ay) {
                d1 += del1; idx += s1; --rem1;
            } else {
                d2 += del2; idx += s2; --rem2;
            }
            importance += val * (cpm_fmin(1.0f, dt1) - dt0);
        }
    }
    float dx = x2.x - x1.x, dy = x2.y - x1.y, dz = x2.z - x1.z;
    float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    return importance * len;
}

__device__ __forceinline__ uint32_t convert_uint_sat_rtp(float v) {
    if (!(v > 0.0f)) return 0u;
    float c = ceilf(v);
    if (c >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)c;
}

struct DetectArgs {
    GridArgs grid;
    Mat4 tex2idx;
    const float4* photons;
    int photon_offset;
    const float4* light_samples;
    const float2* isect;
    int n_light_samples;
    int max_interactions;
    int total_photons;
    uint32_t* importances;
    int equal_importance, percentage, iteration, fix_exit;
};

__global__ void __launch_bounds__(128) detect_kernel(const DetectArgs A) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.n_light_samples) return;
    float imp = 0.0f;
    if (A.equal_importance) {
        int pid = A.photon_offset + tid;
        int div = (A.percentage > 0 && A.percentage <= 100) ? 100 / A.percentage : 1;
        if ((pid + A.iteration) % div == 0) imp = 1.0f;
    } else {
        float4 l0 = A.light_samples[2 * (size_t)tid], l1 = A.light_samples[2 * (size_t)tid + 1];
        float3_ origin = {l0.x, l0.y, l0.z};
        float3_ dir = decode_direction(l1.z, l1.w);
        float2 ip = A.isect[tid];
        float tStart = ip.x, tEnd = ip.y;
        if (tStart < tEnd) {
            float3_ entry = ray_at(origin, tStart, dir);
            for (int k = 0; k < A.max_interactions; ++k) {
                size_t pid = (size_t)A.photon_offset + (size_t)k * A.total_photons + tid;
                float4 p0 = A.photons[2 * pid], p1 = A.photons[2 * pid + 1];
                float3_ exit = {p0.x, p0.y, p0.z};
                if (p0.x == CPM_FLT_MAX || p0.y == CPM_FLT_MAX || p0.z == CPM_FLT_MAX) {
                    if (k == 0) {
                        if (A.fix_exit)
                            exit = ray_at(origin, tEnd, dir);
                        else
                            exit = {tEnd * dir.x, tEnd * dir.y, tEnd * dir.z};  // reference quirk (:128)
                    } else if (entry.x == CPM_FLT_MAX || entry.y == CPM_FLT_MAX || entry.z == CPM_FLT_MAX) {
                        break;
                    } else {
                        const float bmin[3] = {0.f, 0.f, 0.f}, bmax[3] = {1.f, 1.f, 1.f};
                        float t0 = 0.0f, t1 = CPM_FLT_MAX;
                        float3_ pd = decode_direction(p1.z, p1.w);
                        if (p0.w != CPM_FLT_MAX && ray_box(bmin, bmax, entry, pd, t0, t1)) {
                            if (A.fix_exit) {
                                exit = ray_at(entry, t1, pd);   // the evidently intended exit point
                            } else {
                                // reference (:137): `exit += photonDirection*tEnd` on exit == (FLT_MAX, FLT_MAX, FLT_MAX):
                                // x2 becomes +inf on every axis, the DDA sums val * 0 and scales by length(x2 - x1) == inf,
                                // i.e. NaN, and convert_uint_sat_rtp(NaN) == 0: the photon is not flagged (see orc_grid.c)
                                imp = __uint_as_float(0x7fc00000u);
                                break;
                            }
                        } else {
                            break;
                        }
                    }
                }
                float3_ x1 = transform_point(A.tex2idx, entry), x2 = transform_point(A.tex2idx, exit);
                x1 = {x1.x + 0.5f, x1.y + 0.5f, x1.z + 0.5f};
                x2 = {x2.x + 0.5f, x2.y + 0.5f, x2.z + 0.5f};
                imp += uniform_grid_importance(A.grid, x1, x2);
                entry = {p0.x, p0.y, p0.z};
            }
        }
    }
    uint32_t v = convert_uint_sat_rtp(100.0f * imp);
    if (v > 2147483647u) v = 2147483647u;
    A.importances[A.photon_offset + tid] -= v;
}

}  // namespace

extern "C" int cpm_detect_invalid(cpm_ctx* ctx, const float* importance_grid, const int grid_dims[3],
                                  const float cell_size[3], const float texture_to_index[16], const float* photons,
                                  int photon_offset, const float* light_samples, const float* intersections,
                                  int n_light_samples, int max_interactions, int total_photons, uint32_t* importances,
                                  int equal_importance, int percentage, int iteration, uint32_t flags) {
    if (!ctx) return CPM_E_INVALID;
    CPM_REQUIRE(ctx, n_light_samples >= 0, "negative n_light_samples");
    if (n_light_samples == 0) return CPM_OK;
    CPM_REQUIRE(ctx, importances != nullptr, "importances is NULL");
    CPM_REQUIRE(ctx, equal_importance || (importance_grid && grid_dims && cell_size && texture_to_index && photons &&
                                          light_samples && intersections),
                "null argument");
    CPM_REQUIRE(ctx, max_interactions >= 1 && photon_offset >= 0 && total_photons >= photon_offset + n_light_samples,
                "inconsistent photon counts");
    DetectArgs a;
    memset(&a, 0, sizeof(a));
    if (!equal_importance) {
        CPM_REQUIRE(ctx, grid_dims[0] > 0 && grid_dims[1] > 0 && grid_dims[2] > 0, "grid dims must be positive");
        a.grid.grid = importance_grid;
        for (int k = 0; k < 3; ++k) {
            a.grid.dims[k] = grid_dims[k];
            a.grid.cell[k] = cell_size[k];
        }
        for (int k = 0; k < 16; ++k) a.tex2idx.m[k] = texture_to_index[k];
    }
    a.photons = (const float4*)photons;
    a.photon_offset = photon_offset;
    a.light_samples = (const float4*)light_samples;
    a.isect = (const float2*)intersections;
    a.n_light_samples = n_light_samples;
    a.max_interactions = max_interactions;
    a.total_photons = total_photons;
    a.importances = importances;
    a.equal_importance = equal_importance;
    a.percentage = percentage;
    a.iteration = iteration;
    a.fix_exit = (flags & CPM_DETECT_FIX_EXIT) ? 1 : 0;
    CPM_LAUNCH(ctx, detect_kernel, cpm_div_up(n_light_samples, 128), 128, 0, a);
    return CPM_OK;
}
