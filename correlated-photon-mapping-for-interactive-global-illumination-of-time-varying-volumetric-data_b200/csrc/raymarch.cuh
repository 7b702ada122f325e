// raymarch.cuh -- pieces shared by the view ray marchers (gather.cu: photon gather; raycast.cu: light volume).
#pragma once
#include "sampling.cuh"

// read_imagef(tf, smpNormClampEdgeLinear, (v, 0.5)) on a width x 1 image: all four channels
__device__ __forceinline__ float4 sample_tf_rgba(const float4* __restrict__ tf, int width, float fwidth, float v) {
    float u = fmaf(v, fwidth, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fwidth - 1.0f);
    int i1 = min(i0 + 1, width - 1);
    i0 = max(i0, 0);
    float4 p = tf[i0], q = tf[i1];
    return make_float4(lerpf(p.x, q.x, a), lerpf(p.y, q.y, a), lerpf(p.z, q.z, a), lerpf(p.w, q.w, a));
}

__device__ __forceinline__ float sample_tf_alpha4(const float4* __restrict__ tf, int width, float fwidth, float v) {
    float u = fmaf(v, fwidth, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fwidth - 1.0f);
    int i1 = min(i0 + 1, width - 1);
    i0 = max(i0, 0);
    return lerpf(tf[i0].w, tf[i1].w, a);
}

// Leave all-transparent cells (opacity bound <= 0) in one jump each.  k is the index of the next sample,
// t_k = t0 + (k + 1/2) step; returns the index of the first sample that lies in a cell with a positive bound
// (or past t1) and leaves its parameter in t.  The bound grid pads cells by one voxel, i.e. two samples at the
// usual half-voxel step, which covers the rounding of the exit point.
__device__ __forceinline__ int skip_transparent(const BoundGrid& B, const CellRay& R, float ix, float iy, float iz,
                                                float t0, float t1, float step, float inv_step, int k, float& t) {
    while (t < t1) {
        float wx = fminf(fmaxf(fmaf(t, R.dx, R.ox), 0.0f), B.mx[0]);
        float wy = fminf(fmaxf(fmaf(t, R.dy, R.oy), 0.0f), B.mx[1]);
        float wz = fminf(fmaxf(fmaf(t, R.dz, R.oz), 0.0f), B.mx[2]);
        float cx = floorf(wx), cy = floorf(wy), cz = floorf(wz);
        int ci = (int)cx + (int)cy * B.nx + (int)cz * B.nxy;
        if (__ldg(B.g + ci) > 0.0f) break;   // <= 0: transparent (negative: annotated with its clearance)
        float ex = ((R.dx > 0.0f ? cx + 1.0f : cx) - R.ox) * ix;
        float ey = ((R.dy > 0.0f ? cy + 1.0f : cy) - R.oy) * iy;
        float ez = ((R.dz > 0.0f ? cz + 1.0f : cz) - R.oz) * iz;
        float te = fminf(fminf(R.dx != 0.0f ? ex : CPM_FLT_MAX_, R.dy != 0.0f ? ey : CPM_FLT_MAX_),
                         R.dz != 0.0f ? ez : CPM_FLT_MAX_);
        // first sample at or past the exit; at least one step forward
        float kf = ceilf(fmaf(te - t0, inv_step, -0.5f));
        int kn = (kf < 1.0e9f) ? (int)kf : 1000000000;
        k = max(k + 1, kn);
        t = fmaf((float)k + 0.5f, step, t0);
    }
    return k;
}

// camera row of local image row py when the call renders strips of 4 rows (cpm_gather_params::strip_first/stride)
__device__ __forceinline__ int camera_row(const cpm_gather_params& P, int py) {
    const int stride = P.strip_stride > 1 ? P.strip_stride : 1;
    return 4 * (P.strip_first + (py >> 2) * stride) + (py & 3);
}

// camera ray of pixel (px, py): normalize(dir00 + (px + 1/2) du + (py + 1/2) dv)
__device__ __forceinline__ float3_ camera_ray(const cpm_gather_params& P, int px, int py) {
    float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    float dx = fmaf(fy, P.cam_dv[0], fmaf(fx, P.cam_du[0], P.cam_dir00[0]));
    float dy = fmaf(fy, P.cam_dv[1], fmaf(fx, P.cam_du[1], P.cam_dir00[1]));
    float dz = fmaf(fy, P.cam_dv[2], fmaf(fx, P.cam_du[2], P.cam_dir00[2]));
    float inv = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    return {dx * inv, dy * inv, dz * inv};
}
