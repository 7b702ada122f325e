// sampling.cuh -- exact (bit-reproducible) volume / transfer-function sampling and the small
// geometric helpers of the photon path, device side.
//
// Restates what the reference pulls from un-vendored Inviwo OpenCL headers (samplers.cl,
// transformations.cl, intersection/rayboxintersection.cl, shading/shading.cl), following the
// OpenCL 1.2 specification section 8.2 for CLK_NORMALIZED_COORDS_TRUE |
// CLK_ADDRESS_CLAMP_TO_EDGE | CLK_FILTER_LINEAR.  The operation order here is the contract
// that oracle/cpm_oracle.c follows op for op; do not reorder without changing both.
#pragma once
#include <string.h>

#include "common.cuh"

struct float3_ {
    float x, y, z;
};

// `origin + t*direction` as the reference's kernels write it (ppm/cl/transmittance.cl:136, photontracer.cl:165,
// photonrecomputationdetector.cl:122 ...): a rounded product and a rounded sum per component, never a fused
// multiply-add -- what the reference's .cl files give under strict IEEE evaluation (oracle/_ref/libcl_ref.so), which
// the oracle and these kernels match bit for bit.  The intrinsics keep the two roundings whatever -fmad says.
__device__ __forceinline__ float3_ ray_at(const float3_& o, float t, const float3_& d) {
    return {__fadd_rn(o.x, __fmul_rn(t, d.x)), __fadd_rn(o.y, __fmul_rn(t, d.y)), __fadd_rn(o.z, __fmul_rn(t, d.z))};
}
// `t += -native_log(u) * invTauMaxSampleBaseInterval` (ppm/cl/transmittance.cl:135), as written
__device__ __forceinline__ float advance_t(float t, float log_u, float inv) { return __fadd_rn(t, __fmul_rn(-log_u, inv)); }

struct VolumeView {
    const void* lin;
    cudaTextureObject_t tex;
    int nx, ny, nz;
    float fx, fy, fz;  // dims as float
    float scale, offset;
};

static inline VolumeView make_view(const cpm_volume* v) {
    VolumeView w;
    w.lin = v->linear;
    w.tex = v->tex;
    w.nx = v->dims[0];
    w.ny = v->dims[1];
    w.nz = v->dims[2];
    w.fx = (float)v->dims[0];
    w.fy = (float)v->dims[1];
    w.fz = (float)v->dims[2];
    w.scale = v->scale;
    w.offset = v->offset;
    return w;
}

// v/255 and v/65535, correctly rounded, as one FMUL + one FFMA.  Equality with the IEEE
// division for every u8/u16 input is checked in tests (tools/fit_detmath.py documents why:
// the quotients are never within 2^-40 of a rounding boundary).
__device__ __forceinline__ float unorm8(float v) { return fmaf(v, 0x1.010102p-8f, v * -0x1.fdfdfep-33f); }
__device__ __forceinline__ float unorm16(float v) { return fmaf(v, 0x1.0001p-16f, v * 0x1.0001p-48f); }

template <int FMT>
__device__ __forceinline__ float load_linear(const void* base, size_t idx) {
    if (FMT == CPM_FMT_U8) return unorm8((float)__ldg((const unsigned char*)base + idx));
    if (FMT == CPM_FMT_U16) return unorm16((float)__ldg((const unsigned short*)base + idx));
    return __ldg((const float*)base + idx);
}

// Four texels of the bilinear footprint {i0,i0+1} x {j0,j0+1} of layer k, unfiltered, with
// clamp-to-edge addressing done by the texture unit.  The coordinate (i0+1, j0+1) sits exactly
// in the middle of the footprint, far from the 1/256 fixed-point selection boundaries.
// Result order (PTX tld4): x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0).
template <int FMT>
__device__ __forceinline__ float4 gather_layer(cudaTextureObject_t tex, int k, float cx, float cy) {
    float4 r;
    if (FMT == CPM_FMT_F32) {
        asm("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
            : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
            : "l"(tex), "r"(k), "f"(cx), "f"(cy));
    } else {
        unsigned a, b, c, d;
        asm("tld4.r.a2d.v4.u32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
            : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
            : "l"(tex), "r"(k), "f"(cx), "f"(cy));
        if (FMT == CPM_FMT_U8) {
            r.x = unorm8((float)a); r.y = unorm8((float)b); r.z = unorm8((float)c); r.w = unorm8((float)d);
        } else {
            r.x = unorm16((float)a); r.y = unorm16((float)b); r.z = unorm16((float)c); r.w = unorm16((float)d);
        }
    }
    return r;
}

__device__ __forceinline__ float lerpf(float p, float q, float a) { return fmaf(a, q - p, p); }

// getNormalizedVoxel(volume, params, pos).x  -- trilinear, normalised coordinates -- in two steps so that
// the tracer can request the taps of its NEXT sample before it consumes the current one:
// fetch_taps issues the loads (2 x tld4 or 8 x LDG) and keeps the blend weights, blend_taps does the
// fp32 arithmetic of OpenCL 1.2 section 8.2 in the order the oracle follows.
struct Taps {
    float4 g0, g1;   // TEXTURE: tld4 results of layers k0, k1 (raw texel values); LINEAR: t000,t100,t010,t110 / t001,...
    float a, b, c;
};

template <int FMT>
__device__ __forceinline__ float4 gather_layer_raw(cudaTextureObject_t tex, int k, float cx, float cy) {
    float4 r;
    if (FMT == CPM_FMT_F32) {
        asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "l"(tex), "r"(k), "f"(cx), "f"(cy));
    } else {
        unsigned a, b, c, d;
        asm volatile("tld4.r.a2d.v4.u32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                     : "l"(tex), "r"(k), "f"(cx), "f"(cy));
        r.x = __uint_as_float(a); r.y = __uint_as_float(b); r.z = __uint_as_float(c); r.w = __uint_as_float(d);
    }
    return r;
}

template <int FMT>
__device__ __forceinline__ float raw_to_float(float raw) {
    if (FMT == CPM_FMT_U8) return unorm8((float)__float_as_uint(raw));
    if (FMT == CPM_FMT_U16) return unorm16((float)__float_as_uint(raw));
    return raw;
}

template <int FMT>
__device__ __forceinline__ float load_linear_raw(const void* base, size_t idx) {
    if (FMT == CPM_FMT_U8) return __uint_as_float((unsigned)__ldg((const unsigned char*)base + idx));
    if (FMT == CPM_FMT_U16) return __uint_as_float((unsigned)__ldg((const unsigned short*)base + idx));
    return __ldg((const float*)base + idx);
}

template <int FMT, int LAYOUT>
__device__ __forceinline__ Taps fetch_taps(const VolumeView& V, float px, float py, float pz) {
    Taps T;
    float u = fmaf(px, V.fx, -0.5f);
    float v = fmaf(py, V.fy, -0.5f);
    float w = fmaf(pz, V.fz, -0.5f);
    float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    T.a = u - fu; T.b = v - fv; T.c = w - fw;
    // NaN/inf safe conversion: clamp in float first (NaN -> -1)
    int i0 = (int)cpm_clamp(fu, -1.0f, V.fx - 1.0f);
    int j0 = (int)cpm_clamp(fv, -1.0f, V.fy - 1.0f);
    int k0 = (int)cpm_clamp(fw, -1.0f, V.fz - 1.0f);
    int k1 = min(k0 + 1, V.nz - 1);
    k0 = max(k0, 0);
    if (LAYOUT == CPM_VOLUME_TEXTURE) {
        // (i0+1, j0+1) is the centre of the 2x2 footprint; PTX tld4 order: x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0)
        float cx = (float)(i0 + 1), cy = (float)(j0 + 1);
        T.g0 = gather_layer_raw<FMT>(V.tex, k0, cx, cy);
        T.g1 = gather_layer_raw<FMT>(V.tex, k1, cx, cy);
    } else {
        int i1 = min(i0 + 1, V.nx - 1);
        int j1 = min(j0 + 1, V.ny - 1);
        i0 = max(i0, 0);
        j0 = max(j0, 0);
        size_t sy = (size_t)V.nx, sz = (size_t)V.nx * V.ny;
        size_t b00 = (size_t)k0 * sz + (size_t)j0 * sy, b10 = (size_t)k0 * sz + (size_t)j1 * sy;
        size_t b01 = (size_t)k1 * sz + (size_t)j0 * sy, b11 = (size_t)k1 * sz + (size_t)j1 * sy;
        // same slots as the tld4 result: x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0)
        T.g0.w = load_linear_raw<FMT>(V.lin, b00 + i0); T.g0.z = load_linear_raw<FMT>(V.lin, b00 + i1);
        T.g0.x = load_linear_raw<FMT>(V.lin, b10 + i0); T.g0.y = load_linear_raw<FMT>(V.lin, b10 + i1);
        T.g1.w = load_linear_raw<FMT>(V.lin, b01 + i0); T.g1.z = load_linear_raw<FMT>(V.lin, b01 + i1);
        T.g1.x = load_linear_raw<FMT>(V.lin, b11 + i0); T.g1.y = load_linear_raw<FMT>(V.lin, b11 + i1);
    }
    return T;
}

template <int FMT>
__device__ __forceinline__ float blend_taps(const VolumeView& V, const Taps& T) {
    float t000 = raw_to_float<FMT>(T.g0.w), t100 = raw_to_float<FMT>(T.g0.z);
    float t010 = raw_to_float<FMT>(T.g0.x), t110 = raw_to_float<FMT>(T.g0.y);
    float t001 = raw_to_float<FMT>(T.g1.w), t101 = raw_to_float<FMT>(T.g1.z);
    float t011 = raw_to_float<FMT>(T.g1.x), t111 = raw_to_float<FMT>(T.g1.y);
    float x00 = lerpf(t000, t100, T.a), x10 = lerpf(t010, t110, T.a);
    float x01 = lerpf(t001, t101, T.a), x11 = lerpf(t011, t111, T.a);
    float y0 = lerpf(x00, x10, T.b), y1 = lerpf(x01, x11, T.b);
    float val = lerpf(y0, y1, T.c);
    return (val + V.offset) * V.scale;
}

template <int FMT, int LAYOUT>
__device__ __forceinline__ float sample_volume(const VolumeView& V, float px, float py, float pz) {
    Taps T = fetch_taps<FMT, LAYOUT>(V, px, py, pz);
    return blend_taps<FMT>(V, T);
}

// ---- per-cell opacity bound grid (bound.cu) ----------------------------------------------------------------
// Cell of a sample p, per axis: floor(p * fc + hc) = floor((p * dim + 0.5) / cell), clamped to [0, gd - 1].
struct BoundGrid {
    const float* g;   // null: no grid
    float fc[3], hc, mx[3];   // p -> cell coordinate p * fc + hc, last cell per axis (the view ray marchers: raymarch.cuh)
    float fn[3], hn[3];   // p -> NORMALISED cell coordinate p * fn + hn = (p * dim / cell + 0.5 / cell) / gd (the tracer)
    float gs[3];          // normalised -> cell: gd * (1 - 2^-22), so that 1.0 still lands in the last cell
    unsigned bias;    // (1 + nx + nxy) * 0x4B400000 mod 2^32: removes the float-bit biases of the three cell coordinates
    int nx, nxy;
};
static inline bool make_bound_grid(BoundGrid& B, const float* g, const int dims[3], int cell_log2) {
    memset(&B, 0, sizeof(B));
    if (!g) return true;
    if (cell_log2 < 0 || cell_log2 > 8) return false;
    const float cell = (float)(1 << cell_log2);
    int gd[3];
    for (int k = 0; k < 3; ++k) {
        gd[k] = (dims[k] >> cell_log2) + 1;
        B.fc[k] = (float)dims[k] / cell;   // exact: cell is a power of two
        B.mx[k] = (float)(gd[k] - 1);
        B.fn[k] = (float)dims[k] / cell / (float)gd[k];
        B.hn[k] = 0.5f / cell / (float)gd[k];
        B.gs[k] = (float)gd[k] * (1.0f - 0x1p-22f);
    }
    if ((double)gd[0] * gd[1] * gd[2] >= 2147483648.0) return false;
    B.g = g;
    B.hc = 0.5f / cell;
    B.nx = gd[0];
    B.nxy = gd[0] * gd[1];
    // indices are formed from raw float bits (cell + 0x4B400000 per axis) with wrapping 32-bit arithmetic
    B.bias = 0x4B400000u * (1u + (uint32_t)B.nx + (uint32_t)B.nxy);
    return true;
}
// Opacity bound of the cell that holds the trilinear footprint of the sample at parameter t of the ray
// w(t) = wo + t * wd, the ray in NORMALISED cell coordinates (cell / cells per axis; set up once per walk).  Two
// instructions per axis: an FFMA that saturates to [0, 1] (the clamp to the grid is the .SAT modifier, and NaN becomes
// 0), and an FFMA rounded down that scales to cells and adds 1.5 * 2^23, which leaves floor(cell) + 0x4B400000 in the
// bits -- no min / max, no conversion instructions; the three biases leave the index with one wrapping subtraction.
// The arithmetic differs from fetch_taps' i0 = floor(p * dim - 0.5) by a few 1e-4 voxels at most (the grid has at most
// a few hundred cells per axis, every step is one rounding); bound.cu pads every cell by one voxel, which absorbs it.
struct CellRay {
    float ox, oy, oz, dx, dy, dz;
};
// the ray in cell coordinates (raymarch.cuh: skip_transparent works in cells)
__device__ __forceinline__ CellRay cell_ray(const BoundGrid& A, float3_ o, float3_ d) {
    return {fmaf(o.x, A.fc[0], A.hc), fmaf(o.y, A.fc[1], A.hc), fmaf(o.z, A.fc[2], A.hc),
            d.x * A.fc[0], d.y * A.fc[1], d.z * A.fc[2]};
}
// the ray in normalised cell coordinates (bound_at)
__device__ __forceinline__ CellRay cell_ray_n(const BoundGrid& A, float3_ o, float3_ d) {
    return {fmaf(o.x, A.fn[0], A.hn[0]), fmaf(o.y, A.fn[1], A.hn[1]), fmaf(o.z, A.fn[2], A.hn[2]),
            d.x * A.fn[0], d.y * A.fn[1], d.z * A.fn[2]};
}
// `magic` = 12582912.0f held in a register the compiler cannot see through (tracer.cu: ScanRegs): an FFMA takes ONE operand
// that is not a register, and with the constant as an immediate every gs[] would cost a constant-bank load per test.
__device__ __forceinline__ unsigned cell_bits(float t, float wd, float wo, float gs, float magic) {
    return __float_as_uint(__fmaf_rd(__saturatef(fmaf(t, wd, wo)), gs, magic));
}
__device__ __forceinline__ float bound_at(const BoundGrid& A, const CellRay& R, float t, float magic) {
    unsigned bx = cell_bits(t, R.dx, R.ox, A.gs[0], magic);
    unsigned by = cell_bits(t, R.dy, R.oy, A.gs[1], magic);
    unsigned bz = cell_bits(t, R.dz, R.oz, A.gs[2], magic);
    return __ldg(A.g + (bx + by * (unsigned)A.nx + bz * (unsigned)A.nxy - A.bias));
}

// read_imagef(tf, smpNormClampEdgeLinear, (v, 0.5)).w on a width x 1 image: 1-D linear.
__device__ __forceinline__ float sample_tf_alpha(const float* __restrict__ alpha, int width, float fwidth, float v) {
    float u = fmaf(v, fwidth, -0.5f);
    float fu = floorf(u);
    float a = u - fu;
    int i0 = (int)cpm_clamp(fu, -1.0f, fwidth - 1.0f);
    int i1 = min(i0 + 1, width - 1);
    i0 = max(i0, 0);
    return lerpf(alpha[i0], alpha[i1], a);
}

#define CPM_FLT_MAX_ 3.402823466e+38f

// decodeDirection / encodeDirection (Inviwo transformations.cl; host twin
// ppm/photondata.cpp:100-117): theta = acos(clamp(z)), phi = atan2(y, x).
__device__ __forceinline__ float3_ decode_direction(float theta, float phi) {
    float st, ct, sp, cp;
    cpm_sincosf(theta, &st, &ct);
    cpm_sincosf(phi, &sp, &cp);
    return {st * cp, st * sp, ct};
}
__device__ __forceinline__ float2 encode_direction(float3_ d) {
    return make_float2(cpm_acosf(cpm_clamp(d.z, -1.0f, 1.0f)), cpm_atan2f(d.y, d.x));
}

// rayBoxIntersection (Inviwo intersection/rayboxintersection.cl): slab test, tightens [t0,t1].
__device__ __forceinline__ bool ray_box(const float* bmin, const float* bmax, float3_ o, float3_ d, float& t0, float& t1) {
    float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    float ax = (bmin[0] - o.x) * ix, bx = (bmax[0] - o.x) * ix;
    float ay = (bmin[1] - o.y) * iy, by = (bmax[1] - o.y) * iy;
    float az = (bmin[2] - o.z) * iz, bz = (bmax[2] - o.z) * iz;
    float n = cpm_fmax(cpm_fmax(cpm_fmin(ax, bx), cpm_fmin(ay, by)), cpm_fmin(az, bz));
    float f = cpm_fmin(cpm_fmin(cpm_fmax(ax, bx), cpm_fmax(ay, by)), cpm_fmax(az, bz));
    t0 = cpm_fmax(t0, n);
    t1 = cpm_fmin(t1, f);
    return t0 < t1;
}

// uniformSampleSphere (Inviwo shading/shadingmath.cl, pbrt form): z = 1-2u, phi = 2 pi v.
__device__ __forceinline__ float3_ uniform_sample_sphere(float u1, float u2) {
    float z = fmaf(-2.0f, u1, 1.0f);
    float r = sqrtf(cpm_fmax(0.0f, fmaf(-z, z, 1.0f)));
    float s, c;
    cpm_sincosf(CPM_2PI_F * u2, &s, &c);
    return {r * c, r * s, z};
}

// Henyey-Greenstein sampling around the incoming direction `wi` (restated; the reference's
// version lives in un-vendored shading.cl -- parity unpinned, benchmarks run isotropic).
__device__ __forceinline__ float3_ sample_henyey_greenstein(float3_ wi, float g, float u1, float u2) {
    float ct;
    if (fabsf(g) < 1e-3f) {
        ct = fmaf(-2.0f, u1, 1.0f);
    } else {
        float q = (1.0f - g * g) / fmaf(2.0f * g, u1, 1.0f - g);
        ct = (1.0f + g * g - q * q) / (2.0f * g);
    }
    ct = cpm_clamp(ct, -1.0f, 1.0f);
    float st = sqrtf(cpm_fmax(0.0f, fmaf(-ct, ct, 1.0f)));
    float sp, cp;
    cpm_sincosf(CPM_2PI_F * u2, &sp, &cp);
    // orthonormal basis (pbrt coordinateSystem)
    float3_ v2;
    if (fabsf(wi.x) > fabsf(wi.y)) {
        float inv = 1.0f / sqrtf(fmaf(wi.x, wi.x, wi.z * wi.z));
        v2 = {-wi.z * inv, 0.0f, wi.x * inv};
    } else {
        float inv = 1.0f / sqrtf(fmaf(wi.y, wi.y, wi.z * wi.z));
        v2 = {0.0f, wi.z * inv, -wi.y * inv};
    }
    float3_ v3 = {fmaf(wi.y, v2.z, -(wi.z * v2.y)), fmaf(wi.z, v2.x, -(wi.x * v2.z)), fmaf(wi.x, v2.y, -(wi.y * v2.x))};
    float a = st * cp, b = st * sp;
    return {fmaf(a, v2.x, fmaf(b, v3.x, ct * wi.x)), fmaf(a, v2.y, fmaf(b, v3.y, ct * wi.y)),
            fmaf(a, v2.z, fmaf(b, v3.z, ct * wi.z))};
}
